"""Aggregate the SASS page of an .ncu-rep by opcode: python tools_ncu_sass.py rep [kernel-name-substring]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main(rep, which=None, top=22):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = dict(name=r[1], rows=[]); kernels.append(cur)
        elif r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and r and r[0].startswith('0x'):
            cur['rows'].append(r)
    for k in kernels:
        if which and which not in k['name']:
            continue
        h = k['hdr']; si = h.index('Source'); ei = h.index('Instructions Executed'); ni = h.index('# Samples')
        ops = defaultdict(lambda: [0, 0])
        tot_e = tot_s = 0
        for r in k['rows']:
            toks = r[si].split()
            op = toks[1] if toks and toks[0].startswith('@') else (toks[0] if toks else '?')
            op = op.rstrip(';')
            e = int(r[ei] or 0); s = int(r[ni] or 0)
            ops[op][0] += e; ops[op][1] += s; tot_e += e; tot_s += s
        print(k['name'][:100], 'SASS lines', len(k['rows']), 'warp-instr', tot_e, 'samples', tot_s)
        for op, (e, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f'   {op:28s} exec {e:12d} ({100*e/max(tot_e,1):5.1f}%)  samples {s:8d} ({100*s/max(tot_s,1):5.1f}%)')
        hot = sorted(k['rows'], key=lambda r: -int(r[ni] or 0))[:10]
        print('   hottest SASS lines by stall samples:')
        for r in hot:
            print(f'     {int(r[ni] or 0):7d}  {r[si].strip()[:90]}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
