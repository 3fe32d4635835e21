"""Executed-instruction / stall-sample profile of a kernel by SASS address region:
python tools/ncu_regions.py rep [kernel-substring] [chunk]"""
import csv
import subprocess
import sys
from collections import Counter


def main(rep, which=None, chunk=64):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    cur, hdr, body = None, None, []
    for r in rows:
        if r and r[0] == 'Kernel Name':
            if body and (not which or which in cur):
                break
            cur, body = r[1], []
        elif r and r[0] == 'Address':
            hdr = r
        elif r and r[0].startswith('0x'):
            body.append(r)
    si, ei, ni = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot_e = sum(int(r[ei] or 0) for r in body); tot_s = sum(int(r[ni] or 0) for r in body)
    print(cur, 'instructions', len(body), 'executed', tot_e, 'samples', tot_s)
    for c in range(0, len(body), chunk):
        part = body[c:c + chunk]
        e = sum(int(r[ei] or 0) for r in part); s = sum(int(r[ni] or 0) for r in part)
        ops = Counter()
        for r in part:
            t = r[si].split()
            op = (t[1] if t and t[0].startswith('@') else (t[0] if t else '?')).split('.')[0]
            ops[op] += int(r[ei] or 0)
        top = ', '.join(f'{k}:{v * 100 // max(e, 1)}%' for k, v in ops.most_common(5))
        print(f'{c:5d}-{c + len(part):5d}  exec {100 * e / tot_e:5.1f}%  samples {100 * s / max(tot_s, 1):5.1f}%   {top}')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 64)
