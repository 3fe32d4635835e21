#!/bin/bash
# C2 device-timed pass for every library variant under socialmediamodel_b200/variants/
cp socialmediamodel_b200/libsmm_b200.so /tmp/libsmm_orig.so
for v in socialmediamodel_b200/variants/libsmm_*.so; do
  cp $v socialmediamodel_b200/libsmm_b200.so
  echo -n "$(basename $v) "; python bench.py --no-extra --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'])"
done
cp /tmp/libsmm_orig.so socialmediamodel_b200/libsmm_b200.so
