#!/bin/bash
run() { python tools/profile_workloads.py pop_long 2>/dev/null | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read().strip()); print('$1', 'ms/step', round(d['ms_per_step'],4), d.get('gather_roofline'))"; }
SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_head.so run head
run new
SMM_B200_LIB=$PWD/socialmediamodel_b200/libsmm_b200_head.so run head
run new
