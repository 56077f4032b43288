#!/bin/bash
# A/B the build variants under cmcd_b200/variants on the bench workload (dev tool)
for v in "" A B C; do
  if [ -z "$v" ]; then unset CMCD_B200_LIB; name=base; else export CMCD_B200_LIB=$PWD/cmcd_b200/variants/lib$v.so; name=$v; fi
  python bench.py --particles-per-gpu ${1:-262144} --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'ms', round(d['ms_per_step'],1), 'bwd', round(d['roofline']['avg_launch_ms'],1), 'fwd', round(d['roofline']['fwd_kernel']['avg_launch_ms'],1))"
done
