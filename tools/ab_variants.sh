#!/bin/bash
# Timing experiments on the tensor-core adjoint (dev tool): build variants of libcmcd_b200.so with phases of
# bridge_bwd_tc_kernel compiled out (-DBT_X_*; their RESULTS ARE WRONG) and time the kernel on the bench workload, to see which
# phase sits on the per-warp critical path.  Usage: tools/ab_variants.sh build | run [particles]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
V=$ROOT/cmcd_b200/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
NAMES="base NOG1 NOREDUCE NOSPLIT NOSTAGE NOSCORE NOGELU ALL"
if [ "$1" = build ]; then
  mkdir -p $V
  OBJS=$(ls $ROOT/cmcd_b200/build/*.o | grep -v bridge_bwd_tc.o)
  for n in $NAMES; do
    case $n in
      base) D="" ;;
      ALL) D="-DBT_X_NOG1 -DBT_X_NOREDUCE -DBT_X_NOSPLIT -DBT_X_NOSTAGE -DBT_X_NOSCORE -DBT_X_NOGELU" ;;
      *) D="-DBT_X_$n" ;;
    esac
    nvcc $FLAGS $D -c $ROOT/cmcd_b200/csrc/bridge_bwd_tc.cu -o $V/bwd_tc_$n.o 2>/dev/null &
  done
  wait
  for n in $NAMES; do nvcc -shared -o $V/lib$n.so $OBJS $V/bwd_tc_$n.o -lcudart; done
  ls -la $V/*.so
else
  for n in $NAMES; do
    CMCD_B200_LIB=$V/lib$n.so python $ROOT/bench.py --particles-global ${2:-262144} --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); b=d['roofline']['step_breakdown_ms']; print('$n', 'step', round(d['ms_per_step'],2), 'bwd', round(b['bwd_kernel'],2), 'fwd', round(b['fwd_kernel'],2))"
  done
fi
