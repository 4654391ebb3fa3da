#!/bin/bash
# A/B: step times (and optionally GEMM timings) for each library variant:  ab_run.sh "cfgs" variant...   ("base" = the default build)
cfgs="$1"; shift
for v in "$@"; do
  if [ "$v" = base ]; then unset MARL_B200_LIB; else export MARL_B200_LIB=$PWD/marl_b200/lib/ab/libmarl_$v.so; fi
  echo "== $v"
  timeout 300 python tools/step_times.py $cfgs 2>&1 | tail -n $(echo $cfgs | wc -w)
  if [ -n "$AB_GEMM" ]; then timeout 200 python tools/gemm_check.py 2>&1 | head -5; fi
done
