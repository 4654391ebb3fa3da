#!/bin/bash
# every variant in its own process (a faulting descriptor must not take the others down), bounded by timeout
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
out=../../gpurun_out/tma_probe.txt
: > $out
for v in "kk 64 64 3" "kk 64 64 4" "kk 64 96 4" "km 64 32 1" "km 64 32 1 1" "km 64 64 1" "km 64 64 3" "km 96 64 4" \
         "mm 64 32 1" "mm 64 32 1 1" "mm 64 64 1" "mm 64 64 3" "mm 96 64 4" "mm 192 64 4" "ts 192 64 4"; do
  timeout 30 ./tma_probe $v >> $out 2>&1 || echo "variant $v: rc=$?" >> $out
done
cat $out
