set -x
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 500 python bench.py --steps 200 --warmup 20 > gpurun_out/r1d_bench_1gpu.json 2> gpurun_out/r1d_bench.err; tail -c 600 gpurun_out/r1d_bench_1gpu.json
timeout 300 python tools/step_times.py 2>&1 | tail -6 > gpurun_out/r1d_step_times.txt; cat gpurun_out/r1d_step_times.txt
timeout 100 python tools/timeline.py > gpurun_out/r1d_timeline.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1d_launches.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r1d_launches_warm.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"qmix_mix_kernel|vdn_td_fused|gru_unroll" -s 4 -c 3 -o gpurun_out/r1d_full python tools/prof_step.py qmix 3 > gpurun_out/ncu_full.log 2>&1
