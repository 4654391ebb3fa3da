# Round-2 (end) evidence run on one B200: tests, bench (both arms), step times, timeline, phase traces, ncu launch lists and
# ncu --set full captures of the kernels that changed (fused input layers, MN-major weight gradients, one-row BPTT CTAs).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2c_gpu_tests.txt; cat gpurun_out/r2c_gpu_tests.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2c_bench_reference_arm.json 2> gpurun_out/r2c_bench_ref.err; tail -c 300 gpurun_out/r2c_bench_reference_arm.json
timeout 700 python bench.py --steps 200 --warmup 20 > gpurun_out/r2c_bench_1gpu.json 2> gpurun_out/r2c_bench.err; tail -c 300 gpurun_out/r2c_bench_1gpu.json
timeout 300 python tools/step_times.py 2>&1 | tail -6 > gpurun_out/r2c_step_times.txt; cat gpurun_out/r2c_step_times.txt
timeout 100 python tools/timeline.py > gpurun_out/r2c_timeline.txt 2>&1
timeout 100 python tools/replay_overhead.py --pinned > gpurun_out/r2c_replay_overhead.txt 2>&1
timeout 100 python tools/replay_overhead.py >> gpurun_out/r2c_replay_overhead.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2c_launches.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/r2c_launches_warm.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"agent_front_kernel|qmix_mix_kernel|gru_unroll|linear_wgrad|linear_dgrad" -s 9 -c 9 -o gpurun_out/r2c_full python tools/prof_step.py qmix 3 > gpurun_out/ncu_full.log 2>&1
timeout 200 python tools/early_exit_bench.py > gpurun_out/r2c_early_exit.txt 2>&1
MARL_EARLY=1 timeout 100 python tools/config_kernel_times.py 2s3z > gpurun_out/r2c_cfg2_early_kernels.txt 2>&1
timeout 100 python tools/host_gap.py > gpurun_out/r2c_host_gap.txt 2>&1
ls -la gpurun_out | tail -30
