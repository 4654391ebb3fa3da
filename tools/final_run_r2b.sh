# Round-2 (second half) evidence run on one B200: tests, bench (both arms), step times, timeline, phase traces, ncu launch lists and
# ncu --set full captures of the kernels that changed (fused input layers, MN-major weight gradients, one-row BPTT CTAs).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r2b_gpu_tests.txt; cat gpurun_out/r2b_gpu_tests.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2b_bench_reference_arm.json 2> gpurun_out/r2b_bench_ref.err; tail -c 300 gpurun_out/r2b_bench_reference_arm.json
timeout 700 python bench.py --steps 200 --warmup 20 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench.err; tail -c 300 gpurun_out/r2b_bench_1gpu.json
timeout 300 python tools/step_times.py 2>&1 | tail -6 > gpurun_out/r2b_step_times.txt; cat gpurun_out/r2b_step_times.txt
MARL_B200_FRONT=0 timeout 300 python tools/step_times.py 2s3z 3s5z 27m_vs_30m 2>&1 | tail -3 > gpurun_out/r2b_step_times_unfused_front.txt; cat gpurun_out/r2b_step_times_unfused_front.txt
timeout 100 python tools/timeline.py > gpurun_out/r2b_timeline.txt 2>&1
timeout 100 python tools/front_trace.py 400 > gpurun_out/r2b_front_trace.txt 2>&1
timeout 200 python tools/gemm_check.py > gpurun_out/r2b_gemm_linear.txt 2>&1
timeout 60 tools/micro/mma_rate2 148 > gpurun_out/r2b_mma_rate2.txt 2>&1
timeout 100 bash tools/micro/run_tma_probe.sh > /dev/null 2>&1; cp gpurun_out/tma_probe.txt gpurun_out/r2b_tma_probe.txt
timeout 100 python tools/replay_overhead.py --pinned > gpurun_out/r2b_replay_overhead.txt 2>&1
timeout 100 python tools/replay_overhead.py >> gpurun_out/r2b_replay_overhead.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/r2b_launches_warm.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list2.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"agent_front_kernel|qmix_mix_kernel|gru_unroll|linear_wgrad|linear_dgrad" -s 9 -c 9 -o gpurun_out/r2b_full python tools/prof_step.py qmix 3 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -30
