# Round-2 evidence run (one B200): tests, bench (both arms), step times, timeline, ncu launch lists, ncu --set full captures.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err; tail -c 300 gpurun_out/r2_bench_reference_arm.json
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench_1gpu.json
timeout 300 python tools/step_times.py 2>&1 | tail -6 > gpurun_out/r2_step_times.txt; cat gpurun_out/r2_step_times.txt
MARL_B200_TGEMM=1 timeout 300 python tools/step_times.py 2s3z 3s5z 27m_vs_30m 2>&1 | tail -3 > gpurun_out/r2_step_times_tgemm.txt; cat gpurun_out/r2_step_times_tgemm.txt
MARL_B200_DETERMINISTIC=1 timeout 300 python tools/step_times.py 2s3z 3s5z 27m_vs_30m 2>&1 | tail -3 > gpurun_out/r2_step_times_deterministic.txt; cat gpurun_out/r2_step_times_deterministic.txt
timeout 100 python tools/timeline.py > gpurun_out/r2_timeline.txt 2>&1
timeout 200 python tools/gemm_check.py > gpurun_out/r2_gemm_linear.txt 2>&1
MARL_B200_TGEMM=1 timeout 200 python tools/gemm_check.py > gpurun_out/r2_gemm_tgemm.txt 2>&1
timeout 60 tools/micro/mma_rate 148 > gpurun_out/r2_mma_rate.txt 2>&1
timeout 100 bash tools/micro/run_tma_probe.sh > /dev/null 2>&1; cp gpurun_out/tma_probe.txt gpurun_out/r2_tma_probe.txt
timeout 100 python tools/early_exit_bench.py > gpurun_out/r2_early_exit.txt 2>&1
timeout 100 python tools/replay_overhead.py > gpurun_out/r2_replay_overhead.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/r2_launches_warm.csv python tools/prof_step.py qmix 3 > gpurun_out/ncu_list2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"qmix_mix_kernel|gru_unroll|linear_fwd|linear_wgrad" -s 8 -c 8 -o gpurun_out/r2_full python tools/prof_step.py qmix 3 > gpurun_out/ncu_full.log 2>&1
MARL_B200_TGEMM=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tgemm_kernel -s 4 -c 4 -o gpurun_out/r2_full_tgemm python tools/prof_step.py qmix 2 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -30
