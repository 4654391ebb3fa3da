for r in 1 2; do
echo "== unset"; timeout 300 python tools/step_times.py 2s3z 2>&1 | tail -n 1
for v in 256 64; do echo "== swap rows=$v"; MARL_B200_WGRAD_SWAP_ROWS=$v timeout 300 python tools/step_times.py 2s3z 2>&1 | tail -n 1; done; done
MARL_B200_WGRAD_SWAP_ROWS=64 timeout 300 python tools/step_times.py 2s3z 3s5z 27m_vs_30m 2>&1 | tail -n 3
MARL_B200_WGRAD_SWAP_ROWS=64 timeout 100 python tools/timeline.py 2>&1 | tail -8
