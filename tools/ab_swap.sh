for r in 1 2; do for v in 0 1; do echo "== swap=$v"; MARL_B200_WGRAD_SWAP=$v timeout 300 python tools/step_times.py 2s3z 3s5z 27m_vs_30m 2>&1 | tail -n 3; done; done
