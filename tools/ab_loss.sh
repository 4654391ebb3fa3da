for r in 1 2 3; do for v in 0 1; do echo "== loss zero-copy=$v"; MARL_B200_LOSS_ZEROCOPY=$v timeout 300 python tools/step_times.py 2s3z 2>&1 | tail -n 1; done; done
