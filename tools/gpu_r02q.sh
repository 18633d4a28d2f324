set -x
timeout 600 python tools/rollout_time.py 2>&1 | tail -5
timeout 300 python tools/mlp_time.py 2>&1 | tail -2
