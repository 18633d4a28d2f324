set -x
timeout 600 python tools/two_groups.py 2>&1 | tail -4
