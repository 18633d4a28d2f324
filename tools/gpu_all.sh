set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -5
timeout 300 python bench.py 2>&1 | tail -3
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -2
