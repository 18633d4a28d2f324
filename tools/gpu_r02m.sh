# round 2, call m: ABA child gather through shared scratch + U in registers: parity tests, bench; policy network v2 (warp-specialised bulk-TMA pipeline): tests + timing
set -x
timeout 900 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_mlp_gpu.py 2>&1 | tail -6
B() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH $1', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'])"; }
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>gpurun_out/bench_r02m.err | B cur
timeout 300 python tools/mlp_time.py 2>&1 | tail -3
timeout 600 python -m pytest tests/test_mlp_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -15
