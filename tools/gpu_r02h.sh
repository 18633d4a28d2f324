# round 2, call h: context-refactored step kernel (no struct parameters) + first run of the tcgen05 policy network
set -x
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>gpurun_out/bench_r02h.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'])"
timeout 600 python -m pytest tests/test_mlp_gpu.py -m gpu -q --tb=short -s -x 2>&1 | tail -40
timeout 1500 python -m pytest tests -m gpu -q --tb=short --deselect tests/test_mlp_gpu.py 2>&1 | tail -12
