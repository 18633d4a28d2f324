# quick check of a step-kernel change: parity suite (without the mlp tests) + humanoid / dog bench
set -x
timeout 900 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_mlp_gpu.py 2>&1 | tail -6
B() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH $1', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'], d['config']['solver_row_overflows'])"; }
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>gpurun_out/bench_quick.err | B humanoid
timeout 300 python bench.py --steps 96 --no-cpu-baseline --arg-file args/train_dog3d_trot_args.txt 2>>gpurun_out/bench_quick.err | B dog
