# round 2, call p: dog3d row capacity 52 (14 environments per block: one wave), rollout without the per-step host sync: full GPU suite, dog + humanoid bench
set -x
timeout 1200 python -m pytest tests -m gpu -q --tb=short -x -s 2>&1 | grep -E "passed|failed|rollout|Error|error|assert" | tail -12
B() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH $1', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'], d['config']['solver_row_overflows'])"; }
timeout 300 python bench.py --steps 96 --no-cpu-baseline --arg-file args/train_dog3d_trot_args.txt 2>gpurun_out/bench_r02p.err | B dog
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>>gpurun_out/bench_r02p.err | B humanoid
