# round 2, call u: z-loop change (parity suite), barrier policy A/B (DM_SYNC_EVERY_STAGE: 0 = every update (default), -2 / -4 = every 2 / 4 updates, -1000 = never, 1 = every stage)
set -x
timeout 900 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_mlp_gpu.py 2>&1 | tail -4
B() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH $1', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'])"; }
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>gpurun_out/bench_u.err | B sync_every_update
for m in -2 -4 -1000 1; do DM_SYNC_EVERY_STAGE=$m timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>>gpurun_out/bench_u.err | B sync_mode_$m; done
