# round 2, call k: sectioned blocked PGS (B = 2 default; 1, 3, 4 as A/B libraries), pair-lane A build for > W rows: parity tests, bench, section profile
set -x
timeout 900 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_mlp_gpu.py 2>&1 | tail -12
B() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH $1', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'])"; }
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>gpurun_out/bench_r02k.err | B b2
for v in b1 b3 b4; do DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_$v.so timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>>gpurun_out/bench_r02k.err | B $v; done
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_prof.so timeout 300 python tools/section_profile.py 2>&1 | tail -40
