# A/B of two builds of the library: default vs DM_LIB variant(s) given as arguments (names of deepmimic_b200/libdeepmimic_b200_<name>.so)
set -x
timeout 900 python -m pytest tests -m gpu -q --tb=short -x --deselect tests/test_mlp_gpu.py 2>&1 | tail -3
B() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH $1', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'])"; }
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>gpurun_out/bench_ab.err | B default
for v in "$@"; do
  DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_$v.so timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_facade_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -2
  DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_$v.so timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>>gpurun_out/bench_ab.err | B $v
done
timeout 300 python bench.py --steps 128 --no-cpu-baseline 2>>gpurun_out/bench_ab.err | B default_again
