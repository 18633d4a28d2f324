set -x
timeout 300 python bench.py --steps 20 --warmup 4 --no-cpu-baseline 2>gpurun_out/bench_only.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BENCH', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['clocks'], d['roofline']['issue_slot_pct_of_peak'], d['roofline'].get('fp32_tflops'))"
tail -3 gpurun_out/bench_only.err
