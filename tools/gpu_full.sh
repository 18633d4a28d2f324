# the round-end sequence as the driver runs it: full GPU suite, smoke(), bench (b200 arm and reference arm)
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x -s 2>&1 | grep -E "passed|failed|rollout|actor step|Error|error|assert" | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 64 --warmup 4 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_full.json').read()); print('BENCH', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | cut -c1-600
