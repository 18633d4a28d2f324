# round 2, call f: whole default GPU suite on the new kernel (padding frozen, register-resident PGS, un-gated task scenes), smoke, bench lines, variants
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short -s 2>&1 | grep -v "^$" | tail -70
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
run() { env "$@" timeout 300 python bench.py --steps 96 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('VARIANT', '$*', 'value', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), 'overflow', d['config']['solver_row_overflows'], 'step_ms', d['config']['step_ms'])"; }
run A=base
run DM_MAX_ROWS=33
run DM_MAX_ROWS=33 DM_TILES_PER_BLOCK=14
run DM_MAX_ROWS=33 DM_TILES_PER_BLOCK=14 DM_SYNC_EVERY_STAGE=-1000
run DM_MAX_ROWS=33 DM_TILES_PER_BLOCK=14 DM_SYNC_EVERY_STAGE=1
run DM_SYNC_EVERY_STAGE=-1000
run DM_SYNC_EVERY_STAGE=1
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_prof.so timeout 300 python tools/section_profile.py 2>&1 | tail -34
timeout 300 python bench.py > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; tail -c 2500 gpurun_out/bench_r02f.json; tail -3 gpurun_out/bench_r02f.err
for f in train_humanoid3d_walk train_dog3d_trot train_amp_target_humanoid3d_locomotion; do timeout 300 python bench.py --arg-file args/${f}_args.txt --steps 128 --no-cpu-baseline > gpurun_out/bench_${f}_r02f.json 2>> gpurun_out/bench_r02f.err; python -c "import sys,json; d=json.loads(open('gpurun_out/bench_${f}_r02f.json').read()); print('LINE', d['metric'], round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']))"; done
