# occupancy / barrier experiments through environment variables only (no code change): kernel_ms of bench.py --steps 96
set -x
run() { env "$@" timeout 300 python bench.py --steps 96 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('VARIANT', '$*', 'value', round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],4), 'e2e', round(d['e2e']['value']), 'overflow', d['config']['solver_row_overflows'])"; }
run A=base
run DM_MAX_ROWS=33
run DM_MAX_ROWS=33 DM_TILES_PER_BLOCK=14
run DM_MAX_ROWS=33 DM_TILES_PER_BLOCK=14 DM_SYNC_EVERY_STAGE=-1000
run DM_MAX_ROWS=33 DM_TILES_PER_BLOCK=14 DM_SYNC_EVERY_STAGE=1
run DM_MAX_ROWS=33 DM_TILES_PER_BLOCK=14 DM_SYNC_EVERY_STAGE=-4
run DM_SYNC_EVERY_STAGE=-1000
