# bench lines of the four BASELINE.json configurations with the final kernel
set -x
for f in train_humanoid3d_spinkick_args.txt train_humanoid3d_walk_args.txt train_dog3d_trot_args.txt train_amp_target_humanoid3d_locomotion_args.txt; do
  timeout 400 python bench.py --steps 96 --arg-file args/$f > gpurun_out/bench_${f%_args.txt}_r02final.json 2>> gpurun_out/bench_r02final.err
done
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/bench_*_r02final.json')):
    try:
        d=json.loads(open(p).read()); print(p.split('/')[-1], round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value']), 'cpu', round(d.get('cpu_baseline',{}).get('value',0)))
    except Exception as e: print(p, 'ERR', e)
PY
