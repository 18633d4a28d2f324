# round 2, call n: evidence for profiles/ of the current step kernel (section profile, ncu metric pass -> step_metrics json, full-set capture, launch list),
# bench lines of the other BASELINE.json configurations (walk, dog3d trot, target_amp)
set -x
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_prof.so timeout 300 python tools/section_profile.py > gpurun_out/section_profile_r02n.txt 2>&1; tail -34 gpurun_out/section_profile_r02n.txt
M=$(python -c "import tools.ncu_metrics_json as m; print(m.METRICS)")
timeout 600 ncu --metrics $M --clock-control none -k regex:dm_step_kernel -s 56 -c 4 --csv --log-file gpurun_out/step_metrics_r02n.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_metrics_r02n.log 2>&1
python tools/ncu_metrics_json.py gpurun_out/step_metrics_r02n.csv humanoid3d 4096 20 "ncu r02n: bench.py --steps 8 --warmup 4, launches 56-59 of dm_step_kernel<16,0,0>" | tail -20
cp profiles/step_metrics_humanoid3d.json gpurun_out/step_metrics_humanoid3d_r02n.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_step_kernel -s 56 -c 1 -o gpurun_out/prof_step_r02n -f python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full_r02n.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 60 --csv --log-file gpurun_out/launches_r02n.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_list_r02n.log 2>&1
timeout 300 python bench.py --steps 128 > gpurun_out/bench_humanoid_r02n.json 2> gpurun_out/bench_r02n.err
for f in train_humanoid3d_walk_args.txt train_dog3d_trot_args.txt train_amp_target_humanoid3d_locomotion_args.txt; do
  timeout 400 python bench.py --steps 96 --arg-file args/$f > gpurun_out/bench_${f%_args.txt}_r02n.json 2>> gpurun_out/bench_r02n.err
done
python - <<'PY'
import json,glob
for p in sorted(glob.glob('gpurun_out/bench_*_r02n.json')):
    try:
        d=json.loads(open(p).read()); print(p.split('/')[-1], round(d['value']), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value']), 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print(p, 'ERR', e)
PY
tail -5 gpurun_out/bench_r02n.err
