# round 2, call g: tests, bench, section profile, ncu captures of the current kernel
set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -12
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_r02g.json 2> gpurun_out/bench_r02g.err; python -c "import json; d=json.loads(open('gpurun_out/bench_r02g.json').read()); print('BENCH', round(d['value']), d['roofline']['kernel_ms'], round(d['e2e']['value']), d['config']['step_ms'])"
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_prof.so timeout 300 python tools/section_profile.py 2>&1 | tail -34
M=$(python -c "import tools.ncu_metrics_json as m; print(m.METRICS)")
timeout 600 ncu --metrics $M --clock-control none -k regex:dm_step_kernel -s 56 -c 4 --csv --log-file gpurun_out/step_metrics_r02g.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_metrics_r02g.log 2>&1
python tools/ncu_metrics_json.py gpurun_out/step_metrics_r02g.csv humanoid3d 4096 20 "ncu r02g: bench.py --steps 8 --warmup 4, launches 56-59 of dm_step_kernel<16,0,0>" | tail -20
cp profiles/step_metrics_humanoid3d.json gpurun_out/step_metrics_humanoid3d_r02g.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_step_kernel -s 56 -c 1 -o gpurun_out/prof_step_r02g -f python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full_r02g.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 60 --csv --log-file gpurun_out/launches_r02g.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_list_r02g.log 2>&1
ls -la gpurun_out | tail -6
