# round 2, call i: mlp test, section profile, ncu metric pass (-> profiles/step_metrics_humanoid3d.json), full-set capture, launch list
set -x
timeout 600 python -m pytest tests/test_mlp_gpu.py -m gpu -q --tb=short -s -x 2>&1 | tail -8
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_prof.so timeout 300 python tools/section_profile.py 2>&1 | tail -40
M=$(python -c "import tools.ncu_metrics_json as m; print(m.METRICS)")
timeout 600 ncu --metrics $M --clock-control none -k regex:dm_step_kernel -s 56 -c 4 --csv --log-file gpurun_out/step_metrics_r02i.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_metrics_r02i.log 2>&1
python tools/ncu_metrics_json.py gpurun_out/step_metrics_r02i.csv humanoid3d 4096 20 "ncu r02i: bench.py --steps 8 --warmup 4, launches 56-59 of dm_step_kernel<16,0,0>" | tail -20
cp profiles/step_metrics_humanoid3d.json gpurun_out/step_metrics_humanoid3d_r02i.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_step_kernel -s 56 -c 1 -o gpurun_out/prof_step_r02i -f python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full_r02i.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 60 --csv --log-file gpurun_out/launches_r02i.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_list_r02i.log 2>&1
ls -la gpurun_out | tail -6
