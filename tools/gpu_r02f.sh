# round 2, call y: evidence of the final step kernel for profiles/ (ncu metric pass -> step_metrics json for both characters, full-set capture, launch list, section profile)
set -x
M=$(python -c "import tools.ncu_metrics_json as m; print(m.METRICS)")
timeout 600 ncu --metrics $M --clock-control none -k regex:dm_step_kernel -s 56 -c 4 --csv --log-file gpurun_out/step_metrics_r02f.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_metrics_r02f.log 2>&1
python tools/ncu_metrics_json.py gpurun_out/step_metrics_r02f.csv humanoid3d 4096 20 "ncu r02f: bench.py --steps 8 --warmup 4, launches 56-59 of dm_step_kernel<16,0,0>" | tail -16
cp profiles/step_metrics_humanoid3d.json gpurun_out/step_metrics_humanoid3d_r02f.json
timeout 600 ncu --metrics $M --clock-control none -k regex:dm_step_kernel -s 56 -c 4 --csv --log-file gpurun_out/step_metrics_dog_r02f.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline --arg-file args/train_dog3d_trot_args.txt > gpurun_out/ncu_metrics_dog_r02f.log 2>&1
python tools/ncu_metrics_json.py gpurun_out/step_metrics_dog_r02f.csv dog3d 2048 20 "ncu r02f: bench.py --arg-file args/train_dog3d_trot_args.txt --steps 8 --warmup 4, launches 56-59 of dm_step_kernel<32,0,0>" | tail -4
cp profiles/step_metrics_dog3d.json gpurun_out/step_metrics_dog3d_r02f.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_step_kernel -s 56 -c 1 -o gpurun_out/prof_step_r02f -f python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full_r02f.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 60 --csv --log-file gpurun_out/launches_r02f.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_list_r02f.log 2>&1
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_prof.so timeout 300 python tools/section_profile.py > gpurun_out/section_profile_r02f.txt 2>&1; tail -22 gpurun_out/section_profile_r02f.txt | cut -c1-320
