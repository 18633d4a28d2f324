# round 2, call l: policy-network timing, ncu metric pass + full capture + launch list of the current step kernel (blocked PGS)
set -x
timeout 300 python tools/mlp_time.py 2>&1 | tail -3
M=$(python -c "import tools.ncu_metrics_json as m; print(m.METRICS)")
timeout 600 ncu --metrics $M --clock-control none -k regex:dm_step_kernel -s 56 -c 4 --csv --log-file gpurun_out/step_metrics_r02l.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_metrics_r02l.log 2>&1
python tools/ncu_metrics_json.py gpurun_out/step_metrics_r02l.csv humanoid3d 4096 20 "ncu r02l: bench.py --steps 8 --warmup 4, launches 56-59 of dm_step_kernel<16,0,0>" | tail -20
cp profiles/step_metrics_humanoid3d.json gpurun_out/step_metrics_humanoid3d_r02l.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_step_kernel -s 56 -c 1 -o gpurun_out/prof_step_r02l -f python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_full_r02l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 60 --csv --log-file gpurun_out/launches_r02l.csv python bench.py --steps 4 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_list_r02l.log 2>&1
ls -la gpurun_out | tail -4
