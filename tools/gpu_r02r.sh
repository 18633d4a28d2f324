# round 2, call r: rollout on the environment's stream (mlp tests, timing tool), ncu metric pass of the dog3d kernel -> profiles/step_metrics_dog3d.json
set -x
timeout 600 python -m pytest tests/test_mlp_gpu.py tests/test_facade_gpu.py -m gpu -q --tb=short -s 2>&1 | grep -E "passed|failed|rollout|Error|error|assert" | tail -8
timeout 600 python tools/rollout_time.py 2>&1 | tail -4
M=$(python -c "import tools.ncu_metrics_json as m; print(m.METRICS)")
timeout 600 ncu --metrics $M --clock-control none -k regex:dm_step_kernel -s 56 -c 4 --csv --log-file gpurun_out/step_metrics_dog_r02r.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline --arg-file args/train_dog3d_trot_args.txt > gpurun_out/ncu_metrics_dog_r02r.log 2>&1
python tools/ncu_metrics_json.py gpurun_out/step_metrics_dog_r02r.csv dog3d 2048 20 "ncu r02r: bench.py --arg-file args/train_dog3d_trot_args.txt --steps 8 --warmup 4, launches 56-59 of dm_step_kernel<32,0,0>" | tail -16
cp profiles/step_metrics_dog3d.json gpurun_out/step_metrics_dog3d_r02r.json
