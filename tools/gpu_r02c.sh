# round 2, call c: new parity tests, target-scene policy traceback, regular suite, smoke, new bench lines (spinkick / walk / dog)
set -x
timeout 900 python -m pytest tests/test_parity2_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -60
export DM_EXPERIMENTAL_TASK_SCENES=1 DM_EXPERIMENTAL_ROOT_ROT_SYNC=1 DM_RUN_UNVALIDATED_GPU_TESTS=1
timeout 600 python -m pytest tests/test_unvalidated_gpu.py -m gpu -q --tb=short -s -k "not strike and not getup and not args2 and not args3" 2>&1 | tail -80
unset DM_EXPERIMENTAL_TASK_SCENES DM_EXPERIMENTAL_ROOT_ROT_SYNC DM_RUN_UNVALIDATED_GPU_TESTS
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -c 3000 gpurun_out/bench_r02c.json; tail -5 gpurun_out/bench_r02c.err
timeout 300 python bench.py --preroll 0 --episode-seconds 0.5 --no-cpu-baseline > gpurun_out/bench_r02c_short.json 2>> gpurun_out/bench_r02c.err; tail -c 1500 gpurun_out/bench_r02c_short.json
timeout 300 python bench.py --arg-file args/train_humanoid3d_walk_args.txt --steps 128 --no-cpu-baseline > gpurun_out/bench_walk_r02c.json 2>> gpurun_out/bench_r02c.err; tail -c 1500 gpurun_out/bench_walk_r02c.json
timeout 300 python bench.py --arg-file args/train_dog3d_trot_args.txt --steps 128 --no-cpu-baseline > gpurun_out/bench_dog_r02c.json 2>> gpurun_out/bench_r02c.err; tail -c 1500 gpurun_out/bench_dog_r02c.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 | tail -c 1500
