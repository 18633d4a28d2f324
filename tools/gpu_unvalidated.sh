# First GPU call of the next round: run the device code written after round 1's GPU budget was spent (AMP task scenes,
# --sync_char_root_rot, SetSampleCount / clip-end tests), then the regular suite, smoke and a bench line for regression.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_unvalidated.sh > gpurun_out/unvalidated.log 2>&1'
set -x
export DM_EXPERIMENTAL_TASK_SCENES=1 DM_EXPERIMENTAL_ROOT_ROT_SYNC=1 DM_RUN_UNVALIDATED_GPU_TESTS=1
timeout 600 python -m pytest tests/test_timer_anneal_gpu.py tests/test_unvalidated_gpu.py -m gpu -q 2>&1 | tail -40
unset DM_EXPERIMENTAL_TASK_SCENES DM_EXPERIMENTAL_ROOT_ROT_SYNC DM_RUN_UNVALIDATED_GPU_TESTS
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
timeout 300 python bench.py 2>&1 | tail -2
# A/B of the address-space hints: same bench through the library built without them
