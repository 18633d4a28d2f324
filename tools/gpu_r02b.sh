# round 2, call b: full tracebacks of the opt-in device tests + the new parity2 tests
set -x
export DM_EXPERIMENTAL_TASK_SCENES=1 DM_EXPERIMENTAL_ROOT_ROT_SYNC=1 DM_RUN_UNVALIDATED_GPU_TESTS=1
timeout 600 python -m pytest tests/test_unvalidated_gpu.py -m gpu -q --tb=short -k "not fixture_task_policies" 2>&1 | tail -250
unset DM_EXPERIMENTAL_TASK_SCENES DM_EXPERIMENTAL_ROOT_ROT_SYNC DM_RUN_UNVALIDATED_GPU_TESTS
timeout 900 python -m pytest tests/test_parity2_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -150
