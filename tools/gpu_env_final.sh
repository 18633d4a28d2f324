set -x
timeout 900 python -m pytest tests/test_qd_envelope_gpu.py -m gpu -q -s 2>&1 | grep -vE "^\s*$|^\+" | tail -16 > gpurun_out/qd_envelope_final_r02.txt; cat gpurun_out/qd_envelope_final_r02.txt
timeout 300 python tools/parity_stats.py args/run_humanoid3d_spinkick_args.txt 200 2>&1 | tail -6
