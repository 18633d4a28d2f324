# round 2, call t: velocity-envelope calibration with the default build and with IEEE division / square root (make precise); section profile of the current kernel
set -x
timeout 900 python -m pytest tests/test_qd_envelope_gpu.py -m gpu -q -s 2>&1 | grep -vE "^\s*$|^\+" | tail -16 > gpurun_out/qd_envelope_default_r02t.txt; cat gpurun_out/qd_envelope_default_r02t.txt
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_precise.so timeout 900 python -m pytest tests/test_qd_envelope_gpu.py -m gpu -q -s 2>&1 | grep -vE "^\s*$|^\+" | tail -16 > gpurun_out/qd_envelope_precise_r02t.txt; cat gpurun_out/qd_envelope_precise_r02t.txt
DM_LIB=$PWD/deepmimic_b200/libdeepmimic_b200_prof.so timeout 300 python tools/section_profile.py > gpurun_out/section_profile_r02t.txt 2>&1; tail -26 gpurun_out/section_profile_r02t.txt | cut -c1-330
