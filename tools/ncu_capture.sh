# usage: bash tools/ncu_capture.sh <tag>   (on the GPU box; writes gpurun_out/prof_update_<tag>.ncu-rep and launches_<tag>.csv)
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dm_step_kernel -s 3 -c 1 -o gpurun_out/prof_update_$TAG -f python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_list_$TAG.log 2>&1
ls -la gpurun_out | tail -5
