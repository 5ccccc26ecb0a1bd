# mapping / tracking iteration capture on one B200: launch list, ncu --set full of the per-Gaussian tail kernels and the loss kernels,
# memcheck over their parity tests.   usage: bash tools/gpu_slam_iter.sh TAG
TAG=${1:-r02z}; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_slam_launches.csv python tools/slam_step_probe.py 4 4 > /dev/null 2>&1
python tools/summarize.py launches gpurun_out/${TAG}_slam_launches.csv | grep gsb > gpurun_out/${TAG}_slam_launches.txt
ncu --set full --clock-control none --import-source on -k regex:'map_update|pose_gradient|loss_stats|loss_grad|tracking_loss' -s 6 -c 7 -o gpurun_out/${TAG}_slam_prof -f python tools/slam_step_probe.py 3 2 > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_slam_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_slam_prof_raw.csv 2>/dev/null
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "fused_map_update or pose_only or grows_the_binning" > gpurun_out/${TAG}_slam_memcheck.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "fused_map_update and 1.0 or pose_only" > gpurun_out/${TAG}_slam_racecheck.log 2>&1
cat gpurun_out/${TAG}_slam_launches.txt; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_slam_memcheck.log gpurun_out/${TAG}_slam_racecheck.log
