# one full measurement round on one B200 (round 2): parity tests, both bench arms, launch list, ncu --set full of every hot kernel,
# compute-sanitizer over the small parity tests, staging-engine A/B.   usage: bash tools/gpu_round2.sh TAG
TAG=${1:-r02}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --quick --steps 30 > gpurun_out/${TAG}_quick.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --quick --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'blend_|tile_sort_small|preprocess_kernel|gauss_backward|duplicate' -s 24 -c 6 -o gpurun_out/${TAG}_prof -f python bench.py --quick --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
# staging engine A/B on the backward (tune build): LDGSTS ring (default) vs one 48-byte cp.async.bulk (UBLKCP) per record
for v in ldgsts bulk; do echo "== GSB_BLEND_STAGE=$v" >> gpurun_out/${TAG}_stage_ab.txt; GSB_LIB=$PWD/gsorb_slam_b200/libgsb_tune.so GSB_BLEND_STAGE=$v python bench.py --quick --steps 30 2>/dev/null | cut -c1-330 >> gpurun_out/${TAG}_stage_ab.txt; done
# sanitizers over the small parity cases (every kernel of the frame, all staging / barrier / atomic paths)
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_golden_gpu.py -m gpu -x -q -k "small_case and tiny_default or ragged or ties and 6000 or tile_sort and one_plane or blended_pair or alpha_decisions" > gpurun_out/${TAG}_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_golden_gpu.py -m gpu -x -q -k "small_case and tiny_default and True" > gpurun_out/${TAG}_racecheck.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cut -c1-300 gpurun_out/${TAG}_bench_ref.json gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_quick.json; cat gpurun_out/${TAG}_stage_ab.txt | cut -c1-250
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log gpurun_out/${TAG}_racecheck.log
