# one full measurement round on one B200: parity tests, both bench arms, launch list, ncu --set full of every hot kernel
TAG=${1:-r}; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
python bench.py --steps 50 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --quick --steps 30 > gpurun_out/${TAG}_quick.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --quick --steps 4 --warmup 3 > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'blend_|tile_sort_small|preprocess_kernel|gauss_backward|duplicate|tile_scan' -s 27 -c 9 -o gpurun_out/${TAG}_prof python bench.py --quick --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cut -c1-600 gpurun_out/${TAG}_bench_ref.json gpurun_out/${TAG}_bench.json gpurun_out/${TAG}_quick.json
