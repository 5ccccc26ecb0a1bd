mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1_pytest.log
python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/r1_bench_ref.json 2> gpurun_out/r1_bench_ref.err
python bench.py --steps 50 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --quick --steps 30 > gpurun_out/r1_quick.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r1_launches.csv python bench.py --quick --steps 4 --warmup 3 > gpurun_out/r1_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'blend_|tile_sort|preprocess_kernel|gauss_backward|duplicate' -s 30 -c 7 -o gpurun_out/r1_prof python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r1_ncu_full.log 2>&1
tail -3 gpurun_out/r1_pytest.log; cat gpurun_out/r1_bench_ref.json gpurun_out/r1_bench.json gpurun_out/r1_quick.json | cut -c1-1500
