mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_e2e.json 2>gpurun_out/bench_e2e.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e.json').readline()); print('value',round(d['value']),'e2e',d['e2e'])
PY
ncu --set full --clock-control none --import-source on -k regex:'preprocess_kernel|gauss_backward|duplicate|tile_sort_small|tile_scan' -s 18 -c 5 -o gpurun_out/r2_small python bench.py --quick --steps 2 --warmup 3 > gpurun_out/r2_small.log 2>&1
tail -2 gpurun_out/r2_small.log
