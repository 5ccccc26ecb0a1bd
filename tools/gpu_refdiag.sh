mkdir -p gpurun_out; : > gpurun_out/refdiag.txt
python bench.py --impl reference --steps 10 --warmup 3 >> gpurun_out/refdiag.txt 2>&1
OMP_NUM_THREADS=1 python bench.py --impl reference --steps 10 --warmup 3 >> gpurun_out/refdiag.txt 2>&1
python bench.py --impl reference --steps 30 --warmup 5 >> gpurun_out/refdiag.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 30 --warmup 5 >> gpurun_out/refdiag.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/refdiag.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['steps'], d['warmup'], round(d['value'],1), round(d['ms_per_step'],3), d.get('clocks'))
PY
