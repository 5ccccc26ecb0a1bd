# one 8-GPU box, round-2 kernels: keyframe-batch shard at N = 8 (full line) and N = 4 (full line), tile-row shard (pose-only) of config #4 at N = 1 and 8
mkdir -p gpurun_out; OUT=gpurun_out/r02t_multi.txt; : > $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 --steps 30 --warmup 5 2>gpurun_out/r02t_n8.err | grep -E '^\{' > gpurun_out/r02t_bench_n8.json
timeout 400 $TR --nproc-per-node 4 --master-port 29517 bench.py --gpus 4 --steps 30 --warmup 5 2>gpurun_out/r02t_n4.err | grep -E '^\{' > gpurun_out/r02t_bench_n4.json
timeout 300 python bench.py --gpus 1 --shard tile_row --pose-only --workload cfg4_5m --steps 10 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
timeout 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --shard tile_row --pose-only --workload cfg4_5m --steps 10 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
python - <<'PY'
import json
for f in ("gpurun_out/r02t_bench_n8.json", "gpurun_out/r02t_bench_n4.json", "gpurun_out/r02t_multi.txt"):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: print(l.strip()[:300]); continue
        c=d.get("config",{})
        print(d.get("n_gpus"), round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms |", str(c.get("workload", ""))[:40], "|", str(c.get("parallelism",""))[:90], "| e2e", d.get("e2e",{}).get("value"), "| xchk", (d.get("exchange_checked") or {}).get("ok_on_all_ranks"))
PY
tail -3 gpurun_out/r02t_n8.err
