# one 8-GPU box: keyframe-batch shard at N = 8 (full line) and N = 4 (quick), config #3 at N = 8, tile-row shard at N = 8
mkdir -p gpurun_out; OUT=gpurun_out/multi_final.txt; : > $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29516 bench.py --gpus 8 --steps 30 --warmup 5 2>&1 | grep -E '^\{|Error|error' > gpurun_out/bench_n8.json
timeout 300 $TR --nproc-per-node 4 --master-port 29517 bench.py --gpus 4 --quick --steps 30 2>&1 | grep -E '^\{|Error|error' | cut -c1-500 >> $OUT
timeout 300 $TR --nproc-per-node 8 --master-port 29518 bench.py --gpus 8 --quick --steps 20 --workload cfg3_2m 2>&1 | grep -E '^\{|Error|error' | cut -c1-500 >> $OUT
timeout 300 $TR --nproc-per-node 8 --master-port 29519 bench.py --gpus 8 --shard tile_row --steps 20 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
timeout 300 $TR --nproc-per-node 8 --master-port 29520 bench.py --gpus 8 --shard tile_row --pose-only --steps 20 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
timeout 400 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --shard tile_row --pose-only --workload cfg4_5m --steps 10 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
python - <<'PY'
import json
for f in ("gpurun_out/bench_n8.json", "gpurun_out/multi_final.txt"):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: print(l.strip()[:300]); continue
        c=d.get("config",{})
        print(d.get("n_gpus"), round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms |", str(c.get("workload", d.get("exchange","")))[:50], "|", str(c.get("parallelism",""))[:110], "| e2e", d.get("e2e",{}).get("value"))
PY
