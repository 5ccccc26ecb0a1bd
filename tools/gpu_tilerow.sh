# usage: gpu_tilerow.sh N WORKLOAD...  -- tile-row shard (strong scaling) at N GPUs, mapping and pose-only exchange
N=$1; shift; mkdir -p gpurun_out; OUT=gpurun_out/tilerow_n$N.txt; : > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for w in "$@"; do
  python bench.py --shard tile_row --workload $w --steps 20 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
  timeout 300 $TR --master-port 29521 bench.py --gpus $N --shard tile_row --workload $w --steps 20 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
  python bench.py --shard tile_row --pose-only --workload $w --steps 20 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
  timeout 300 $TR --master-port 29522 bench.py --gpus $N --shard tile_row --pose-only --workload $w --steps 20 --warmup 3 2>&1 | grep -E '^\{|Error|error' >> $OUT
done
python - <<PY
import json
for l in open("$OUT"):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d["n_gpus"], round(d["value"],1), "fps", round(d["ms_per_step"],3), "ms |", d["config"]["workload"][:40], "|", d["config"]["parallelism"][:140])
PY
