# usage: gpu_multi.sh N  -- exchange probe + bench (quick) with each exchange variant + full bench line at N GPUs
N=$1; mkdir -p gpurun_out; OUT=gpurun_out/multi_n$N.txt; : > $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29514 tools/exchange_probe.py 2>&1 | grep -E "multimem|p2p|nccl:|Error|error" >> $OUT
for x in auto nccl p2p multimem; do timeout 300 $TR --master-port 29515 bench.py --gpus $N --quick --steps 30 --exchange $x 2>&1 | grep -E '^\{|Error|error' | cut -c1-400 >> $OUT; done
timeout 400 $TR --master-port 29516 bench.py --gpus $N --steps 30 --warmup 5 2>&1 | grep -E '^\{|Error|error' > gpurun_out/bench_n$N.json
cat $OUT; cut -c1-300 gpurun_out/bench_n$N.json
