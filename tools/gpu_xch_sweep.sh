# exchange kernel sweep at N GPUs (tune build): bash tools/gpu_xch_sweep.sh N
N=$1; mkdir -p gpurun_out; OUT=gpurun_out/xch_sweep_n$N.txt; : > $OUT
export GSB_LIB=$PWD/gsorb_slam_b200/libgsb_tune.so
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for c in 1 2 4 6; do for u in 2 4 8; do
  echo "== ctas_per_sm=$c unroll=$u" >> $OUT
  GSB_XCH_CTAS_PER_SM=$c GSB_XCH_UNROLL=$u timeout 200 $TR --master-port 29514 tools/exchange_probe.py 2>&1 | grep -E "^multimem|^p2p:|^nccl:|rror" >> $OUT
done; done
cat $OUT
