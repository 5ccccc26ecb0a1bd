# developer loop on one B200: selected parity tests + per-stage timings.  usage: bash tools/gpu_dev.sh TAG "pytest -k expr" [workloads...]
TAG=${1:-dev}; KEXPR=${2:-"not quantised"}; shift 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
for wl in headline_1m "$@"; do
  timeout 300 python bench.py --quick --steps 30 --workload $wl >> gpurun_out/${TAG}_quick.json 2>> gpurun_out/${TAG}_quick.err
done
cat gpurun_out/${TAG}_quick.json | cut -c1-900
