# knob sweep on the tuning build: bash tools/gpu_sweep.sh TAG VAR=a,b,c [VAR2=...]   (each value: one quick bench line)
TAG=${1:-sweep}; shift
mkdir -p gpurun_out
export GSB_LIB=$PWD/gsorb_slam_b200/libgsb_tune.so
for spec in "$@"; do
  var=${spec%%=*}; vals=${spec#*=}
  for v in ${vals//,/ }; do
    echo "== $var=$v" >> gpurun_out/${TAG}_sweep.txt
    env $var=$v timeout 300 python bench.py --quick --steps 30 2>/dev/null | cut -c1-400 >> gpurun_out/${TAG}_sweep.txt
  done
done
cat gpurun_out/${TAG}_sweep.txt
