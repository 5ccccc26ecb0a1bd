# ncu --set full of selected kernels in the quick bench.  usage: bash tools/gpu_ncu.sh TAG 'regex' [count]
TAG=${1:-ncu}; RE=${2:-blend_}; CNT=${3:-2}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 6 -c $CNT -o gpurun_out/${TAG}_prof -f python bench.py --quick --steps 2 --warmup 3 > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_prof*
