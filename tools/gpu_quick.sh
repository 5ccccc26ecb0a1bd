# developer loop: parity tests + per-stage timings (+ optional knob sweep given as "VAR=a,b,c" arguments)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/q_pytest.log
python bench.py --quick --steps 30 > gpurun_out/q_quick.json 2>&1
for spec in "$@"; do
  var=${spec%%=*}; vals=${spec#*=}
  for v in ${vals//,/ }; do env $var=$v python bench.py --quick --steps 30 >> gpurun_out/q_quick.json 2>&1; done
done
cat gpurun_out/q_pytest.log; cat gpurun_out/q_quick.json
