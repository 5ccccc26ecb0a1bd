# usage: gpu_sweep.sh "ENV1=v ENV2=v" "ENV1=v" ...   (one bench --quick run per argument; "" = defaults)
mkdir -p gpurun_out; : > gpurun_out/sweep.json
for spec in "$@"; do env $spec python bench.py --quick --steps 30 >> gpurun_out/sweep.json 2>&1; done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.json'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(round(d['value']), d['env'], {k:v for k,v in d['stages_us'].items()})
PY
