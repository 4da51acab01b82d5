timeout 900 python bench.py > gpurun_out/r2_b13.json 2> gpurun_out/r2_b13.err
tail -2 gpurun_out/r2_b13.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_b13.json").read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), d['value'], d['e2e'], d['cpu_baseline'], d['clocks'], d['roofline']['frac'])
PY
