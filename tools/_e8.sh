timeout 1200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2_e8_b.json 2> gpurun_out/r2_e8_b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_e8_b.json").read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), d['stage_ms'], d['info']['steps'], d['info']['nmult'], d['info']['knn_hash'], d['info']['edges_hash'], d['info']['edges_total'])
print({k:round(v['ms_total'],1) for k,v in d['kernel_ms_fine_pass'].items()})
PY
