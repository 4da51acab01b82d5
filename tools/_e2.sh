set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_e2_tests.log 2>&1
tail -4 gpurun_out/r2_e2_tests.log
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --workload c5 --cells 200000"
timeout 300 $B > gpurun_out/r2_e2_c5a.json 2> gpurun_out/r2_e2_c5a.err
AB_KNN_MT2=112 timeout 300 $B > gpurun_out/r2_e2_c5b.json 2> gpurun_out/r2_e2_c5b.err
for f in c5a c5b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_e2_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d['ms_per_step'],1), d['kernel_ms']['knn_main'], d['kernel_ms']['knn_rerank'], d['stage_ms'], d['info']['knn_stats'], d['info']['knn_hash'], d['info']['edges_hash'])
except Exception as e: print("$f", "ERR", e)
PY
done
tail -3 gpurun_out/r2_e2_c5b.err
