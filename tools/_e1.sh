set -x
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
AB_KNN_NPROD=3 timeout 300 $B > gpurun_out/r2_e1_np3.json 2> gpurun_out/r2_e1_np3.err
AB_KNN_NPROD=2 timeout 300 $B > gpurun_out/r2_e1_np2.json 2> gpurun_out/r2_e1_np2.err
AB_KNN_NPROD=2 AB_KNN_M=64 timeout 300 $B > gpurun_out/r2_e1_np2m64.json 2> gpurun_out/r2_e1_np2m64.err
AB_KNN_NPROD=2 timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_scale.py -m gpu -x -q -k "knn or c2 or frontend" > gpurun_out/r2_e1_tests.log 2>&1
tail -5 gpurun_out/r2_e1_tests.log
for f in np3 np2 np2m64; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_e1_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d['ms_per_step'],1), d['kernel_ms']['knn_main'], d['kernel_ms']['knn_rerank'], d['stage_ms']['knn'], d['info']['knn_stats'], d['info']['knn_hash'], d['info']['edges_hash'])
except Exception as e: print("$f", "ERR", e)
PY
done
