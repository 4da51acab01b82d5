#!/bin/bash
# One B200: the records kept under profiles/ for round 2.  usage: gpurun -- bash tools/run_n1_final.sh <tag>
T=${1:-r2f}
timeout 900 python bench.py > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err
timeout 300 python bench.py --workload c1 --steps 5 --warmup 3 > gpurun_out/${T}_c1.json 2> gpurun_out/${T}_c1.err
timeout 300 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/${T}_c2.json 2> gpurun_out/${T}_c2.err
timeout 600 python benchmarks/c4_knn_snn.py --ks 10,20,50,100 > gpurun_out/${T}_c4_n1.jsonl 2> gpurun_out/${T}_c4_n1.err
# launch list of one pass (after the same command ran clean above), then one full capture of the dominant kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/${T}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 1 -o gpurun_out/${T}_prof_knn -f \
    python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/${T}_ncu_knn.log 2>&1
for f in n1 c1 c2; do tail -c 300 gpurun_out/${T}_$f.json; echo; done
