#!/bin/bash
# One 8-GPU box: C3 bench line, C5 (4M cells) bench line, C4 kNN+SNN k sweep.  usage: gpurun --gpus 8 -- bash tools/run_n8.sh <tag>
T=${1:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/${T}_n8_c3.json 2> gpurun_out/${T}_n8_c3.err
timeout 900 $TR bench.py --gpus 8 --steps 2 --warmup 2 --workload c5 --no-cpu-baseline > gpurun_out/${T}_n8_c5.json 2> gpurun_out/${T}_n8_c5.err
timeout 900 $TR benchmarks/c4_knn_snn.py --ks 10,20,50,100 > gpurun_out/${T}_c4_n8.jsonl 2> gpurun_out/${T}_c4_n8.err
tail -c 600 gpurun_out/${T}_n8_c3.json; echo; tail -c 400 gpurun_out/${T}_n8_c5.json; echo; tail -n 4 gpurun_out/${T}_c4_n8.jsonl | cut -c1-300
