#!/bin/bash
T=${1:-r2f}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err
tail -c 400 gpurun_out/${T}_n2.json
