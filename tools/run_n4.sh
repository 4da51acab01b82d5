#!/bin/bash
T=${1:-r2g}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n4.json 2> gpurun_out/${T}_n4.err
tail -c 300 gpurun_out/${T}_n4.json
