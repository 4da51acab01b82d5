#!/bin/bash
T=${1:-r2i}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${T}_n8_c3.json 2> gpurun_out/${T}_n8_c3.err
tail -c 300 gpurun_out/${T}_n8_c3.json
