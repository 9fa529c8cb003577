#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29561"
$TR bench.py --gpus 4 --steps 5 --warmup 3 --skip-e2e --skip-cpu > gpurun_out/m4_bench.json 2> gpurun_out/m4_bench.err
tail -c 1500 gpurun_out/m4_bench.json; echo
tail -n 5 gpurun_out/m4_bench.err
