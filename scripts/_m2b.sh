#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_fused_gpu.py -m gpu -q -k "partition" > gpurun_out/m2b_pytest.log 2>&1
tail -3 gpurun_out/m2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551"
VB2_EXCHANGE_SEGMENT_MB=4200 $TR scripts/bench_config5_multi.py --iters 3 > gpurun_out/m2b_c5_p2p.json 2> gpurun_out/m2b_c5_p2p.err
tail -c 1000 gpurun_out/m2b_c5_p2p.json; echo
tail -n 5 gpurun_out/m2b_c5_p2p.err
