#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_multigpu_gpu.py tests/test_multi_driver_gpu.py -m gpu -x -q) > gpurun_out/m2_pytest.log 2>&1
tail -5 gpurun_out/m2_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551"
$TR scripts/bench_config5_multi.py --iters 3 > gpurun_out/m2_c5_nccl.json 2> gpurun_out/m2_c5_nccl.err
tail -c 900 gpurun_out/m2_c5_nccl.json; echo
VB2_EXCHANGE_SEGMENT_MB=4200 $TR scripts/bench_config5_multi.py --iters 3 > gpurun_out/m2_c5_p2p.json 2> gpurun_out/m2_c5_p2p.err
tail -c 900 gpurun_out/m2_c5_p2p.json; echo
tail -5 gpurun_out/m2_c5_nccl.err gpurun_out/m2_c5_p2p.err
