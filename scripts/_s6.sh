#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests/test_slice_agg_gpu.py tests/test_operators_gpu.py -m gpu -x -q) > gpurun_out/s6_pytest.log 2>&1
tail -3 gpurun_out/s6_pytest.log
python scripts/bench_config5.py --iters 3 > gpurun_out/s6_c5_default.json 2> gpurun_out/s6_c5_default.err
head -c 400 gpurun_out/s6_c5_default.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s6_c5_launches.csv python scripts/bench_config5.py --iters 0 > gpurun_out/s6_c5_ncu_launch.log 2>&1
