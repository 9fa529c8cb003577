#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s5_pytest.log 2>&1
tail -3 gpurun_out/s5_pytest.log
python scripts/bench_config5.py --iters 3 > gpurun_out/s5_c5_default.json 2> gpurun_out/s5_c5_default.err
VB2_SLICE_SCATTER_CTAS=3 python scripts/bench_config5.py --iters 3 > gpurun_out/s5_c5_scatter3.json 2> gpurun_out/s5_c5_scatter3.err
head -c 400 gpurun_out/s5_c5_default.json; echo; head -c 300 gpurun_out/s5_c5_scatter3.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s5_c5_launches.csv python scripts/bench_config5.py --iters 0 > gpurun_out/s5_c5_ncu_launch.log 2>&1
