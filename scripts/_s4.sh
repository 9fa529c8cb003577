#!/bin/bash
# one-off measurement batch (round 2, session 3)
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s4_pytest.log 2>&1
tail -3 gpurun_out/s4_pytest.log
python scripts/bench_config5.py --iters 2 > gpurun_out/s4_c5_default.json 2> gpurun_out/s4_c5_default.err
VB2_SLICE_SCATTER_CTAS=2 python scripts/bench_config5.py --iters 2 > gpurun_out/s4_c5_scatter2.json 2> gpurun_out/s4_c5_scatter2.err
VB2_SLICE_GENERIC_AGG=1 python scripts/bench_config5.py --iters 2 > gpurun_out/s4_c5_genericagg.json 2> gpurun_out/s4_c5_genericagg.err
head -c 700 gpurun_out/s4_c5_default.json; echo; head -c 300 gpurun_out/s4_c5_scatter2.json; echo; head -c 300 gpurun_out/s4_c5_genericagg.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4_c5_launches.csv python scripts/bench_config5.py --iters 0 > gpurun_out/s4_c5_ncu_launch.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on -k regex:'slice_aggregate_kernel|part_scatter_kernel|part_hist_kernel' -c 6 -f -o gpurun_out/s4_slice python scripts/bench_config5.py --rows 2.5e8 --keys 2.5e7 --iters 0 > gpurun_out/s4_c5_ncu_full.log 2>&1
ls -la gpurun_out
