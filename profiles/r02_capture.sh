#!/bin/bash
# round-2 profile captures (one GPU)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_launches_default.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:join_tile_tc -s 2 -c 1 -o gpurun_out/r02_join_tc -f python bench.py --workload halfphone --no-cpu > gpurun_out/r02_ncu_join.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:knn_tc_kernel -s 10 -c 2 -o gpurun_out/r02_knn_k50 -f python bench.py --workload halfphone --no-cpu > gpurun_out/r02_ncu_k50.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"rerank_kernel|select_segments|kth_of_rows|viterbi_kernel" -s 12 -c 4 -o gpurun_out/r02_small -f python bench.py --workload halfphone --no-cpu > gpurun_out/r02_ncu_small.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:knn_tc_kernel -s 30 -c 1 -o gpurun_out/r02_knn_greedy -f python bench.py --steps 1 --warmup 1 --no-cpu --no-secondary > gpurun_out/r02_ncu_greedy.log 2>&1
python bench.py > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
tail -c 300 gpurun_out/bench_r2d.err
PROBE_NCU=1 ncu --set full --clock-control none --import-source on -k regex:greedy_one -s 1 -c 1 -f -o gpurun_out/r02_greedy_one python tests/multigpu/probe_single.py > gpurun_out/r02_ncu_greedy_one.log 2>&1
