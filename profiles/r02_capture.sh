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
# --- end of round 2 (single-utterance persistent kernel, selection kernels, final lines)
python tests/multigpu/probe_single.py                      # phase split (SNK_G1_TIMING) + per-CTA scan-end skew, A/B against the batched path
SNK_G1_NO_BALANCE=1 python tests/multigpu/probe_single.py  # the same with equal row slices
compute-sanitizer --tool memcheck  python -m pytest tests/test_gpu_greedy_one.py -x -q -k "equals_batched or small_and_odd or halfphone_epoch or certificate_failure"
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_greedy_one.py -x -q -k "small_and_odd or halfphone_epoch"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kth_of_rows|select_segments|rerank_kernel|knn_tc_kernel" -s 12 -c 36 --csv --log-file gpurun_out/r02_launches_k50_final.csv python bench.py --workload halfphone --no-cpu
bash tests/multigpu/sweep_k50.sh                            # sampling stride / bound slack of the k = 50 search
python bench.py > gpurun_out/bench_r2g.json                 # -> profiles/r02_bench_n1_final.json
# gpurun --gpus 2 / 8:
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3   # -> r02_bench_n2_final.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3   # -> r02_bench_n8_final.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tests/multigpu/probe_sharded_greedy.py   # -> r02_sharded_greedy_n8_single_utterance.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tests/multigpu/run_sharded_greedy.py      # paths / distances equal the replicated search
