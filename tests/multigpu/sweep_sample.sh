#!/bin/bash
# k = 50 search of the halfphone workload for several sampling strides (SNK_TC_SAMPLE): bench lines into gpurun_out/
for s in 8 4 2 1; do
  SNK_TC_SAMPLE=$s python bench.py --workload halfphone --no-cpu > gpurun_out/bench_hp_sample$s.json 2> gpurun_out/bench_hp_sample$s.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_hp_sample$s.json") if l.startswith("{")][0])
h=d["halfphone"]; print("SAMPLE=$s knn ms", round(h["knn_roofline"]["ms"],3), "gemm ms", round(h["knn_roofline"]["gemm_launch_ms"],3), "pipeline ms", round(h["ms_per_step"],3), h["exactness"], h["parity"]["mismatches_outside_1e-6_tie"])
PY
done
