for s in 4 6 8 12; do
  SNK_TC_SAMPLE=$s python bench.py --workload halfphone --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); h=d['halfphone']; print('sample $s', round(h['ms_per_step'],3), round(h['knn_roofline']['ms'],3), h['exactness'], h['parity']['mismatches'])"
done
for k in 0 3 25; do
  SNK_TC_KSLACK=$k python bench.py --workload halfphone --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); h=d['halfphone']; print('kslack $k', round(h['ms_per_step'],3), round(h['knn_roofline']['ms'],3), h['exactness'])"
done
