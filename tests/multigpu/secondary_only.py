import sys, json
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import bench
print(json.dumps(bench.secondary_viterbi(0)))
