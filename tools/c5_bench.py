"""Only the cfg5 sub-record of bench.py (the landmark front end on a batch): python tools/c5_bench.py [bones] [steps]"""
import json, sys
sys.path.insert(0, ".")
import bench
g = bench.Gpu(0, 1, 0)
bones = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
print(json.dumps(bench.landmark_record(g, bones, steps)))
