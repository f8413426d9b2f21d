"""One real create_proof per circuit shape (bench.py's real_flow section alone): python scripts/real_flow.py [names...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import halo2_vectordb_b200 as h
import bench
h.init(0)
names = sys.argv[1:] or ["distances_k13", "query_k13", "kmeans_k16"]
print(json.dumps(bench.bench_real_flow(h, torch, names), indent=1))
