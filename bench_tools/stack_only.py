"""run only the KPConv stack section of bench.py (for ncu launch lists)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
print(json.dumps(bench.bench_kpconv_stack(torch.device("cuda", 0), 1, 0)))
