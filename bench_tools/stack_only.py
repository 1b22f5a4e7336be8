"""KPConv encoder stack section of bench.py on its own:  python bench_tools/stack_only.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print(json.dumps(bench.bench_kpconv_stack(torch.device("cuda", 0), 1, 0), indent=1))
