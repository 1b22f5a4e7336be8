"""KPConv encoder stack section of bench.py under torchrun (DDP):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 bench_tools/stack_ddp.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = bench.bench_kpconv_stack(dev, world, rank)
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
