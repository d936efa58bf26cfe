"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): runs tests/multi_gpu_checks.py."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import multi_gpu_checks  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
res = multi_gpu_checks.run_all(dev, rank, world)
if rank == 0:
    print("multi-gpu worker ok: world", world, res)
dist.destroy_process_group()
