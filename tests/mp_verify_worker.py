"""torchrun worker of tests/test_multigpu_torchrun.py: bench.verify_launch under one process per GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
v = bench.verify_launch(dist, rank, world, local)
# a second handle in the same processes: the cached communicator is reused, results must not change
v2 = bench.verify_launch(dist, rank, world, local)
# Phase A in place (AA pattern) across the slabs
v3 = bench.verify_launch(dist, rank, world, local, in_place=True)
if rank == 0:
    v["second_run_ok"] = bool(v2["ok"])
    v["in_place_ok"] = bool(v3["ok"])
    v["ok"] = bool(v["ok"] and v2["ok"] and v3["ok"])
    print(json.dumps(v))
dist.barrier()
dist.destroy_process_group()
