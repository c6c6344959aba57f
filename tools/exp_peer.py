"""Dev experiment (2 GPUs, torchrun): is the peer-read grouping phase slowed by the remote reads themselves or by both
ranks hammering each other's L2 at the same time?  Times sibgpu_dist_group_peer alone on each rank, then concurrently."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sibelia_b200 as sb
from sibelia_b200 import synth

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
ndev = torch.cuda.device_count()
dev = int(os.environ["LOCAL_RANK"]) % ndev
torch.cuda.set_device(dev)
if world <= ndev:
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
else:
    dist.init_process_group("gloo")                  # ranks share a GPU: timing only meaningful in the "alone" column
ctx = sb.Context(dev)
MB = int(os.environ.get("EXP_MBASES", "100"))
chrs = [synth.random_genome(MB * 1_000_000, 12345 + c) for c in range(world)]
ctx.dist_upload(chrs, rank, world)
nparts, cnt, cap, ovf = ctx.dist_scatter_local(25)
handle = ctx.dist_export_send()
mine = np.concatenate([cnt.astype(np.int64), np.array([cap, int(ovf)], dtype=np.int64), handle.view(np.int64)])
cdev = "cuda" if dist.get_backend() == "nccl" else "cpu"
allm = torch.empty(world * len(mine), dtype=torch.int64, device=cdev)
dist.all_gather_into_tensor(allm, torch.from_numpy(mine).to(cdev))
allm = allm.cpu().numpy().reshape(world, len(mine))
ctx.dist_import_peers(np.ascontiguousarray(allm[:, nparts + 2:]).view(np.uint8).reshape(world, 64))
counts, caps = allm[:, :nparts].astype(np.uint64), allm[:, nparts].astype(np.uint64)


def timed():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ctx.dist_group_peer(counts, caps)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3


for it in range(3):
    dist.barrier()
    both = timed()
    dist.barrier()
    alone = []
    for r in range(world):
        if r == rank:
            alone.append(timed())
        dist.barrier()
    print("rank %d iter %d: concurrent %.2f ms, alone %.2f ms" % (rank, it, both, alone[0]), flush=True)
ctx.set_profiling(True)
for r in range(world):
    if r == rank:
        timed()
        print("rank %d kernel stats:" % rank, [(x["name"], x["launches"], round(x["ms"], 3)) for x in ctx.kernel_stats()], flush=True)
    dist.barrier()
dist.barrier()
dist.destroy_process_group()
