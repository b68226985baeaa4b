"""Raw device->host ceiling of the box: N processes (one per GPU, torchrun) copy from HBM into page-locked host memory at the
same time — no library code, plain torch pinned tensors and cudaMemcpyAsync. Establishes what `e2e` (bench.py: six targets back
to host buffers, 12.29 MB per 640x480 frame) can reach at N GPUs.

  python tools/d2h_probe.py                                   (one GPU)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/d2h_probe.py

Rank 0 prints one JSON line: per-rank GB/s (copies of all ranks overlapping in time), the aggregate, and the same with each
rank copying ALONE (ranks take turns), so contention on the host side shows up as the difference."""
import json
import os
import time

import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = 256
src = torch.empty(MB << 20, dtype=torch.uint8, device="cuda").random_(0, 255)
dst = torch.empty(MB << 20, dtype=torch.uint8).pin_memory()
REPS = 12


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def copy_gbs():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return REPS * (MB << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9


dst.copy_(src, non_blocking=True)
barrier()
together = copy_gbs()                      # every rank at once
barrier()
alone = 0.0
for r in range(world):                     # ranks take turns
    if r == rank:
        alone = copy_gbs()
    barrier()
vals = torch.tensor([together, alone], device="cuda", dtype=torch.float64)
if world > 1:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    tg = [float(v[0]) for v in allv]
    al = [float(v[1]) for v in allv]
    frame_mb = 640 * 480 * 40 / 1e6
    print(json.dumps({"n_gpus": world, "copy_mb": MB, "reps": REPS, "together_gbs_per_rank": [round(x, 2) for x in tg], "together_gbs_total": round(sum(tg), 2),
                      "alone_gbs_per_rank": [round(x, 2) for x in al], "e2e_ceiling_frames_per_s_640x480_six_targets": round(sum(tg) * 1e3 / frame_mb, 1),
                      "cpu_count": os.cpu_count(), "time": time.strftime("%Y-%m-%d %H:%M:%S")}))
if world > 1:
    dist.destroy_process_group()
