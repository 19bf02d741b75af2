"""Host-side copy ceiling of the box with ALL ranks copying at once (the limit of the end-to-end
number at N = 8: every image crosses host memory and PCIe, nothing else is shared between ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29512 tools/host_ceiling.py --out profiles/r02_host_ceiling_nN.json

Per direction (H2D, D2H, both at once): every rank copies 64 MiB pinned buffers back to back for
`--reps` copies between two barriers; per-rank GB/s from CUDA events, aggregate = sum."""
import argparse
import json
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=40)
ap.add_argument("--out", default=None)
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nb = 64 << 20
hx, hy = (torch.empty(nb, dtype=torch.uint8).pin_memory() for _ in range(2))
dx, dy = (torch.empty(nb, dtype=torch.uint8, device="cuda") for _ in range(2))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h):
    for _ in range(3):
        if h2d:
            dx.copy_(hx, non_blocking=True)
        if d2h:
            hy.copy_(dy, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(cur)
    s2.wait_stream(cur)
    for _ in range(a.reps):
        if h2d:
            with torch.cuda.stream(s1):
                dx.copy_(hx, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
    cur.wait_stream(s1)
    cur.wait_stream(s2)
    e1.record()
    torch.cuda.synchronize()
    gbs = a.reps * nb / 1e9 / (e0.elapsed_time(e1) / 1e3)       # per direction
    t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
    if world > 1:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [float(p.item()) for p in parts]
    return [gbs]


out = {"n_gpus": world, "bytes_per_copy": nb, "reps": a.reps, "host_cores": os.cpu_count()}
for name, h2d, d2h in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
    per = run(h2d, d2h)
    out[name] = {"per_rank_gbs_each_direction": [round(v, 2) for v in per],
                 "aggregate_gbs_each_direction": round(sum(per), 1),
                 "aggregate_gbs_total": round(sum(per) * (2 if (h2d and d2h) else 1), 1)}
if rank == 0:
    print(json.dumps(out))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(out, open(a.out, "w"), indent=1)
if world > 1:
    dist.destroy_process_group()
