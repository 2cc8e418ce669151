"""Host-side probe for the end-to-end leg at N > 1 (run under torchrun): where do the ranks, their pinned buffers and
their GPUs sit, and how much host<->device bandwidth does a rank get alone versus all ranks together?  Explains the
e2e numbers of bench.py at N = 8 (aggregate host traffic fell from 172 GB/s at N = 4 to 120 GB/s at N = 8).
Prints one JSON line on rank 0."""
import glob
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def status(key):
    for line in open("/proc/self/status"):
        if line.startswith(key):
            return line.split(":", 1)[1].strip()
    return None


def gpu_numa():
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else bdf).lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        p = f"/sys/bus/pci/devices/{bdf}/numa_node"
        return open(p).read().strip() if os.path.exists(p) else "?"
    except Exception as e:
        return f"? ({type(e).__name__})"


N_IN, N_OUT = 1_250_016, 1_250_016 * 4          # bytes per timestep and rank of bench.py's e2e leg
h_in = torch.empty(N_IN, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N_OUT, dtype=torch.uint8).pin_memory()
d_in = torch.empty(N_IN, dtype=torch.uint8, device=dev)
d_out = torch.empty(N_OUT, dtype=torch.uint8, device=dev)
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def copy_loop(reps, both=True):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
        if both:
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def barrier():
    if world > 1:
        dist.barrier()


copy_loop(20)
alone = []
for r in range(world):            # one rank at a time
    barrier()
    if r == rank:
        alone.append(copy_loop(200))
    barrier()
barrier()
together = copy_loop(400)         # everybody at once
barrier()
d2h_only = copy_loop(400, both=False)
res = {"rank": rank, "cpus_allowed": status("Cpus_allowed_list"), "mems_allowed": status("Mems_allowed_list"),
       "gpu_numa_node": gpu_numa(), "alone_us": round(alone[0] * 1e6, 1), "together_us": round(together * 1e6, 1),
       "together_d2h_only_us": round(d2h_only * 1e6, 1)}
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    tot = (N_IN + N_OUT) * world
    print(json.dumps({"world": world, "host_numa_nodes": len(nodes), "cpu_count": os.cpu_count(),
                      "bytes_per_timestep_per_rank": N_IN + N_OUT,
                      "aggregate_GBps_together": round(tot / max(r["together_us"] for r in out) / 1e3, 1),
                      "per_rank_GBps_alone": [round((N_IN + N_OUT) / r["alone_us"] / 1e3, 1) for r in out],
                      "ranks": out}), flush=True)
if world > 1:
    dist.destroy_process_group()
