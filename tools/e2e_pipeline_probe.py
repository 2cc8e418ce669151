"""Host-side cost of the multi-GPU end-to-end entry (dist.HostStepPipeline) on ONE GPU: wall and CPU time per
timestep with spinning vs blocking slot events, plus the CPU set the box gives the process.  Round-1 result on a
1-GPU box (16 CPUs): spinning 95 us wall / 95 us CPU per timestep (PCIe-bound, one core fully busy), blocking
108-116 us wall / 70 us CPU (51 us after submit() dropped its stream context managers) -- i.e. the Python submit
path itself costs 50-70 us of CPU per timestep and rank."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    m = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    print("gpu0 nvml cpu affinity bits", sum(bin(x).count("1") for x in m), "gpus visible", pynvml.nvmlDeviceGetCount(), flush=True)
except Exception as e:
    print("nvml", e)
import brainevent_b200 as be
from brainevent_b200 import dist as bd
dev = torch.device("cuda", 0)
n, c = 1250016, 1000
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, n, (n, c), device=dev, dtype=torch.int32, generator=g)
w = torch.rand((n, c), device=dev, generator=g)
x = bd.SpikeExchange(n, device=dev)
proj = bd.PostShardedProjection("fcn", n, n, x, weights=w, indices=idx)
hin = [(torch.rand(n) < 0.01).to(torch.uint8).pin_memory() for _ in range(8)]
hout = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(4)]
for blocking in ("0", "1"):
    os.environ["BE_B200_BLOCKING_EVENTS"] = blocking
    pipe = bd.HostStepPipeline(proj, n, slots=4)
    for t in range(40):
        pipe.submit(hin[t % 8], hout[t % 4])
    pipe.wait()
    t0 = time.perf_counter(); c0 = time.process_time()
    for t in range(400):
        pipe.submit(hin[t % 8], hout[t % 4])
    pipe.wait()
    dt = time.perf_counter() - t0; ct = time.process_time() - c0
    last = (400 - 1)
    ref = be.binary_fcnmv(w, idx, hin[last % 8].to(dev), shape=(n, n), transpose=True).cpu()
    assert torch.allclose(hout[last % 4], ref, rtol=1e-5, atol=2e-5), "pipeline output differs from the direct call"
    print(f"blocking={blocking}: {dt / 400 * 1e6:.1f} us/timestep wall, {ct / 400 * 1e6:.1f} us/timestep CPU", flush=True)
    pipe.close()
