"""Quick A/B for the CSR SpMM config (BASELINE configs[2]) at several spike rates."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brainevent_b200 as be
from brainevent_b200 import _lib as L
import bench

dev = torch.device("cuda", 0)
peak, src = bench.peaks()
for rate in (0.002, 0.01, 0.05):
    wl = bench.CsrSpmm(torch, be, dev, rate=rate)
    r = bench.measure_workload(torch, None, L, wl, 3, 3, 4, 1, peak, src)
    print(json.dumps({"rate": rate, "events_per_s": r["value"], "ms_per_call": r["ms_per_step"] / 4,
                      "kernel_ms": r["roofline"]["kernel_ms_avg"], "GBs": r["roofline"]["achieved"]}))
    # gather direction on the same matrix (what a CSC-stored matrix would run): csr @ B
    del wl
    torch.cuda.empty_cache()

# gather direction: csr @ B (same traversal a CSC mirror of the transposed product would run)
import time
n, nnz, batch = 1_000_000, 100, 64
g = torch.Generator(device=dev).manual_seed(7)
indptr = torch.arange(0, (n + 1) * nnz, nnz, device=dev, dtype=torch.int32)
indices = torch.randint(0, n, (n * nnz,), device=dev, dtype=torch.int32, generator=g)
weights = torch.randn(n * nnz, device=dev, generator=g)
perm = torch.randperm(n * nnz, device=dev, generator=g).to(torch.int32)
for rate in (0.002, 0.01, 0.05):
    B = torch.rand((n, batch), device=dev, generator=g) < rate
    for name, fn in (("gather", lambda: be.binary_csrmm(weights, indices, indptr, B, shape=(n, n), transpose=False)),
                     ("gather_indexed", lambda: be.binary_csrmm_indexed(weights, indices, indptr, perm, B, shape=(n, n), transpose=False))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ev = float(B.sum().item()) * nnz
        print(json.dumps({"kernel": name, "rate": rate, "ms_per_call": ms, "events_per_s": ev / ms * 1e3}))
