"""Matvec micro-benchmarks on the 1M x 1M, 100 nnz/row CSR (BASELINE configs[2] structure):
scatter (spk @ csr) at several rates, gather (csr @ spk), homo/hetero; achieved algorithmic GB/s."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brainevent_b200 as be

dev = "cuda"
n, nnz = 1_000_000, 100
g = torch.Generator(device=dev).manual_seed(7)
indptr = torch.arange(0, (n + 1) * nnz, nnz, device=dev, dtype=torch.int32)
indices = torch.randint(0, n, (n * nnz,), device=dev, dtype=torch.int32, generator=g)
weights = torch.randn(n * nnz, device=dev, generator=g)
w1 = torch.tensor([0.5], device=dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for rate in (0.01, 0.05, 0.2):
    pool = [torch.rand(n, device=dev, generator=g) < rate for _ in range(8)]
    it = iter(range(10**9))
    ev = float(sum(p.sum().item() for p in pool)) / len(pool) * nnz
    for name, w, bpe in (("scatter_hetero", weights, 8), ("scatter_homo", w1, 4)):
        ms = timeit(lambda: be.binary_csrmv(w, indices, indptr, pool[next(it) % 8], shape=(n, n), transpose=True))
        print(json.dumps({"op": name, "rate": rate, "ms": round(ms, 4), "events_per_s": ev / ms * 1e3,
                          "alg_GBs": round(ev * bpe / ms / 1e6, 1)}))
    for name, w, bytes_ in (("gather_hetero", weights, 8 * n * nnz), ("gather_homo", w1, 4 * n * nnz)):
        ms = timeit(lambda: be.binary_csrmv(w, indices, indptr, pool[next(it) % 8], shape=(n, n), transpose=False), 5)
        print(json.dumps({"op": name, "rate": rate, "ms": round(ms, 4), "events_per_s": ev / ms * 1e3,
                          "alg_GBs_full_matrix": round(bytes_ / ms / 1e6, 1),
                          "note": "hetero reads a weight quad only when one of its 4 sources fired"}))
