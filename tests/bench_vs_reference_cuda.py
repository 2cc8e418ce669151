"""Head-to-head on one B200: the reference's OWN CUDA kernels (unmodified sources, recompiled for
sm_100a into oracle/_ref by `make -C oracle/ref_cuda`) against brainevent_b200's kernels, same device
buffers, same spike vectors, both replayed from a CUDA graph of POOL calls so host launch overhead
is out of the picture.  Prints one JSON line per case: per-call time of each side, the ratio, and the
largest output difference (a GPU-side parity check against the reference itself).

    python tests/bench_vs_reference_cuda.py [case ...]     cases: fcn csr spmm jits jitn jitu dense compact
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import brainevent_b200 as be
from oracle import ref_cuda as R

dev = torch.device("cuda", 0)
POOL = 8


def graph_ms(fn, reps=5):
    """fn(i) enqueues call i; capture POOL calls, replay `reps` times, return ms per call."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for i in range(2):
            fn(i)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for i in range(POOL):
            fn(i)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * POOL)


def report(case, events, ours_fn, ref_fn, check, reps=5, **extra):
    a, b = check()
    err = (a.double() - b.double()).abs().max().item()
    scale = b.double().abs().max().item()
    ms_o = graph_ms(ours_fn, reps)
    ms_r = graph_ms(ref_fn, reps)
    print(json.dumps({"case": case, "ours_ms": round(ms_o, 5), "reference_cuda_ms": round(ms_r, 5),
                      "speedup": round(ms_r / ms_o, 2), "ours_events_per_s": events / ms_o * 1e3,
                      "reference_events_per_s": events / ms_r * 1e3, "max_abs_diff": err,
                      "max_abs_ref": scale, **extra}), flush=True)


def spikes(n, rate, g, shape=None):
    return [torch.rand(shape or (n,), device=dev, generator=g) < rate for _ in range(POOL)]


def case_fcn():
    """BASELINE configs[4], one GPU's shard of the headline bench: 1.25 M x 1000."""
    n, conn = 1_250_016, 1000
    g = torch.Generator(device=dev).manual_seed(11)
    idx = torch.randint(0, n, (n, conn), device=dev, dtype=torch.int32, generator=g)
    w = torch.rand((n, conn), device=dev, generator=g)
    w1 = torch.tensor([0.5], device=dev)
    out = torch.empty(n, device=dev)
    for rate in (0.001, 0.01):
        pool = spikes(n, rate, g)
        ev = sum(int(p.sum()) for p in pool) / POOL * conn
        for name, ww in (("hetero", w), ("homo", w1)):
            report(f"fcn_scatter_{name} 1.25M x 1000, {rate:.1%}", ev,
                   lambda i: be.binary_fcnmv(ww, idx, pool[i % POOL], shape=(n, n), transpose=True),
                   lambda i: R.binary_fcnmv_scatter(ww, idx, pool[i % POOL], n, out),
                   lambda: (be.binary_fcnmv(ww, idx, pool[0], shape=(n, n), transpose=True),
                            R.binary_fcnmv_scatter(ww, idx, pool[0], n).clone()))


def _csr(n, nnz, g):
    indptr = torch.arange(0, (n + 1) * nnz, nnz, device=dev, dtype=torch.int32)
    indices = torch.randint(0, n, (n * nnz,), device=dev, dtype=torch.int32, generator=g)
    weights = torch.randn(n * nnz, device=dev, generator=g)
    return indptr, indices, weights


def case_csr():
    """matvec on the BASELINE configs[2] matrix (1 M x 1 M, 100 nnz/row) + configs[0] (COBA 4 k)."""
    g = torch.Generator(device=dev).manual_seed(7)
    for n, nnz, rates in ((1_000_000, 100, (0.01, 0.05)), (4000, 80, (0.005,))):
        indptr, indices, weights = _csr(n, nnz, g)
        w1 = torch.tensor([0.5], device=dev)
        ws = R.HybridWorkspace(indptr)
        out = torch.empty(n, device=dev)
        for rate in rates:
            pool = spikes(n, rate, g)
            ev = sum(int(p.sum()) for p in pool) / POOL * nnz
            for name, ww in (("hetero", weights), ("homo", w1)):
                report(f"spk@csr scatter_{name} {n}x{nnz}, {rate:.1%}", ev,
                       lambda i: be.binary_csrmv(ww, indices, indptr, pool[i % POOL], shape=(n, n), transpose=True),
                       lambda i: R.binary_csrmv_t_hybrid(ww, indices, indptr, pool[i % POOL], n, ws, out),
                       lambda: (be.binary_csrmv(ww, indices, indptr, pool[0], shape=(n, n), transpose=True),
                                R.binary_csrmv_t_hybrid(ww, indices, indptr, pool[0], n, ws).clone()))
                report(f"csr@spk gather_{name} {n}x{nnz}, {rate:.1%}", ev,
                       lambda i: be.binary_csrmv(ww, indices, indptr, pool[i % POOL], shape=(n, n), transpose=False),
                       lambda i: R.binary_csrmv_nt(ww, indices, indptr, pool[i % POOL], n, out),
                       lambda: (be.binary_csrmv(ww, indices, indptr, pool[0], shape=(n, n), transpose=False),
                                R.binary_csrmv_nt(ww, indices, indptr, pool[0], n).clone()), reps=3)


def case_spmm():
    """BASELINE configs[2]: B[64, 1 M] @ csr, via binary_csrmm(transpose=True) on B^T [1 M, 64]."""
    n, nnz, batch = 1_000_000, 100, 64
    g = torch.Generator(device=dev).manual_seed(7)
    indptr, indices, weights = _csr(n, nnz, g)
    ws = R.HybridWorkspace(indptr)
    out = torch.empty(batch, n, device=dev)
    global POOL
    keep, POOL = POOL, 2
    for rate in (0.01, 0.05):
        pool = spikes(n, rate, g, (n, batch))
        ev = sum(int(p.sum()) for p in pool) / POOL * nnz
        report(f"B[64,1M]@csr hetero, {rate:.0%}", ev,
               lambda i: be.binary_csrmm(weights, indices, indptr, pool[i % POOL], shape=(n, n), transpose=True),
               lambda i: R.binary_csrmm_t_hybrid(weights, indices, indptr, pool[i % POOL], n, ws, out),
               lambda: (be.binary_csrmm(weights, indices, indptr, pool[0], shape=(n, n), transpose=True),
                        R.binary_csrmm_t_hybrid(weights, indices, indptr, pool[0], n, ws).t()), reps=2)
    POOL = keep


def _case_jit(tag, cls, ref, params, shape, prob, seed, rates, gather_reps=2):
    g = torch.Generator(device=dev).manual_seed(5)
    n_pre, n_post = shape
    ps = [torch.tensor([p], device=dev) for p in params]
    out = torch.empty(n_post, device=dev)
    for rate in rates:
        pool = spikes(n_pre, rate, g)
        ev = sum(int(p.sum()) for p in pool) / POOL * n_post * prob
        # spk @ M is the operator call with transpose=True and corder = not M.corder
        # (`_jit_scalar/main.py:963-1011`).  Operator corder=False -> reference `trans` kernel (scatter
        # over active rows); corder=True -> `notrans` (every output walks all its sources).
        for corder, kind, reps in ((False, "scatter", 5), (True, "gather", gather_reps)):
            m = cls((*params, prob, seed), shape=shape, corder=not corder)
            report(f"spk@{tag} {n_pre}x{n_post} prob {prob:g} {kind} kernel, {rate:.2%}", ev,
                   lambda i: be.BinaryArray(pool[i % POOL]) @ m,
                   lambda i: ref(*ps, prob, pool[i % POOL], seed, shape=shape, transpose=True, corder=corder, out=out),
                   lambda: (be.BinaryArray(pool[0]) @ m,
                            ref(*ps, prob, pool[0], seed, shape=shape, transpose=True, corder=corder).clone()),
                   reps=reps, events_note="expected events = active sources x n_post x prob")


def case_jits():
    """BASELINE configs[1]: CUBA 400 k, excitatory projection 320 k x 400 k, prob 2e-4."""
    _case_jit("JITCScalarR", be.JITCScalarR, R.binary_jitsmv, (1.62,), (320_000, 400_000), 2e-4, 42, (0.0024,))


def case_jitn():
    """BASELINE configs[3]: 1 M x 1 M, prob 1e-3, Normal(1, 0.1)."""
    _case_jit("JITCNormalR", be.JITCNormalR, R.binary_jitnmv, (1.0, 0.1), (1_000_000, 1_000_000), 1e-3, 123,
              (0.01,), gather_reps=1)


def case_jitu():
    _case_jit("JITCUniformR", be.JITCUniformR, R.binary_jitumv, (0.0, 1.0), (1_000_000, 1_000_000), 1e-3, 123,
              (0.01,), gather_reps=1)


def case_dense():
    """BinaryArray[64, 16384] @ W[16384, 16384] (binary_densemm transpose=True)."""
    k = m = 16384
    batch = 64
    g = torch.Generator(device=dev).manual_seed(3)
    W = torch.randn((k, m), device=dev, generator=g)
    out = torch.empty(batch, m, device=dev)
    for rate in (0.01, 0.05):
        pool = spikes(k, rate, g, (k, batch))
        ev = sum(int(p.sum()) for p in pool) / POOL * m
        report(f"dense B[64,16384]@W, {rate:.0%}", ev,
               lambda i: be.binary_densemm(W, pool[i % POOL], transpose=True),
               lambda i: R.binary_densemm_transpose(W, pool[i % POOL], out),
               lambda: (be.binary_densemm(W, pool[0], transpose=True),
                        R.binary_densemm_transpose(W, pool[0]).t()), reps=3)


def case_compact():
    g = torch.Generator(device=dev).manual_seed(2)
    for n in (1_250_016, 10_000_000):
        pool = spikes(n, 0.01, g)
        ids_o, cnt_o = be.binary_1d_array_index_p_call(pool[0])
        ids_r, cnt_r = R.compact_1d(pool[0])
        k = int(cnt_r.item())
        same = int(cnt_o.reshape(-1)[0].item()) == k and torch.equal(ids_o[:k], torch.sort(ids_r[:k]).values)
        ms_o = graph_ms(lambda i: be.binary_1d_array_index_p_call(pool[i % POOL]))
        ms_r = graph_ms(lambda i: R.compact_1d(pool[i % POOL]))
        print(json.dumps({"case": f"compact_1d n={n}, 1%", "ours_ms": round(ms_o, 5), "reference_cuda_ms": round(ms_r, 5),
                          "speedup": round(ms_r / ms_o, 2), "same_id_set": bool(same),
                          "note": "ours ascending; the reference's is unordered (sorted for the comparison)"}), flush=True)


CASES = {"fcn": case_fcn, "csr": case_csr, "spmm": case_spmm, "jits": case_jits, "jitn": case_jitn,
         "jitu": case_jitu, "dense": case_dense, "compact": case_compact}

if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "pool": POOL,
                      "timing": "CUDA events around CUDA-graph replays of POOL calls each"}), flush=True)
    for nm in names:
        try:
            CASES[nm]()
        except Exception as e:  # keep going: one failing reference entry must not hide the others
            print(json.dumps({"case": nm, "error": f"{type(e).__name__}: {e}"}), flush=True)
        torch.cuda.empty_cache()
