"""Timing of the SURVEY 8(f) "next" rows on one B200 (CUDA events on the launching stream, inputs
larger than L2 or rotated).  Prints one line per case: time, algorithmic bytes, GB/s, fraction of
the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brainevent_b200 as be  # noqa: E402

dev = "cuda"
PEAK = 6550.4
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"{name:<64s} {ms * 1e3:9.1f} us  {nbytes / 1e6:9.1f} MB  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of HBM peak", flush=True)


# ---- float CSR / FCN ops at the configs[2] matrix (1M x 1M, 100 nnz/row) ---------------------------
m = k = 1_000_000
nnz = 100
ix = torch.randint(0, k, (m, nnz), dtype=torch.int32, device=dev, generator=g)
w = torch.rand((m, nnz), device=dev, generator=g)
indptr = torch.arange(0, (m + 1) * nnz, nnz, dtype=torch.int32, device=dev)
x = torch.randn(k, device=dev, generator=g)
nse = m * nnz
report("csrmv  A @ x   (gather, hetero)", timeit(lambda: be.csrmv(w.view(-1), ix.view(-1), indptr, x, shape=(m, k))), nse * 8 + m * 8 + k * 4)
report("csrmv  A^T @ x (scatter, hetero, dense x)", timeit(lambda: be.csrmv(w.view(-1), ix.view(-1), indptr, x, shape=(m, k), transpose=True)), nse * 8 + m * 8 + k * 4)
report("fcnmv  A @ x   (gather, hetero)", timeit(lambda: be.fcnmv(w, ix, x, shape=(m, k))), nse * 8 + m * 4 + k * 4)
report("fcnmv  A @ x   (gather, homo)", timeit(lambda: be.fcnmv(w[:1, :1].reshape(1), ix, x, shape=(m, k))), nse * 4 + m * 4 + k * 4)
X = torch.randn((k, 32), device=dev, generator=g)
report("csrmm  A @ X[.,32] (gather)", timeit(lambda: be.csrmm(w.view(-1), ix.view(-1), indptr, X, shape=(m, k)), reps=5), nse * 8 + (m + k) * 32 * 4)
report("csrmm  A^T @ X[.,32] (scatter)", timeit(lambda: be.csrmm(w.view(-1), ix.view(-1), indptr, X, shape=(m, k), transpose=True), reps=5), nse * 8 + (m + k) * 32 * 4)

# ---- plasticity on the same matrix -----------------------------------------------------------------
pre = torch.rand(m, device=dev, generator=g) < 0.01
trace = torch.randn(k, device=dev, generator=g)
wf = w.view(-1).clone()
n_act = int(pre.sum())
report("update_csr_on_binary_pre (1% rows, in place)", timeit(lambda: be.update_csr_on_binary_pre(wf, ix.view(-1), indptr, pre, trace, shape=(m, k), inplace=True)), n_act * nnz * 12 + m)
report("update_fixed_post_conn_on_binary_pre (1% rows, in place)", timeit(lambda: be.update_fixed_post_conn_on_binary_pre(w, ix, pre, trace, shape=(m, k), inplace=True)), n_act * nnz * 12 + m)
cp, ci, perm = be.csr_to_csc_index(indptr, ix.view(-1), shape=(m, k))
post = torch.rand(k, device=dev, generator=g) < 0.01
report("update_csr_on_binary_post (1% cols via mirror, in place)", timeit(lambda: be.update_csr_on_binary_post(wf, ci, cp, perm, trace, post, shape=(m, k), inplace=True)), int(post.sum()) * nnz * (4 + 4 + 8 + 32) + k)
del cp, ci, perm

# ---- JITC float + materialisation ------------------------------------------------------------------
n = 1_000_000
xv = torch.randn(n, device=dev, generator=g)
for corder in (True, False):
    ms = timeit(lambda: be.jitnmv(0.5, 0.25, 1e-3, xv, 7, shape=(n, n), corder=corder), reps=5)
    report(f"jitnmv 1M x 1M prob 1e-3 corder={corder} (CSR-equivalent 8 B/conn)", ms, int(n * n * 1e-3) * 8)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
for corder in (True, False):
    be.jitn_to_csr(0.5, 0.25, 1e-3, 7, shape=(200_000, 200_000), corder=corder, matrix_mode="mv")
    torch.cuda.synchronize(); t0.record()
    c = be.jitn_to_csr(0.5, 0.25, 1e-3, 7, shape=(n, n), corder=corder, matrix_mode="mv")
    t1.record(); torch.cuda.synchronize()
    report(f"jitn_to_csr 1M x 1M prob 1e-3 corder={corder} (nnz {c.indices.numel() / 1e6:.0f}M, output bytes)", t0.elapsed_time(t1), c.indices.numel() * 8 + n * 8)
    del c

# ---- encoders ----------------------------------------------------------------------------------------
S = torch.rand((1_000_000, 64), device=dev, generator=g) < 0.01
nb = S.numel()
report("binary_2d_csr_encode 1M x 64 bool, 1%", timeit(lambda: be.binary_2d_csr_encode_p_call(S)), nb + nb * 4 + int(S.sum()) * 8)
report("binary_2d_pair_stream_encode 1M x 64 bool, 1%", timeit(lambda: be.binary_2d_pair_stream_encode_p_call(S)), nb + nb * 8 + int(S.sum()) * 12)
report("binary_2d_csc_encode 1M x 64 bool, 1%", timeit(lambda: be.binary_2d_csc_encode_p_call(S)), 3 * nb + nb * 4 + int(S.sum()) * 8)
report("binary_2d_row_sparse_encode 1M x 64 -> row_size 16", timeit(lambda: be.binary_2d_row_sparse_encode_p_call(S, row_size=16)), nb + 1_000_000 * 16 * 4)
