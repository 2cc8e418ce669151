"""A/B of the dense event matmul at batch 64: tcgen05 path vs SIMT path (vs cuBLAS fp32 for context)."""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import brainevent_b200 as be
    tag = sys.argv[2]
    for (k, m, n) in ((16384, 16384, 64), (8192, 32768, 128), (4096, 4096, 64)):
        g = torch.Generator(device="cuda").manual_seed(0)
        W = torch.randn((k, m), device="cuda", generator=g)
        for rate in (0.01, 0.1, 0.5):
            S = torch.rand((k, n), device="cuda", generator=g) < rate
            fn = (lambda: be.binary_densemm(W, S, transpose=True)) if tag != "cublas" else (lambda: W.t() @ S.float())
            for _ in range(3):
                out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            ref = W.t().double() @ S.double()
            err = float(((out.double() - ref).abs() / (W.abs().t().double() @ S.double() + 1e-6)).max())
            print(json.dumps({"impl": tag, "k": k, "m": m, "n": n, "rate": rate, "ms": round(ms, 4),
                              "W_GBs": round(4 * k * m / ms / 1e6, 1), "TFLOPs_dense_equiv": round(2 * k * m * n / ms / 1e9, 1),
                              "max_err_over_sum_abs_w": err}))
else:
    for tag, env in (("tcgen05", {"BE_B200_DENSE_TC": "1"}), ("simt", {"BE_B200_DENSE_TC": "0"}), ("cublas", {})):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, __file__, "child", tag], env=e)
