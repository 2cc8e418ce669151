"""A/B of the tcgen05 dense event matmul's knobs: BE_B200_TC_PREFETCH (L2 prefetch distance in k-blocks, 0 = off) and
BE_B200_TC_LANEMAP (lane -> patch element map).  Usage: bench_dense_pf.py [pf=0,6] [map=0,2]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brainevent_b200 as be

opts = dict(a.split("=") for a in sys.argv[1:])
dists = [int(a) for a in opts.get("pf", "0").split(",")]
maps = [int(a) for a in opts.get("map", "0").split(",")]
splits = opts.get("splitk", "0").split(",")
for (k, m, n) in ((16384, 16384, 64), (16384, 16384, 256)):
    g = torch.Generator(device="cuda").manual_seed(0)
    W = torch.randn((k, m), device="cuda", generator=g)
    for rate in (0.01, 0.05):
        pool = [torch.rand((k, n), device="cuda", generator=g) < rate for _ in range(4)]
        ref = W.t().double() @ pool[0].double()
        for d, lm, sk in [(d, lm, sk) for d in dists for lm in maps for sk in splits]:
            os.environ["BE_B200_TC_PREFETCH"] = str(d)
            os.environ["BE_B200_TC_LANEMAP"] = str(lm)
            if sk != "0": os.environ["BE_B200_TC_SPLITK"] = sk
            else: os.environ.pop("BE_B200_TC_SPLITK", None)
            out = be.binary_densemm(W, pool[0], transpose=True)
            err = float(((out.double() - ref).abs() / (W.abs().t().double() @ pool[0].double() + 1e-6)).max())
            for i in range(3):
                be.binary_densemm(W, pool[i % 4], transpose=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20):
                be.binary_densemm(W, pool[i % 4], transpose=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            live = float(pool[0].any(1).float().mean())
            print(json.dumps({"k": k, "m": m, "n": n, "rate": rate, "pf_dist": d, "lane_map": lm, "splitk": sk, "ms_per_call_incl_pack": round(ms, 4),
                              "live_rows": round(live, 3), "W_live_GBs": round(4 * k * m * live / ms / 1e6, 1), "rel_err": err}), flush=True)
