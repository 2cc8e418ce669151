"""Does row alignment decide the fused exchange + scatter kernel's time?  One GPU, 8 emulated ranks, rank 0's shard of a
10 M-neuron network at 1 % spikes, three layouts of the same ~128 synapses per pre row: ELL (128-bit loads, no masks),
the same rows as CSR (masked quads), ragged CSR (binomial row lengths)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brainevent_b200 as be
from brainevent_b200 import dist as bd

dev = torch.device("cuda", 0)
W, n_loc, conn = 8, 1_250_016, 128
N = W * n_loc
g = torch.Generator(device=dev).manual_seed(1)
idx = torch.randint(0, n_loc, (N, conn), device=dev, dtype=torch.int32, generator=g)
w = torch.rand((N, conn), device=dev, generator=g)
xs = bd.P2PSpikeExchange.emulate(N, W, dev)
pools = [[torch.rand(b[1] - b[0], device=dev, generator=g) < 0.01 for _ in range(8)] for b in xs[0].bounds]


def run(name, proj, reps=40):
    def step(t):
        for r in range(1, W):
            xs[r].publish_only(pools[r][t % 8])
        return proj.step(pools[0][t % 8])
    for t in range(5):
        step(t)
    torch.cuda.synchronize()
    # time only rank 0's consume kernel: publish everything first, then event-bracket the consume
    tot = 0.0
    for t in range(reps):
        for r in range(1, W):
            xs[r].publish_only(pools[r][t % 8])
        out = proj.publish(pools[0][t % 8])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        proj.consume(out)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f"{name}: {tot / reps * 1e3:.1f} us per fused scatter", flush=True)


run("ELL 128/row (VEC=2)", bd.PostShardedProjection("fcn", N, N, xs[0], weights=w, indices=idx))
ip = torch.arange(0, (N + 1) * conn, conn, device=dev, dtype=torch.int32)
run("same rows as CSR, aligned starts (VEC=1 masked)", bd.PostShardedProjection("csr", N, N, xs[0], data=w.view(-1), indices=idx.view(-1), indptr=ip))
lens = torch.randint(100, 151, (N,), device=dev, generator=g)
keep = torch.arange(conn + 32, device=dev)[None, :] < lens[:, None]
del keep
lens = torch.clamp(lens, max=conn)
mask = torch.arange(conn, device=dev)[None, :] < lens[:, None]
ip2 = torch.zeros(N + 1, device=dev, dtype=torch.int64)
torch.cumsum(lens, 0, out=ip2[1:])
run("ragged CSR 100..128/row (VEC=1 masked)", bd.PostShardedProjection("csr", N, N, xs[0], data=w[mask], indices=idx[mask], indptr=ip2.to(torch.int32)))
