"""Multi-GPU correctness check (run under torchrun, N >= 2): the fused pack + peer-memory all-gather must
deliver exactly the words NCCL's all-gather delivers, and every post-sharded product (CSR ragged shards
behind the exchange fused into the scatter, the unfused three-kernel chain, FixedNumPerPre, dense, JITC in
both orders, the C host pipeline) must equal the oracle's UNSHARDED product on every rank's block.
Prints one `parity OK` line per rank; bench.py's own parity block covers the full-size networks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import brainevent_b200 as be
from brainevent_b200 import dist as bd
import oracle as O

n_local, n_conn = 4096 + 64, 24
N = n_local * world
rng = np.random.default_rng(1)
idx = rng.integers(0, N, (N, n_conn)).astype(np.int32)
w = rng.standard_normal((N, n_conn)).astype(np.float32)
shards = bd.partition_fcn_by_post(idx, w, N, world)
ip, ix, dw, (lo, hi) = shards[rank]
t = lambda a: torch.as_tensor(a, device=dev)
xn = bd.SpikeExchange(N, device=dev)
xp = bd.P2PSpikeExchange(N, device=dev)
assert (lo, hi) == xp.local == xn.local
proj_n = bd.PostShardedProjection("csr", N, N, xn, data=t(dw), indices=t(ix), indptr=t(ip))
proj_p = bd.PostShardedProjection("csr", N, N, xp, data=t(dw), indices=t(ix), indptr=t(ip))     # fused (default)
proj_u = bd.PostShardedProjection("csr", N, N, xp, data=t(dw), indices=t(ix), indptr=t(ip))
proj_u.set_option("fused", 0)
for step in range(7):          # odd count on purpose: parity lives on the device
    spikes = np.random.default_rng(100 + step).random(N) < 0.03
    mine = t(spikes[lo:hi])
    words_n = xn.exchange(proj_n.pack_local(mine)).clone()
    words_p = xp.exchange_spikes(mine).clone()
    assert torch.equal(words_n, words_p), f"rank {rank} step {step}: gathered words differ"
    np.testing.assert_array_equal(words_p.cpu().numpy().view(np.uint32), O.bitpack(spikes, 0))
    ref = O.binary_fcnmv(w, idx, spikes, shape=(N, N), transpose=True)[lo:hi]
    np.testing.assert_allclose(proj_p.step(mine).cpu().numpy(), ref, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(proj_u.step(mine).cpu().numpy(), ref, rtol=1e-5, atol=2e-5)
    scratch = torch.ones(hi - lo, device=dev)
    _, ids_t = xp.exchange_compact(mine, scratch)     # the wait kernel's id list of the gathered vector + output clear
    assert float(scratch.abs().max().item()) == 0.0
    ids = ids_t.cpu().numpy()
    assert ids[N] == spikes.sum(), f"rank {rank} step {step}: n_active {ids[N]} != {spikes.sum()}"
    np.testing.assert_array_equal(np.sort(ids[: ids[N]]), np.nonzero(spikes)[0])
    np.testing.assert_allclose(proj_n.step(mine).cpu().numpy(), ref, rtol=1e-5, atol=2e-5)
xp.check()
# FixedNumPerPre shard: every pre neuron has c_loc targets inside this rank's block
n_loc, c_loc = hi - lo, 6
rl = np.random.default_rng(50 + rank)
idx_f = rl.integers(0, n_loc, (N, c_loc)).astype(np.int32)
w_f = rl.standard_normal((N, c_loc)).astype(np.float32)
proj_f = bd.PostShardedProjection("fcn", N, N, xp, weights=t(w_f), indices=t(idx_f))
for step, rate in enumerate([0.03, 0.0, 0.5, 0.01, 0.2]):
    spikes = np.random.default_rng(200 + step).random(N) < rate
    out = proj_f.step(t(spikes[lo:hi]))
    ref = O.binary_fcnmv(w_f, idx_f, spikes, shape=(N, n_loc), transpose=True)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=2e-5)
xp.check()
# C host pipeline (be_b200_sharded_submit_host / _run_host): the fused step inside one CUDA graph per slot
pipe_f = bd.HostStepPipeline(proj_f, n_loc, slots=4)
assert pipe_f.native
hf_in = [torch.empty(n_loc, dtype=torch.uint8).pin_memory() for _ in range(7)]
hf_out = [torch.empty(n_loc, dtype=torch.float32).pin_memory() for _ in range(7)]
refs_f = []
for q in range(7):
    sp = np.random.default_rng(400 + q).random(N) < 0.05
    hf_in[q].copy_(torch.from_numpy(sp[lo:hi].astype(np.uint8)))
    refs_f.append(O.binary_fcnmv(w_f, idx_f, sp, shape=(N, n_loc), transpose=True))
for q in range(7):
    pipe_f.submit(hf_in[q], hf_out[q])
pipe_f.wait()
for q in range(7):
    np.testing.assert_allclose(hf_out[q].numpy(), refs_f[q], rtol=1e-5, atol=2e-5)
for o in hf_out:
    o.zero_()
pipe_f.run(7, hf_in, hf_out)
for q in range(7):
    np.testing.assert_allclose(hf_out[q].numpy(), refs_f[q], rtol=1e-5, atol=2e-5)
pipe_f.close()
del pipe_f
# host-buffer pipeline over the NCCL exchange (Python-driven, one CUDA graph per slot) == the plain step
pipe = bd.HostStepPipeline(proj_n, hi - lo, slots=4)
h_in = [torch.empty(hi - lo, dtype=torch.uint8).pin_memory() for _ in range(9)]
h_out = [torch.empty(hi - lo, dtype=torch.float32).pin_memory() for _ in range(9)]
refs = []
for q in range(9):
    sp = np.random.default_rng(300 + q).random(N) < 0.04
    h_in[q].copy_(torch.from_numpy(sp[lo:hi].astype(np.uint8)))
    refs.append(O.binary_fcnmv(w, idx, sp, shape=(N, N), transpose=True)[lo:hi])
for q in range(9):
    pipe.submit(h_in[q], h_out[q])
pipe.wait()
for q in range(9):
    np.testing.assert_allclose(h_out[q].numpy(), refs[q], rtol=1e-5, atol=2e-5)
pipe.close()
del pipe
# dense event matvec, column (post) split of W
Wd = np.random.default_rng(9).standard_normal((N, N)).astype(np.float32) if N <= 20_000 else None
if Wd is not None:
    pd = bd.PostShardedProjection("dense", N, N, xp, W=t(np.ascontiguousarray(Wd[:, lo:hi])))
    for step, rate in enumerate([0.05, 0.0, 0.4]):
        spikes = np.random.default_rng(600 + step).random(N) < rate
        out = pd.step(t(spikes[lo:hi])).cpu().numpy()
        np.testing.assert_allclose(out, O.binary_densemv(Wd, spikes, transpose=True)[lo:hi], rtol=1e-4, atol=2e-3)
# JITC post-sharding (column window for scatter, row window for gather) against the oracle
for corder in (False, True):
    pj = bd.PostShardedProjection("jitc", N, N, xp, family="normal", params=(1.0, 0.1), prob=0.01, seed=5, corder=corder)
    spikes = np.random.default_rng(7).random(N) < 0.05
    out = pj.step(t(spikes[lo:hi])).cpu().numpy()
    ref = O.binary_jitnmv(1.0, 0.1, 0.01, spikes, 5, shape=(N, N), transpose=True, corder=corder)[lo:hi]
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=2e-5)
xp.check()
torch.cuda.synchronize()
dist.barrier()
# a rank that stalls past the spin limit: everyone must notice.  Separate exchange, 50 ms limit,
# rank `world-1` sleeps 0.5 s before its step.
xt = bd.P2PSpikeExchange(N, device=dev)
xt.set_spin_limit_ms(50)
pt = bd.PostShardedProjection("csr", N, N, xt, data=t(dw), indices=t(ix), indptr=t(ip))
spikes = np.random.default_rng(1).random(N) < 0.03
pt.step(t(spikes[lo:hi]))
torch.cuda.synchronize()
dist.barrier()
if rank == world - 1:
    time.sleep(0.5)
pt.step(t(spikes[lo:hi]))
torch.cuda.synchronize()
try:
    xt.check()
    raise SystemExit(f"rank {rank}: the stalled exchange was not reported")
except bd.ExchangeTimeout:
    pass
dist.barrier()
print(f"rank {rank}/{world}: exchange + post-sharded csr(fused/unfused)/fcn/dense/jitc + host pipelines + timeout: parity OK", flush=True)
dist.destroy_process_group()
