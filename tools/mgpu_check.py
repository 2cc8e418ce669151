"""Multi-GPU correctness check (run under torchrun): the fused pack + peer-memory all-gather must
deliver exactly the words NCCL's all-gather delivers, and the post-sharded product must equal the
oracle's unsharded product on every rank's block.  Prints one OK line per rank."""
import os, sys
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
proj_p = bd.PostShardedProjection("csr", N, N, xp, data=t(dw), indices=t(ix), indptr=t(ip))
for step in range(7):          # odd count on purpose: parity lives on the device
    spikes = np.random.default_rng(100 + step).random(N) < 0.03
    mine = t(spikes[lo:hi])
    words_n = xn.exchange(proj_n.pack_local(mine)).clone()
    words_p = xp.exchange_spikes(mine).clone()
    assert torch.equal(words_n, words_p), f"rank {rank} step {step}: gathered words differ"
    np.testing.assert_array_equal(words_p.cpu().numpy().view(np.uint32), O.bitpack(spikes, 0))
    out = proj_p.step(mine).cpu().numpy()
    ref = O.binary_fcnmv(w, idx, spikes, shape=(N, N), transpose=True)[lo:hi]
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=2e-5)
assert not xp.timed_out()
# FixedNumPerPre shard behind the FUSED exchange (wait + pull + compaction + output clear in one kernel,
# consumed as SPK_COMPACT_CLEARED): every pre neuron has c_loc targets inside this rank's block
n_loc, c_loc = hi - lo, 6
rl = np.random.default_rng(50 + rank)
idx_f = rl.integers(0, n_loc, (N, c_loc)).astype(np.int32)
w_f = rl.standard_normal((N, c_loc)).astype(np.float32)
proj_f = bd.PostShardedProjection("fcn", N, N, xp, weights=t(w_f), indices=t(idx_f))
for step, rate in enumerate([0.03, 0.0, 0.5, 0.01, 0.2]):
    spikes = np.random.default_rng(200 + step).random(N) < rate
    out = proj_f.step(t(spikes[lo:hi]))
    ids = xp.ids.cpu().numpy()
    assert ids[N] == spikes.sum(), f"rank {rank} step {step}: n_active {ids[N]} != {spikes.sum()}"
    np.testing.assert_array_equal(np.sort(ids[: ids[N]]), np.nonzero(spikes)[0])
    ref = O.binary_fcnmv(w_f, idx_f, spikes, shape=(N, n_loc), transpose=True)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=2e-5)
assert not xp.timed_out()
pipe_f = bd.HostStepPipeline(proj_f, n_loc, slots=4)            # the fused step inside one CUDA graph per slot
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
pipe_f.close()
del pipe_f
assert not xp.timed_out()
# host-buffer pipeline (4 timesteps in flight, one CUDA graph per slot) == the plain step
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
# JITC post-sharding (column window for scatter, row window for gather) against the oracle
for corder in (False, True):
    pj = bd.PostShardedProjection("jitc", N, N, xp, family="normal", params=(1.0, 0.1), prob=0.01, seed=5, corder=corder)
    spikes = np.random.default_rng(7).random(N) < 0.05
    out = pj.step(t(spikes[lo:hi])).cpu().numpy()
    ref = O.binary_jitnmv(1.0, 0.1, 0.01, spikes, 5, shape=(N, N), transpose=True, corder=corder)[lo:hi]
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=2e-5)
torch.cuda.synchronize()
dist.barrier()
print(f"rank {rank}/{world}: exchange + post-sharded fcn/jitc parity OK", flush=True)
dist.destroy_process_group()
