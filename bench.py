#!/usr/bin/env python
"""bench.py -- synaptic events / second of the event-driven propagation path on B200(s).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A *step* is one block of T simulated timesteps (default 100) of ONE projection:
every timestep takes a fresh synthetic spike vector through the hot path
(N = 1: bit-pack + compaction -> active-row scatter; N > 1: bit-pack + publish -> spike exchange
over NVLink peer memory fused into the active-row scatter).

Default workload `fcn` (BASELINE.json configs[4], weak scaling): FixedNumPerPre, 1.25 M neurons per
GPU x 1000 synapses per neuron, fp32 heterogeneous weights, 1 % spikes per timestep.  The matrix is
the SURVEY 8(d) generator: every pre neuron draws its 1000 targets uniformly over ALL post neurons;
it is generated shard-wise with counter-based seeds (every rank draws the same chunk stream and keeps
the synapses ending in its post block), so a rank's shard is a RAGGED CSR (binomial ~1000/N per
row) and the union of the shards is the single-GPU matrix.  At N = 8 this is the named
10 M-neuron x 1000 network.  The line also carries
  * `parity` (N > 1): every rank's output block checked before timing against an independent path
    (NCCL all-gather + the plain operator) and an oracle slice; the run fails if it differs;
  * `strong_scaling`: the FULL 10 000 128 x 1000 network on these N GPUs (80 GB on one GPU at N = 1);
  * `other_workloads`: N = 1: the other BASELINE configs with `reference_cuda_us` = the reference's own
    CUDA kernel (oracle/_ref, recompiled for sm_100a) timed on the same buffers; N > 1: configs[3]
    (JITCNormalR 1 M, post-sharded, both orders).

`--impl reference` times the CPU arm: the C restatement of the reference's numba kernels
(oracle/, kind "port") on the host cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FCN_NEURONS_PER_GPU = 1_250_016   # multiple of 32 (packed all-gather slots), ~1.25 M
FCN_CONN = 1000
RATE = 0.01
POOL = 16                          # distinct spike vectors cycled through the timesteps
DEFAULT_EXCHANGE = "p2p"           # spike exchange at N > 1: "nccl" (all_gather_into_tensor) or "p2p" (csrc/exchange.cu); env BE_B200_EXCHANGE overrides


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread while the GPU runs the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap", 0x2: "applications_clocks_setting", 0x1: "gpu_idle"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        if self._t is not None:
            self._stop.set()
            self._t.join()
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------------
class Workload:
    """build() -> resident tensors; timestep(t) enqueues one pass of the hot path; events_per_step."""
    name = "?"
    dominant_kernel = "?"
    launches_per_timestep = 2
    profiled_kernels_per_timestep = 1
    bytes_per_event = 8.0

    def algorithmic_bytes_per_timestep(self):
        return self.events_per_timestep * self.bytes_per_event + self.extra_bytes_per_timestep


def spike_pool(torch, n, rate, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.rand((POOL, n), device=device, generator=g) < rate


C5_NEURONS = 8 * FCN_NEURONS_PER_GPU   # 10 000 128: the named 10 M network
C5_SEED = 1234
C5_CHUNK_ROWS = 65_536


P2P_TEXT = ("NVLink peer-memory exchange fused into the scatter (csrc/sharded.cu: pack + publish, then one kernel pulls each "
            "peer's words as its flag arrives and scatters them; own rows first)")


def fcn_workload_text(n_total, n_conn, rate, shards):
    return (f"FixedNumPerPre {n_total} neurons x {n_conn} synapses/neuron, fp32 hetero weights, "
            f"{rate:.0%} spikes/timestep, post-sharded over {shards} GPU(s) "
            f"({n_total // shards} post neurons per GPU; every pre neuron draws {n_conn} targets uniformly over ALL "
            f"{n_total} posts, so a GPU holds a ragged CSR of ~{n_conn // shards} synapses/pre-neuron), "
            "BASELINE configs[4] at 8 GPUs")


def fcn_config(n_total, n_conn, rate, world, T):
    """The `config` object of the line: identical for the B200 arm and the CPU (reference) arm."""
    return {"workload": fcn_workload_text(n_total, n_conn, rate, world),
            "neurons": n_total, "synapses_per_neuron": n_conn, "spike_rate": rate,
            "generator": f"SURVEY 8(d): targets = randint(0, {n_total}) per (pre, slot), weights U(0,1), chunks of "
                         f"{C5_CHUNK_ROWS} pre rows seeded {C5_SEED} * 1000003 + chunk on every rank",
            "spike_exchange": P2P_TEXT if world > 1 else "none (single GPU)",
            "timesteps_per_step": T,
            "l2_policy": f"inputs larger than L2: {POOL} distinct spike vectors cycle, so the active-row working set "
                         "of one cycle exceeds the 126 MB L2",
            "parallelism": f"post-sharded x{world}" if world > 1 else "single GPU"}


def c5_chunks(torch, device, n_total, n_conn, seed, with_weights=True, out_idx=None, out_w=None):
    """Yields (row0, targets int32[rows, n_conn], weights f32[rows, n_conn] | None) of the SURVEY 8(d) network.
    Chunk c is drawn from its own generator seeded from (seed, c): every rank, and the single-GPU build,
    produce the identical matrix without anybody materialising more than one chunk."""
    for c, r0 in enumerate(range(0, n_total, C5_CHUNK_ROWS)):
        rows = min(C5_CHUNK_ROWS, n_total - r0)
        g = torch.Generator(device=device).manual_seed(seed * 1_000_003 + c)
        # in-place draws (what torch.randint / torch.rand do internally): identical streams whether the chunk is a
        # fresh tensor or a row slice of the preallocated single-GPU matrix
        tgt = out_idx[r0:r0 + rows] if out_idx is not None else torch.empty((rows, n_conn), device=device, dtype=torch.int32)
        tgt.random_(0, n_total, generator=g)
        wv = None
        if with_weights:
            wv = out_w[r0:r0 + rows] if out_w is not None else torch.empty((rows, n_conn), device=device, dtype=torch.float32)
            wv.uniform_(0.0, 1.0, generator=g)
        yield r0, tgt, wv


def build_c5_ell(torch, device, n_total, n_conn, seed=C5_SEED):
    """The whole network on one GPU: indices int32 [n_total, n_conn], weights f32 [n_total, n_conn] (FixedNumPerPre)."""
    idx = torch.empty((n_total, n_conn), device=device, dtype=torch.int32)
    w = torch.empty((n_total, n_conn), device=device, dtype=torch.float32)
    for _ in c5_chunks(torch, device, n_total, n_conn, seed, out_idx=idx, out_w=w):
        pass
    return idx, w


def build_c5_shard(torch, device, n_total, n_conn, lo, hi, seed=C5_SEED):
    """One rank's share: the synapses ending in post block [lo, hi) as a ragged CSR over all n_total pre rows with
    LOCAL post ids (two passes over the chunk stream: count, then fill exact-size arrays)."""
    counts = torch.empty(n_total, device=device, dtype=torch.int64)
    for r0, tgt, _ in c5_chunks(torch, device, n_total, n_conn, seed, with_weights=False):
        counts[r0:r0 + tgt.shape[0]] = ((tgt >= lo) & (tgt < hi)).sum(1)
    indptr = torch.zeros(n_total + 1, device=device, dtype=torch.int64)
    torch.cumsum(counts, 0, out=indptr[1:])
    nse = int(indptr[-1].item())
    indices = torch.empty(nse, device=device, dtype=torch.int32)
    weights = torch.empty(nse, device=device, dtype=torch.float32)
    for r0, tgt, wv in c5_chunks(torch, device, n_total, n_conn, seed):
        m = (tgt >= lo) & (tgt < hi)
        p0, p1 = int(indptr[r0].item()), int(indptr[r0 + tgt.shape[0]].item())
        indices[p0:p1] = tgt[m] - lo
        weights[p0:p1] = wv[m]
    if nse < 2**31:
        indptr = indptr.to(torch.int32)
    return indptr, indices, weights, counts


class FcnSharded(Workload):
    """FixedNumPerPre network of n_total neurons x n_conn synapses, post-sharded over the ranks.
    world == 1: the FixedNumPerPre operator on the ELL matrix; world > 1: ragged CSR shards behind the
    peer-memory spike exchange fused into the scatter (csrc/sharded.cu), or NCCL all-gather + operator."""
    name = "fcn_perpre_postsharded"
    dominant_kernel = "scatter_mv_kernel"

    def __init__(self, torch, be, rank, world, device, n_total, n_conn=FCN_CONN, rate=RATE, exchange=None, T=100,
                 scaling="weak", emulate_world=0):
        from brainevent_b200 import dist as bd
        self.emulate = emulate_world      # (profiling) rank 0's shard of an `emulate_world`-GPU run on ONE GPU
        if emulate_world:
            world, rank = emulate_world, 0
        self.torch, self.be, self.world, self.rank, self.device = torch, be, world, rank, device
        self.n_total, self.n_conn, self.rate, self.scaling = n_total, n_conn, rate, scaling
        self.bounds = bd.post_block_bounds(n_total, world)
        self.lo, self.hi = self.bounds[rank]
        self.n_local = self.hi - self.lo
        # spikes of the neurons I own (every rank can regenerate every rank's pool: seeds 99 + r)
        self.pool = spike_pool(torch, self.n_local, rate, 99 + rank, device)
        gpool = (self.pool if world == 1 else
                 torch.cat([spike_pool(torch, b[1] - b[0], rate, 99 + r, device) for r, b in enumerate(self.bounds)], 1))
        self.active_per_timestep = float(gpool.sum(1).double().mean().item())
        self.exchange_kind = "none (single GPU)"
        self.x = self.proj = None
        if world == 1:
            self.indices, self.weights = build_c5_ell(torch, device, n_total, n_conn)
            self.events_per_timestep = self.active_per_timestep * n_conn
            self.launches_per_timestep = 2
            indptr_bytes = 0.0
        else:
            self.indptr, self.indices, self.weights, counts = build_c5_shard(torch, device, n_total, n_conn, self.lo, self.hi)
            self.events_per_timestep = float((gpool.double() @ counts.double()).mean().item())     # this rank's share
            del counts
            exchange = exchange or os.environ.get("BE_B200_EXCHANGE", DEFAULT_EXCHANGE)
            if emulate_world:
                # all ranks' exchange buffers on this GPU: per timestep every emulated peer publishes its block,
                # then rank 0's fused kernel runs exactly as on the 8-GPU box (peer reads become local reads)
                self.peers = bd.P2PSpikeExchange.emulate(n_total, world, device)
                self.x = self.peers[0]
                self.peer_pools = [spike_pool(torch, b[1] - b[0], rate, 99 + r, device) for r, b in enumerate(self.bounds)]
                self.exchange_kind = f"emulated: {world} ranks' buffers on one GPU (profiling)"
                self.dominant_kernel = "exchange_scatter_kernel"
                self.launches_per_timestep = 1 + world
            elif exchange == "p2p":
                try:
                    self.x = bd.P2PSpikeExchange(n_total, device=device)
                    self.exchange_kind = P2P_TEXT
                    self.dominant_kernel = "exchange_scatter_kernel"
                    self.launches_per_timestep = 2
                except Exception as e:  # pragma: no cover - symmetric memory unavailable: NCCL path
                    sys.stderr.write(f"[bench] P2P exchange unavailable ({type(e).__name__}: {e}); using NCCL\n")
            if self.x is None:
                self.x = bd.SpikeExchange(n_total, device=device)
                self.exchange_kind = "bit-pack + NCCL all_gather_into_tensor + operator"
                self.launches_per_timestep = 4
            self.proj = bd.PostShardedProjection("csr", n_total, n_total, self.x, data=self.weights, indices=self.indices,
                                                 indptr=self.indptr)
            if os.environ.get("BE_B200_SHARDED_FUSED") == "0" and self.proj._h is not None:     # A/B: three-kernel form
                self.proj.set_option("fused", 0)
                self.dominant_kernel = "scatter_mv_kernel"
            indptr_bytes = 8.0 * self.active_per_timestep
        del gpool
        # 8 B/event + packed spikes + indptr pair per active row (CSR) + the output vector
        self.extra_bytes_per_timestep = n_total / 8.0 + indptr_bytes + 4.0 * self.n_local
        self.config = fcn_config(n_total, n_conn, rate, world, T)
        if world > 1 and self.exchange_kind != P2P_TEXT:
            self.config["spike_exchange"] = self.exchange_kind

    def timestep(self, t):
        if self.world == 1:
            return self.be.binary_fcnmv(self.weights, self.indices, self.pool[t % POOL],
                                        shape=(self.n_total, self.n_local), transpose=True)
        if self.emulate:
            for r in range(1, self.world):
                self.peers[r].publish_only(self.peer_pools[r][t % POOL])
        return self.proj.step(self.pool[t % POOL])

    def reference_cuda_timestep(self, R, t):
        if self.world != 1:
            return None
        if getattr(self, "_ref_out", None) is None:
            self._ref_out = self.torch.empty(self.n_local, device=self.device)
        return R.binary_fcnmv_scatter(self.weights, self.indices, self.pool[t % POOL], self.n_local, self._ref_out)

    def total_events_per_timestep(self):
        return self.active_per_timestep * self.n_conn          # every synapse lands on exactly one rank

    def release(self):
        self.proj = self.x = None
        for a in ("indptr", "indices", "weights"):
            if hasattr(self, a):
                delattr(self, a)
        self.torch.cuda.empty_cache()


def parity_block(torch, dist, be, wl):
    """N > 1, before any timing: every rank checks its output block against an independent result and rank 0 gets
    the verdict for the JSON line.  (1) three timesteps of the measured step vs NCCL all-gather of the bit-packed
    spikes + the plain single-GPU operator (its own compaction + scatter kernels); (2) the gathered words of a second
    peer-memory exchange vs NCCL's, bit for bit; (3) a timestep whose only active pre rows are a 2000-row slice owned by
    rank 1, against the CPU oracle on this rank's shard of those rows.  Raises if anything differs."""
    import oracle as O
    from brainevent_b200 import dist as bd
    dev, n, n_loc = wl.device, wl.n_total, wl.n_local
    xn = bd.SpikeExchange(n, device=dev)
    pack = bd.PostShardedProjection("csr", n, n, xn, data=wl.weights, indices=wl.indices, indptr=wl.indptr)
    xp2 = bd.P2PSpikeExchange(n, device=dev) if isinstance(wl.x, bd.P2PSpikeExchange) else None
    max_err, max_ref, ids_equal = 0.0, 0.0, True
    for t in range(3):
        mine = wl.pool[t]
        out = wl.proj.step(mine).clone()
        words = xn.exchange(pack.pack_local(mine)).clone()
        ref = be.binary_csrmv(wl.weights, wl.indices, wl.indptr, be.BitPackedBinary(packed=words, shape=(n,)),
                              shape=(n, n_loc), transpose=True)
        max_err = max(max_err, float((out - ref).abs().max().item()))
        max_ref = max(max_ref, float(ref.abs().max().item()))
        if xp2 is not None:
            ids_equal = ids_equal and bool(torch.equal(xp2.exchange_spikes(mine), words))
    # oracle slice: rows [r0, r0 + 2000) of the GLOBAL vector, owned by rank 1, every third one active
    r0 = wl.bounds[1][0] + 4096
    rows = np.arange(r0, r0 + 2000)
    active = rows[(rows - r0) % 3 == 0]
    mine = torch.zeros(n_loc, dtype=torch.bool, device=dev)
    if wl.rank == 1:
        mine[torch.as_tensor(active - wl.lo, device=dev)] = True
    out = wl.proj.step(mine).cpu().numpy()
    ip = wl.indptr[r0:r0 + 2001].cpu().numpy().astype(np.int64)
    sub_ip = (ip - ip[0]).astype(np.int32)
    sub_ix = wl.indices[int(ip[0]):int(ip[-1])].cpu().numpy()
    sub_w = wl.weights[int(ip[0]):int(ip[-1])].cpu().numpy()
    sl = np.zeros(2000, bool)
    sl[active - r0] = True
    ref = O.binary_csrmv(sub_w, sub_ix, sub_ip, sl, shape=(2000, n_loc), transpose=True)
    oracle_err = float(np.abs(out - ref).max())
    timed_out = bool(wl.x.timed_out()) if hasattr(wl.x, "timed_out") else False
    ok = (max_err <= 1e-5 * max(max_ref, 1.0) + 1e-5) and oracle_err <= 2e-5 and ids_equal and not timed_out
    v = torch.tensor([max_err, oracle_err, 0.0 if ok else 1.0, 0.0 if ids_equal else 1.0], device=dev, dtype=torch.float64)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    res = {"checked": True, "ok": bool(v[2].item() == 0.0), "timesteps": 3, "max_abs_err": float(v[0].item()),
           "against": "NCCL all_gather of bit-packed spikes + binary_csrmv (separate compaction + scatter kernels) on the same shard",
           "ids_equal_nccl": bool(v[3].item() == 0.0) if xp2 is not None else None,
           "oracle_slice": {"rows": 2000, "active": int(active.size), "owner_rank": 1, "max_abs_err": float(v[1].item())}}
    del pack, xn, xp2
    if not res["ok"]:
        raise RuntimeError(f"multi-GPU parity check failed: {res}")
    return res


class Coba(Workload):
    name = "coba_4000_csr"
    dominant_kernel = "scatter_mv_small_kernel"
    bytes_per_event = 4.0   # homogeneous weight: indices only

    def __init__(self, torch, be, device):
        self.be = be
        rng = np.random.default_rng(0)
        self.mats = []
        ev = 0.0
        act = 0.0
        for n_pre, w in ((3200, 0.6), (800, 6.7)):
            indptr = torch.as_tensor((np.arange(n_pre + 1) * 80).astype(np.int32), device=device)
            indices = torch.as_tensor(rng.integers(0, 4000, n_pre * 80).astype(np.int32), device=device)
            csr = be.CSR((torch.tensor([w], device=device), indices, indptr), shape=(n_pre, 4000))
            pool = spike_pool(torch, n_pre, 0.005, n_pre, device)
            self.mats.append((csr, pool))
            a = float(pool.sum(1).double().mean().item())
            act += a
            ev += a * 80
        self.events_per_timestep = ev
        self.extra_bytes_per_timestep = 12.0 * act
        self.launches_per_timestep = 2          # one single-cluster launch per projection (compaction + clear + scatter)
        self.profiled_kernels_per_timestep = 2
        self.config = {"workload": "COBA_2005: 4000-neuron E/I network, 80 synapses/neuron CSR, homogeneous weights, "
                                   "0.5% spikes/timestep, BinaryArray @ CSR (BASELINE configs[0])"}

    def timestep(self, t):
        out = None
        for csr, pool in self.mats:
            out = self.be.BinaryArray(pool[t % POOL]) @ csr
        return out

    def reference_cuda_timestep(self, R, t):
        # the reference's GPU default for spk @ csr: the hybrid scatter kernel (_csr/binary.py:564-621)
        if getattr(self, "_ref", None) is None:
            self._ref = [(R.HybridWorkspace(csr.indptr), csr.data.new_empty(4000)) for csr, _ in self.mats]
        out = None
        for (csr, pool), (ws, o) in zip(self.mats, self._ref):
            out = R.binary_csrmv_t_hybrid(csr.data, csr.indices, csr.indptr, pool[t % POOL], 4000, ws, o)
        return out


class CubaJitc(Workload):
    name = "cuba_400k_jitc_scalar"
    bytes_per_event = 8.0   # CSR-equivalent bytes (SURVEY 8d ii): JITC reads no connectivity from HBM

    def __init__(self, torch, be, device, corder):
        self.be = be
        self.corder = corder
        self.dominant_kernel = "jit_mv_scatter_kernel" if corder else "jit_mv_gather_kernel"
        self.mats = []
        ev = 0.0
        for n_pre, w in ((320_000, 1.62), (80_000, -9.0)):
            m = be.JITCScalarR((w, 2e-4, 42), shape=(n_pre, 400_000), corder=corder)
            pool = spike_pool(torch, n_pre, 0.0024, n_pre, device)
            self.mats.append((m, pool))
            ev += float(pool.sum(1).double().mean().item()) * 400_000 * 2e-4
        self.events_per_timestep = ev
        self.extra_bytes_per_timestep = 0.0
        self.launches_per_timestep = 4
        self.profiled_kernels_per_timestep = 2
        self.config = {"workload": f"CUBA_2005 scaled to 400k neurons, spk @ JITCScalarR(prob 2e-4, corder={corder}) "
                                   f"({'scatter' if corder else 'gather'} kernel), 0.24% spikes/timestep (BASELINE configs[1])"}

    def timestep(self, t):
        out = None
        for m, pool in self.mats:
            out = self.be.BinaryArray(pool[t % POOL]) @ m
        return out

    def reference_cuda_timestep(self, R, t):
        # spk @ M = operator call with transpose=True and corder = not M.corder (_jit_scalar/main.py:963-1011)
        if getattr(self, "_ref", None) is None:
            import torch
            dev = self.mats[0][1].device
            self._ref = [(torch.tensor([w], device=dev), torch.empty(400_000, device=dev)) for w in (1.62, -9.0)]
        out = None
        for (m, pool), (w, o), n_pre in zip(self.mats, self._ref, (320_000, 80_000)):
            out = R.binary_jitsmv(w, 2e-4, pool[t % POOL], 42, shape=(n_pre, 400_000), transpose=True,
                                  corder=not self.corder, out=o)
        return out


class CsrSpmm(Workload):
    name = "csr_1m_spmm_b64"
    dominant_kernel = "scatter_mm_bm_kernel"

    def __init__(self, torch, be, device, rate=0.01, batch=64):
        self.be = be
        self.batch = batch
        n, nnz = 1_000_000, 100
        g = torch.Generator(device=device).manual_seed(7)
        self.indptr = torch.arange(0, (n + 1) * nnz, nnz, device=device, dtype=torch.int32)
        self.indices = torch.randint(0, n, (n * nnz,), device=device, dtype=torch.int32, generator=g)
        self.weights = torch.randn(n * nnz, device=device, generator=g)
        self.pool = torch.rand((4, n, batch), device=device, generator=g) < rate
        self.n = n
        ev = float(self.pool.sum((1, 2)).double().mean().item()) * nnz
        rows_any = float(self.pool.any(2).sum(1).double().mean().item())
        self.events_per_timestep = ev
        # matrix bytes counted once per call: rows active in any column
        self.bytes_per_event = 0.0
        self.extra_bytes_per_timestep = 8.0 * nnz * rows_any + 4.0 * n * batch + n * (batch / 8)
        self.config = {"workload": f"BinaryArray @ CSR SpMM: 1M x 1M CSR, 100 nnz/row, batch {batch}, {rate:.0%} spikes "
                                   "(BASELINE configs[2])"}

    def timestep(self, t):
        return self.be.binary_csrmm(self.weights, self.indices, self.indptr, self.pool[t % 4], shape=(self.n, self.n),
                                    transpose=True)

    def reference_cuda_timestep(self, R, t):
        if getattr(self, "_ref", None) is None:
            self._ref = (R.HybridWorkspace(self.indptr), self.weights.new_empty((self.batch, self.n)))
        return R.binary_csrmm_t_hybrid(self.weights, self.indices, self.indptr, self.pool[t % 4], self.n, *self._ref)


class CsrMv1M(Workload):
    """1M x 1M CSR, 100 nnz/row matvec (configs[2] structure): scatter `spk @ csr` or the
    deterministic post-major gather `csr @ spk`."""
    name = "csr_1m_matvec"

    def __init__(self, torch, be, device, gather, homo=False, rate=0.01):
        self.be = be
        n, nnz = 1_000_000, 100
        g = torch.Generator(device=device).manual_seed(7)
        self.n = n
        self.gather = gather
        self.indptr = torch.arange(0, (n + 1) * nnz, nnz, device=device, dtype=torch.int32)
        self.indices = torch.randint(0, n, (n * nnz,), device=device, dtype=torch.int32, generator=g)
        self.weights = torch.tensor([0.5], device=device) if homo else torch.randn(n * nnz, device=device, generator=g)
        self.pool = spike_pool(torch, n, rate, 5, device)
        act = float(self.pool.sum(1).double().mean().item())
        self.events_per_timestep = act * nnz
        if gather:
            self.dominant_kernel = "gather_mv_kernel"
            self.bytes_per_event = 0.0
            # whole index stream + output + packed spikes; hetero adds the weight quads actually touched
            touched = 0.0 if homo else 4.0 * n * nnz * (1.0 - (1.0 - rate) ** 4)
            self.extra_bytes_per_timestep = 4.0 * n * nnz + 4.0 * n + n / 8 + touched
        else:
            self.dominant_kernel = "scatter_mv_kernel"
            self.bytes_per_event = 4.0 if homo else 8.0
            self.extra_bytes_per_timestep = 12.0 * act
        self.config = {"workload": f"{'csr @ spk (post-major gather)' if gather else 'spk @ csr (scatter)'}: 1M x 1M CSR, "
                                   f"100 nnz/row, {'homogeneous' if homo else 'fp32 hetero'} weights, {rate:.0%} spikes "
                                   "(matvec on the BASELINE configs[2] matrix)"}

    def timestep(self, t):
        return self.be.binary_csrmv(self.weights, self.indices, self.indptr, self.pool[t % POOL], shape=(self.n, self.n),
                                    transpose=not self.gather)

    def reference_cuda_timestep(self, R, t):
        if getattr(self, "_ref", None) is None:
            import torch
            self._ref = (R.HybridWorkspace(self.indptr), torch.empty(self.n, device=self.indices.device))
        ws, out = self._ref
        if self.gather:
            return R.binary_csrmv_nt(self.weights, self.indices, self.indptr, self.pool[t % POOL], self.n, out)
        return R.binary_csrmv_t_hybrid(self.weights, self.indices, self.indptr, self.pool[t % POOL], self.n, ws, out)


class DenseMm(Workload):
    """BinaryArray @ dense at batch 64: the tcgen05 path (3 x bf16 split of the fp32 weights)."""
    name = "dense_mm_b64"
    dominant_kernel = "dense_mm_t_tc_kernel"

    def __init__(self, torch, be, device, k=16384, m=16384, batch=64, rate=0.05):
        self.be = be
        g = torch.Generator(device=device).manual_seed(3)
        self.W = torch.randn((k, m), device=device, generator=g)
        self.pool = torch.rand((4, k, batch), device=device, generator=g) < rate
        self.events_per_timestep = float(self.pool.sum((1, 2)).double().mean().item()) * m
        self.bytes_per_event = 0.0
        # 4 B x m x (rows that fired in at least one batch column) + output; silent rows of W are never read
        live_rows = float(self.pool.any(2).sum(1).double().mean().item())
        self.extra_bytes_per_timestep = 4.0 * live_rows * m + 4.0 * m * batch
        self.config = {"workload": f"BinaryArray[{batch},{k}] @ dense W[{k},{m}] fp32, {rate:.0%} spikes "
                                   "(binary_densemm transpose=True on tcgen05; events = active (source, column) pairs x m)"}

    def timestep(self, t):
        return self.be.binary_densemm(self.W, self.pool[t % 4], transpose=True)

    def reference_cuda_timestep(self, R, t):
        if getattr(self, "_ref", None) is None:
            self._ref = self.W.new_empty((self.pool.shape[2], self.W.shape[1]))
        return R.binary_densemm_transpose(self.W, self.pool[t % 4], self._ref)


class Jitc1M(Workload):
    name = "jitc_normal_1m"
    bytes_per_event = 8.0

    def __init__(self, torch, be, device, corder, rate=0.01):
        self.be = be
        n = 1_000_000
        self.corder = corder
        self.dominant_kernel = "jit_mv_scatter_kernel" if corder else "jit_mv_gather_kernel"
        self.m = be.JITCNormalR((1.0, 0.1, 1e-3, 123), shape=(n, n), corder=corder)
        self.pool = spike_pool(torch, n, rate, 5, device)
        self.events_per_timestep = float(self.pool.sum(1).double().mean().item()) * n * 1e-3
        self.extra_bytes_per_timestep = 0.0
        self.config = {"workload": f"spk @ JITCNormalR 1M x 1M, prob 1e-3, corder={corder} "
                                   f"({'scatter' if corder else 'gather'} kernel), {rate:.0%} spikes (BASELINE configs[3])"}

    def timestep(self, t):
        return self.be.BinaryArray(self.pool[t % POOL]) @ self.m

    def reference_cuda_timestep(self, R, t):
        if getattr(self, "_ref", None) is None:
            import torch
            dev = self.pool.device
            self._ref = (torch.tensor([1.0], device=dev), torch.tensor([0.1], device=dev), torch.empty(1_000_000, device=dev))
        a, b, out = self._ref
        return R.binary_jitnmv(a, b, 1e-3, self.pool[t % POOL], 123, shape=(1_000_000, 1_000_000), transpose=True,
                               corder=not self.corder, out=out)


class Jitc1MSharded(Workload):
    """BASELINE configs[3]: spk @ JITCNormalR 1 M x 1 M, prob 1e-3, post-sharded over the ranks with the
    spike exchange (STRONG scaling: the network is fixed, each rank regenerates only its post block --
    a column window of the scatter order or a row window of the gather order, the GLOBAL shape keys the
    RNG so the union of the shards is index-for-index the single-GPU matrix)."""
    name = "jitc_normal_1m_sharded"
    bytes_per_event = 8.0
    scaling = "strong"

    def __init__(self, torch, be, rank, world, device, gather=False, rate=0.01, emulate_world=0):
        from brainevent_b200 import dist as bd
        self.emulate = emulate_world      # (profiling) rank 0's window of an N-GPU run on ONE GPU
        if emulate_world:
            world, rank = emulate_world, 0
        self.world, n = world, 1_000_000
        self.gather = gather
        self.dominant_kernel = "jit_mv_gather_kernel" if gather else "jit_mv_scatter_kernel"
        bounds = bd.post_block_bounds(n, world)
        lo, hi = bounds[rank]
        self.n_local = hi - lo
        self.pool = spike_pool(torch, self.n_local, rate, 5 + rank, device)       # spikes of the neurons I own
        self.x = None
        if emulate_world:
            self.peers = bd.P2PSpikeExchange.emulate(n, world, device)
            self.x = self.peers[0]
            self.peer_pools = [spike_pool(torch, b[1] - b[0], rate, 5 + r, device) for r, b in enumerate(bounds)]
        elif world > 1 and os.environ.get("BE_B200_EXCHANGE", DEFAULT_EXCHANGE) == "p2p":
            try:
                self.x = bd.P2PSpikeExchange(n, device=device)
            except Exception as e:  # pragma: no cover
                sys.stderr.write(f"[bench] P2P exchange unavailable ({type(e).__name__}: {e}); using NCCL\n")
        if self.x is None:
            self.x = bd.SpikeExchange(n, device=device)
        # operator corder: True = gather order (rows = post), False = scatter order (cols = post)
        self.proj = bd.PostShardedProjection("jitc", n, n, self.x, family="normal", params=(1.0, 0.1), prob=1e-3,
                                             seed=123, corder=gather)
        act = sum(float(spike_pool(torch, b[1] - b[0], rate, 5 + r, device).sum(1).double().mean().item())
                  for r, b in enumerate(bounds))
        self.events_per_timestep = act * self.n_local * 1e-3                       # this rank's share
        self._total = act * n * 1e-3
        self.extra_bytes_per_timestep = 0.0
        self.launches_per_timestep = 3 if world > 1 else 2
        self.config = {"workload": f"spk @ JITCNormalR 1M x 1M, prob 1e-3, {'gather' if gather else 'scatter'} order, "
                                   f"{rate:.0%} spikes, post-sharded over {world} GPU(s) with the spike exchange "
                                   "(BASELINE configs[3]; strong scaling)"}

    def timestep(self, t):
        if self.emulate:
            for r in range(1, self.world):
                self.peers[r].publish_only(self.peer_pools[r][t % POOL])
        return self.proj.step(self.pool[t % POOL])

    def total_events_per_timestep(self):
        return self._total


class DenseSharded(Workload):
    """SURVEY 8(e) dense row: BinaryArray[k] @ W[k, m] (binary_densemv, transpose=True), W split by post columns over
    the ranks (each rank holds W[:, lo:hi]) with the spike exchange; strong scaling (the matrix is fixed, every rank
    draws its own column block from a generator seeded by the block's first column)."""
    name = "dense_mv_sharded"
    bytes_per_event = 4.0
    scaling = "strong"
    dominant_kernel = "dense_mv_t_kernel"

    def __init__(self, torch, be, rank, world, device, n=65_536, rate=0.01, emulate_world=0):
        from brainevent_b200 import dist as bd
        self.emulate = emulate_world
        if emulate_world:
            world, rank = emulate_world, 0
        self.world, self.rank, self.n = world, rank, n
        bounds = bd.post_block_bounds(n, world)
        lo, hi = bounds[rank]
        self.lo, self.hi, self.n_local = lo, hi, hi - lo
        self.pool = spike_pool(torch, self.n_local, rate, 11 + rank, device)
        self.x = None
        if emulate_world:
            self.peers = bd.P2PSpikeExchange.emulate(n, world, device)
            self.x = self.peers[0]
            self.peer_pools = [spike_pool(torch, b[1] - b[0], rate, 11 + r, device) for r, b in enumerate(bounds)]
        elif world > 1 and os.environ.get("BE_B200_EXCHANGE", DEFAULT_EXCHANGE) == "p2p":
            try:
                self.x = bd.P2PSpikeExchange(n, device=device)
            except Exception as e:  # pragma: no cover
                sys.stderr.write(f"[bench] P2P exchange unavailable ({type(e).__name__}: {e}); using NCCL\n")
        if self.x is None:
            self.x = bd.SpikeExchange(n, device=device)
        g = torch.Generator(device=device).manual_seed(7_000_003 + lo)
        self.W = torch.randn((n, self.n_local), device=device, generator=g)
        self.proj = bd.PostShardedProjection("dense", n, n, self.x, W=self.W)
        act = sum(float(spike_pool(torch, b[1] - b[0], rate, 11 + r, device).sum(1).double().mean().item())
                  for r, b in enumerate(bounds))
        self.events_per_timestep = act * self.n_local                              # this rank's share
        self._total = act * n
        self.extra_bytes_per_timestep = n / 8.0 + 4.0 * self.n_local
        self.launches_per_timestep = 3 if world > 1 else 2
        self.config = {"workload": f"BinaryArray[{n}] @ dense W[{n},{n}] fp32 (binary_densemv transpose=True), {rate:.0%} spikes, "
                                   f"W split by post columns over {world} GPU(s) with the spike exchange (SURVEY 8e; strong scaling)"}

    def timestep(self, t):
        if self.emulate:
            for r in range(1, self.world):
                self.peers[r].publish_only(self.peer_pools[r][t % POOL])
        return self.proj.step(self.pool[t % POOL])

    def total_events_per_timestep(self):
        return self._total


def parity_sharded_other(torch, dist, be, w2):
    """N > 1, before the JITC / dense sharded lines are timed: two timesteps of the measured step against NCCL
    all-gather of the bit-packed spikes + the UNSHARDED operator on the global shape (JITC: the 1 M x 1 M matrix
    without any window, sliced to this rank's post block afterwards -- this also proves the shard windows are
    index-for-index the single-GPU matrix; dense: binary_densemv on this rank's column block)."""
    from brainevent_b200 import dist as bd
    dev = w2.pool.device
    n = w2.proj.n_pre
    lo, hi = w2.proj.lo, w2.proj.hi
    xn = bd.SpikeExchange(n, device=dev)
    packer = bd.PostShardedProjection(w2.proj.kind, n, n, xn, **w2.proj.conn)
    max_err, max_ref = 0.0, 0.0
    for t in range(2):
        mine = w2.pool[t]
        out = w2.proj.step(mine).clone()
        words = xn.exchange(packer.pack_local(mine)).clone()
        ev = be.BitPackedBinary(packed=words, shape=(n,))
        if isinstance(w2, Jitc1MSharded):
            full = be.JITCNormalR((1.0, 0.1, 1e-3, 123), shape=(n, n), corder=not w2.gather)
            ref = (ev @ full)[lo:hi]
        else:
            ref = be.binary_densemv(w2.W, ev, transpose=True)
        max_err = max(max_err, float((out - ref).abs().max().item()))
        max_ref = max(max_ref, float(ref.abs().max().item()))
    timed_out = bool(w2.x.timed_out()) if hasattr(w2.x, "timed_out") else False
    ok = max_err <= 1e-5 * max(max_ref, 1.0) + 2e-5 and not timed_out
    v = torch.tensor([max_err, 0.0 if ok else 1.0], device=dev, dtype=torch.float64)
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    res = {"checked": True, "ok": bool(v[1].item() == 0.0), "timesteps": 2, "max_abs_err": float(v[0].item()),
           "against": "NCCL all_gather of bit-packed spikes + the unsharded operator, sliced to this rank's post block"}
    del packer, xn
    if not res["ok"]:
        raise RuntimeError(f"multi-GPU parity check failed ({w2.name}): {res}")
    return res


# ------------------------------------------------------------------------------------------------
# timing helpers
# ------------------------------------------------------------------------------------------------
def make_step(torch, wl, T, use_graph=True):
    """Returns (callable running one step = T timesteps, used_graph)."""
    def eager():
        out = None
        for t in range(T):
            out = wl.timestep(t)
        return out
    if not use_graph:
        return eager, False
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for t in range(2):
                wl.timestep(t)          # warm up allocations / workspaces on the side stream
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            eager()
        return g.replay, True
    except Exception as e:  # pragma: no cover - falls back to eager launches
        sys.stderr.write(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); eager launches\n")
        torch.cuda.synchronize()
        return eager, False


def time_steps(torch, dist, step, K, W, world, sampler=None):
    for _ in range(W):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), clocks


def kernel_profile(torch, L, wl, T, kernels_per_timestep_filter=None):
    """Average duration (ms) of the dominant kernel, from CUDA event pairs recorded around it by
    the library on the launching stream (eager launches, after warm-up)."""
    for t in range(3):
        wl.timestep(t)
    torch.cuda.synchronize()
    cap = 8 * T + 64
    L.check(L.lib().be_b200_profile_begin(cap), "profile_begin")
    for t in range(T):
        wl.timestep(t)
    torch.cuda.synchronize()
    buf = (ctypes.c_float * cap)()
    n = ctypes.c_int(0)
    L.check(L.lib().be_b200_profile_end(buf, cap, ctypes.byref(n)), "profile_end")
    ms = np.array(buf[: n.value], dtype=np.float64)
    return ms


def traffic_key(wl):
    k = wl.name
    if isinstance(wl, FcnSharded):
        if wl.world > 1:
            k += f"_w{wl.world}"
        if wl.scaling == "strong":
            k += "_full"
    elif isinstance(wl, (CubaJitc, Jitc1M)):
        k += "_scatter" if wl.corder else "_gather"
    elif isinstance(wl, CsrMv1M):
        k += ("_gather" if wl.gather else "_scatter") + ("_homo" if wl.weights.numel() == 1 else "_hetero")
    elif isinstance(wl, Jitc1MSharded):
        k += ("_gather" if wl.gather else "_scatter") + f"_w{wl.world}"
    elif isinstance(wl, DenseSharded):
        k += f"_w{wl.world}"
    return k


def measured_traffic(wl):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the dominant kernel from the committed
    ncu --set full capture of this workload (profiles/traffic.json, written by tools/summarize_ncu.py; multi-GPU
    shards are captured on one GPU with the peers' exchange buffers emulated locally).  None if not captured."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(traffic_key(wl))
        return e["bytes_per_launch"] if e else None
    except Exception:
        return None


def roofline(wl, kernel_ms, peak_gbs, peak_src, kernels_per_timestep=1, timestep_ms=None):
    """`timestep_ms`: the timed region's time per timestep (CUDA events around K graph-replayed steps).  A kernel cannot
    take longer than the timestep that contains it, so when the per-launch event pairs of the eager profiling pass (which
    carry the event-record gaps of an eager launch) exceed it, the timestep bounds the kernel's duration from above."""
    per_launch_bytes = wl.algorithmic_bytes_per_timestep() / kernels_per_timestep
    avg_ms = float(np.mean(kernel_ms)) if len(kernel_ms) else float("nan")
    src = "CUDA event pair around every launch (eager pass after the timed region)"
    if timestep_ms is not None and timestep_ms / kernels_per_timestep < avg_ms:
        avg_ms = timestep_ms / kernels_per_timestep
        src = ("the timed region itself: time per graph-replayed timestep (an upper bound of the kernel's duration; the eager "
               f"event pairs read {float(np.mean(kernel_ms)):.5f} ms)")
    achieved = per_launch_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else float("nan")
    return {"bound": "hbm", "kernel": wl.dominant_kernel, "achieved": round(achieved, 1), "peak": peak_gbs,
            "unit": "GB/s", "frac": round(achieved / peak_gbs, 4), "traffic": measured_traffic(wl),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": round(per_launch_bytes),
            "kernel_ms_avg": round(avg_ms, 5), "kernel_launches_timed": int(len(kernel_ms)), "kernel_ms_source": src}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's numba kernels on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_fcn_sample(n_pre=100_000, n_post=FCN_NEURONS_PER_GPU, n_conn=FCN_CONN, rate=RATE, reps=10):
    import oracle as O
    rng = np.random.default_rng(0)
    idx = rng.integers(0, n_post, (n_pre, n_conn), dtype=np.int32)
    w = rng.random((n_pre, n_conn), dtype=np.float32)
    spikes = [rng.random(n_pre) < rate for _ in range(4)]
    O.binary_fcnmv(w, idx, spikes[0], shape=(n_pre, n_post), transpose=True)
    times = []
    for r in range(reps):
        s = spikes[r % 4]
        t0 = time.perf_counter()
        O.binary_fcnmv(w, idx, s, shape=(n_pre, n_post), transpose=True)
        times.append((time.perf_counter() - t0, int(s.sum()) * n_conn))
    return times, (f"{n_pre} of the pre rows x {n_conn} synapses into {n_post} post neurons, {rate:.0%} spikes, "
                   f"{reps} calls (scatter kernels are serial in the reference: _fcn/binary.py:167-175)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    O.build()
    K, W = args.steps, args.warmup
    world = max(args.gpus, 1)
    n_total = FCN_NEURONS_PER_GPU * world
    reps_per_step = 40      # ~0.15 s of CPU work per step
    times, sample = cpu_fcn_sample(n_post=n_total, reps=(K + W) * reps_per_step)
    times = times[W * reps_per_step:]
    total_t = sum(t for t, _ in times)
    total_ev = sum(e for _, e in times)
    value = total_ev / total_t
    cfg = fcn_config(n_total, FCN_CONN, RATE, world, args.timesteps)
    line = {"impl": "reference", "metric": "synaptic events/sec", "value": value, "unit": "events/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * total_t / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "reference_arm": "reference algorithm (oracle port of _fcn/binary.py:156-253, the numba scatter kernel) on the "
                             "host CPU, one core (the kernel is serial in the reference), bounded sample of the pre rows",
            "cpu_baseline": {"value": value, "unit": "events/s", "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
def measure_workload(torch, dist, L, wl, K, W, T, world, peak, peak_src, sampler=None, use_graph=True):
    step, graphed = make_step(torch, wl, T, use_graph)
    ms, clocks = time_steps(torch, dist, step, K, W, world, sampler)
    per_step_ms = ms / K
    total_events = getattr(wl, "total_events_per_timestep", lambda: wl.events_per_timestep)() * T
    value = total_events / (per_step_ms * 1e-3)
    kms = kernel_profile(torch, L, wl, min(T, 50))
    # keep only the dominant kernel's records when an op launches several profiled kernels
    rl = roofline(wl, kms, peak, peak_src, wl.profiled_kernels_per_timestep, per_step_ms / T)
    return {"value": value, "ms_per_step": per_step_ms, "graph": graphed, "roofline": rl, "clocks": clocks,
            "events_per_step": total_events}


def reference_cuda_us(torch, wl, T):
    """Per-timestep time (us) of the reference's OWN CUDA kernels for this workload (unmodified sources recompiled
    for sm_100a into oracle/_ref; test/bench infrastructure), same device buffers and spike vectors, replayed from a
    CUDA graph of T timesteps like our side.  None when the library is absent or the workload has no twin."""
    if not hasattr(wl, "reference_cuda_timestep"):
        return None
    try:
        from oracle import ref_cuda as R
        if wl.reference_cuda_timestep(R, 0) is None:
            return None
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            wl.reference_cuda_timestep(R, 1)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for t in range(T):
                wl.reference_cuda_timestep(R, t)
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) * 1e3 / (reps * T), 2)
    except Exception as e:  # pragma: no cover - the comparison library is optional
        sys.stderr.write(f"[bench] reference CUDA timing unavailable for {wl.name} ({type(e).__name__}: {e})\n")
        torch.cuda.synchronize()
        return None


_REAL_STDOUT = None


def _claim_stdout():
    """Keep fd 1 for the single JSON line: libraries that print to stdout (NCCL's version banner,
    torchrun notices) are redirected to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def e2e_leg(torch, dist, L, wl, world, device, K, T, numa_bound):
    """Same metric end to end: per timestep this rank's spike block comes from pinned host memory and its output block
    goes back to pinned host memory, through the C ABI (N = 1: be_b200_projection_submit_host/_wait; N > 1:
    be_b200_sharded_run_host, csrc/sharded.cu -- the whole block of timesteps is ONE C call per rank)."""
    n_in, n_out = wl.n_local, wl.n_local
    host_spk = [torch.empty(n_in, dtype=torch.uint8).pin_memory() for _ in range(POOL)]
    for i in range(POOL):
        host_spk[i].copy_(wl.pool[i].to(torch.uint8))
    n_slots = int(os.environ.get("BE_B200_E2E_SLOTS", "8"))       # timesteps in flight at N > 1 (N = 2: 137 us with 4, 126 us with 8)
    host_out = [torch.empty(n_out, dtype=torch.float32).pin_memory() for _ in range(max(4, n_slots))]
    Te = min(T, 50)
    sync_value = None
    lib = L.lib()
    if world == 1:
        proj = lib.be_b200_projection_adopt_fcn(ctypes.c_void_p(wl.weights.data_ptr()), 0,
                                                ctypes.c_void_p(wl.indices.data_ptr()), wl.n_total, wl.n_conn, wl.n_local)
        assert proj, lib.be_b200_last_error()
        proj = ctypes.c_void_p(proj)

        def e2e_step_sync():
            for t in range(Te):
                rc = lib.be_b200_projection_apply_host(proj, ctypes.c_void_p(host_spk[t % POOL].data_ptr()),
                                                       ctypes.c_void_p(host_out[0].data_ptr()))
                assert rc == 0

        def e2e_step():
            # pipelined C-ABI entry: copy-in / hot path / copy-out of neighbouring timesteps overlap;
            # every timestep still moves its own inputs in and its own result out
            for t in range(Te):
                rc = lib.be_b200_projection_submit_host(proj, ctypes.c_void_p(host_spk[t % POOL].data_ptr()),
                                                        ctypes.c_void_p(host_out[t % 4].data_ptr()))
                assert rc == 0
            assert lib.be_b200_projection_wait(proj) == 0

        for _ in range(2):
            e2e_step_sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            e2e_step_sync()
        sync_value = wl.total_events_per_timestep() * Te * 3 / (time.perf_counter() - t0)
        entry = ("be_b200_projection_submit_host/_wait (C ABI, 4 calls in flight; `sync_value` is the blocking "
                 "be_b200_projection_apply_host)")
    else:
        from brainevent_b200 import dist as bd
        pipe = bd.HostStepPipeline(wl.proj, n_in, slots=n_slots)

        def e2e_step():
            pipe.run(Te, host_spk, host_out)
        entry = (f"be_b200_sharded_run_host (C ABI, csrc/sharded.cu: one call per block of timesteps, {n_slots} in flight, a CUDA "
                 "graph of the fused step per slot)" if pipe.native else
                 "HostStepPipeline.submit/wait over PostShardedProjection.step (Python-driven, NCCL exchange)")
    for _ in range(2):
        e2e_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    Ke = max(3, min(K, 10))
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_events = wl.total_events_per_timestep() * Te * Ke
    # the host link's own floor: the same bytes per timestep and rank, every rank at once, copies only (no kernels)
    d_i = torch.empty(n_in, dtype=torch.uint8, device=device)
    d_o = torch.empty(n_out, dtype=torch.float32, device=device)
    s_i, s_o = torch.cuda.Stream(), torch.cuda.Stream()

    def copies(reps):
        for r in range(reps):
            with torch.cuda.stream(s_i):
                d_i.copy_(host_spk[r % POOL], non_blocking=True)
            with torch.cuda.stream(s_o):
                host_out[r % 4].copy_(d_o, non_blocking=True)
        torch.cuda.synchronize()
    copies(20)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    copies(200)
    ct = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ct, op=dist.ReduceOp.MAX)
    copy_floor_us = float(ct.item()) * 1e6 / 200
    e2e = {"value": e2e_events / float(dt.item()), "unit": "events/s",
           "h2d_bytes_per_step": int(n_in * T), "d2h_bytes_per_step": int(4 * n_out * T),
           "us_per_timestep": round(float(dt.item()) * 1e6 / (Te * Ke), 2),
           "copy_only_us_per_timestep": round(copy_floor_us, 2),
           "host_link_GBps_all_ranks": round((n_in + 4 * n_out) * world / copy_floor_us / 1e3, 1),
           "numa_bound": bool(numa_bound),
           "host_cpus": len(os.sched_getaffinity(0)),
           "note": f"per timestep and rank: pinned host spikes (uint8[{n_in}]) -> device, hot path, fp32[{n_out}] result -> "
                   f"pinned host; timed over {Ke}x{Te} timesteps with the host clock (max over ranks), bytes scaled to "
                   f"{T} timesteps/step; copy_only_us_per_timestep = the same host<->device bytes with no kernel at all, every "
                   f"rank at once (the host link's floor for this leg; tools/e2e_host_probe.py); entry point: " + entry}
    if sync_value is not None:
        e2e["sync_value"] = sync_value
    if world == 1:
        lib.be_b200_projection_destroy(proj)
    else:
        pipe.close()        # also raises on every rank if a spike-exchange wait ever timed out
        del pipe
    return e2e


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fcn", choices=["fcn", "fcn_full", "coba", "cuba_jitc", "cuba_jitc_gather", "csr_spmm", "jitc_1m",
                                                           "csr_gather", "csr_gather_hetero", "csr_scatter", "dense_mm", "jitc_1m_gather",
                                                           "jitc_1m_sharded", "jitc_1m_sharded_gather", "dense_sharded"])
    ap.add_argument("--timesteps", type=int, default=100, help="simulated timesteps per step")
    ap.add_argument("--no-extra", action="store_true", help="skip the strong-scaling leg and the other BASELINE configs")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="(profiling) rank 0's shard of an N-GPU post-sharded fcn run on ONE GPU, the peers' exchange buffers emulated locally")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    import brainevent_b200 as be
    from brainevent_b200 import _lib as L
    L.check(L.lib().be_b200_check_device(), "check_device")
    if os.environ.get("BE_B200_NUMA_BIND", "1") != "0":
        # pinned host buffers of the e2e leg and the copy-issuing thread go to the GPU's NUMA node
        from brainevent_b200.dist import bind_to_gpu_numa
        numa_bound = bind_to_gpu_numa(local_rank)
    else:
        numa_bound = False
    peak, peak_src = peaks()
    K, W, T = args.steps, args.warmup, args.timesteps

    is_fcn = args.workload in ("fcn", "fcn_full")
    if args.workload == "fcn":
        n_total = FCN_NEURONS_PER_GPU * (args.emulate_world or world)
        wl = FcnSharded(torch, be, rank, world, device, n_total, T=T, emulate_world=args.emulate_world)
    elif args.workload == "fcn_full":
        wl = FcnSharded(torch, be, rank, world, device, C5_NEURONS, T=T, scaling="strong", emulate_world=args.emulate_world)
    elif args.workload == "coba":
        wl = Coba(torch, be, device)
    elif args.workload in ("cuba_jitc", "cuba_jitc_gather"):
        wl = CubaJitc(torch, be, device, corder=not args.workload.endswith("gather"))
    elif args.workload == "csr_spmm":
        wl = CsrSpmm(torch, be, device)
    elif args.workload in ("csr_gather", "csr_gather_hetero"):
        wl = CsrMv1M(torch, be, device, gather=True, homo=not args.workload.endswith("hetero"))
    elif args.workload == "csr_scatter":
        wl = CsrMv1M(torch, be, device, gather=False)
    elif args.workload == "dense_mm":
        wl = DenseMm(torch, be, device)
    elif args.workload == "jitc_1m_gather":
        wl = Jitc1M(torch, be, device, corder=False)
    elif args.workload in ("jitc_1m_sharded", "jitc_1m_sharded_gather"):
        wl = Jitc1MSharded(torch, be, rank, world, device, gather=args.workload.endswith("gather"), emulate_world=args.emulate_world)
    elif args.workload == "dense_sharded":
        wl = DenseSharded(torch, be, rank, world, device, emulate_world=args.emulate_world)
    else:
        wl = Jitc1M(torch, be, device, corder=True)

    # ---- multi-GPU parity: checked BEFORE anything is timed; the run fails if a block differs ----------
    parity = None
    if world > 1 and is_fcn and not args.emulate_world:
        parity = parity_block(torch, dist, be, wl)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    res = measure_workload(torch, dist, L, wl, K, W, T, world, peak, peak_src, sampler, not args.no_graph)
    if world > 1 and hasattr(getattr(wl, "x", None), "check"):
        wl.x.check()        # a flag wait that gave up is fatal on every rank (results would be stale)
    ref_main_us = None
    if world == 1 and not args.no_extra and not args.emulate_world and not (is_fcn and wl.n_total > 2_000_000):
        ref_main_us = reference_cuda_us(torch, wl, 16)     # (not on the 1e10-synapse network: the reference indexes in int32)

    # ---- e2e: host buffers in, host buffers out, every timestep -------------------------------
    e2e = None
    if is_fcn and not args.emulate_world and not args.no_e2e:
        e2e = e2e_leg(torch, dist, L, wl, world, device, K, T, numa_bound)

    # ---- strong scaling: the FULL 10 M x 1000 network on these N GPUs (all ranks take part) -----------
    strong = None
    if args.workload == "fcn" and not args.no_extra and not args.emulate_world:
        if wl.n_total == C5_NEURONS:
            strong = {"same_as_main_line": True, "value": res["value"], "us_per_timestep": round(res["ms_per_step"] * 1e3 / T, 2)}
        else:
            wl.release()
            try:
                w_full = FcnSharded(torch, be, rank, world, device, C5_NEURONS, T=T, scaling="strong")
                Ts = min(T, 50)
                r_full = measure_workload(torch, dist, L, w_full, max(3, min(K, 5)), 3, Ts, world, peak, peak_src)
                if world > 1 and hasattr(w_full.x, "check"):
                    w_full.x.check()
                strong = {"workload": w_full.config["workload"], "value": r_full["value"], "unit": "events/s",
                          "us_per_timestep": round(r_full["ms_per_step"] * 1e3 / Ts, 2), "roofline": r_full["roofline"],
                          "bytes_resident_per_gpu": int(w_full.indices.numel() * 8),
                          "note": "strong scaling: the same 10 000 128 x 1000 matrix at every N (counter-based chunk seeds); "
                                  "N = 8 is the main line itself"}
                w_full.release()
                del w_full
            except Exception as e:  # pragma: no cover
                strong = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.synchronize()
            torch.cuda.empty_cache()

    # ---- other BASELINE configs ---------------------------------------------------------------------------
    others = []
    if args.workload == "fcn" and not args.no_extra and not args.emulate_world:
        if hasattr(wl, "release"):
            wl.release()
        if world == 1:
            makers = (lambda: Coba(torch, be, device),
                      lambda: CubaJitc(torch, be, device, corder=True),
                      lambda: CubaJitc(torch, be, device, corder=False),
                      lambda: Jitc1M(torch, be, device, corder=True),
                      lambda: Jitc1M(torch, be, device, corder=False),
                      lambda: CsrMv1M(torch, be, device, gather=False),
                      lambda: CsrMv1M(torch, be, device, gather=True, homo=True),
                      lambda: CsrMv1M(torch, be, device, gather=True, homo=False),
                      lambda: CsrSpmm(torch, be, device),
                      lambda: DenseMm(torch, be, device))
        else:
            # BASELINE configs[3] at N GPUs: the fixed 1 M x 1 M JITCNormalR matrix, post-sharded (strong scaling)
            makers = (lambda: Jitc1MSharded(torch, be, rank, world, device, gather=False),
                      lambda: Jitc1MSharded(torch, be, rank, world, device, gather=True),
                      lambda: DenseSharded(torch, be, rank, world, device))
        for make in makers:
            try:
                w2 = make()
                heavy = isinstance(w2, (CsrSpmm, DenseMm)) or (isinstance(w2, Jitc1M) and not w2.corder) or \
                    (isinstance(w2, Jitc1MSharded) and w2.gather)
                T2 = 4 if heavy else 50
                par2 = parity_sharded_other(torch, dist, be, w2) if world > 1 else None
                r2 = measure_workload(torch, dist, L, w2, 3 if heavy else 5, 3, T2, world, peak, peak_src)
                if world > 1 and hasattr(getattr(w2, "x", None), "check"):
                    w2.x.check()
                entry = {"workload": w2.config["workload"], "value": r2["value"], "unit": "events/s",
                         "ms_per_timestep": r2["ms_per_step"] / T2, "cuda_graph": r2["graph"], "roofline": r2["roofline"]}
                if par2 is not None:
                    entry["parity"] = par2
                if world == 1:
                    ref_us = reference_cuda_us(torch, w2, T2 if heavy else 16)
                    entry["reference_cuda_us"] = ref_us
                    if ref_us:
                        entry["vs_reference_cuda"] = round(ref_us / (entry["ms_per_timestep"] * 1e3), 2)
                others.append(entry)
                del w2
                torch.cuda.empty_cache()
            except Exception as e:  # pragma: no cover
                others.append({"workload": "failed", "error": f"{type(e).__name__}: {e}"})
                torch.cuda.synchronize()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu and args.workload == "fcn":
        times, sample = cpu_fcn_sample(reps=3000)      # ~10 s of CPU work
        tt, ee = sum(t for t, _ in times), sum(e for _, e in times)
        cpu = {"value": ee / tt, "unit": "events/s", "cores": 1, "kind": "port", "sample": sample}

    cfg = dict(wl.config)
    if "timesteps_per_step" not in cfg:
        cfg.update({"timesteps_per_step": T,
                    "l2_policy": f"inputs larger than L2: {POOL} distinct spike vectors cycle through the timesteps",
                    "parallelism": f"post-sharded x{world}" if world > 1 else "single GPU"})
    line = {"metric": "synaptic events/sec", "value": res["value"], "unit": "events/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cuda_graph": res["graph"], "roofline": res["roofline"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(wl.launches_per_timestep * T * K), "clocks": res["clocks"],
            "events_per_step": res["events_per_step"]}
    if ref_main_us:
        line["reference_cuda_us"] = ref_main_us
        line["vs_reference_cuda"] = round(ref_main_us / (res["ms_per_step"] * 1e3 / T), 2)
    if parity is not None:
        line["parity"] = parity
    if strong is not None:
        line["strong_scaling"] = strong
    if others:
        line["other_workloads"] = others
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
