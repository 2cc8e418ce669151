#!/usr/bin/env python
"""bench.py -- synaptic events / second of the event-driven propagation path on B200(s).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line on rank 0.  A *step* is one block of T simulated timesteps (default 100) of ONE projection:
every timestep takes a fresh synthetic spike vector through the hot path
(bit-pack + ordered compaction -> [all-gather of packed words when N > 1] -> active-row scatter).

Default workload `fcn` (BASELINE.json configs[4], weak scaling): FixedNumPerPre, 1.25 M neurons per
GPU x 1000 synapses per neuron, fp32 heterogeneous weights, 1 % spikes per timestep, post-sharded:
at N = 8 this is the named 10 M-neuron x 1000 network.  Other configs: --workload
{coba, cuba_jitc, csr_spmm, jitc_1m}; at N = 1 their short results ride along under "other_workloads".

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


def fcn_workload_text(n_total, n_conn, rate, shards):
    return (f"FixedNumPerPre {n_total} neurons x {n_conn} synapses/neuron, fp32 hetero weights, "
            f"{rate:.0%} spikes/timestep, post-sharded over {shards} GPU(s) "
            f"({n_total // shards} post neurons and {n_conn // shards} synapses/pre-neuron per GPU, "
            "targets stratified evenly over post blocks), BASELINE configs[4] at 8 GPUs")


class FcnWeak(Workload):
    """FixedNumPerPre, post-sharded, weak scaling (configs[4] at N = 8)."""
    name = "fcn_perpre_postsharded"
    dominant_kernel = "scatter_mv_kernel"

    def __init__(self, torch, be, rank, world, device, neurons_per_gpu=FCN_NEURONS_PER_GPU, n_conn=FCN_CONN, rate=RATE,
                 emulate_world=0, exchange=None):
        from brainevent_b200 import dist as bd
        self.torch, self.be, self.world, self.rank = torch, be, world, rank
        self.emulate = emulate_world      # profile one rank's shard of an `emulate_world`-GPU run on a single GPU
        shards = emulate_world if emulate_world else world
        self.n_local = neurons_per_gpu
        self.n_total = neurons_per_gpu * shards
        self.conn_local = n_conn // shards          # targets stratified evenly over the post blocks
        g = torch.Generator(device=device).manual_seed(1234 + rank)
        # this rank's incoming synapses: for every pre neuron of the whole network, conn_local
        # targets inside the local post block (local ids) with U(0,1) fp32 weights
        self.indices = torch.randint(0, self.n_local, (self.n_total, self.conn_local), device=device,
                                     dtype=torch.int32, generator=g)
        self.weights = torch.rand((self.n_total, self.conn_local), device=device, generator=g)
        self.pool = spike_pool(torch, self.n_local, rate, 99 + rank, device)     # spikes of the neurons I own
        self.exchange_kind = "nccl all_gather"
        self.x = None
        exchange = exchange or os.environ.get("BE_B200_EXCHANGE", DEFAULT_EXCHANGE)
        if world > 1 and exchange == "p2p":
            try:
                self.x = bd.P2PSpikeExchange(self.n_total, device=device)
                self.exchange_kind = "fused bit-pack + NVLink peer-memory all-gather (csrc/exchange.cu: pack into the own symmetric buffer, flag the peers, pull their words)"
            except Exception as e:  # pragma: no cover - symmetric memory unavailable: NCCL path
                sys.stderr.write(f"[bench] P2P exchange unavailable ({type(e).__name__}: {e}); using NCCL\n")
        if self.x is None:
            self.x = bd.SpikeExchange(self.n_total, device=device)
        self.proj = bd.PostShardedProjection("fcn", self.n_total, self.n_total, self.x,
                                             weights=self.weights, indices=self.indices)
        self.launches_per_timestep = 2 if world == 1 else 3
        # events: every rank sees ALL spikes; count from the global pool (same seeds on every rank)
        if emulate_world:
            # no exchange: the global packed spike vector of every timestep is precomputed
            gpool = torch.cat([spike_pool(torch, self.n_local, rate, 99 + r, device) for r in range(shards)], 1)
            self.packed_pool = [be.BitPackedBinary(gpool[i]) for i in range(POOL)]
            per_t = gpool.sum(1).double()
        elif world == 1:
            per_t = self.pool.sum(1).double()
        else:
            per_t = sum(spike_pool(torch, self.n_local, rate, 99 + r, device).sum(1).double() for r in range(world))
        self.active_per_timestep = float(per_t.mean().item())
        self.events_per_timestep = self.active_per_timestep * self.conn_local          # per rank
        self.extra_bytes_per_timestep = 4.0 * self.active_per_timestep
        self.config = {"workload": fcn_workload_text(self.n_total, n_conn, rate, shards),
                       "neurons": self.n_total, "synapses_per_neuron": n_conn, "spike_rate": rate,
                       "spike_exchange": self.exchange_kind if world > 1 else "none (single GPU)"}

    def timestep(self, t):
        if self.emulate:
            return self.be.binary_fcnmv(self.weights, self.indices, self.packed_pool[t % POOL],
                                        shape=(self.n_total, self.n_local), transpose=True)
        if self.world == 1:
            return self.be.binary_fcnmv(self.weights, self.indices, self.pool[t % POOL],
                                        shape=(self.n_total, self.n_local), transpose=True)
        return self.proj.step(self.pool[t % POOL])

    def total_events_per_timestep(self):
        return self.events_per_timestep * self.world


class Coba(Workload):
    name = "coba_4000_csr"
    dominant_kernel = "scatter_mv_kernel"
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
        self.launches_per_timestep = 4
        self.profiled_kernels_per_timestep = 2
        self.config = {"workload": "COBA_2005: 4000-neuron E/I network, 80 synapses/neuron CSR, homogeneous weights, "
                                   "0.5% spikes/timestep, BinaryArray @ CSR (BASELINE configs[0])"}

    def timestep(self, t):
        out = None
        for csr, pool in self.mats:
            out = self.be.BinaryArray(pool[t % POOL]) @ csr
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


class CsrSpmm(Workload):
    name = "csr_1m_spmm_b64"
    dominant_kernel = "scatter_mm_kernel"

    def __init__(self, torch, be, device, rate=0.01, batch=64):
        self.be = be
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


class Jitc1MSharded(Workload):
    """BASELINE configs[3]: spk @ JITCNormalR 1 M x 1 M, prob 1e-3, post-sharded over the ranks with the
    spike exchange (STRONG scaling: the network is fixed, each rank regenerates only its post block --
    a column window of the scatter order or a row window of the gather order, the GLOBAL shape keys the
    RNG so the union of the shards is index-for-index the single-GPU matrix)."""
    name = "jitc_normal_1m_sharded"
    bytes_per_event = 8.0
    scaling = "strong"

    def __init__(self, torch, be, rank, world, device, gather=False, rate=0.01):
        from brainevent_b200 import dist as bd
        self.world, n = world, 1_000_000
        self.dominant_kernel = "jit_mv_gather_kernel" if gather else "jit_mv_scatter_kernel"
        bounds = bd.post_block_bounds(n, world)
        lo, hi = bounds[rank]
        self.n_local = hi - lo
        self.pool = spike_pool(torch, self.n_local, rate, 5 + rank, device)       # spikes of the neurons I own
        self.x = None
        if world > 1 and os.environ.get("BE_B200_EXCHANGE", DEFAULT_EXCHANGE) == "p2p":
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
        return self.proj.step(self.pool[t % POOL])

    def total_events_per_timestep(self):
        return self._total


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


def measured_traffic(wl):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if this
    workload has not been captured)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        key = wl.name + ("_shard8" if getattr(wl, "emulate", 0) == 8 else "")
        if getattr(wl, "world", 1) > 1:
            return None
        return t[key]["bytes_per_launch"] if key in t else None
    except Exception:
        return None


def roofline(wl, kernel_ms, peak_gbs, peak_src, kernels_per_timestep=1):
    per_launch_bytes = wl.algorithmic_bytes_per_timestep() / kernels_per_timestep
    avg_ms = float(np.mean(kernel_ms)) if len(kernel_ms) else float("nan")
    achieved = per_launch_bytes / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else float("nan")
    return {"bound": "hbm", "kernel": wl.dominant_kernel, "achieved": round(achieved, 1), "peak": peak_gbs,
            "unit": "GB/s", "frac": round(achieved / peak_gbs, 4), "traffic": measured_traffic(wl),
            "peak_source": peak_src, "algorithmic_bytes_per_launch": round(per_launch_bytes),
            "kernel_ms_avg": round(avg_ms, 5), "kernel_launches_timed": int(len(kernel_ms))}


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
    reps_per_step = 4
    times, sample = cpu_fcn_sample(reps=(K + W) * reps_per_step)
    times = times[W * reps_per_step:]
    total_t = sum(t for t, _ in times)
    total_ev = sum(e for _, e in times)
    value = total_ev / total_t
    line = {"impl": "reference", "metric": "synaptic events/sec", "value": value, "unit": "events/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * total_t / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": fcn_workload_text(FCN_NEURONS_PER_GPU * max(args.gpus, 1), FCN_CONN, RATE, max(args.gpus, 1)),
                       "neurons": FCN_NEURONS_PER_GPU * max(args.gpus, 1), "synapses_per_neuron": FCN_CONN, "spike_rate": RATE,
                       "reference_arm": "reference algorithm (oracle port of _fcn/binary.py:156-253) on the host CPU, bounded "
                                        "sample of the pre rows of the same projection"},
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
    rl = roofline(wl, kms, peak, peak_src, wl.profiled_kernels_per_timestep)
    return {"value": value, "ms_per_step": per_step_ms, "graph": graphed, "roofline": rl, "clocks": clocks,
            "events_per_step": total_events}


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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fcn", choices=["fcn", "coba", "cuba_jitc", "csr_spmm", "jitc_1m", "csr_gather", "csr_scatter", "dense_mm", "jitc_1m_gather",
                                                           "jitc_1m_sharded", "jitc_1m_sharded_gather"])
    ap.add_argument("--timesteps", type=int, default=100, help="simulated timesteps per step")
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configs at N=1")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--emulate-world", type=int, default=0,
                    help="(profiling) run rank 0's shard of an N-GPU post-sharded fcn run on one GPU, no exchange")
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

    if args.workload == "fcn":
        wl = FcnWeak(torch, be, rank, world, device, emulate_world=args.emulate_world)
    elif args.workload == "coba":
        wl = Coba(torch, be, device)
    elif args.workload == "cuba_jitc":
        wl = CubaJitc(torch, be, device, corder=True)
    elif args.workload == "csr_spmm":
        wl = CsrSpmm(torch, be, device)
    elif args.workload == "csr_gather":
        wl = CsrMv1M(torch, be, device, gather=True, homo=True)
    elif args.workload == "csr_scatter":
        wl = CsrMv1M(torch, be, device, gather=False)
    elif args.workload == "dense_mm":
        wl = DenseMm(torch, be, device)
    elif args.workload == "jitc_1m_gather":
        wl = Jitc1M(torch, be, device, corder=False)
    elif args.workload in ("jitc_1m_sharded", "jitc_1m_sharded_gather"):
        wl = Jitc1MSharded(torch, be, rank, world, device, gather=args.workload.endswith("gather"))
    else:
        wl = Jitc1M(torch, be, device, corder=True)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    res = measure_workload(torch, dist, L, wl, K, W, T, world, peak, peak_src, sampler, not args.no_graph)
    if world > 1 and hasattr(getattr(wl, "x", None), "timed_out"):
        # the peer-memory exchange bounds its flag waits; a wait that gave up means a peer never
        # published (results would be stale): fall back to NCCL on every rank and measure again
        bad = torch.tensor([int(wl.x.timed_out())], device=device)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()) and args.workload != "fcn":
            raise RuntimeError("peer-memory spike exchange timed out (a peer never published); rerun with BE_B200_EXCHANGE=nccl")
        if int(bad.item()):
            sys.stderr.write("[bench] peer-memory exchange timed out; re-measuring with the NCCL all-gather\n")
            del wl
            torch.cuda.empty_cache()
            wl = FcnWeak(torch, be, rank, world, device, emulate_world=args.emulate_world, exchange="nccl")
            sampler = ClockSampler(local_rank) if rank == 0 else None
            res = measure_workload(torch, dist, L, wl, K, W, T, world, peak, peak_src, sampler, not args.no_graph)

    # ---- e2e: host buffers in, host buffers out, every timestep -------------------------------
    e2e = None
    if args.workload == "fcn" and not args.emulate_world:
        n_in, n_out = wl.n_local, wl.n_local
        host_spk = [torch.empty(n_in, dtype=torch.uint8).pin_memory() for _ in range(POOL)]
        for i in range(POOL):
            host_spk[i].copy_(wl.pool[i].to(torch.uint8))
        host_out = [torch.empty(n_out, dtype=torch.float32).pin_memory() for _ in range(4)]
        Te = min(T, 50)
        sync_value = None
        if world == 1:
            lib = L.lib()
            proj = lib.be_b200_projection_adopt_fcn(ctypes.c_void_p(wl.weights.data_ptr()), 0,
                                                    ctypes.c_void_p(wl.indices.data_ptr()), wl.n_total,
                                                    wl.conn_local, wl.n_local)
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
        else:
            from brainevent_b200 import dist as bd
            pipe = bd.HostStepPipeline(wl.proj, n_in, slots=4)

            def e2e_step():
                # pipelined like the single-GPU C-ABI entry: four timesteps in flight per rank, each moving
                # its own spike block in and its own output block out
                for t in range(Te):
                    pipe.submit(host_spk[t % POOL], host_out[t % 4])
                pipe.wait()
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
        e2e = {"value": e2e_events / float(dt.item()), "unit": "events/s",
               "h2d_bytes_per_step": int(n_in * T), "d2h_bytes_per_step": int(4 * n_out * T),
               "numa_bound": bool(numa_bound),
               "note": f"per timestep: pinned host spikes (uint8[{n_in}]) -> device, hot path, fp32[{n_out}] result -> "
                       f"pinned host, stream sync; timed over {Ke}x{Te} timesteps with the host clock, bytes scaled to "
                       f"{T} timesteps/step; entry point: " +
                       ("be_b200_projection_submit_host/_wait (C ABI, 4 calls in flight; `sync_value` is the "
                        "blocking be_b200_projection_apply_host)" if world == 1 else
                        "HostStepPipeline.submit/wait over PostShardedProjection.step (per rank: its own spike block in, its "
                        "own output block out, 4 timesteps in flight)")}
        if sync_value is not None:
            e2e["sync_value"] = sync_value
        if world == 1:
            L.lib().be_b200_projection_destroy(proj)
        else:
            pipe.close()        # graphs that captured NCCL collectives must go before the group is destroyed
            del pipe, e2e_step

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu and args.workload == "fcn":
        times, sample = cpu_fcn_sample(reps=12)
        tt, ee = sum(t for t, _ in times), sum(e for _, e in times)
        cpu = {"value": ee / tt, "unit": "events/s", "cores": 1, "kind": "port", "sample": sample}

    others = []
    if world == 1 and not args.no_extra and args.workload == "fcn":
        del wl.weights, wl.indices, wl.proj
        torch.cuda.empty_cache()
        for make in (lambda: Coba(torch, be, device),
                     lambda: CubaJitc(torch, be, device, corder=True),
                     lambda: CubaJitc(torch, be, device, corder=False),
                     lambda: Jitc1M(torch, be, device, corder=True),
                     lambda: Jitc1M(torch, be, device, corder=False),
                     lambda: CsrMv1M(torch, be, device, gather=False),
                     lambda: CsrMv1M(torch, be, device, gather=True, homo=True),
                     lambda: CsrMv1M(torch, be, device, gather=True, homo=False),
                     lambda: CsrSpmm(torch, be, device),
                     lambda: DenseMm(torch, be, device)):
            try:
                w2 = make()
                heavy = isinstance(w2, (CsrSpmm, DenseMm)) or (isinstance(w2, Jitc1M) and not w2.corder)
                r2 = measure_workload(torch, dist, L, w2, 3 if heavy else 5, 3, 4 if heavy else 50, 1, peak, peak_src)
                others.append({"workload": w2.config["workload"], "value": r2["value"], "unit": "events/s",
                               "ms_per_timestep": r2["ms_per_step"] / (4 if heavy else 50), "roofline": r2["roofline"]})
                del w2
                torch.cuda.empty_cache()
            except Exception as e:  # pragma: no cover
                others.append({"workload": "failed", "error": f"{type(e).__name__}: {e}"})

    cfg = dict(wl.config)
    cfg.update({"timesteps_per_step": T, "cuda_graph": res["graph"],
                "l2_policy": f"inputs larger than L2: {POOL} distinct spike vectors cycle, so the active-row "
                             "working set of one cycle (~1.6 GB of a 10 GB matrix) exceeds the 126 MB L2",
                "parallelism": f"post-sharded x{world}" if world > 1 else "single GPU"})
    line = {"metric": "synaptic events/sec", "value": res["value"], "unit": "events/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": getattr(wl, "scaling", "weak"), "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": res["roofline"], "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(wl.launches_per_timestep * T * K), "clocks": res["clocks"],
            "events_per_step": res["events_per_step"]}
    if others:
        line["other_workloads"] = others
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
