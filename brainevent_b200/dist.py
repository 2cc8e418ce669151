"""Post-synaptic sharding across the GPUs of one box (net-new: the reference has no distributed
code).

Layout: rank g owns the contiguous block of post-synaptic neurons [lo_g, hi_g) and stores only the
synapses that END in its block (pre-major, local post ids), so outputs are disjoint and need no
reduction.  Per timestep every rank bit-packs the spikes of the neurons it owns and ONE all-gather
of uint32 words delivers the global spike vector (N/32 words in total); propagation then runs on
the packed words directly (event kind PACKED of the C ABI).  Block sizes are multiples of 32 so
the concatenation of per-rank words IS the packing of the concatenated spike vector.
"""
import os
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

ALIGN = 32


def post_block_bounds(n_post: int, world: int, align: int = ALIGN) -> List[Tuple[int, int]]:
    """Contiguous blocks, every boundary but the last a multiple of `align`."""
    per = -(-n_post // world)
    per = -(-per // align) * align
    out = []
    for g in range(world):
        lo = min(g * per, n_post)
        hi = min(lo + per, n_post)
        out.append((lo, hi))
    return out


def partition_csr_by_post(indptr, indices, data, n_post: int, world: int):
    """Split a pre-major CSR (rows = pre, cols = post) into per-rank CSRs holding the synapses that
    end in each rank's post block, with LOCAL column ids.  Works on numpy arrays (setup time).
    Returns [(indptr_g, indices_g, data_g, (lo, hi)), ...].  Order within a row is preserved, so
    each shard sums its synapses in the same order as the unsharded matrix."""
    indptr = np.asarray(indptr)
    indices = np.asarray(indices)
    data = np.asarray(data)
    homo = data.size == 1
    m = indptr.shape[0] - 1
    row_of = np.repeat(np.arange(m, dtype=np.int64), np.diff(indptr))
    shards = []
    for lo, hi in post_block_bounds(n_post, world):
        sel = (indices >= lo) & (indices < hi)
        counts = np.bincount(row_of[sel], minlength=m)
        ip = np.zeros(m + 1, indptr.dtype)
        np.cumsum(counts, out=ip[1:])
        shards.append((ip, (indices[sel] - lo).astype(np.int32), data if homo else data[sel], (lo, hi)))
    return shards


def partition_fcn_by_post(indices, weights, n_post: int, world: int):
    """FixedNumPerPre [n_pre, n_conn] -> per-rank ragged CSR shards (the per-rank fan-out of a pre
    neuron is binomial, so shards are CSR, not ELL)."""
    indices = np.asarray(indices)
    n_pre, n_conn = indices.shape
    indptr = (np.arange(n_pre + 1, dtype=np.int64) * n_conn)
    w = np.asarray(weights)
    return partition_csr_by_post(indptr if indptr[-1] >= 2**31 else indptr.astype(np.int32),
                                 indices.reshape(-1), w if w.size == 1 else w.reshape(-1), n_post, world)


def bind_to_gpu_numa(device_index: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to the GPU, so that pinned host
    buffers allocated afterwards (first touch) and the copy-issuing thread sit on the GPU's NUMA
    node.  With 8 ranks streaming host<->device every timestep this keeps the traffic off the
    inter-socket link.  Returns False (and changes nothing) when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


class SpikeExchange:
    """All-gather of bit-packed spike words over the process group (NCCL on GPUs, gloo in tests)."""

    def __init__(self, n_neurons: int, group=None, device="cuda"):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.bounds = post_block_bounds(n_neurons, self.world)
        self.n_neurons = n_neurons
        self.words_per_rank = (self.bounds[0][1] - self.bounds[0][0] + 31) // 32
        self.local = self.bounds[self.rank]
        self.gathered = torch.zeros(self.words_per_rank * self.world, dtype=torch.int32, device=device)
        self.n_words = (n_neurons + 31) // 32

    def exchange(self, packed_local: torch.Tensor) -> torch.Tensor:
        """packed_local: int32[words_per_rank] (zero-padded) -> int32[ceil(N/32)] global words."""
        assert packed_local.numel() == self.words_per_rank and packed_local.dtype == torch.int32
        if self.world == 1:
            return packed_local
        dist.all_gather_into_tensor(self.gathered, packed_local, group=self.group)
        return self.gathered[: self.n_words]


class P2PSpikeExchange(SpikeExchange):
    """Fused bit-pack + all-gather over NVLink peer memory (csrc/exchange.cu): every rank packs
    into its own symmetric buffer, flags completion on every peer, and pulls the peers' words
    over NVLink; no NCCL call on the per-timestep path.  Peer pointers come from a torch symmetric-memory rendezvous
    (torch is used for the memory mapping only; the kernels are ours)."""

    def __init__(self, n_neurons: int, group=None, device="cuda"):
        super().__init__(n_neurons, group, device)
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        W, wpr = self.world, self.words_per_rank
        self.n_buf = 2 * W * wpr
        self.symm = symm.empty(self.n_buf + 64, dtype=torch.int32, device=device)
        self.symm.zero_()
        grp = group if group is not None else dist.group.WORLD
        self.handle = symm.rendezvous(self.symm, grp)
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.peer_bufs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.peer_flags = torch.tensor([p + 4 * self.n_buf for p in ptrs], dtype=torch.int64, device=device)
        self.ctl = torch.zeros(max(4, _lib.lib().be_b200_exchange_ctl_bytes() // 4), dtype=torch.int32, device=device)
        self.local = self.bounds[self.rank]
        torch.cuda.synchronize(device)
        dist.barrier(group=group)        # every rank's buffer is zeroed before anyone publishes

    def exchange_spikes(self, local_spikes: torch.Tensor) -> torch.Tensor:
        """bool/uint8/float32 spikes of the neurons this rank owns -> int32[ceil(N/32)] global words."""
        from . import _lib
        from ._runtime import ptr, stream_ptr
        s = local_spikes
        if s.dtype in (torch.bool, torch.uint8, torch.int8):
            kind = _lib.SPK_BOOL
        elif s.dtype == torch.float32:
            kind = _lib.SPK_F32
        else:
            s, kind = (s > 0), _lib.SPK_BOOL
        s = s.contiguous()
        L = _lib.lib()
        dev = s.device
        me = self.symm
        _lib.check(L.be_b200_pack_publish(ptr(s), kind, s.numel(), ptr(self.peer_bufs), ptr(self.peer_flags), self.world,
                                          self.rank, self.words_per_rank, ptr(self.ctl), stream_ptr(dev)), "pack_publish")
        import ctypes
        _lib.check(L.be_b200_wait_exchange(ctypes.c_void_p(me.data_ptr() + 4 * self.n_buf), ptr(self.peer_bufs), self.world,
                                           self.words_per_rank, ptr(self.gathered), ptr(self.ctl), stream_ptr(dev)),
                   "wait_exchange")
        return self.gathered[: self.n_words]

    def exchange_compact(self, local_spikes: torch.Tensor, out: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Exchange for a SCATTER consumer: like exchange_spikes, but the wait kernel also turns the
        gathered words into the active-id list and clears `out` (the consumer's output vector) while it
        waits for the peers, so no separate compaction kernel or memset runs on the per-timestep chain.
        Returns (words int32[ceil(N/32)], ids_and_count int32[N + 1]); consume with event kind
        SPK_COMPACT_CLEARED (spikes = words, aux = ids_and_count)."""
        import ctypes
        from . import _lib
        from ._runtime import ptr, stream_ptr
        s = local_spikes
        if s.dtype in (torch.bool, torch.uint8, torch.int8):
            kind = _lib.SPK_BOOL
        elif s.dtype == torch.float32:
            kind = _lib.SPK_F32
        else:
            s, kind = (s > 0), _lib.SPK_BOOL
        s = s.contiguous()
        if getattr(self, "ids", None) is None:
            self.ids = torch.zeros(self.n_neurons + 1, dtype=torch.int32, device=s.device)
        L = _lib.lib()
        dev = s.device
        counter = ctypes.c_void_p(self.ids.data_ptr() + 4 * self.n_neurons)
        _lib.check(L.be_b200_pack_publish_reset(ptr(s), kind, s.numel(), ptr(self.peer_bufs), ptr(self.peer_flags),
                                                self.world, self.rank, self.words_per_rank, ptr(self.ctl), counter,
                                                stream_ptr(dev)), "pack_publish_reset")
        _lib.check(L.be_b200_wait_exchange_compact(ctypes.c_void_p(self.symm.data_ptr() + 4 * self.n_buf),
                                                   ptr(self.peer_bufs), self.world, self.words_per_rank, self.n_neurons,
                                                   ptr(self.gathered), ptr(self.ids), ptr(out),
                                                   out.numel() * out.element_size(), ptr(self.ctl), stream_ptr(dev)),
                   "wait_exchange_compact")
        return self.gathered[: self.n_words], self.ids

    def timed_out(self) -> bool:
        return bool(self.ctl[2].item())


class PostShardedProjection:
    """One rank's share of a projection pre[N] -> post[N], partitioned by post block.

    `kind`: 'fcn' (weights/indices [n_pre_total, n_conn_local], local post ids), 'csr'
    (data, indices, indptr with local post ids) or 'jitc' (parameters only; the generated matrix is
    sliced by a column window for the scatter orientation or a row window for the gather
    orientation, with the GLOBAL shape keying the RNG walk)."""

    def __init__(self, kind: str, n_pre: int, n_post: int, exchange: SpikeExchange, **conn):
        self.kind, self.n_pre, self.n_post, self.x = kind, n_pre, n_post, exchange
        self.lo, self.hi = post_block_bounds(n_post, exchange.world)[exchange.rank]
        self.conn = conn

    def pack_local(self, local_spikes: torch.Tensor) -> torch.Tensor:
        """Bit-pack the spikes of the neurons this rank owns into its all-gather slot."""
        from ._event import bitpack
        words = bitpack(local_spikes, axis=0)
        if words.numel() < self.x.words_per_rank:
            words = torch.nn.functional.pad(words, (0, self.x.words_per_rank - words.numel()))
        return words

    def propagate(self, packed_all: torch.Tensor) -> torch.Tensor:
        """packed global spikes -> this rank's block of the output."""
        from ._csr import binary_csrmv
        from ._event import BitPackedBinary
        from ._fcn import binary_fcnmv
        from . import _jitc
        ev = BitPackedBinary(packed=packed_all, shape=(self.n_pre,))
        n_loc = self.hi - self.lo
        c = self.conn
        if self.kind == "fcn":
            return binary_fcnmv(c["weights"], c["indices"], ev, shape=(self.n_pre, n_loc), transpose=True)
        if self.kind == "csr":
            return binary_csrmv(c["data"], c["indices"], c["indptr"], ev, shape=(self.n_pre, n_loc), transpose=True)
        if self.kind == "jitc":
            prim = {"scalar": _jitc.binary_jitsmv_p, "normal": _jitc.binary_jitnmv_p,
                    "uniform": _jitc.binary_jitumv_p}[c["family"]]
            corder = c["corder"]  # passed corder: True = gather (rows = post), False = scatter (cols = post)
            extra = dict(row_window=(self.lo, n_loc)) if corder else dict(col_window=(self.lo, self.hi))
            if getattr(self, "_jit_dev", None) is None:
                # parameters, clen and seed live on the device once: per-step calls then issue no host
                # copy and the whole step can be captured in a CUDA graph
                dev = packed_all.device
                self._jit_dev = ([torch.as_tensor([float(p)], dtype=torch.float32, device=dev) for p in c["params"]],
                                 _jitc._initialize_conn_length(c["prob"], dev), _jitc._initialize_seed(c["seed"], dev))
            params, clen, seed = self._jit_dev
            return _jitc._jit_p_call(prim, params, clen, ev, seed, (self.n_pre, self.n_post), True, corder, None, False,
                                     **extra)[0]
        raise ValueError(self.kind)

    def _step_fcn_fused(self, local_spikes: torch.Tensor) -> torch.Tensor:
        """FixedNumPerPre scatter behind the fused exchange: three kernels per timestep (pack + publish,
        wait + pull + compaction + output clear, scatter) instead of five launches."""
        from . import _lib
        from ._runtime import ptr, stream_ptr, workspace, wt_code
        w, idx = self.conn["weights"], self.conn["indices"]
        n_loc = self.hi - self.lo
        out = torch.empty(n_loc, dtype=w.dtype, device=w.device)
        words, ids = self.x.exchange_compact(local_spikes, out)
        ws = workspace(w.device, self.n_pre, 1)
        rc = _lib.lib().be_b200_binary_fcnmv(ptr(w), wt_code(w.dtype), int(w.numel() == 1), ptr(idx), idx.shape[0],
                                            idx.shape[1], ptr(words), ptr(ids), _lib.SPK_COMPACT_CLEARED, self.n_pre,
                                            n_loc, 1, ptr(out), ptr(ws), ws.numel(), stream_ptr(w.device))
        _lib.check(rc, "binary_fcnmv")
        return out

    def step(self, local_spikes: torch.Tensor) -> torch.Tensor:
        if (self.kind == "fcn" and hasattr(self.x, "exchange_compact")
                and os.environ.get("BE_B200_FUSED_COMPACT", "1") != "0"):
            return self._step_fcn_fused(local_spikes)
        if hasattr(self.x, "exchange_spikes"):           # fused pack + peer-memory all-gather
            return self.propagate(self.x.exchange_spikes(local_spikes))
        return self.propagate(self.x.exchange(self.pack_local(local_spikes)))


class HostStepPipeline:
    """Host-buffer entry of the post-sharded step: per timestep this rank's spike block comes from
    pinned host memory and its output block goes back to pinned host memory (the multi-GPU twin of
    be_b200_projection_submit_host / _wait).  `slots` timesteps are in flight: copy-in, the device
    step (bit-pack, all-gather, compaction, scatter -- one CUDA graph per slot, so the host issues
    three calls per timestep) and copy-out run on three streams and overlap across neighbouring
    timesteps.  Every rank must call submit() the same number of times (the step has a collective).
    """

    def __init__(self, proj: "PostShardedProjection", n_in: int, slots: int = 4, use_graph: bool = True):
        self.proj, self.slots = proj, slots
        dev = proj.x.gathered.device
        self.dev = dev
        self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.main = torch.cuda.Stream(device=dev)
        self.d_in = [torch.zeros(n_in, dtype=torch.uint8, device=dev) for _ in range(slots)]
        mk = lambda: [torch.cuda.Event() for _ in range(slots)]
        self.ev_in, self.ev_done = mk(), mk()
        # the host waits on ev_out when it reuses a slot: a blocking event lets that thread sleep instead
        # of spinning on a core, which matters when 8 ranks (plus their driver threads) share the host
        blocking = os.environ.get("BE_B200_BLOCKING_EVENTS", "0") != "0"
        self.ev_out = [torch.cuda.Event(blocking=blocking) for _ in range(slots)]
        self.inflight = [False] * slots
        self.i = 0
        self.graphs, self.d_out = [None] * slots, [None] * slots
        with torch.cuda.stream(self.main):
            for k in range(slots):                       # warm-up: allocations, workspaces, NCCL channels
                self.d_out[k] = proj.step(self.d_in[k])
        self.main.synchronize()
        if use_graph:
            for k in range(slots):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.main):
                    self.d_out[k] = proj.step(self.d_in[k])
                self.graphs[k] = g
        self.main.synchronize()

    def submit(self, spikes_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """spikes_host: pinned uint8/bool [n_in]; out_host: pinned [n_out] in the weight dtype.  Both must
        stay untouched until wait() (or the `slots`-th next submit) returns."""
        i = self.i
        if self.inflight[i]:
            self.ev_out[i].synchronize()
        # plain set_stream calls instead of three `with torch.cuda.stream(...)` blocks: the host side of a
        # timestep is ~70 us of Python (tools/e2e_pipeline_probe.py), close to what the PCIe copies take
        prev = torch.cuda.current_stream(self.dev)
        try:
            torch.cuda.set_stream(self.s_in)
            self.d_in[i].copy_(spikes_host.view(torch.uint8), non_blocking=True)
            self.ev_in[i].record(self.s_in)
            self.main.wait_event(self.ev_in[i])
            torch.cuda.set_stream(self.main)
            if self.graphs[i] is not None:
                self.graphs[i].replay()
            else:
                self.d_out[i] = self.proj.step(self.d_in[i])
            self.ev_done[i].record(self.main)
            self.s_out.wait_event(self.ev_done[i])
            torch.cuda.set_stream(self.s_out)
            out_host.copy_(self.d_out[i], non_blocking=True)
            self.ev_out[i].record(self.s_out)
        finally:
            torch.cuda.set_stream(prev)
        self.inflight[i] = True
        self.i = (i + 1) % self.slots

    def wait(self) -> None:
        for k in range(self.slots):
            if self.inflight[k]:
                self.ev_out[k].synchronize()
                self.inflight[k] = False

    def close(self) -> None:
        """Drain and drop the captured graphs.  Call before dist.destroy_process_group(): tearing the
        NCCL communicator down while graphs that captured its collectives are alive can block."""
        self.wait()
        self.main.synchronize()
        self.graphs = [None] * self.slots
        self.d_out = [None] * self.slots
        torch.cuda.synchronize(self.dev)
