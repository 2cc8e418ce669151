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

    def __init__(self, n_neurons: int, group=None, device="cuda", world: Optional[int] = None,
                 rank: Optional[int] = None):
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
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
    (torch is used for the memory mapping only; the kernels are ours).  With world == 1 (or
    `local_only=True`) a plain device buffer stands in for the symmetric one, so the same kernels run on
    a single GPU."""

    def __init__(self, n_neurons: int, group=None, device="cuda", world: Optional[int] = None,
                 rank: Optional[int] = None, _emulated_bufs=None):
        super().__init__(n_neurons, group, device, world, rank)
        from . import _lib
        W, wpr = self.world, self.words_per_rank
        self.n_buf = 2 * W * wpr
        if _emulated_bufs is not None:          # emulate(): all ranks' buffers live on this one GPU
            self.symm = _emulated_bufs[self.rank]
            ptrs = [int(b.data_ptr()) for b in _emulated_bufs]
        elif W == 1:
            self.symm = torch.zeros(self.n_buf + 64, dtype=torch.int32, device=device)
            ptrs = [int(self.symm.data_ptr())]
        else:
            import torch.distributed._symmetric_memory as symm
            self.symm = symm.empty(self.n_buf + 64, dtype=torch.int32, device=device)
            self.symm.zero_()
            grp = group if group is not None else dist.group.WORLD
            self.handle = symm.rendezvous(self.symm, grp)
            ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.peer_ptrs = ptrs
        self.peer_bufs = torch.tensor(ptrs, dtype=torch.int64, device=device)
        self.peer_flags = torch.tensor([p + 4 * self.n_buf for p in ptrs], dtype=torch.int64, device=device)
        self.ctl = torch.zeros(max(4, _lib.lib().be_b200_exchange_ctl_bytes() // 4), dtype=torch.int32, device=device)
        self.local = self.bounds[self.rank]
        self.ids = None
        torch.cuda.synchronize(device)
        if W > 1 and _emulated_bufs is None:
            dist.barrier(group=group)        # every rank's buffer is zeroed before anyone publishes

    @classmethod
    def emulate(cls, n_neurons: int, world: int, device="cuda") -> List["P2PSpikeExchange"]:
        """`world` ranks of one exchange on ONE GPU (tests / profiling without a second device): the
        "peer" buffers are local allocations.  Nothing may wait on a flag that a later launch on the
        same GPU would publish, so drive the ranks in two rounds per timestep: every rank's publish,
        then every rank's consume (PostShardedProjection.publish / .consume)."""
        wpr = (post_block_bounds(n_neurons, world)[0][1] + 31) // 32
        bufs = [torch.zeros(2 * world * wpr + 64, dtype=torch.int32, device=device) for _ in range(world)]
        return [cls(n_neurons, device=device, world=world, rank=r, _emulated_bufs=bufs) for r in range(world)]

    def set_spin_limit_ms(self, ms: int) -> None:
        """Bound of one flag wait (default 20 s).  A wait that gives up is fatal: see timed_out()."""
        self.ctl[3] = int(ms)

    def _spikes_arg(self, local_spikes: torch.Tensor):
        from . import _lib
        s = local_spikes
        n_own = self.local[1] - self.local[0]
        assert s.numel() == n_own, (f"rank {self.rank} owns neurons [{self.local[0]}, {self.local[1]}) but got "
                                    f"{s.numel()} spikes (a short vector would shift every id behind it)")
        if s.dtype in (torch.bool, torch.uint8, torch.int8):
            kind = _lib.SPK_BOOL
        elif s.dtype == torch.float32:
            kind = _lib.SPK_F32
        else:
            s, kind = (s > 0), _lib.SPK_BOOL
        return s.contiguous(), kind

    def publish_only(self, local_spikes: torch.Tensor) -> None:
        """First half of exchange_spikes(): bit-pack + flag the peers.  With emulate(), publishing every
        emulated peer's block lets ONE rank's consumer kernel be profiled on a single GPU."""
        from . import _lib
        from ._runtime import ptr, stream_ptr
        s, kind = self._spikes_arg(local_spikes)
        _lib.check(_lib.lib().be_b200_pack_publish(ptr(s) if s.numel() else None, kind, s.numel(), ptr(self.peer_bufs),
                                                   ptr(self.peer_flags), self.world, self.rank, self.words_per_rank,
                                                   ptr(self.ctl), stream_ptr(self.symm.device)), "pack_publish")

    def exchange_spikes(self, local_spikes: torch.Tensor) -> torch.Tensor:
        """bool/uint8/float32 spikes of the neurons this rank owns -> int32[ceil(N/32)] global words."""
        import ctypes
        from . import _lib
        from ._runtime import ptr, stream_ptr
        s, kind = self._spikes_arg(local_spikes)
        L = _lib.lib()
        dev = self.symm.device
        _lib.check(L.be_b200_pack_publish(ptr(s) if s.numel() else None, kind, s.numel(), ptr(self.peer_bufs),
                                          ptr(self.peer_flags), self.world, self.rank, self.words_per_rank,
                                          ptr(self.ctl), stream_ptr(dev)), "pack_publish")
        _lib.check(L.be_b200_wait_exchange(ctypes.c_void_p(self.symm.data_ptr() + 4 * self.n_buf), ptr(self.peer_bufs),
                                           self.world, self.words_per_rank, ptr(self.gathered), ptr(self.ctl),
                                           stream_ptr(dev)), "wait_exchange")
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
        s, kind = self._spikes_arg(local_spikes)
        dev = self.symm.device
        if self.ids is None:
            self.ids = torch.zeros(self.n_neurons + 1, dtype=torch.int32, device=dev)
        L = _lib.lib()
        counter = ctypes.c_void_p(self.ids.data_ptr() + 4 * self.n_neurons)
        _lib.check(L.be_b200_pack_publish_reset(ptr(s) if s.numel() else None, kind, s.numel(), ptr(self.peer_bufs),
                                                ptr(self.peer_flags), self.world, self.rank, self.words_per_rank,
                                                ptr(self.ctl), counter, stream_ptr(dev)), "pack_publish_reset")
        _lib.check(L.be_b200_wait_exchange_compact(ctypes.c_void_p(self.symm.data_ptr() + 4 * self.n_buf),
                                                   ptr(self.peer_bufs), self.world, self.words_per_rank, self.n_neurons,
                                                   ptr(self.gathered), ptr(self.ids), ptr(out),
                                                   out.numel() * out.element_size(), ptr(self.ctl), stream_ptr(dev)),
                   "wait_exchange_compact")
        return self.gathered[: self.n_words], self.ids

    def timed_out(self) -> bool:
        """True when a flag wait of this rank ever gave up (sticky; reads the device, i.e. synchronises)."""
        return bool(self.ctl[2].item())

    def check(self, group=None) -> None:
        """Raise on every rank if any rank's exchange timed out (collective)."""
        bad = torch.tensor([int(self.timed_out())], device=self.ctl.device)
        if self.world > 1 and dist.is_initialized():
            dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group if group is not None else self.group)
        if int(bad.item()):
            raise ExchangeTimeout("peer-memory spike exchange: a flag wait timed out on some rank (a peer never "
                                  "published); every result since then is invalid")


class ExchangeTimeout(RuntimeError):
    """A bounded flag wait of the peer-memory exchange gave up; the projection must be rebuilt."""


class PostShardedProjection:
    """One rank's share of a projection pre[N] -> post[N], partitioned by post block.

    `kind`: 'fcn' (weights/indices [n_pre_total, n_conn_local], local post ids), 'csr'
    (data, indices, indptr with local post ids), 'dense' (W [n_pre_total, n_local]: this rank's post
    columns of the dense matrix) or 'jitc' (parameters only; the generated matrix is sliced by a
    column window for the scatter orientation or a row window for the gather orientation, with the
    GLOBAL shape keying the RNG walk).

    On a peer-memory exchange (P2PSpikeExchange) the step runs through the C object of
    csrc/sharded.cu (be_b200_sharded_*): for stored connectivity that is two kernels per timestep with the
    all-gather fused into the scatter.  On the NCCL exchange it is pack -> all_gather -> operator."""

    def __init__(self, kind: str, n_pre: int, n_post: int, exchange: SpikeExchange, **conn):
        assert kind in ("fcn", "csr", "dense", "jitc"), kind
        assert n_pre == exchange.n_neurons, (f"the exchange carries {exchange.n_neurons} neurons but the projection has "
                                             f"{n_pre} pre neurons (n_active would be read from the wrong slot)")
        self.kind, self.n_pre, self.n_post, self.x = kind, n_pre, n_post, exchange
        self.lo, self.hi = post_block_bounds(n_post, exchange.world)[exchange.rank]
        self.conn = conn
        self._h = None
        self._jit_dev = None
        if isinstance(exchange, P2PSpikeExchange):
            self._h = self._make_handle()

    # ---- C object ---------------------------------------------------------------------------------
    def _make_handle(self):
        import ctypes
        from . import _lib
        from ._runtime import ptr, wt_code
        x, c = self.x, self.conn
        L = _lib.lib()
        n_loc = self.hi - self.lo
        arr = (ctypes.c_void_p * x.world)(*x.peer_ptrs)
        with torch.cuda.device(x.symm.device):
            h = L.be_b200_sharded_create(x.world, x.rank, self.n_pre, n_loc, x.words_per_rank, arr, ptr(x.ctl))
            if not h:
                raise _lib.KernelExecutionError("sharded_create: " + L.be_b200_last_error().decode())
            h = ctypes.c_void_p(h)
            if self.kind == "fcn":
                w, idx = c["weights"], c["indices"]
                assert idx.dtype == torch.int32 and idx.is_contiguous() and w.is_contiguous()
                rc = L.be_b200_sharded_set_fcn(h, ptr(w), wt_code(w.dtype), int(w.numel() == 1), ptr(idx), idx.shape[0],
                                               idx.shape[1])
                self._out_dtype = w.dtype
            elif self.kind == "csr":
                w, idx, ip = c["data"], c["indices"], c["indptr"]
                assert idx.dtype == torch.int32 and ip.dtype in (torch.int32, torch.int64)
                assert ip.numel() == self.n_pre + 1
                rc = L.be_b200_sharded_set_csr(h, ptr(w), wt_code(w.dtype), int(w.numel() == 1), ptr(idx), ptr(ip),
                                               int(ip.dtype == torch.int64), self.n_pre, idx.numel())
                self._out_dtype = w.dtype
            elif self.kind == "dense":
                W = c["W"]
                assert W.shape == (self.n_pre, n_loc) and W.is_contiguous()
                rc = L.be_b200_sharded_set_dense(h, ptr(W), wt_code(W.dtype))
                self._out_dtype = W.dtype
            else:
                from . import _jitc
                params, clen, seed = self._jit_params(x.symm.device)
                fam = {"scalar": _lib.JIT_SCALAR, "normal": _lib.JIT_NORMAL, "uniform": _lib.JIT_UNIFORM}[c["family"]]
                # the operator is called with transpose=True on the logical (n_pre, n_post) matrix, so chunk_size
                # derives from the GLOBAL n_post exactly as in _jitc._jit_kernel
                chunk = _jitc._normalize_chunk_size(self.n_post)
                rc = L.be_b200_sharded_set_jitc(h, fam, ptr(params[0]), ptr(params[1]) if len(params) > 1 else None,
                                                wt_code(params[0].dtype), ptr(clen), ptr(seed), self.n_post, chunk,
                                                int(bool(c["corder"])))
                self._out_dtype = params[0].dtype
            _lib.check(rc, "sharded_set_" + self.kind)
        return h

    def set_option(self, name: str, value: int) -> None:
        from . import _lib
        assert self._h is not None, "options belong to the peer-memory (C object) path"
        _lib.check(_lib.lib().be_b200_sharded_set_option(self._h, name.encode(), int(value)), "sharded_set_option")

    def _jit_params(self, dev):
        from . import _jitc
        if self._jit_dev is None:
            # parameters, clen and seed live on the device once: per-step calls then issue no host
            # copy and the whole step can be captured in a CUDA graph
            c = self.conn
            self._jit_dev = ([torch.as_tensor([float(p)], dtype=torch.float32, device=dev) for p in c["params"]],
                             _jitc._initialize_conn_length(c["prob"], dev), _jitc._initialize_seed(c["seed"], dev))
        return self._jit_dev

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None:
            try:
                from . import _lib
                _lib.lib().be_b200_sharded_destroy(h)
            except Exception:
                pass

    # ---- NCCL / generic path ------------------------------------------------------------------------
    def pack_local(self, local_spikes: torch.Tensor) -> torch.Tensor:
        """Bit-pack the spikes of the neurons this rank owns into its all-gather slot."""
        from ._event import bitpack
        n_own = self.x.local[1] - self.x.local[0]
        assert local_spikes.numel() == n_own, f"rank owns {n_own} neurons, got {local_spikes.numel()} spikes"
        words = bitpack(local_spikes, axis=0)
        if words.numel() < self.x.words_per_rank:
            words = torch.nn.functional.pad(words, (0, self.x.words_per_rank - words.numel()))
        return words

    def propagate(self, packed_all: torch.Tensor) -> torch.Tensor:
        """packed global spikes -> this rank's block of the output."""
        from ._csr import binary_csrmv
        from ._dense import binary_densemv
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
        if self.kind == "dense":
            return binary_densemv(c["W"], ev, transpose=True)
        if self.kind == "jitc":
            prim = {"scalar": _jitc.binary_jitsmv_p, "normal": _jitc.binary_jitnmv_p,
                    "uniform": _jitc.binary_jitumv_p}[c["family"]]
            corder = c["corder"]  # passed corder: True = gather (rows = post), False = scatter (cols = post)
            extra = dict(row_window=(self.lo, n_loc)) if corder else dict(col_window=(self.lo, self.hi))
            params, clen, seed = self._jit_params(packed_all.device)
            return _jitc._jit_p_call(prim, params, clen, ev, seed, (self.n_pre, self.n_post), True, corder, None, False,
                                     **extra)[0]
        raise ValueError(self.kind)

    def publish(self, local_spikes: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """First phase of step() on the peer-memory path: bit-pack + flag the peers (+ clear `out`)."""
        from . import _lib
        from ._runtime import ptr, stream_ptr
        assert self._h is not None, "publish/consume belong to the peer-memory (C object) path"
        s, kind = self.x._spikes_arg(local_spikes)
        dev = self.x.symm.device
        if out is None:
            out = torch.empty(self.hi - self.lo, dtype=self._out_dtype, device=dev)
        _lib.check(_lib.lib().be_b200_sharded_publish(self._h, ptr(s) if s.numel() else None, kind, ptr(out),
                                                      stream_ptr(dev)), "sharded_publish")
        return out

    def consume(self, out: torch.Tensor) -> torch.Tensor:
        """Second phase: pull the peers' words and propagate them into `out` (from publish())."""
        from . import _lib
        from ._runtime import ptr, stream_ptr
        _lib.check(_lib.lib().be_b200_sharded_consume(self._h, ptr(out), stream_ptr(self.x.symm.device)),
                   "sharded_consume")
        return out

    def step(self, local_spikes: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One timestep: this rank's spike block in, this rank's output block out (collective)."""
        if self._h is not None:
            from . import _lib
            from ._runtime import ptr, stream_ptr
            s, kind = self.x._spikes_arg(local_spikes)
            dev = self.x.symm.device
            if out is None:
                out = torch.empty(self.hi - self.lo, dtype=self._out_dtype, device=dev)
            _lib.check(_lib.lib().be_b200_sharded_step(self._h, ptr(s) if s.numel() else None, kind, ptr(out),
                                                       stream_ptr(dev)), "sharded_step")
            return out
        return self.propagate(self.x.exchange(self.pack_local(local_spikes)))


class HostStepPipeline:
    """Host-buffer entry of the post-sharded step: per timestep this rank's spike block comes from
    pinned host memory and its output block goes back to pinned host memory (the multi-GPU twin of
    be_b200_projection_submit_host / _wait).  `slots` timesteps are in flight: copy-in, the device
    step (one CUDA graph per slot) and copy-out run on three streams and overlap across neighbouring
    timesteps.  Every rank must call submit() the same number of times (the step has a collective).

    On the peer-memory exchange the whole pipeline lives in C (be_b200_sharded_submit_host / _wait /
    _run_host, csrc/sharded.cu): one call per timestep, or one call for a whole block of timesteps with
    run().  On the NCCL exchange it is driven from Python (torch streams and graphs)."""

    def __init__(self, proj: "PostShardedProjection", n_in: int, slots: int = 4, use_graph: bool = True):
        self.proj, self.slots = proj, slots
        dev = proj.x.gathered.device
        self.dev = dev
        self.native = proj._h is not None
        if self.native:
            from . import _lib
            assert n_in == proj.x.local[1] - proj.x.local[0]
            with torch.cuda.device(dev):
                if not (proj.kind in ("fcn", "csr")):
                    # unfused consumers allocate their gathered / id buffers on the first step: run it (collectively)
                    # before the per-slot graphs are captured
                    proj.step(torch.zeros(n_in, dtype=torch.uint8, device=dev))
                    torch.cuda.synchronize(dev)
                _lib.check(_lib.lib().be_b200_sharded_host_init(proj._h, slots, int(use_graph)), "sharded_host_init")
            return
        self.s_in, self.s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self.main = torch.cuda.Stream(device=dev)
        self.d_in = [torch.zeros(n_in, dtype=torch.uint8, device=dev) for _ in range(slots)]
        mk = lambda: [torch.cuda.Event() for _ in range(slots)]
        self.ev_in, self.ev_done = mk(), mk()
        # the host waits on ev_out when it reuses a slot: a blocking event lets that thread sleep instead
        # of spinning on a core, which matters when 8 ranks (plus their driver threads) share the host
        blocking = os.environ.get("BE_B200_BLOCKING_EVENTS", "1") != "0"
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
        if self.native:
            import ctypes
            from . import _lib
            rc = _lib.lib().be_b200_sharded_submit_host(self.proj._h, ctypes.c_void_p(spikes_host.data_ptr()),
                                                        ctypes.c_void_p(out_host.data_ptr()))
            _lib.check(rc, "sharded_submit_host")
            return
        i = self.i
        if self.inflight[i]:
            self.ev_out[i].synchronize()
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

    def run(self, n_steps: int, spikes_host: List[torch.Tensor], out_host: List[torch.Tensor]) -> None:
        """n_steps timesteps back to back: timestep t reads spikes_host[t % len] and writes
        out_host[t % len]; returns after the last result has landed.  Native path: ONE C call, no
        interpreter work between timesteps."""
        if self.native:
            import ctypes
            from . import _lib
            ins = (ctypes.c_void_p * len(spikes_host))(*[t.data_ptr() for t in spikes_host])
            outs = (ctypes.c_void_p * len(out_host))(*[t.data_ptr() for t in out_host])
            with torch.cuda.device(self.dev):
                rc = _lib.lib().be_b200_sharded_run_host(self.proj._h, int(n_steps), ins, len(spikes_host), outs,
                                                         len(out_host))
            _lib.check(rc, "sharded_run_host")
            return
        for t in range(n_steps):
            self.submit(spikes_host[t % len(spikes_host)], out_host[t % len(out_host)])
        self.wait()

    def wait(self) -> None:
        if self.native:
            from . import _lib
            _lib.check(_lib.lib().be_b200_sharded_wait(self.proj._h), "sharded_wait")
            return
        for k in range(self.slots):
            if self.inflight[k]:
                self.ev_out[k].synchronize()
                self.inflight[k] = False

    def close(self) -> None:
        """Drain, agree across ranks that no exchange wait timed out (raises ExchangeTimeout otherwise) and
        drop the captured graphs.  Call before dist.destroy_process_group(): tearing the NCCL
        communicator down while graphs that captured its collectives are alive can block."""
        try:
            self.wait()
        finally:
            if not self.native:
                self.main.synchronize()
                self.graphs = [None] * self.slots
                self.d_out = [None] * self.slots
            torch.cuda.synchronize(self.dev)
        if hasattr(self.proj.x, "check"):
            self.proj.x.check()
