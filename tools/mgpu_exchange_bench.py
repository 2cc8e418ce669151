"""Per-timestep cost of the spike exchange alone (run under torchrun): bit-pack of this rank's
1,250,016 neurons + all-gather of the packed words, captured in a CUDA graph of 100 exchanges,
NCCL vs the peer-memory kernels.  Prints max-over-ranks microseconds per exchange."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import brainevent_b200 as be
from brainevent_b200 import dist as bd

n_local = 1_250_016
N = n_local * world
T = 100
tag = os.environ.get("TAG", "default")
spikes = [(torch.rand(n_local, device=dev) < 0.01) for _ in range(4)]


def time_graph(fn):
    for s in spikes:
        fn(s)
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            for t in range(T):
                fn(spikes[t % 4])
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record(); torch.cuda.synchronize()
    dt = torch.tensor([a.elapsed_time(b) / 5 / T * 1e3], device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt.item())


xn = bd.SpikeExchange(N, device=dev)
pn = bd.PostShardedProjection("fcn", N, N, xn)
t_pack = time_graph(lambda s: pn.pack_local(s))
t_nccl = time_graph(lambda s: xn.exchange(pn.pack_local(s)))
msg = f"[{tag}] world={world} pack {t_pack:.1f} us, pack+NCCL all-gather {t_nccl:.1f} us"
if os.environ.get("NO_P2P") != "1":
    xp = bd.P2PSpikeExchange(N, device=dev)
    t_p2p = time_graph(lambda s: xp.exchange_spikes(s))
    msg += f", fused pack+P2P exchange {t_p2p:.1f} us (timed_out={xp.timed_out()})"
    st = xp.ctl.view(torch.int64)[2:8].tolist()     # %globaltimer stamps of the last exchange (ns)
    print(f"[{tag}] rank {rank} last exchange: pack {st[1]-st[0]} ns, fence+flags {st[2]-st[1]} ns, "
          f"gap to wait kernel {st[3]-st[2]} ns, poll {st[4]-st[3]} ns, copy {st[5]-st[4]} ns", flush=True)
if rank == 0:
    print(msg, flush=True)
dist.barrier()
dist.destroy_process_group()
