"""The peer-memory M-step alone, against all-reduce + M-step: torchrun --nproc-per-node N profiles/exp_peer.py"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from poccala_b200 import _native as nat
from poccala_b200.engine import Engine, _p, _stream
from poccala_b200.distributed import PeerExchange
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["NCCL_DEBUG"] = "WARN"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = Engine(local)
dev = eng.device
U, D = 57, 39
for mix in (16, 64):
    G = U * 3 * mix
    peer = PeerExchange(eng, U, G, dist.group.WORLD)
    g = torch.Generator(device=dev); g.manual_seed(1 + rank)
    for w in (0, 1):
        acc, tsum, tmax = peer.views(w)
        acc.copy_(torch.rand(acc.shape, generator=g, device=dev, dtype=torch.float64) + 0.5)
        acc[:, 40:79] += 2.0
        tsum.fill_(1.0); tmax.fill_(-5.0)
    mean = torch.zeros((G, D), dtype=torch.float64, device=dev); var = torch.ones_like(mean)
    alpha = torch.ones((G,), dtype=torch.float64, device=dev); tm = torch.zeros((U, 5, 5), dtype=torch.float64, device=dev)
    flat = torch.rand((G * 80 + U * 9,), dtype=torch.float64, device=dev)
    tmx = torch.zeros((U, 9), dtype=torch.float64, device=dev)

    def peer_step():
        nat.call("pc_update_params_peer", eng.h, mix, D, None, None, 1e-3, 0, _p(mean), _p(var), _p(alpha), _p(tm), _stream())

    def nccl_step():
        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        dist.all_reduce(flat)
        nat.call("pc_update_params", eng.h, U, mix, D, _p(flat), _p(tmx), _p(flat[G * 80:]), None, None, 1e-3, 0,
                 _p(mean), _p(var), _p(alpha), _p(tm), _stream())

    for name, fn in (("peer", peer_step), ("nccl", nccl_step)):
        for _ in range(4):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print("mix", mix, name, "us per call %.1f" % (e0.elapsed_time(e1) * 1e3 / 20), "bytes", G * 80 * 8, flush=True)
        # the same after 5 ms of local work that evicts L2 and leaves the links idle (what an EM iteration does)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for dirty in (False, True):
            ts = []
            for _ in range(8):
                flush.zero_()
                torch.cuda._sleep(8_000_000)
                if dirty and name == "peer":
                    a_cur = peer.current()[0]
                    a_cur.add_(1e-9)  # the statistics were just written (K3's atomics leave them dirty in L2)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record(); fn(); a1.record(); torch.cuda.synchronize()
                ts.append(a0.elapsed_time(a1) * 1e3)
            if rank == 0:
                print("mix", mix, name, "after idle + L2 flush%s: us" % (" + fresh statistics" if dirty else ""), ["%.0f" % t for t in ts], flush=True)
        del flush
    if rank == 0:
        print("   timeouts", peer.timeouts())
    del acc, tsum, tmax
    peer.close()
dist.destroy_process_group()
