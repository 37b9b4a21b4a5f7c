"""Debug aid: the host entry point with the reduce hook installed, against the same call without it.
   python profiles/exp_hook.py            (world 1, NCCL group of one rank)
   torchrun --nproc-per-node 2 ... profiles/exp_hook.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poccala_b200 import synth  # noqa: E402
from poccala_b200.engine import Corpus, Engine, HostReduceHook, em_iteration_host, frame_moments_host  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
os.environ["NCCL_DEBUG"] = "WARN"
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
group = dist.group.WORLD
eng = Engine(local)
for opt in ('k1_kernel', 'k2_kernel', 'k3_kernel', 'host_chunks', 'tensor_core'):
    if os.environ.get(opt.upper()):
        eng.set_option(opt, int(os.environ[opt.upper()]))
U, M, T, L, n_all = 40, 16, 300, 5, int(os.environ.get('N_ALL', '400'))
truth, init0, labels, x = synth.torch_corpus(n_all, T, L, U, M, 2, eng.device, 10)
tm0 = synth.default_transmat(U)
x = x.view(n_all, T, 39)


def run(sel, hook_group):
    xs = x[sel].reshape(-1, 39).contiguous().cpu().numpy()
    corpus = Corpus(eng, labels[sel], np.full(len(sel), T, dtype=np.int32), U)
    shift, isc = frame_moments_host(eng, xs, group=hook_group)
    hook = HostReduceHook(eng, U, U * 3 * M, hook_group) if hook_group is not None else None
    hp = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in init0] + [tm0.copy()]
    ll = em_iteration_host(eng, corpus, xs, *hp, c_covariance=1e-6, shift=shift, inv_scale=isc)
    if hook is not None:
        if hook.error is not None:
            raise hook.error
        hook.remove()
        # the reduced statistics against the device-resident path on the same shard
        from poccala_b200.engine import EStep, Model
        model = Model(eng, init0[0], init0[1], init0[2], tm0)
        es = EStep(eng, corpus, model)
        g2 = None if os.environ.get('PC_HOOK_MODE') == 'noop' else hook_group
        es.load_frames(torch.as_tensor(xs).to(eng.device), group=hook_group)
        es.estep(0, g2)
        torch.cuda.synchronize()
        a, b2 = hook.flat[:es.acc.numel()].view_as(es.acc), es.acc
        d = (a - b2).abs()
        print(rank, "acc diff", float(d.max()), float(b2.abs().max()), "rows off", int((d.view(-1, 80).max(dim=1).values > 1e-6 * b2.abs().max()).sum()),
              "of", d.numel() // 80, "tsum diff", float((hook.flat[es.acc.numel():] - es.tsum.view(-1)).abs().max()), flush=True)
    return hp, ll


ref, ll0 = run(np.arange(n_all), None)
if os.environ.get('OWN_STREAM'):
    with torch.cuda.stream(torch.cuda.Stream()):
        got, ll1 = run(np.arange(rank, n_all, world), group)
    torch.cuda.synchronize()
else:
    got, ll1 = run(np.arange(rank, n_all, world), group)
for name, a, b in zip(("mean", "var", "alpha", "transmat"), got, ref):
    print(rank, name, float(np.abs(a - b).max()), float(np.abs(b).max()), flush=True)
print(rank, "logp", ll0, ll1, flush=True)
if world == 1 and os.environ.get("HALF"):
    # the host entry against the device-resident path on a half corpus (the shard shape of a 2-rank run)
    from poccala_b200.engine import EStep, Model
    sel = np.arange(0, n_all, 2)
    got, ll = run(sel, None)
    corpus = Corpus(eng, labels[sel], np.full(len(sel), T, dtype=np.int32), U)
    model = Model(eng, init0[0], init0[1], init0[2], tm0)
    es = EStep(eng, corpus, model)
    es.load_frames(x[sel].reshape(-1, 39).contiguous())
    es.em_iteration(c_covariance=1e-6)
    torch.cuda.synchronize()
    for name, a, b in zip(("mean", "var", "alpha", "transmat"), got, (model.mean, model.var, model.alpha, model.transmat)):
        b = b.cpu().numpy()
        print("half", name, float(np.abs(a - b).max()), float(np.abs(b).max()), flush=True)
    print("half logp", ll, float(es.utt_logp.sum().item()))
dist.destroy_process_group()
