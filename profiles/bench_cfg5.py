"""BASELINE.json configs[4] per-GPU shard: 100k utterances / 8 GPUs = 12 500 utterances x 300 frames x 10
IF units, 64-mix GMMs; uniform segmentation -> per-state k-means (K = 64) -> 5 Baum-Welch iterations.
    python profiles/bench_cfg5.py [n_utt]            (one GPU; torchrun for N GPUs: accumulators allreduced)
Prints one JSON line (rank 0): k-means init time, per-iteration milliseconds, EM frames/s."""
import json, os, random, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model, kmeans_run, kmeans_seed_points

N_UNITS, N_INITIALS, MIX, T, L = 57, 22, 64, 300, 10
n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
group = None
if world > 1:
    import torch.distributed as dist
    os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    group = dist.group.WORLD
eng = Engine(local)
dev = eng.device
truth, _, labels, x = synth.torch_corpus(n_utt, T, L, N_UNITS, MIX, 5, dev, N_INITIALS, data_seed=5 + rank)
# ---- uniform segmentation (AcousticModel.py:605-626): frame t of an utterance -> (position, state)
chunk = T // L
pos = torch.arange(T, device=dev).div(chunk, rounding_mode="floor").clamp_(max=L - 1)
within = torch.arange(T, device=dev) - pos * chunk
c3 = chunk // 3
st = within.div(c3, rounding_mode="floor").clamp_(max=2)
valid = torch.arange(T, device=dev) < chunk * L
lab_t = torch.as_tensor(labels.astype(np.int64), device=dev)
gstate = (lab_t[:, pos] * 3 + st[None, :])[:, valid].reshape(-1)           # global state of every used frame
frames = x.view(n_utt, T, -1)[:, valid].reshape(-1, x.shape[1]).double()
order = torch.argsort(gstate, stable=True)
counts = torch.bincount(gstate, minlength=N_UNITS * 3).cpu().numpy()
# k-means on at most 4096 points per state (the reference's O(N^2) greedy passes; the device kernel is
# bit-exact with it, its cost grows with N^2 / K as well)
cap = 4096
off, sel = [0], []
start = np.concatenate([[0], np.cumsum(counts)])
for s in range(N_UNITS * 3):
    n_s = min(int(counts[s]), cap)
    sel.append(order[start[s]:start[s] + n_s])
    off.append(off[-1] + n_s)
pts = frames[torch.cat(sel)].contiguous()
x0 = pts[:, 0].cpu().numpy()
seeds = [kmeans_seed_points(np.ascontiguousarray(x0[off[s]:off[s + 1]]), MIX, random.Random(s)) for s in range(N_UNITS * 3)]
torch.cuda.synchronize(); t0 = time.perf_counter()
out = kmeans_run(eng, pts, np.array(off), MIX, np.array(seeds, dtype=np.int32))
torch.cuda.synchronize(); t_km = time.perf_counter() - t0
mean = out["mean"].cpu().numpy().reshape(N_UNITS, 3, MIX, -1)
var = np.maximum(out["var"].cpu().numpy().reshape(N_UNITS, 3, MIX, -1), 1e-2)
alpha = out["alpha"].cpu().numpy().reshape(N_UNITS, 3, MIX); alpha = alpha / alpha.sum(-1, keepdims=True)
# ---- EM
corpus = Corpus(eng, labels, np.full(n_utt, T, dtype=np.int32), N_UNITS)
model = Model(eng, mean, var, alpha, synth.default_transmat(N_UNITS))
es = EStep(eng, corpus, model)
es.load_frames(x, group=group)
ms, ll = [], []
for it in range(5):
    torch.cuda.synchronize()
    if group is not None: dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); es.em_iteration(c_covariance=1e-3, group=group); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1)); ll.append(float(es.utt_logp.sum()))
# stage split of one more iteration (events on the launch stream)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
ev[0].record(); es.score(); ev[1].record(); es.forward_backward(); ev[2].record(); es.accumulate(); ev[3].record()
es.reduce_statistics(group); es.mstep(c_covariance=1e-3); ev[4].record(); torch.cuda.synchronize()
stage_ms = {"K1_score": ev[0].elapsed_time(ev[1]), "K2_forward_backward": ev[1].elapsed_time(ev[2]),
            "K3_accumulate": ev[2].elapsed_time(ev[3]), "reduce_mstep": ev[3].elapsed_time(ev[4])}
if rank == 0:
    from poccala_b200 import _native as nat
    print(json.dumps({"workload": "cfg5 per-GPU shard: %d utt x %d frames x %d units, %d-mix, k-means init (<= %d points per state) + 5 EM iterations" % (n_utt, T, L, MIX, cap),
                      "n_gpus": world, "kmeans_s": t_km, "kmeans_points": int(off[-1]), "em_ms": ms,
                      "em_frames_per_s": world * n_utt * T / (min(ms[1:]) * 1e-3), "stage_ms": stage_ms, "sum_logp": ll,
                      "k3_active_pair_frac": nat.lib().pc_corpus_active_tiles(corpus.c) / nat.lib().pc_corpus_total_tiles(corpus.c)}))
if group is not None: dist.destroy_process_group()
