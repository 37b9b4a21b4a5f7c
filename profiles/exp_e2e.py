"""Where the end-to-end step goes: python profiles/exp_e2e.py
Times pc_em_iteration_host (pinned frames + pinned model) for several chunk counts, next to the plain copy."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, em_iteration_host, frame_moments_host
n_utt, T, L, mix = 1000, 300, 10, 16
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(n_utt, T, L, 57, mix, 2, eng.device, 22)
corpus = Corpus(eng, labels, np.full(n_utt, T, dtype=np.int32), 57)
tm0 = synth.default_transmat(57)
host_x = torch.empty((n_utt * T, 39), dtype=torch.float32).pin_memory()
host_x.copy_(x.cpu())
hx = host_x.numpy()
hp_t = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).clone().pin_memory() for a in list(init0) + [tm0]]
hp = [t.numpy() for t in hp_t]
shift, isc = frame_moments_host(eng, hx)
dst = torch.empty_like(x)
for chunks in (0, 1, 2, 4, 8):
    eng.set_option("host_chunks", chunks)
    for _ in range(3):
        em_iteration_host(eng, corpus, hx, *hp, c_covariance=1e-6, shift=shift, inv_scale=isc)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        em_iteration_host(eng, corpus, hx, *hp, c_covariance=1e-6, shift=shift, inv_scale=isc)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("host_chunks", chunks, "ms/step %.3f" % (dt * 1e3), "M frames/s %.1f" % (n_utt * T / dt / 1e6), flush=True)
for _ in range(3):
    dst.copy_(host_x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    dst.copy_(host_x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("plain pinned copy ms %.3f  GB/s %.1f" % (dt * 1e3, hx.nbytes / dt / 1e9))
