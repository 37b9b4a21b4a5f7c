"""End-to-end entry point (pc_em_iteration_host): transfer pipeline depth sweep + phase split."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth
from poccala_b200.engine import Corpus, Engine, EStep, Model, em_iteration_host
N_UNITS, N_INITIALS, MIX, N_UTT, T, L, DIM = 57, 22, 16, 1000, 300, 10, 39
eng = Engine(0)
truth, init0, labels, x = synth.torch_corpus(N_UTT, T, L, N_UNITS, MIX, 2, eng.device, N_INITIALS)
corpus = Corpus(eng, labels, np.full(N_UTT, T, dtype=np.int32), N_UNITS)
tm0 = synth.default_transmat(N_UNITS)
host_x = torch.empty((N_UTT * T, DIM), dtype=torch.float32).pin_memory()
host_x.copy_(x.cpu())
hp = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in init0] + [tm0.copy()]
# raw copy speed
d = torch.empty_like(x)
for _ in range(3): d.copy_(host_x, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): d.copy_(host_x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("pinned H2D %.1f MB: %.3f ms = %.1f GB/s" % (host_x.numel() * 4 / 1e6, dt * 1e3, host_x.numel() * 4 / dt / 1e9))
for chunks in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    eng.set_option("host_chunks", chunks)
    for _ in range(3):
        p = [a.copy() for a in hp]; em_iteration_host(eng, corpus, host_x.numpy(), *p, c_covariance=1e-6)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        p = [a.copy() for a in hp]; ll = em_iteration_host(eng, corpus, host_x.numpy(), *p, c_covariance=1e-6)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print("host_chunks %d: %.3f ms per EM iteration (%.1f M frames/s), sum logp %.4f" % (chunks, dt * 1e3, N_UTT * T / dt / 1e6, ll), flush=True)
