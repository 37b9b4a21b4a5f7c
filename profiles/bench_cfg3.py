"""BASELINE.json configs[2]: GMM scoring sweep, 39-dim diagonal Gaussians, 64 mixtures per state,
F frames x G Gaussians on one B200 through the tcgen05 scoring kernel (Engine.score_dense_tc).
    python profiles/bench_cfg3.py [F [G,G,...]]   (default 10 000 000 frames for G = 4096, scaled down for larger G)
Prints one line per G: milliseconds, algorithmic TFLOP/s (158 flop per (frame, Gaussian) pair) and
the fraction of the measured bf16 peak."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from poccala_b200.engine import Engine

F_MAX = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
MIX, D = 64, 39
eng = Engine(0)
dev = eng.device
peak = 1400.0
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["bf16_tflops_sustained"])
gen = torch.Generator(device=dev); gen.manual_seed(3)
GS = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [4096, 8192, 16384, 32768, 65536]
for G in GS:
    F = min(F_MAX, F_MAX * 4096 // G * 2)  # keeps the emission buffer and the run time bounded
    x = torch.randn((F, D), generator=gen, device=dev, dtype=torch.float32)
    mean = torch.randn((G, D), generator=gen, device=dev, dtype=torch.float64)
    var = torch.rand((G, D), generator=gen, device=dev, dtype=torch.float64) + 0.5
    alpha = torch.full((G,), 1.0 / MIX, device=dev, dtype=torch.float64)
    # build once (descriptor tables, operand images), then time the kernel alone
    from poccala_b200 import _native as nat
    from poccala_b200.engine import Corpus, _p, _stream, EMIT
    S = G // MIX; U = (S + EMIT - 1) // EMIT
    pad = U * EMIT * MIX - G
    if pad:
        mean = torch.cat([mean, torch.zeros((pad, D), dtype=mean.dtype, device=dev)])
        var = torch.cat([var, torch.ones((pad, D), dtype=var.dtype, device=dev)])
        alpha = torch.cat([alpha, torch.zeros((pad,), dtype=alpha.dtype, device=dev)])
    n_frames = np.full((F + 383) // 384, 384, dtype=np.int32)
    if F % 384: n_frames[-1] = F % 384
    labels = np.ascontiguousarray(np.broadcast_to(np.arange(U, dtype=np.int32), (len(n_frames), U)))
    t0 = time.perf_counter()
    corpus = Corpus(eng, labels, n_frames, U)
    W = eng.pack_gmm(mean, var, alpha, mix=MIX)
    X = eng.prepare_frames(corpus, x)
    b = eng.empty((corpus.emis_floats,), torch.float32)
    torch.cuda.synchronize(); setup = time.perf_counter() - t0
    def run(): nat.call("pc_gmm_score", eng.h, corpus.c, _p(X), _p(W), MIX, _p(b), _stream())
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    tf = 158.0 * F * G / (ms * 1e-3) / 1e12
    print(json.dumps({"workload": "cfg3 GMM scoring sweep", "frames": F, "gaussians": G, "mix": MIX, "ms": ms,
                      "tflops_algorithmic": tf, "frac_of_bf16_peak": tf / peak, "frames_per_s": F / (ms * 1e-3),
                      "setup_s": setup}), flush=True)
    del corpus, W, X, b, x, mean, var, alpha
    torch.cuda.empty_cache()
