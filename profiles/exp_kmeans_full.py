"""k-means (K5) on ONE state at the full configs[4] size (~175 000 frames per state) next to the capped size the bench
uses: python profiles/exp_kmeans_full.py [n_points]"""
import random, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200.engine import Engine, kmeans_run, kmeans_seed_points
eng = Engine(0)
K, D = 64, 39
rng = np.random.default_rng(5)
centres = rng.normal(0, 1.0, size=(K, D))
for n in [4096, 20480, int(sys.argv[1]) if len(sys.argv) > 1 else 175000]:
    data = centres[rng.integers(0, K, size=n)] + rng.normal(size=(n, D)) * 0.7
    random.seed(7)
    seeds = kmeans_seed_points(np.ascontiguousarray(data[:, 0]), K, random)
    x = torch.as_tensor(data).to(eng.device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = kmeans_run(eng, x, np.array([0, n], dtype=np.int64), K, np.array([seeds], dtype=np.int32))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    passes, moves = int(out["passes"][0]), int(out["moves"][0])
    print("points %7d  passes %6d  moves %8d  %.3f s  (%.1f GB/s of 9 B per point and arg-min)" % (n, passes, moves, dt, passes * K * n * 9.0 / dt / 1e9), flush=True)
