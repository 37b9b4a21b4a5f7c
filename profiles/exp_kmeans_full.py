"""k-means (K5) on ONE state at the full configs[4] size (~175 000 frames per state) next to the capped size the bench
uses, single CTA against thread-block cluster: python profiles/exp_kmeans_full.py [n_points] [compare_up_to]"""
import random, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200.engine import Engine, kmeans_run, kmeans_seed_points
eng = Engine(0)
K, D = 64, 39
rng = np.random.default_rng(5)
centres = rng.normal(0, 1.0, size=(K, D))
n_big = int(sys.argv[1]) if len(sys.argv) > 1 else 175000
cmp_to = int(sys.argv[2]) if len(sys.argv) > 2 else 70000
for n in [4096, 20480, 30000, 65000, n_big]:
    data = centres[rng.integers(0, K, size=n)] + rng.normal(size=(n, D)) * 0.7
    if n == 30000:
        data[::7, 0] = np.round(data[::7, 0], 1)  # duplicate metric coordinates: ties go to the lowest index
    random.seed(7)
    seeds = kmeans_seed_points(np.ascontiguousarray(data[:, 0]), K, random)
    x = torch.as_tensor(data).to(eng.device)
    res = {}
    for mode in (1, 0):
        if mode == 0 and (n > cmp_to or n <= 20480):
            continue
        eng.set_option("kmeans_cluster", mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = kmeans_run(eng, x, np.array([0, n], dtype=np.int64), K, np.array([seeds], dtype=np.int32))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        passes, moves = int(out["passes"][0]), int(out["moves"][0])
        res[mode] = out
        print("points %7d  %s  passes %6d  moves %8d  %.3f s  (%.1f GB/s of 9 B per point and arg-min)" % (
            n, "cluster   " if mode and n > 20480 else "single CTA", passes, moves, dt, passes * K * n * 9.0 / dt / 1e9), flush=True)
    eng.set_option("kmeans_cluster", 1)
    if 0 in res and 1 in res:
        same = all(torch.equal(res[0][k], res[1][k]) for k in ("owner", "member_list", "member_count", "passes", "moves", "mean", "var", "alpha"))
        print("   cluster == single CTA (owner, member lists, counts, passes, moves, mean, var, alpha): %s" % same, flush=True)
