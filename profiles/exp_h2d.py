"""Pinned host -> device copy bandwidth on this box: streams in parallel, chunk sizes, NUMA hints."""
import os, time, torch
n = 46_800_000 // 4
host = torch.empty(n, dtype=torch.float32).pin_memory()
host.normal_()
dev = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
def one(): dev.copy_(host, non_blocking=True)
print("1 stream : %.3f ms  %.1f GB/s" % (t(one) * 1e3, n * 4 / t(one) / 1e9))
for k in (2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    def multi():
        step = n // k
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dev[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)
    dt = t(multi)
    print("%d streams: %.3f ms  %.1f GB/s" % (k, dt * 1e3, n * 4 / dt / 1e9))
def d2h(): host.copy_(dev, non_blocking=True)
dt = t(d2h); print("D2H      : %.3f ms  %.1f GB/s" % (dt * 1e3, n * 4 / dt / 1e9))
os.system("nvidia-smi topo -m 2>/dev/null | head -6; nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv 2>/dev/null; numactl -H 2>/dev/null | head -4; nproc")
