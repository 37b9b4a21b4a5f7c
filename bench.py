#!/usr/bin/env python
"""Headline benchmark: EM frames/sec (GMM score + forward-backward + accumulate [+ allreduce] +
M-step) on BASELINE.json configs[1] per GPU: Mandarin initial/final unit set (57 units), 3-state
HMMs, 16-mix GMMs, 1k synthetic utterances x 300 frames x 10 units, one embedded Baum-Welch
iteration per step.  Weak scaling: every rank owns its own 1k utterances; the only exchange step
is the NCCL allreduce of the accumulators.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--check] [--cfg5-utt U]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput, `e2e` the same iteration
through the host-buffer C-ABI entry point (pc_em_iteration_host) with pinned host inputs (at N > 1 with
the two NCCL collectives inside the call).  Further legs in the same line: `cfg5` = BASELINE.json
configs[4] (100k utterances over the N ranks, 64-mix, k-means init + 5 EM iterations), at N = 1 also
`viterbi` (configs[3]) and `scoring_sweep` (configs[2] at 10M frames x 4 096 and 65 536 Gaussians).
`--check`: the same 2 000 utterances on one rank and on N ranks, models compared (no timing).
`--impl reference` times the CPU port of the reference (oracle/ref_port.py, the reference's own
cost structure: /root/reference does not exist on the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UNITS, N_INITIALS, MIX, N_UTT, T, L, DIM = 57, 22, 16, 1000, 300, 10, 39
WORKLOAD = "cfg2: IF units(57) x 3 states x 16-mix, 1000 utt x 300 frames x 10 units per GPU, 1 EM iteration/step"
METRIC = "EM frames/sec (GMM score+fwd-bwd+accum)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def ncu_traffic(stage):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the stage's kernel from the
    committed `ncu --set full` excerpt of this workload (profiles/current/, see profiles/README.md);
    None when the excerpt is missing."""
    name = {"K1_score": "score_tc_wide_kernel", "K2_forward_backward": "fwdbwd_warp_kernel",
            "K3_accumulate": "accumulate_tcx_kernel"}.get(stage)
    path = os.path.join(ROOT, "profiles", "current", "%s.csv" % name)
    if not name or not os.path.exists(path):
        return None
    total, scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    with open(path) as f:
        for line in f:
            parts = line.strip().split(",")
            if len(parts) >= 3 and parts[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(parts[1]) * scale.get(parts[2], 1.0)
    return total or None


# ------------------------------------------------------------------------------------ CPU arm
_CPU = {}


def _cpu_init(seed, mix=None):
    from oracle import ref_port as rp
    from poccala_b200 import synth

    MIX = mix or globals()["MIX"]
    truth = synth.make_truth(N_UNITS, MIX, 1000 * seed + 7)
    init = synth.perturb(*truth, seed=1000 * seed + 11)
    names = [str(i) for i in range(N_UNITS)]
    units = {}
    for i, u in enumerate(names):
        d = rp.new_unit(5, MIX, DIM)
        for r, g in enumerate(d["gmms"]):
            g["mean"], g["var"], g["alpha"] = init[0][i, r].copy(), init[1][i, r].copy(), init[2][i, r].copy()
        units[u] = d
    _CPU.update(rp=rp, synth=synth, truth=truth, units=units, names=names,
                labels=synth.random_labels(N_UTT, L, N_UNITS, 1000 * seed + 17, N_INITIALS), seed=seed)


def _cpu_task(i):
    rp, synth = _CPU["rp"], _CPU["synth"]
    lab = _CPU["labels"][i % N_UTT]
    X = synth.make_utterance(lab, T, _CPU["truth"], 1000 * _CPU["seed"] + 100 + i)
    t0 = time.perf_counter()
    r = rp.estep_utterance(_CPU["units"], [_CPU["names"][k] for k in lab], X)
    return time.perf_counter() - t0, float(r["logp"])


def cpu_arm(n_utt_sample, cores, seed=2, steps=1, mix=None):
    """E-step of the reference's CPU path on `n_utt_sample` utterances of the bench workload per
    step, one utterance per task on `cores` processes (AcousticModel.py:861-870).
    Returns (frames/s over all steps, [wall seconds per step])."""
    ctx = mp.get_context("spawn")
    walls = []
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(seed, mix)) as pool:
        pool.map(_cpu_task, range(cores))  # warm-up: imports, page-in (1 utterance per core)
        for k in range(steps):
            t0 = time.perf_counter()
            pool.map(_cpu_task, range(k * n_utt_sample, (k + 1) * n_utt_sample), chunksize=1)
            walls.append(time.perf_counter() - t0)
    return steps * n_utt_sample * T / sum(walls), walls


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = cores  # one utterance per core and step: a few seconds of wall per step at cfg2 shape
    steps = max(1, min(args.steps, 8))
    value, walls = cpu_arm(per_step, cores, steps=steps)
    sample = ("%d utterances x %d frames per step (%d steps; 1 warm-up utterance per process), one utterance "
              "per process" % (per_step, T, steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled through NVML every ~2 ms while the timed region runs
    (nvidia-smi is too slow for a region of tens of milliseconds).  NVML is initialised before the
    region starts (round 1 lost every sample to a slow nvmlInit)."""

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)
        self.max_mhz = None
        self.err = None
        self.h = self.nv = self.get_reasons = None
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = index
            if vis:
                try:
                    idx = int(vis.split(",")[index])
                except ValueError:
                    idx = index
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self._sample()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _sample(self):
        nv = self.nv
        self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(self.get_reasons(self.h)),
                          time.perf_counter()))

    def _run(self):
        try:
            while self.h is not None and not self.stop.is_set():
                self._sample()
                self.stop.wait(0.002)
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        if self.h is not None and not self.err:
            try:
                self._sample()
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self, t0=None, t1=None):
        if t0 is not None:
            inside = [r for r in self.rows if t0 <= r[2] <= t1 + 0.01]
            self.rows = inside if inside else self.rows[-1:]
        bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                "hw_power_brake_slowdown": 0x80, "sw_power_cap": 0x4}
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for r in self.rows for n, b in bits.items() if r[1] & b})
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
               "samples": len(sm)}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------ GPU arm
def _bind_to_gpu_numa_node(local):
    """Pin this process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated: the
    first touch then places the frames on the GPU's own NUMA node, and 8 ranks copying at once do not cross the socket
    interconnect.  Returns the number of CPUs bound to (None when NVML / the affinity call are not available)."""
    if os.environ.get("PC_BENCH_NO_AFFINITY"):
        return None
    try:
        import pynvml as nv

        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local
        if vis:
            try:
                idx = int(vis.split(",")[local])
            except ValueError:
                idx = local
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        n_cpu = os.cpu_count() or 1
        words = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


def _init_group(world, local):
    """NCCL process group of the launch (torchrun environment), or None at world size 1."""
    import torch
    import torch.distributed as dist

    if world <= 1:
        return None
    # the image exports NCCL_DEBUG=VERSION, whose banner goes to stdout: rank 0 must print ONE JSON line
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    # ... and whatever NCCL still writes while the communicator comes up goes to stderr
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
        warm = torch.zeros(1, device=torch.device("cuda", local))
        dist.all_reduce(warm, group=group)
        dist.all_reduce(warm, op=dist.ReduceOp.MAX, group=group)
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    return group


def run_check(args):
    """`--check`: the same 2 000 utterances (configs[1] shape) trained for two EM iterations (a) on every
    rank alone, whole corpus, and (b) sharded over the N ranks (utterance u -> rank u mod N) with the NCCL
    reductions - through the device-resident path and through the host-buffer entry point with its reduce
    hook.  Asserts: replicas bit-identical, sharded == single-rank model (5e-5 after one iteration).  Prints one
    JSON line; exit code 1 on failure."""
    import torch
    import torch.distributed as dist

    from poccala_b200 import synth
    from poccala_b200.engine import (Corpus, Engine, EStep, HostReduceHook, Model, em_iteration_host,
                                     frame_moments_host)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    group = _init_group(world, local)
    eng = Engine(local)
    dev = eng.device
    n_all, iters = 2000, 4  # four full iterations (both statistic sets of the peer exchange are reused once) + two partial ones
    truth, init0, labels, x = synth.torch_corpus(n_all, T, L, N_UNITS, MIX, 2, dev, N_INITIALS)  # identical on every rank
    tm0 = synth.default_transmat(N_UNITS)

    def train(sel, grp, peer=None):
        xs = x.view(n_all, T, DIM)[sel].reshape(-1, DIM).contiguous()
        corpus = Corpus(eng, labels[sel], np.full(len(sel), T, dtype=np.int32), N_UNITS)
        model = Model(eng, init0[0], init0[1], init0[2], tm0)
        es = EStep(eng, corpus, model)
        es.load_frames(xs, group=grp)
        if peer is not None:
            es.use_peer(peer)
        ll, snaps = [], []
        # the fifth and sixth iteration keep the GMMs / the transitions fixed (fix_code 2 / 4, LHMM.py:66-75): the
        # branches of the reductions that skip one of the two statistic sets
        for it in range(iters + 2):
            fix = 0 if it < iters else (2 if it == iters else 4)
            es.em_iteration(c_covariance=1e-6, fix_code=fix, group=grp)
            s_ll = es.utt_logp.sum().reshape(1)
            if grp is not None:
                dist.all_reduce(s_ll, group=grp)
            ll.append(float(s_ll.item()))
            snaps.append([model.mean.clone(), model.var.clone(), model.alpha.clone(), model.transmat.clone()])
        torch.cuda.synchronize()
        return snaps, ll, corpus, xs

    single_all, ll1, _, _ = train(np.arange(n_all), None)
    mine = np.arange(rank, n_all, world)
    sharded_all, llN, corpus_s, xs = train(mine, group)
    # the same shards with the reduction over peer memory (no collective call inside the iterations)
    peer = peer_all = None
    if group is not None:
        from poccala_b200.distributed import PeerExchange

        peer = PeerExchange(eng, N_UNITS, N_UNITS * 3 * MIX, group)
        peer_all, llP, _, _ = train(mine, group, peer)

    def diff(got, ref):
        """Worst deviation of a model from a reference model: means in units of max(|mean|, standard deviation),
        variances relative, weights and transition probabilities relative above 1e-3."""
        sd = ref[1].sqrt()
        return {"mean": float(((got[0] - ref[0]).abs() / torch.maximum(ref[0].abs(), sd)).max().item()),
                "var": float(((got[1] - ref[1]).abs() / ref[1]).max().item()),
                "alpha": float(((got[2] - ref[2]).abs() / ref[2].clamp_min(1e-3)).max().item()),
                "transmat": float(((got[3] - ref[3]).abs() / ref[3].clamp_min(1e-3)).max().item())}

    # One iteration from the same model: the shards change only the order of the fp32 partial sums inside the
    # accumulation kernel (a work item covers other tiles) - 5e-5 (measured on a variance: 1.6e-5 at 2 ranks,
    # 3.0e-5 at 8; 7e-7 on the means).  The second iteration starts from models that
    # differ by that much, and components with little occupancy amplify it - 1e-3.
    diffs1, diffs = diff(sharded_all[0], single_all[0]), diff(sharded_all[1], single_all[1])
    sharded = sharded_all[-1]
    identical = True
    peer_diffs = peer_host_diffs = None
    if peer is not None:
        # peer-memory reduction against the NCCL reduction of the same shards: the same sums up to the order of the
        # N terms (and the same at N = 2)
        peer_diffs = diff(peer_all[-1], sharded_all[-1])
        sharded = list(sharded) + list(peer_all[-1])
    if group is not None:
        for t in sharded:
            hi, lo = t.clone(), t.clone()
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            identical = identical and bool((hi == lo).all().item())
    # host-buffer entry point, one iteration, with the collectives inside the call
    host_x = xs.cpu().numpy()
    shift, isc = frame_moments_host(eng, host_x, group=group)
    hook = HostReduceHook(eng, N_UNITS, N_UNITS * 3 * MIX, group) if group is not None else None
    hp = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in init0] + [tm0.copy()]
    em_iteration_host(eng, corpus_s, host_x, *hp, c_covariance=1e-6, shift=shift, inv_scale=isc)
    if hook is not None:
        if hook.error is not None:
            raise hook.error
        hook.remove()
    # against the single-rank model after ONE iteration
    host_diffs = diff([torch.as_tensor(a).to(dev) for a in hp], single_all[0])
    ok = identical and max(diffs1.values()) <= 5e-5 and max(host_diffs.values()) <= 5e-5 and \
        max(diffs.values()) <= 1e-3 and abs(llN[1] - ll1[1]) <= 1e-8 * abs(ll1[1])
    peer_timeouts = None
    if peer is not None:
        # the host entry point without a hook: the connected exchange block makes it reduce over peer memory
        hp2 = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in init0] + [tm0.copy()]
        em_iteration_host(eng, corpus_s, host_x, *hp2, c_covariance=1e-6, shift=shift, inv_scale=isc)
        peer_host_diffs = diff([torch.as_tensor(a).to(dev) for a in hp2], single_all[0])
        peer_timeouts = peer.timeouts()
        ok = ok and max(peer_diffs.values()) <= 1e-7 and max(peer_host_diffs.values()) <= 5e-5 and peer_timeouts == 0
        peer.close()
    flag = torch.tensor([0 if ok else 1], device=dev)
    if group is not None:
        dist.all_reduce(flag, group=group)
    ok_all = int(flag.item()) == 0
    if rank == 0:
        print(json.dumps({"check": "ok" if ok_all else "FAILED", "n_gpus": world, "utterances": n_all, "iterations": iters,
                          "replicas_bit_identical": identical, "diff_vs_single_rank_iteration_1": diffs1,
                          "diff_vs_single_rank_iteration_2": diffs,
                          "host_entry_diff_vs_single_rank_iteration_1": host_diffs,
                          "peer_memory_vs_nccl_iteration_6": peer_diffs,
                          "peer_memory_host_entry_diff_vs_single_rank_iteration_1": peer_host_diffs,
                          "peer_timeouts": peer_timeouts,
                          "sum_logp_single": ll1, "sum_logp_sharded": llN}), flush=True)
    if group is not None:
        dist.destroy_process_group()
    if not ok_all:
        raise SystemExit(1)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from poccala_b200 import synth
    from poccala_b200.engine import Corpus, Engine, EStep, Model, em_iteration_host

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU fallback")
    torch.cuda.set_device(local)
    numa_cpus = _bind_to_gpu_numa_node(local)
    group = _init_group(world, local)
    eng = Engine(local)
    dev = eng.device
    # every rank starts from the SAME model (replicated parameters) and owns its own utterances
    truth, init0, labels, x = synth.torch_corpus(N_UTT, T, L, N_UNITS, MIX, 2, dev, N_INITIALS, data_seed=2 + rank)
    corpus = Corpus(eng, labels, np.full(N_UTT, T, dtype=np.int32), N_UNITS)
    tm0 = synth.default_transmat(N_UNITS)
    model = Model(eng, init0[0], init0[1], init0[2], tm0)
    es = EStep(eng, corpus, model)
    es.load_frames(x, group=group)  # corpus-wide standardisation: identical on every rank
    peer = None
    peer_error = None
    if group is not None and args.collective in ("peer", "auto"):
        from poccala_b200.distributed import PeerExchange

        try:
            peer = PeerExchange(eng, N_UNITS, N_UNITS * 3 * MIX, group)  # (raises on every rank or on none)
            es.use_peer(peer)
        except RuntimeError as e:
            if args.collective == "peer":
                raise
            peer_error = str(e)  # auto: NCCL it is
    frames = corpus.total_frames
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def reset_model():
        model.mean.copy_(torch.as_tensor(init0[0]).to(dev))
        model.var.copy_(torch.as_tensor(init0[1]).to(dev))
        model.alpha.copy_(torch.as_tensor(init0[2]).to(dev))
        model.transmat.copy_(torch.as_tensor(tm0).to(dev))

    def sync_all():
        torch.cuda.synchronize()
        if group is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def step(ev=None):
        if ev is None:
            es.em_iteration(c_covariance=1e-6, group=group)
            return
        ev[0].record()
        if getattr(es, "peer", None) is not None:
            es._peer_bind()  # this iteration's statistic set of the exchange block
        if not os.environ.get("PC_NO_BANDS_ASYNC"):
            es.log_bands_async()  # log(transmat) bands on the side stream, beside K1
        es.score(); ev[1].record(); es.forward_backward(); ev[2].record()
        es.reduce_transitions_async(group)  # side stream: transition log-sum-exp + the MAX collective, under K3
        es.accumulate(); ev[3].record()
        es.reduce_statistics(group)  # join + the SUM collective over the flat statistics buffer
        es.mstep(c_covariance=1e-6)
        ev[4].record()

    trial_ms = None
    if peer is not None and args.collective == "auto":
        # both reductions for a few steps each (device time, max over ranks: every rank takes the same decision)
        trial_ms = {}
        for mode in ("nccl", "peer", "nccl", "peer"):
            es.use_peer(peer) if mode == "peer" else es.use_nccl()
            for _ in range(2):
                flush.zero_()
                step()
            sync_all()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(6):
                step()
            t1.record()
            torch.cuda.synchronize()
            tm = torch.tensor([t0.elapsed_time(t1) / 6], dtype=torch.float64, device=dev)
            dist.all_reduce(tm, op=dist.ReduceOp.MAX, group=group)
            trial_ms[mode] = min(trial_ms.get(mode, 1e9), float(tm.item()))
        if trial_ms["peer"] <= trial_ms["nccl"]:
            es.use_peer(peer)
        else:
            es.use_nccl()
            peer.close()
            peer = None
    for _ in range(max(args.warmup, 3)):
        flush.zero_()
        step()
    reset_model()
    sync_all()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    l0 = eng.launches
    with ClockSampler(local) as clocks:
        sync_all()
        w0 = time.perf_counter()
        for k in range(args.steps):
            flush.zero_()  # L2 flush between timed steps (outside the per-step events)
            step(evs[k])
        sync_all()
        wall = time.perf_counter() - w0
    launches = eng.launches - l0
    from poccala_b200 import _native as nat
    active_tiles = int(nat.lib().pc_corpus_active_tiles(corpus.c))
    total_tiles = int(nat.lib().pc_corpus_total_tiles(corpus.c))
    step_ms = [e[0].elapsed_time(e[4]) for e in evs]
    k1 = [e[0].elapsed_time(e[1]) for e in evs]
    k2 = [e[1].elapsed_time(e[2]) for e in evs]
    k3 = [e[2].elapsed_time(e[3]) for e in evs]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX, group=group)
    total_ms = float(total_ms.item())
    value = world * frames * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (pinned host frames, H2D + D2H inside).
    # The corpus' standardisation constants are computed once (pc_frame_moments_host + all-reduce of the
    # moments); at N > 1 the reduce hook puts the two NCCL collectives inside the call, so an e2e step is
    # one distributed EM iteration that leaves identical models on every rank.
    from poccala_b200.engine import HostReduceHook, frame_moments_host

    host_x = torch.empty((frames, DIM), dtype=torch.float32).pin_memory()
    host_x.copy_(x.cpu())
    # the model lives in pinned host buffers too (updated in place by every call, as a training loop would keep it)
    hp_t = [torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).clone().pin_memory() for a in list(init0) + [tm0]]
    hp = [t.numpy() for t in hp_t]
    shift_h, isc_h = frame_moments_host(eng, host_x.numpy(), group=group)
    # N > 1: the connected exchange block makes the call reduce over peer memory; with --collective nccl the reduce
    # hook queues the two all-reduces instead
    hook = HostReduceHook(eng, N_UNITS, N_UNITS * 3 * MIX, group) if group is not None and peer is None else None
    e2e_steps = max(3, min(args.steps, 10))
    host_frames = host_x.numpy()
    for _ in range(2):
        em_iteration_host(eng, corpus, host_frames, *hp, c_covariance=1e-6, shift=shift_h, inv_scale=isc_h)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        em_iteration_host(eng, corpus, host_frames, *hp, c_covariance=1e-6, shift=shift_h, inv_scale=isc_h)
    sync_all()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX, group=group)
    e2e_val = world * frames * e2e_steps / float(e2e_s.item())
    if hook is not None:
        if hook.error is not None:
            raise hook.error
        hook.remove()
    G = N_UNITS * 3 * MIX
    h2d = frames * DIM * 4 + (2 * G * DIM + G + N_UNITS * 25 + 2 * DIM) * 8
    d2h = (2 * G * DIM + G + N_UNITS * 25) * 8 + 8
    # the plain pinned copy of the same frames, for scale (the e2e step is transfer-bound)
    dst = torch.empty((frames, DIM), dtype=torch.float32, device=dev)
    dst.copy_(host_x, non_blocking=True)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(3):
        dst.copy_(host_x, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = 3 * frames * DIM * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del dst, host_x
    pk, pk_src = peaks()
    kern_ms = {"K1_score": statistics.mean(k1), "K2_forward_backward": statistics.mean(k2),
               "K3_accumulate": statistics.mean(k3)}
    peer_used = peer is not None
    peer_timeouts = None
    if peer is not None:
        peer_timeouts = peer.timeouts()
        del es
        peer.close()
        peer = None
    del model, corpus, x
    es = None
    torch.cuda.empty_cache()

    # ---- BASELINE.json configs[4]: every rank takes part (the job is split over the ranks)
    cfg5 = None
    if args.cfg5_utt > 0:
        cfg5 = cfg5_leg(eng, pk, group, world, rank, args.cfg5_utt, sync_all,
                        use_peer=("auto" if args.collective == "auto" and peer_error is None else peer_used))

    if rank != 0:
        if group is not None:
            dist.destroy_process_group()
        return
    vit = sweep = None
    if world == 1 and not os.environ.get("PC_BENCH_SHORT"):  # single-GPU configurations of BASELINE.json (PC_BENCH_SHORT: profiler runs)
        vit = viterbi_leg(eng, pk)
        sweep = [scoring_sweep_leg(eng, pk, F=10_000_000, G=g) for g in (4096, 65536)]
    # dominant kernel and its roofline: K1/K3 are contractions, 158 flops per
    # (frame, Gaussian) pair; K2 moves 8 B per (emitting state, frame)
    pairs = frames * 3 * L * MIX
    kern = {"K1_score": (kern_ms["K1_score"], "tensor", 158.0 * pairs),
            "K2_forward_backward": (kern_ms["K2_forward_backward"], "hbm", 8.0 * frames * 3 * L),
            "K3_accumulate": (kern_ms["K3_accumulate"], "tensor", 158.0 * pairs)}
    dom = max(kern, key=lambda k: kern[k][0])
    ms, bound, work = kern[dom]
    if bound == "tensor":
        ach, peak, unit = work / (ms * 1e-3) / 1e12, float(pk["bf16_tflops_sustained"]), "TFLOP/s"
    else:
        ach, peak, unit = work / (ms * 1e-3) / 1e9, float(pk["hbm_gbs"]), "GB/s"
    roofline = {"kernel": dom, "bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                "traffic": ncu_traffic(dom), "peak_source": pk_src + (" (bf16 sustained)" if bound == "tensor" else ""),
                "ms_per_launch": ms,
                "all_ms": {k: v[0] for k, v in kern.items()},
                "note": ("achieved = 158 flop per (frame, Gaussian) pair of the corpus / event time of the stage "
                         "(K1 stage includes the model packing launch; the 3-product fp16 split executes 3x these "
                         "flops on the tensor pipe); K3 contracts only (tile, unit) pairs with posterior mass: "
                         "%d of %d active" % (active_tiles, total_tiles))}
    roofline["all_frac"] = {k: (v[2] / (v[0] * 1e-3) / (1e12 * float(pk["bf16_tflops_sustained"]) if v[1] == "tensor"
                                                       else 1e9 * float(pk["hbm_gbs"]))) for k, v in kern.items()}
    cores = os.cpu_count() or 1
    cpu = {}
    if world == 1 and not os.environ.get("PC_BENCH_NO_CPU"):
        n_sample = 2 * cores
        cpu_v, cpu_walls = cpu_arm(n_sample, cores)
        cpu = {"value": cpu_v, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "%d utterances x %d frames of the same workload (E-step), %.1f s" % (n_sample, T, sum(cpu_walls))}
        if cfg5 is not None:  # the reference CPU path at the configs[4] shape (64 mixtures), for the north-star ratio
            c5_v, c5_walls = cpu_arm(cores, cores, mix=64)
            cfg5["cpu_baseline"] = {"value": c5_v, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "%d utterances x %d frames, 64-mix (E-step), %.1f s" % (cores, T, sum(c5_walls))}
            cfg5["speedup_vs_cpu"] = cfg5["value"] / c5_v
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "l2": "256 MiB memset between timed steps (outside the per-step events)",
                   "k3_active_pair_frac": active_tiles / max(total_tiles, 1),
                   "wall_s_timed_region": wall, "parallelism": "dp%d" % world},
        "clocks": clocks.summary(w0, w0 + wall),
        "e2e": {"value": e2e_val, "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "api": "pc_em_iteration_host", "pinned_h2d_gbs": h2d_gbs, "cpus_bound_to_gpu_numa_node": numa_cpus,
                "collectives_inside": world > 1},
        "collective": (None if world == 1 else ("peer-memory reduction inside the M-step kernels (CUDA IPC, NVLink loads)"
                                                if peer_used else "NCCL all-reduce (MAX + SUM)")),
        "peer_timeouts": peer_timeouts, "collective_trial_ms_per_step": trial_ms, "peer_unavailable": peer_error,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cfg5": cfg5,
        "viterbi": vit,
        "scoring_sweep": sweep,
        "cpu_baseline": cpu or None,
    }
    print(json.dumps(line), flush=True)
    if group is not None:
        dist.destroy_process_group()


def scoring_sweep_leg(eng, pk, F=10_000_000, G=4096, mix=64, slab=2_500_000):
    """BASELINE.json configs[2] (GMM scoring sweep) on one GPU: F frames x G 39-dim diagonal Gaussians, 64
    mixtures per state, through the tcgen05 scoring kernel (Engine.score_dense_tc's layout), in slabs of
    `slab` frames so that the emission buffer stays bounded (F x G / 64 floats: 41 GB at 65 536 Gaussians).
    Every slab's frames are prepared before its timed launch (inputs resident in HBM, larger than L2).
    Tensor roofline: 158 flop per (frame, Gaussian) pair against the measured bf16 peak (the 3-product
    fp16 split executes 3x these flops)."""
    import torch

    from poccala_b200 import _native as nat
    from poccala_b200.engine import EMIT, Corpus, _p, _stream

    dev = eng.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    mean = torch.randn((G, DIM), generator=gen, device=dev, dtype=torch.float64)
    var = torch.rand((G, DIM), generator=gen, device=dev, dtype=torch.float64) + 0.5
    alpha = torch.full((G,), 1.0 / mix, device=dev, dtype=torch.float64)
    U = (G // mix + EMIT - 1) // EMIT
    pad = U * EMIT * mix - G
    if pad:
        mean = torch.cat([mean, torch.zeros((pad, DIM), dtype=mean.dtype, device=dev)])
        var = torch.cat([var, torch.ones((pad, DIM), dtype=var.dtype, device=dev)])
        alpha = torch.cat([alpha, torch.zeros((pad,), dtype=alpha.dtype, device=dev)])
    slab = min(slab, F)
    n_slabs = (F + slab - 1) // slab
    n_frames = np.full((slab + 383) // 384, 384, dtype=np.int32)
    if slab % 384:
        n_frames[-1] = slab % 384
    labels = np.ascontiguousarray(np.broadcast_to(np.arange(U, dtype=np.int32), (len(n_frames), U)))
    corpus = Corpus(eng, labels, n_frames, U)
    W = eng.pack_gmm(mean, var, alpha, mix=mix)
    b = eng.empty((corpus.emis_floats,), torch.float32)
    X = None
    total_ms = 0.0
    for k in range(n_slabs + 1):  # slab 0 runs twice: the first pass is the warm-up
        x = torch.randn((slab, DIM), generator=gen, device=dev, dtype=torch.float32)
        X = eng.prepare_frames(corpus, x, out=X)
        del x
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.call("pc_gmm_score", eng.h, corpus.c, _p(X), _p(W), mix, _p(b), _stream())
        e1.record()
        torch.cuda.synchronize()
        if k > 0:
            total_ms += e0.elapsed_time(e1)
    frames = n_slabs * slab
    tf = 158.0 * frames * G / (total_ms * 1e-3) / 1e12
    peak = float(pk["bf16_tflops_sustained"])
    del corpus, W, X, b
    torch.cuda.empty_cache()
    return {"value": frames / (total_ms * 1e-3), "unit": "frames/s", "ms": total_ms,
            "workload": "cfg3: %d frames x %d Gaussians (39-dim diag, %d-mix) in %d launches of %d frames, inputs larger "
                        "than L2" % (frames, G, mix, n_slabs, slab),
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak}}


def cfg5_leg(eng, pk, group, world, rank, n_total, sync_all, T5=300, L5=10, mix=64, em_iters=5, kmeans_points=4096,
             use_peer=False):
    """BASELINE.json configs[4]: data-parallel EM on `n_total` synthetic utterances (100k) split over the
    ranks, 64-mix IF HMMs: uniform segmentation -> per-state k-means (K = 64, the 171 states sharded over the
    ranks, parameters all-gathered) -> 5 Baum-Welch iterations with the NCCL accumulator all-reduce.  The
    k-means runs on the first `kmeans_points` frames of every state pooled over the ranks: the reference's
    greedy algorithm (and the bit-exact device kernel) costs O(N^2 / K) per state, 175k frames per state
    would take minutes - the cap is part of the workload string.  Device times, max over ranks."""
    import torch
    import torch.distributed as dist

    from poccala_b200 import _native as nat
    from poccala_b200 import synth
    from poccala_b200.engine import (EMIT, Corpus, EStep, Model, gather_state_data, group_frames, kmeans_states,
                                     segment_keys)

    dev = eng.device
    n_utt = n_total // world
    truth, _, labels, x = synth.torch_corpus(n_utt, T5, L5, N_UNITS, mix, 5, dev, N_INITIALS, data_seed=5 + rank)
    corpus = Corpus(eng, labels, np.full(n_utt, T5, dtype=np.int32), N_UNITS)
    S = N_UNITS * EMIT

    def max_over_ranks(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        if group is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    # ---- initialisation (AcousticModel.process_data(mode=1, init=True) + __cal_gmm's k-means)
    sync_all()
    t0 = time.perf_counter()
    key, _ = segment_keys(eng, corpus, None)  # uniform segmentation, three-way split
    grouped = group_frames(eng, key, S, x.double())
    data, key_off = gather_state_data(eng, grouped["data"], grouped["key_off"], group, max_points=kmeans_points)
    del grouped
    torch.cuda.synchronize()
    t_seg = max_over_ranks(time.perf_counter() - t0)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    km = kmeans_states(eng, data, key_off, mix, group)
    e1.record()
    torch.cuda.synchronize()
    t_km = max_over_ranks(time.perf_counter() - t0)
    km_dev_ms = max_over_ranks(e0.elapsed_time(e1))  # includes the host-side seeding draws between the launches
    pts = np.diff(key_off)
    passes = km["passes"].cpu().numpy()
    # K5 traffic: every pass runs up to K masked arg-mins over the state's points, 8 B coordinate + 1 B owner each
    scan_bytes = float(np.sum(passes[km["states"]] * mix * pts[km["states"]] * 9.0)) if km["states"] else 0.0
    mean = km["mean"].cpu().numpy().reshape(N_UNITS, EMIT, mix, -1)
    var = np.maximum(km["var"].cpu().numpy().reshape(N_UNITS, EMIT, mix, -1), 1e-2)
    alpha = km["alpha"].cpu().numpy().reshape(N_UNITS, EMIT, mix)
    alpha = alpha / alpha.sum(-1, keepdims=True)
    del data, km

    # ---- EM
    model = Model(eng, mean, var, alpha, synth.default_transmat(N_UNITS))
    es = EStep(eng, corpus, model)
    es.load_frames(x, group=group)
    del x
    peer = None
    trial = None
    if use_peer and group is not None:
        from poccala_b200.distributed import PeerExchange

        try:
            peer = PeerExchange(eng, N_UNITS, N_UNITS * EMIT * mix, group)  # (raises on every rank or on none)
            es.use_peer(peer)
        except RuntimeError:
            if use_peer != "auto":
                raise
            peer = None
    es.em_iteration(c_covariance=1e-3, group=group)  # warm-up (first launches, NCCL buffers); its update is kept
    if peer is not None and use_peer == "auto":
        trial = {}
        for mode in ("nccl", "peer"):  # one iteration each after a warm-up of its own; the updates are kept
            es.use_peer(peer) if mode == "peer" else es.use_nccl()
            es.em_iteration(c_covariance=1e-3, group=group)
            sync_all()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            es.em_iteration(c_covariance=1e-3, group=group)
            t1.record()
            torch.cuda.synchronize()
            trial[mode] = max_over_ranks(t0.elapsed_time(t1))
        if trial["peer"] <= trial["nccl"]:
            es.use_peer(peer)
        else:
            es.use_nccl()
            peer.close()
            peer = None
    ms, ll = [], []
    for it in range(em_iters):
        sync_all()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        es.em_iteration(c_covariance=1e-3, group=group)
        a1.record()
        torch.cuda.synchronize()
        ms.append(max_over_ranks(a0.elapsed_time(a1)))
        s_ll = es.utt_logp.sum().reshape(1)
        if group is not None:
            dist.all_reduce(s_ll, group=group)
        ll.append(float(s_ll.item()))
    # stage split of one more iteration (events on the launch stream)
    sync_all()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    if getattr(es, "peer", None) is not None:
        es._peer_bind()
    es.log_bands_async(); es.score(); ev[1].record(); es.forward_backward(); ev[2].record()
    es.reduce_transitions_async(group); es.accumulate(); ev[3].record()
    es.reduce_statistics(group); es.mstep(c_covariance=1e-3); ev[4].record()
    torch.cuda.synchronize()
    stage = {"K1_score": ev[0].elapsed_time(ev[1]), "K2_forward_backward": ev[1].elapsed_time(ev[2]),
             "K3_accumulate": ev[2].elapsed_time(ev[3]), "reduce_mstep": ev[3].elapsed_time(ev[4])}
    if os.environ.get("PC_BENCH_RANK_STAGES"):
        print("rank %d cfg5 stages %s" % (rank, {k: round(v, 4) for k, v in stage.items()}), file=sys.stderr, flush=True)
    stage = {k: max_over_ranks(v) for k, v in stage.items()}
    active = nat.lib().pc_corpus_active_tiles(corpus.c) / max(nat.lib().pc_corpus_total_tiles(corpus.c), 1)
    frames_rank = n_utt * T5
    pairs = frames_rank * EMIT * L5 * mix
    tpeak, hpeak = 1e12 * float(pk["bf16_tflops_sustained"]), 1e9 * float(pk["hbm_gbs"])
    it_ms = statistics.mean(ms)
    out = {
        "workload": "cfg5: %d utterances x %d frames x %d IF units over %d GPU(s) (%d per rank), %d-mix, uniform "
                    "segmentation + per-state k-means (K = %d, <= %d pooled frames per state, states sharded over the "
                    "ranks) + %d EM iterations with the cross-rank reduction" % (n_utt * world, T5, L5, world, n_utt, mix, mix,
                                                                        kmeans_points, em_iters),
        "n_gpus": world, "value": world * frames_rank / (it_ms * 1e-3), "unit": "frames/s",
        "ms_per_iteration": it_ms, "em_ms": ms, "stage_ms": stage,
        "roofline_frac": {"K1_score": 158.0 * pairs / (stage["K1_score"] * 1e-3) / tpeak,
                          "K2_forward_backward": 8.0 * frames_rank * EMIT * L5 / (stage["K2_forward_backward"] * 1e-3) / hpeak,
                          "K3_accumulate": 158.0 * pairs / (stage["K3_accumulate"] * 1e-3) / tpeak},
        "kmeans": {"wall_s": t_km, "device_ms": km_dev_ms, "states": S, "points": int(pts.sum()),
                   "passes_max": int(passes.max()) if len(passes) else 0,
                   "scan_gbs": scan_bytes / max(km_dev_ms * 1e-3, 1e-9) / 1e9,
                   "bound": "latency: K dependent masked arg-mins per pass over shared-memory resident points"},
        "segmentation_s": t_seg, "sum_logp": ll, "k3_active_pair_frac": active,
        "collective": None if group is None else ("peer memory" if peer is not None else "nccl"),
        "collective_trial_ms_per_iteration": trial,
    }
    del es, model, corpus
    if peer is not None:
        peer.close()
    torch.cuda.empty_cache()
    return out


def viterbi_leg(eng, pk, n_utt=10000, T_v=1000, L_v=20, reps=5):
    """Second half of BASELINE.json's metric: Viterbi forced alignment frames/s on the configs[3]
    shape (1000-frame utterances against 20 concatenated IF unit HMMs, N = 62 states, 16-mix
    emissions scored by K1) at its full size, 10 000 utterances, emissions resident in HBM.  HBM roofline:
    4 B per (emitting state, frame) read + 4 B per frame written."""
    import torch

    from poccala_b200 import synth
    from poccala_b200.engine import Corpus, EStep, Model, host_log_bands, viterbi

    dev = eng.device
    truth, init0, labels, x = synth.torch_corpus(n_utt, T_v, L_v, N_UNITS, MIX, 4, dev, N_INITIALS)
    corpus = Corpus(eng, labels, np.full(n_utt, T_v, dtype=np.int32), N_UNITS)
    tm0 = synth.default_transmat(N_UNITS)
    model = Model(eng, init0[0], init0[1], init0[2], tm0)
    es = EStep(eng, corpus, model)
    es.load_frames(x)
    es.score()
    ls, ln = host_log_bands(tm0, dev)
    n_states = 3 * L_v + 2
    logpi = torch.full((n_utt,), float(np.log(np.ones(n_states) / n_states)[0]), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(2):
        viterbi(eng, corpus, es.b, ls, ln, utt_logpi=logpi)
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        score, path, units = viterbi(eng, corpus, es.b, ls, ln, utt_logpi=logpi)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = statistics.mean(ms) * 1e-3
    frames = n_utt * T_v
    bytes_alg = 4.0 * frames * 3 * L_v + 4.0 * frames
    ach = bytes_alg / t / 1e9
    if os.environ.get("PC_TRACE"):
        import ctypes as C
        from poccala_b200 import _native as nat
        buf = (C.c_longlong * 8)()
        nat.lib().pc_debug_read_vit.argtypes = [C.c_void_p]
        nat.lib().pc_debug_read_vit(buf)
        print("viterbi block 0: recurrence %d clk, traceback %d clk" % (buf[1] - buf[0], buf[2] - buf[1]), file=sys.stderr)
    # sanity: every path is a monotone walk
    p = path.view(n_utt, T_v)
    d = p[:, 1:] - p[:, :-1]
    ok = bool(((d == 0) | (d == 1)).all().item())
    return {"value": frames / t, "unit": "frames/s", "ms": t * 1e3,
            "workload": "cfg4: %d utt x %d frames x %d units (N=%d), 16-mix emissions, bit-exact fp64 recurrence"
                        % (n_utt, T_v, L_v, n_states),
            "roofline": {"bound": "hbm", "achieved": ach, "peak": float(pk["hbm_gbs"]), "unit": "GB/s",
                         "frac": ach / float(pk["hbm_gbs"])},
            "paths_monotone": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--check", action="store_true", help="N-rank vs 1-rank model agreement instead of timing")
    ap.add_argument("--collective", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: two NCCL all-reduces (the MAX one hidden under K3), the reduction over peer memory inside "
                         "the M-step kernels, or (default) whichever is faster in a trial during the warm-up - NCCL's "
                         "all-reduce of the 1.75 MB of statistics varies from box to box (profiles/README.md)")
    ap.add_argument("--cfg5-utt", type=int, default=int(os.environ.get("PC_BENCH_CFG5_UTT", "100000")),
                    help="utterances of the configs[4] leg over all ranks (0 = skip the leg)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.check:
        run_check(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
