"""K2 timing and block 0 phase clocks on the bench workload, both kernels, three corpus sizes:
    python profiles/exp_k2.py"""
import ctypes as C
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from poccala_b200 import synth, _native as nat
from poccala_b200.engine import Corpus, Engine, EStep, Model, _p, _stream
N_UNITS, N_INITIALS, MIX, T, L = 57, 22, 16, 300, 10
eng = Engine(0)
lib = nat.lib()
for n_utt in (148, 1000, 12500):
    truth, init0, labels, x = synth.torch_corpus(n_utt, T, L, N_UNITS, MIX, 2, eng.device, N_INITIALS)
    corpus = Corpus(eng, labels, np.full(n_utt, T, dtype=np.int32), N_UNITS)
    model = Model(eng, *init0, synth.default_transmat(N_UNITS))
    es = EStep(eng, corpus, model)
    es.load_frames(x)
    es.score()
    ls, ln = model.log_bands()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    for k2 in (1, 0):
        eng.set_option("k2_kernel", k2)
        for _ in range(3):
            es.forward_backward()
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            flush.zero_()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            ev[0].record()
            nat.call("pc_forward_backward", eng.h, corpus.c, _p(es.b), _p(ls), _p(ln), _p(es.lgam), _p(es.utt_logp),
                     _p(es.utt_iters), _p(es.pair_trans), _stream())
            ev[1].record()
            torch.cuda.synchronize()
            ms.append(ev[0].elapsed_time(ev[1]))
        buf = (C.c_longlong * 16)()
        fn = lib.pc_debug_read_fw if k2 else lib.pc_debug_read_fb
        fn.argtypes = [C.c_void_p]
        fn(buf)
        a = list(buf)
        gbs = 8.0 * n_utt * T * 3 * L / (min(ms) * 1e-3) / 1e9
        print("n_utt %5d k2_kernel %d: %.1f us (min %.1f), %.0f GB/s algorithmic; block 0: backward %d clk (%.0f / frame), pi %d, "
              "forward %d (%.0f / frame); sum logp %.6f" % (n_utt, k2, 1e3 * sum(ms) / len(ms), 1e3 * min(ms), gbs, a[1] - a[0],
              (a[1] - a[0]) / (T - 1), a[2] - a[1], a[3] - a[2], (a[3] - a[2]) / (T - 1), es.utt_logp.sum().item()), flush=True)
    eng.set_option("k2_kernel", 1)
    del es, model, corpus, x
