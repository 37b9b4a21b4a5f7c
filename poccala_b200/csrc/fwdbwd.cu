// K2: log-space forward-backward over the banded sentence HMM (one warp per utterance).
//
// Restates LHMM.baulm_welch on the HMM that AcousticModel.embedded assembles
// (AcousticModel.py:957-1014; LHMM.py:335-366 forward/backward, :426-471 ksai/gamma/pi,
// :412-422 likelihood, :526-544 pi iteration, :486-500 per-frame normalised log gamma), using the
// structure SURVEY A.2/A.3 verified against the executed reference:
//   * composite states 0 (entry, emission log 1), 1..3L (emitting), 3L+1 (exit, emission log 0);
//     the transition matrix is bidiagonal (self, next);
//   * beta never depends on pi, so one backward pass gives w = B[:,0] + beta_0 and the
//     pi-iteration (threshold 0.64, Q5) runs on w alone; then one forward pass with the final pi;
//   * ksai/gamma are NOT normalised by P(O) (Q6): we emit log expected counts relative to logP
//     and the utterance's logP in fp64; the cross-utterance log-sum-exp is done in fp64 (reduce.cu).
// Arithmetic: fp32 recurrences renormalised every frame (max subtracted; the cumulative offset is
// carried in fp64), so |alpha_hat|, |beta_hat| stay O(100) and the fp32 ulp stays ~1e-5 or below
// (SURVEY §7 hard part 2b).  The exit state never carries mass and is not materialised.
// Lane l owns states [l*SPL, (l+1)*SPL); the j-1 predecessor comes from a warp shuffle.
#include "common.cuh"

#define FB_WARPS 4

template <int SPL>
struct FbState {
    float ls[SPL], ln[SPL];
    int kind[SPL];    // 0 entry, 1 emitting, 2 inactive
    int64_t row[SPL];  // float offset of the emitting row inside the utterance block
};

template <int SPL>
__device__ __forceinline__ float fb_emis(const FbState<SPL> &s, const float *__restrict__ bu, int q,
                                         int t) {
    return s.kind[q] == 1 ? bu[s.row[q] + t] : (s.kind[q] == 0 ? 0.f : PC_NEG_INF);
}

// online log-sum-exp accumulator: sum * exp(mx) += exp(x)
__device__ __forceinline__ void acc_lse(float &mx, float &sum, float x) {
    if (x == PC_NEG_INF) return;
    if (x > mx) {
        sum *= __expf(mx - x);
        mx = x;
    }
    sum += __expf(x - mx);
}

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int SPL>
__global__ void __launch_bounds__(FB_WARPS * 32)
fwdbwd_kernel(CorpusView v, const float *__restrict__ b, const double *__restrict__ log_self,
              const double *__restrict__ log_next, float *__restrict__ lgam,
              float *__restrict__ scratch0, double *__restrict__ utt_logp,
              int32_t *__restrict__ utt_iters, float *__restrict__ pair_trans) {
    const int lane = threadIdx.x & 31;
    const int wid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (wid >= v.n_utt) return;
    const int u = v.fb_order[wid];
    const int64_t f0 = v.frame_off[u];
    const int T = (int)(v.frame_off[u + 1] - f0);
    const int64_t p0 = v.pair_off[u];
    const int L = (int)(v.pair_off[u + 1] - p0);
    const int NE = PC_EMIT * L;  // states 0..NE are materialised; N = NE + 2
    const int tp = pc_tpad(T);
    const float *bu = b + v.emis_off[u];
    float *gu = lgam + v.emis_off[u];
    float *s0 = scratch0 + f0;

    FbState<SPL> st;
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const int s = lane * SPL + q;
        st.row[q] = 0;
        if (s == 0) {
            const int unit = v.labels[p0];
            st.kind[q] = 0;
            st.ls[q] = (float)log_self[unit * PC_STATES];
            st.ln[q] = (float)log_next[unit * PC_STATES];
        } else if (s <= NE) {
            const int p = (s - 1) / PC_EMIT, r = (s - 1) - p * PC_EMIT;
            const int unit = v.labels[p0 + p];
            st.kind[q] = 1;
            st.row[q] = (int64_t)(s - 1) * tp;
            st.ls[q] = (float)log_self[unit * PC_STATES + 1 + r];
            // the last emitting state's successor is the exit state (emission log 0): no mass
            st.ln[q] = (s == NE) ? PC_NEG_INF : (float)log_next[unit * PC_STATES + 1 + r];
        } else {
            st.kind[q] = 2;
            st.ls[q] = PC_NEG_INF;
            st.ln[q] = PC_NEG_INF;
        }
    }

    // ---------------------------------------------------------------- backward (LHMM.py:353-366)
    float bh[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) bh[q] = (st.kind[q] == 2) ? PC_NEG_INF : 0.f;
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        if (st.kind[q] == 1) gu[st.row[q] + T - 1] = bh[q];
        if (st.kind[q] == 0) s0[T - 1] = bh[q];
    }
    for (int t = T - 2; t >= 0; --t) {
        float nb[SPL], raw[SPL];
#pragma unroll
        for (int q = 0; q < SPL; ++q) nb[q] = bh[q] + fb_emis(st, bu, q, t + 1);
        float up = __shfl_down_sync(0xffffffffu, nb[0], 1);
        if (lane == 31) up = PC_NEG_INF;
        float mx = PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            const float nxt = (q + 1 < SPL) ? nb[(q + 1) % SPL] : up;
            raw[q] = logadd_f(st.ls[q] + nb[q], st.ln[q] + nxt);
            mx = fmaxf(mx, raw[q]);
        }
        mx = warp_max(mx);
        if (mx == PC_NEG_INF) mx = 0.f;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            bh[q] = raw[q] - mx;
            if (st.kind[q] == 1) gu[st.row[q] + t] = bh[q];
            if (st.kind[q] == 0) s0[t] = bh[q];
        }
    }

    // ------------------------------------------------- pi iteration (LHMM.py:447-452,526-544; A.3)
    double w[SPL], lp[SPL], lp_used[SPL];
    const double log_uniform = log(1.0 / (double)(NE + 2));
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        w[q] = (double)bh[q] + (double)fb_emis(st, bu, q, 0);
        lp[q] = (st.kind[q] == 2) ? (double)PC_NEG_INF : log_uniform;
        lp_used[q] = lp[q];
    }
    int iters = 0;
    double q_prev = (double)PC_NEG_INF;
    for (int guard = 0; guard < 100000; ++guard) {
        double m = (double)PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) m = fmax(m, lp[q] + w[q]);
        m = warp_max_d(m);
        double qn = m;
        if (m != (double)PC_NEG_INF) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < SPL; ++q) s += exp(lp[q] + w[q] - m);
            qn = m + log(warp_sum_d(s));
        }
        ++iters;
#pragma unroll
        for (int q = 0; q < SPL; ++q) lp_used[q] = lp[q];
        if (!((qn - q_prev) > 0.64)) break;
        q_prev = qn;
#pragma unroll
        for (int q = 0; q < SPL; ++q) lp[q] = log(exp(lp[q] + w[q] - qn));  // linear-space pi (A.3)
    }

    // ---------------------------------------------------------------- forward (LHMM.py:335-351)
    // transition counters are kept as (max, scaled sum) pairs so that counts far below the fp32
    // range (states the alignment never visits) keep a finite log like the fp64 reference
    float ah[SPL], cs[SPL], cn[SPL], cg[SPL], ms[SPL], mn[SPL], mg[SPL], bcur[SPL];
    double Ca;
    {
        float mx = PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            ah[q] = (float)lp_used[q] + fb_emis(st, bu, q, 0);
            mx = fmaxf(mx, ah[q]);
            cs[q] = cn[q] = cg[q] = 0.f;
            ms[q] = mn[q] = mg[q] = PC_NEG_INF;
        }
        mx = warp_max(mx);
        if (mx == PC_NEG_INF) mx = 0.f;
#pragma unroll
        for (int q = 0; q < SPL; ++q) ah[q] -= mx;
        Ca = (double)mx;
    }
#pragma unroll
    for (int q = 0; q < SPL; ++q)
        bcur[q] = st.kind[q] == 1 ? gu[st.row[q]] : (st.kind[q] == 0 ? s0[0] : PC_NEG_INF);
    double logp = 0.0;
    for (int t = 0; t < T; ++t) {
        // per-frame normalised log gamma (LHMM.py:486-500)
        float vv[SPL], lg[SPL];
        float m = PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            vv[q] = ah[q] + bcur[q];
            m = fmaxf(m, vv[q]);
        }
        m = warp_max(m);
        float Z = PC_NEG_INF;
        if (m != PC_NEG_INF) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < SPL; ++q) s += __expf(vv[q] - m);
            Z = m + __logf(warp_sum(s));
        }
#pragma unroll
        for (int q = 0; q < SPL; ++q) lg[q] = (m == PC_NEG_INF) ? PC_NEG_INF : vv[q] - Z;
        if (t == T - 1) logp = Ca + (double)Z;

        float bnxt[SPL];
        if (t < T - 1) {
            // expected transition counts over t < T-1 (LHMM.py:431-445), split of gamma_t(i) into
            // self / next by the backward recurrence (the backward normaliser cancels)
            float nb[SPL];
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                bnxt[q] = st.kind[q] == 1 ? gu[st.row[q] + t + 1]
                                          : (st.kind[q] == 0 ? s0[t + 1] : PC_NEG_INF);
                nb[q] = bnxt[q] + fb_emis(st, bu, q, t + 1);
            }
            float up = __shfl_down_sync(0xffffffffu, nb[0], 1);
            if (lane == 31) up = PC_NEG_INF;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                const float nxt = (q + 1 < SPL) ? nb[(q + 1) % SPL] : up;
                const float sv = st.ls[q] + nb[q], mv = st.ln[q] + nxt;
                // log of the self / next shares of gamma_t(i): -softplus(mv-sv), (mv-sv)-softplus
                float lfs, lfn;
                if (sv == PC_NEG_INF) {
                    lfs = PC_NEG_INF;
                    lfn = (mv == PC_NEG_INF) ? PC_NEG_INF : 0.f;
                } else {
                    const float d = mv - sv;  // -inf when the successor carries no mass
                    const float sp = fmaxf(d, 0.f) + __logf(1.f + __expf(-fabsf(d)));
                    lfs = -sp;
                    lfn = (d == PC_NEG_INF) ? PC_NEG_INF : d - sp;
                }
                acc_lse(mg[q], cg[q], lg[q]);
                acc_lse(ms[q], cs[q], lg[q] + lfs);
                acc_lse(mn[q], cn[q], lg[q] + lfn);
            }
        }
        // overwrite beta_hat_t with log gamma_t (beta_hat_{t+1} is already in registers)
#pragma unroll
        for (int q = 0; q < SPL; ++q)
            if (st.kind[q] == 1) gu[st.row[q] + t] = lg[q];
        if (t < T - 1) {
            float left = __shfl_up_sync(0xffffffffu, ah[SPL - 1] + st.ln[SPL - 1], 1);
            if (lane == 0) left = PC_NEG_INF;
            float raw[SPL];
            float mx = PC_NEG_INF;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                const float from_left = (q > 0) ? ah[(q + SPL - 1) % SPL] + st.ln[(q + SPL - 1) % SPL] : left;
                raw[q] = logadd_f(ah[q] + st.ls[q], from_left) + fb_emis(st, bu, q, t + 1);
                mx = fmaxf(mx, raw[q]);
            }
            mx = warp_max(mx);
            if (mx == PC_NEG_INF) mx = 0.f;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                ah[q] = raw[q] - mx;
                bcur[q] = bnxt[q];
            }
            Ca += (double)mx;
        }
    }
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        if (st.kind[q] == 1) {
            const int s = lane * SPL + q;
            float *o = pair_trans + (size_t)(p0 + (s - 1) / PC_EMIT) * PC_TRANS_SLOTS +
                       ((s - 1) % PC_EMIT) * 3;
            o[0] = (ms[q] == PC_NEG_INF) ? PC_NEG_INF : ms[q] + __logf(cs[q]);
            o[1] = (mn[q] == PC_NEG_INF) ? PC_NEG_INF : mn[q] + __logf(cn[q]);
            o[2] = (mg[q] == PC_NEG_INF) ? PC_NEG_INF : mg[q] + __logf(cg[q]);
        }
    }
    if (lane == 0) {
        utt_logp[u] = logp;
        utt_iters[u] = iters;
    }
}

template <int SPL>
static int launch_fb(pc_handle h, const CorpusView &v, const float *b, const double *log_self,
                     const double *log_next, float *lgam, float *scratch0, double *utt_logp,
                     int32_t *utt_iters, float *pair_trans, cudaStream_t st) {
    int blocks = (v.n_utt + FB_WARPS - 1) / FB_WARPS;
    fwdbwd_kernel<SPL><<<blocks, FB_WARPS * 32, 0, st>>>(v, b, log_self, log_next, lgam, scratch0,
                                                        utt_logp, utt_iters, pair_trans);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int launch_forward_backward(pc_handle h, const CorpusView &v, const float *b,
                            const double *log_self, const double *log_next, float *lgam,
                            float *scratch0, double *utt_logp, int32_t *utt_iters,
                            float *pair_trans, cudaStream_t st) {
    if (v.n_utt == 0) return PC_OK;
    const int states = PC_EMIT * v.max_labels + 1;
    if (states <= 32) return launch_fb<1>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 64) return launch_fb<2>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 128) return launch_fb<4>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 256) return launch_fb<8>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    pc_set_error("pc_forward_backward: %d labels per utterance exceeds the limit of 85", v.max_labels);
    return PC_ERR_UNSUPPORTED;
}
