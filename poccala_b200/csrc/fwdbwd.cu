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
//
// Arithmetic: fp32 log2-domain recurrences.  Every frame's emissions are shifted by g_t = max_j b_t(j)
// (computed ahead of the recurrence from the prefetched row, so it is off the dependency chain) and
// every FB_RENORM frames the state vector is renormalised exactly (one warp max on the chain); the
// shifts are accumulated in fp64, so |alpha_hat|, |beta_hat| stay O(100) and the fp32 ulp stays
// ~1e-5 or below (SURVEY §7 hard part 2b).  The per-frame normaliser of gamma needs no reduction:
// sum_j alpha_t(j) beta_t(j) = P(O) for every t, hence Z_t = Z_{t-1} - shift_a(t) + shift_b(t-1)
// (fp64 scalar).  The only cross-lane traffic on the dependency chain is ONE shuffle per frame
// (the j-1 / j+1 neighbour); lane l owns states [l*SPL, (l+1)*SPL).
// Emissions are time-major (b[t][s]), so a warp reads / writes one coalesced row per frame.
#include "common.cuh"

#define FB_RENORM 8

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// log2(2^a + 2^b) with -inf handling
__device__ __forceinline__ float logadd2(float a, float b) {
    const float m = fmaxf(a, b);
    float d = -fabsf(a - b);  // NaN only when a == b == -inf
    d = (m == PC_NEG_INF) ? 0.f : d;
    return m + lg2f(1.f + ex2f(d));
}
// sum * 2^mx += 2^x (x finite), rescaling only when x overtakes the reference by a wide margin
__device__ __forceinline__ void acc_lse2(float &mx, float &sum, float x) {
    if (x > mx + 24.f) {
        sum *= ex2f(mx - x);  // first use: mx = -inf, sum = 0
        mx = x;
    }
    sum += ex2f(x - mx);
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// FB_CH: frames per prefetch chunk (loads are issued one chunk ahead); FB_WARPS: warps per block
template <int SPL, int FB_CH, int FB_WARPS>
__global__ void __launch_bounds__(FB_WARPS * 32)
fwdbwd_kernel(CorpusView v, const float *__restrict__ b, const double *__restrict__ log_self,
              const double *__restrict__ log_next, float *__restrict__ lgam, float4 *scratch,
              double *__restrict__ utt_logp, int32_t *__restrict__ utt_iters,
              float *__restrict__ pair_trans) {
    const int lane = threadIdx.x & 31;
    const int wid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (wid >= v.n_utt) return;
    const int u = v.fb_order[wid];
    const int64_t f0 = v.frame_off[u];
    const int T = (int)(v.frame_off[u + 1] - f0);
    const int64_t p0 = v.pair_off[u];
    const int L = (int)(v.pair_off[u + 1] - p0);
    const int NE = PC_EMIT * L;  // states 0..NE are materialised; N = NE + 2
    const int sp = pc_spad(L);
    const float *bu = b + v.emis_off[u];
    float *gu = lgam + v.emis_off[u];
    // per frame: x = g_t (natural log), y = shift_b(t) (log2), z = beta_hat_t of the entry state (log2)
    float4 *rec = scratch + f0;

    float ls[SPL], ln[SPL];
    int kind[SPL], col[SPL];  // kind: 0 entry, 1 emitting, 2 inactive
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const int s = lane * SPL + q;
        col[q] = 0;
        if (s == 0) {
            const int unit = v.labels[p0];
            kind[q] = 0;
            ls[q] = (float)log_self[unit * PC_STATES] * kLog2e;
            ln[q] = (float)log_next[unit * PC_STATES] * kLog2e;
        } else if (s <= NE) {
            const int p = (s - 1) / PC_EMIT, r = (s - 1) - p * PC_EMIT;
            const int unit = v.labels[p0 + p];
            kind[q] = 1;
            col[q] = s - 1;
            ls[q] = (float)log_self[unit * PC_STATES + 1 + r] * kLog2e;
            // the last emitting state's successor is the exit state (emission log 0): no mass
            ln[q] = (s == NE) ? PC_NEG_INF : (float)log_next[unit * PC_STATES + 1 + r] * kLog2e;
        } else {
            kind[q] = 2;
            ls[q] = PC_NEG_INF;
            ln[q] = PC_NEG_INF;
        }
    }
    // raw emission row (natural log) of frame t for this lane's states; -inf for non-emitting
    auto load_row = [&](int t, float (&e)[SPL]) {
#pragma unroll
        for (int q = 0; q < SPL; ++q) e[q] = (kind[q] == 1) ? __ldg(bu + (size_t)t * sp + col[q]) : PC_NEG_INF;
    };
    auto frame_max = [&](const float (&e)[SPL]) {
        float m = PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) m = fmaxf(m, e[q]);
        m = warp_max(m);
        return (m == PC_NEG_INF) ? 0.f : m;
    };
    // shifted log2 emissions: (b - g)*log2e; the entry state emits log 1 = 0
    auto shift_row = [&](const float (&e)[SPL], float g, float (&es)[SPL]) {
#pragma unroll
        for (int q = 0; q < SPL; ++q)
            es[q] = (kind[q] == 1) ? (e[q] - g) * kLog2e : (kind[q] == 0 ? -g * kLog2e : PC_NEG_INF);
    };

    // ================================================================ backward (LHMM.py:353-366)
    float bh[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        bh[q] = (kind[q] == 2) ? PC_NEG_INF : 0.f;
        if (kind[q] == 1) gu[(size_t)(T - 1) * sp + col[q]] = 0.f;
    }
    if (lane == 0) {
        rec[T - 1].y = 0.f;
        rec[T - 1].z = 0.f;
    }
    {
        // step tau (= T-1 .. 1) consumes emission row tau and produces beta_hat_{tau-1}
        float e_nxt[FB_CH][SPL];
        int tau_hi = T - 1;
#pragma unroll
        for (int k = 0; k < FB_CH; ++k) load_row(max(tau_hi - k, 0), e_nxt[k]);
        int since = 0;
        while (tau_hi >= 1) {
            float es[FB_CH][SPL], g[FB_CH];
#pragma unroll
            for (int k = 0; k < FB_CH; ++k) {  // off the dependency chain
                g[k] = frame_max(e_nxt[k]);
                shift_row(e_nxt[k], g[k], es[k]);
            }
#pragma unroll
            for (int k = 0; k < FB_CH; ++k) load_row(max(tau_hi - FB_CH - k, 0), e_nxt[k]);  // prefetch
#pragma unroll
            for (int k = 0; k < FB_CH; ++k) {
                const int tau = tau_hi - k;
                if (tau >= 1) {
                    float nb[SPL], raw[SPL];
#pragma unroll
                    for (int q = 0; q < SPL; ++q) nb[q] = bh[q] + es[k][q];
                    float up = __shfl_down_sync(0xffffffffu, nb[0], 1);
                    if (lane == 31) up = PC_NEG_INF;
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        const float nxt = (q + 1 < SPL) ? nb[(q + 1) % SPL] : up;
                        raw[q] = logadd2(ls[q] + nb[q], ln[q] + nxt);
                    }
                    float r = 0.f;
                    if (++since == FB_RENORM) {
                        since = 0;
                        float m = PC_NEG_INF;
#pragma unroll
                        for (int q = 0; q < SPL; ++q) m = fmaxf(m, raw[q]);
                        m = warp_max(m);
                        r = (m == PC_NEG_INF) ? 0.f : m;
                    }
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        bh[q] = raw[q] - r;
                        if (kind[q] == 1) gu[(size_t)(tau - 1) * sp + col[q]] = bh[q];
                    }
                    if (lane == 0) {
                        rec[tau].x = g[k];
                        rec[tau - 1].y = g[k] * kLog2e + r;  // shift_b(tau-1)
                        rec[tau - 1].z = bh[0];
                    }
                }
            }
            tau_hi -= FB_CH;
        }
    }
    float e0[SPL];
    load_row(0, e0);
    const float g0 = frame_max(e0);
    if (lane == 0) rec[0].x = g0;
    __syncwarp();  // rec[] was written by lane 0 and is read by every lane below

    // ================================================= pi iteration (LHMM.py:447-452,526-544; A.3)
    // natural-log fp64 on w = B[:,0] + beta_0 (relative to a common offset, which cancels)
    double w[SPL], lp[SPL], lp_used[SPL];
    const double log_uniform = log(1.0 / (double)(NE + 2));
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const double em = (kind[q] == 1) ? (double)e0[q] : (kind[q] == 0 ? 0.0 : (double)PC_NEG_INF);
        w[q] = (double)bh[q] * (double)kLn2 + em;
        lp[q] = (kind[q] == 2) ? (double)PC_NEG_INF : log_uniform;
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

    // ================================================================ forward (LHMM.py:335-351)
    // log2 units; cumulative shift Ca and the gamma normaliser Z in fp64
    float ah[SPL], ms[SPL], cs[SPL], mn[SPL], cn[SPL], lg_prev[SPL];
    double Ca, Z;
    float sb_prev;
    {
        float es0[SPL], bcur[SPL];
        shift_row(e0, g0, es0);
        float m = PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            ah[q] = (kind[q] == 2) ? PC_NEG_INF : (float)(lp_used[q] * (double)kLog2e) + es0[q];
            m = fmaxf(m, ah[q]);
            ms[q] = mn[q] = PC_NEG_INF;
            cs[q] = cn[q] = 0.f;
        }
        m = warp_max(m);
        if (m == PC_NEG_INF) m = 0.f;
#pragma unroll
        for (int q = 0; q < SPL; ++q) ah[q] -= m;
        Ca = (double)g0 * (double)kLog2e + (double)m;
        const float4 r0 = __ldcg(rec);
        sb_prev = r0.y;
#pragma unroll
        for (int q = 0; q < SPL; ++q)
            bcur[q] = kind[q] == 1 ? gu[col[q]] : (kind[q] == 0 ? r0.z : PC_NEG_INF);
        // Z_0 = log2 sum_j 2^(alpha_hat + beta_hat): the only normaliser that needs a reduction
        float mm = PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) mm = fmaxf(mm, ah[q] + bcur[q]);
        mm = warp_max(mm);
        float ssum = 0.f;
#pragma unroll
        for (int q = 0; q < SPL; ++q) ssum += (mm == PC_NEG_INF) ? 0.f : ex2f(ah[q] + bcur[q] - mm);
        ssum = warp_sum(ssum);
        Z = (mm == PC_NEG_INF) ? (double)PC_NEG_INF : (double)mm + (double)lg2f(ssum);
        const float zf = (float)Z;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            lg_prev[q] = ah[q] + bcur[q] - zf;
            if (kind[q] == 1) gu[col[q]] = lg_prev[q] * kLn2;  // frame 0 gamma
        }
    }
    {
        auto load_beta = [&](int t, float (&x)[SPL]) {
#pragma unroll
            for (int q = 0; q < SPL; ++q) x[q] = (kind[q] == 1) ? gu[(size_t)t * sp + col[q]] : PC_NEG_INF;
        };
        float e_nxt[FB_CH][SPL], bn_nxt[FB_CH][SPL];
        float4 r_nxt[FB_CH];
        int tau_lo = 1;
#pragma unroll
        for (int k = 0; k < FB_CH; ++k) {
            const int t = min(tau_lo + k, T - 1);
            load_row(t, e_nxt[k]);
            load_beta(t, bn_nxt[k]);
            r_nxt[k] = __ldcg(rec + t);
        }
        int since = 0;
        while (tau_lo <= T - 1) {
            float e_cur[FB_CH][SPL], bn_cur[FB_CH][SPL];
            float4 r_cur[FB_CH];
#pragma unroll
            for (int k = 0; k < FB_CH; ++k) {
                r_cur[k] = r_nxt[k];
#pragma unroll
                for (int q = 0; q < SPL; ++q) { e_cur[k][q] = e_nxt[k][q]; bn_cur[k][q] = bn_nxt[k][q]; }
            }
            // prefetch the next chunk: its frames still hold beta_hat (gamma is only stored for
            // frames of the current chunk, after these loads in program order)
#pragma unroll
            for (int k = 0; k < FB_CH; ++k) {
                const int t = min(tau_lo + FB_CH + k, T - 1);
                load_row(t, e_nxt[k]);
                load_beta(t, bn_nxt[k]);
                r_nxt[k] = __ldcg(rec + t);
            }
#pragma unroll
            for (int k = 0; k < FB_CH; ++k) {
                const int tau = tau_lo + k;
                if (tau <= T - 1) {
                    const float g = r_cur[k].x;
                    float es[SPL], bnx[SPL];
                    shift_row(e_cur[k], g, es);
#pragma unroll
                    for (int q = 0; q < SPL; ++q) bnx[q] = (kind[q] == 0) ? r_cur[k].z : bn_cur[k][q];
                    // ---- expected transition counts for t = tau-1 (LHMM.py:431-445): gamma_{tau-1}(i)
                    // split into self / next by the backward recurrence (its shift cancels)
                    {
                        float nb[SPL];
#pragma unroll
                        for (int q = 0; q < SPL; ++q) nb[q] = bnx[q] + es[q];
                        float up = __shfl_down_sync(0xffffffffu, nb[0], 1);
                        if (lane == 31) up = PC_NEG_INF;
#pragma unroll
                        for (int q = 0; q < SPL; ++q) {
                            const float nxt = (q + 1 < SPL) ? nb[(q + 1) % SPL] : up;
                            const float sv = ls[q] + nb[q], mv = ln[q] + nxt;
                            if (lg_prev[q] != PC_NEG_INF) {
                                if (sv == PC_NEG_INF) {
                                    if (mv != PC_NEG_INF) acc_lse2(mn[q], cn[q], lg_prev[q]);
                                } else {
                                    const float d = mv - sv;  // -inf when the successor carries no mass
                                    const float spl = fmaxf(d, 0.f) + lg2f(1.f + ex2f(-fabsf(d)));
                                    acc_lse2(ms[q], cs[q], lg_prev[q] - spl);
                                    if (d != PC_NEG_INF) acc_lse2(mn[q], cn[q], lg_prev[q] + d - spl);
                                }
                            }
                        }
                    }
                    // ---- alpha recurrence
                    float left = __shfl_up_sync(0xffffffffu, ah[SPL - 1] + ln[SPL - 1], 1);
                    if (lane == 0) left = PC_NEG_INF;
                    float raw[SPL];
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        const float from_left = (q > 0) ? ah[(q + SPL - 1) % SPL] + ln[(q + SPL - 1) % SPL] : left;
                        raw[q] = logadd2(ah[q] + ls[q], from_left) + es[q];
                    }
                    float r = 0.f;
                    if (++since == FB_RENORM) {
                        since = 0;
                        float m = PC_NEG_INF;
#pragma unroll
                        for (int q = 0; q < SPL; ++q) m = fmaxf(m, raw[q]);
                        m = warp_max(m);
                        r = (m == PC_NEG_INF) ? 0.f : m;
                    }
#pragma unroll
                    for (int q = 0; q < SPL; ++q) ah[q] = raw[q] - r;
                    const double sa = (double)g * (double)kLog2e + (double)r;
                    Ca += sa;
                    Z += (double)sb_prev - sa;  // Z_tau = Z_{tau-1} - shift_a(tau) + shift_b(tau-1)
                    sb_prev = r_cur[k].y;
                    if (since == 0) {
                        // same cadence as the renormalisation: recompute Z exactly so that rounding
                        // drift common to all states cannot accumulate in the gamma normaliser
                        float mm = PC_NEG_INF;
#pragma unroll
                        for (int q = 0; q < SPL; ++q) mm = fmaxf(mm, ah[q] + bnx[q]);
                        mm = warp_max(mm);
                        float ssum = 0.f;
#pragma unroll
                        for (int q = 0; q < SPL; ++q) ssum += (mm == PC_NEG_INF) ? 0.f : ex2f(ah[q] + bnx[q] - mm);
                        ssum = warp_sum(ssum);
                        if (mm != PC_NEG_INF) Z = (double)mm + (double)lg2f(ssum);
                    }
                    // ---- per-frame normalised log gamma (LHMM.py:486-500)
                    const float zf = (float)Z;
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        lg_prev[q] = ah[q] + bnx[q] - zf;
                        if (kind[q] == 1) gu[(size_t)tau * sp + col[q]] = lg_prev[q] * kLn2;
                    }
                }
            }
            tau_lo += FB_CH;
        }
    }
    // log P(O) = Ca_{T-1} + Z_{T-1}   (beta_hat_{T-1} = 0, no backward shift after the last frame)
    const double logp = (Ca + Z) * 0.6931471805599453;
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        if (kind[q] == 1) {
            const int s = lane * SPL + q;
            float *o = pair_trans + (size_t)(p0 + (s - 1) / PC_EMIT) * PC_TRANS_SLOTS +
                       ((s - 1) % PC_EMIT) * 3;
            const float ks = (ms[q] == PC_NEG_INF) ? PC_NEG_INF : ms[q] + lg2f(cs[q]);
            const float kn = (mn[q] == PC_NEG_INF) ? PC_NEG_INF : mn[q] + lg2f(cn[q]);
            o[0] = ks * kLn2;
            o[1] = kn * kLn2;
            o[2] = logadd2(ks, kn) * kLn2;  // occupancy over t < T-1 = self + next
        }
    }
    if (lane == 0) {
        utt_logp[u] = logp;
        utt_iters[u] = iters;
    }
}

template <int SPL, int CH, int WARPS>
int launch_fb_cfg(pc_handle h, const CorpusView &v, const float *b, const double *log_self,
                  const double *log_next, float *lgam, float *scratch0, double *utt_logp,
                  int32_t *utt_iters, float *pair_trans, cudaStream_t st) {
    int blocks = (v.n_utt + WARPS - 1) / WARPS;
    fwdbwd_kernel<SPL, CH, WARPS><<<blocks, WARPS * 32, 0, st>>>(v, b, log_self, log_next, lgam,
                                                                reinterpret_cast<float4 *>(scratch0), utt_logp,
                                                                utt_iters, pair_trans);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

template <int SPL>
int launch_fb(pc_handle h, const CorpusView &v, const float *b, const double *log_self,
              const double *log_next, float *lgam, float *scratch0, double *utt_logp,
              int32_t *utt_iters, float *pair_trans, cudaStream_t st) {
    if constexpr (SPL == 1) {
        switch (h->fb_cfg) {  // option "fb_cfg": tuning experiments (profiles/exp_k2.py)
            case 1: return launch_fb_cfg<1, 4, 4>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
            case 2: return launch_fb_cfg<1, 16, 4>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
            case 3: return launch_fb_cfg<1, 8, 1>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
            case 4: return launch_fb_cfg<1, 16, 1>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
            case 5: return launch_fb_cfg<1, 4, 1>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
            default: return launch_fb_cfg<1, 8, 4>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
        }
    }
    return launch_fb_cfg<SPL, 4, 4>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
}

}  // namespace

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
