// K2, one warp per utterance: log-space forward-backward over the banded sentence HMM with the
// posteriors and the expected transition counts formed in the forward pass itself.
//
// Same mathematics as fwdbwd.cu (LHMM.baulm_welch on the HMM of AcousticModel.embedded:
// LHMM.py:335-366 forward / backward, :426-471 ksai / gamma / pi, :412-422 likelihood, :526-544 pi
// iteration, :486-500 log gamma), restated around two identities that remove the
// per-frame reductions of the three-warp kernel (139 instructions per frame there, ~55 here):
//   * sum_j alpha_t(j) beta_t(j) = P(O) for EVERY t, so log gamma_t(j) = log alpha_t(j) +
//     log beta_t(j) - log P(O): the normaliser is the utterance likelihood the pi iteration
//     produced after the backward pass, not a log-sum-exp over the states of each frame
//     (LHMM.py:486-500 computes that sum numerically; it equals log P(O) to 1e-12 in fp64);
//   * with nb_t(j) = log b_t(j) + log beta_t(j) (one row per frame, written by the backward pass)
//         log gamma_t(j)    = [stay_t(j) (+) move_t(j)] + nb_t(j) - log P
//         log xi_t-1(j, j)  =  stay_t(j) + nb_t(j) - log P,   stay_t(j) = log alpha_t-1(j)   + log a_jj
//         log xi_t-1(j-1,j) =  move_t(j) + nb_t(j) - log P,   move_t(j) = log alpha_t-1(j-1) + log a_j-1,j
//     i.e. the forward recurrence's own two terms, one add each (LHMM.py:431-445).
// The recurrences stay in the log domain: the reference's transition accumulators are weighted by
// the utterance likelihood (Q6), so expected counts of e^-300 can decide a unit's transition row
// and a scaled linear-domain recurrence would flush them.
//
// Arithmetic: fp32 log2 domain.  Every frame's emissions are shifted by g_t = ceil(max_j log2 b_t(j)) and
// every FW_CH frames the state vector is renormalised by the ceiling of its maximum (one CREDUX on the
// chain).  All shifts are INTEGERS, so their running sums are exact in fp32 (|sum| < 2^24) and the scale
// constant of frame t, (forward shifts to t-1) + (backward shifts from t) - log2 P, is an exact small
// integer minus the fractional part of log2 P: no fp64 inside the loops.  Once per chunk the forward pass
// measures the drift of its normaliser (log2 sum_j gamma_t(j), zero in exact arithmetic, ~1e-5 after 300
// frames of MUFU approximations) and folds it into the constant.  Lane l owns states [l*SPL, (l+1)*SPL);
// emissions are time-major, a warp reads / writes one coalesced row per frame.  Scratch: nb rows in the
// corpus' beta scratch (same layout as b), one float4 per frame {backward shift, g_t, nb of the entry
// state}.  K3's activity flags are set from the log gamma rows while they are at hand.
#include <type_traits>

#include "common.cuh"

#define FW_CH 8   // frames per chunk: prefetch depth and renormalisation period (4: measured slower, 593 vs 322 clk per forward frame)
#define FW_WPB 1  // one warp per block: every index below is provably warp-uniform (no divergence checks around the shuffles)

__device__ long long g_fw_dbg[16];

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
__device__ __forceinline__ float redux_max(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}
// log2(2^a + 2^b) with -inf handling
__device__ __forceinline__ float logadd2(float a, float b) {
    const float m = fmaxf(a, b);
    float d = fminf(a, b) - m;  // NaN only when a == b == -inf
    d = (m == PC_NEG_INF) ? 0.f : d;
    return m + lg2f(1.f + ex2f(d));
}
// asynchronous global -> shared copies (LDGSTS): the prefetch of the next chunk's rows neither holds
// registers nor can be scheduled late by the compiler (register prefetches were sunk to their first use:
// a quarter of the kernel's time went into waiting for them, profiles/README.md)
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int SPL>
__global__ void __launch_bounds__(FW_WPB * 32)
fwdbwd_warp_kernel(CorpusView v, const float *__restrict__ b, const double *__restrict__ log_self,
                   const double *__restrict__ log_next, float *__restrict__ lgam, float4 *__restrict__ fscr,
                   double *__restrict__ utt_logp, int32_t *__restrict__ utt_iters,
                   float *__restrict__ pair_trans) {
    constexpr int CH = FW_CH;
    // two-stage ring of prefetched rows: every lane copies (and later reads) its own columns
    // ring of NST chunks of rows: the rows of chunk i + NST - 1 are requested while chunk i is worked on, so a row has
    // (NST - 1) chunks of chain time (~800 clk each) to arrive - with two stages every chunk waited ~1 000 clk for
    // HBM / L2 (measured 244 clk per frame against ~100 of dependent latency)
    constexpr int NST = SPL <= 2 ? 4 : (SPL <= 4 ? 3 : 2);
    __shared__ float sm_b[NST][CH][32 * SPL], sm_n[NST][CH][32 * SPL];
    __shared__ __align__(16) float4 sm_f[NST][CH];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x;
    const int u = v.fb_order[idx];
    const int64_t f0 = v.frame_off[u];
    const int T = (int)(v.frame_off[u + 1] - f0);
    const int64_t p0 = v.pair_off[u];
    const int L = (int)(v.pair_off[u + 1] - p0);
    const int NE = PC_EMIT * L;  // states 0..NE are materialised; N = NE + 2
    const int sp = pc_spad(L);
    const float *bu = b + v.emis_off[u];
    float *gu = lgam + v.emis_off[u];
    float *nbu = v.scratch1 + v.emis_off[u];  // nb rows of the emitting states
    float4 *fs = fscr + f0;                   // per frame: {shift lo, shift hi, g*log2e, nb of the entry state}
    const bool trace = (blockIdx.x == 0 && threadIdx.x == 0);
    if (trace) g_fw_dbg[0] = clock64();

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
    // per-lane views of the [T][SP] blocks: element (t, this state) of b / nb / lgam is base[q][t * sp]
    bool emit[SPL];
    const float *bq[SPL];
    float *nq[SPL], *gq[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        emit[q] = kind[q] == 1;
        bq[q] = bu + col[q];
        nq[q] = nbu + col[q];
        gq[q] = gu + col[q];
    }
    const unsigned spu = (unsigned)sp;
    auto load_row = [&](int t, float (&e)[SPL]) {
        const unsigned off = (unsigned)t * spu;
#pragma unroll
        for (int q = 0; q < SPL; ++q) e[q] = emit[q] ? __ldg(bq[q] + off) : PC_NEG_INF;
    };
    auto frame_max = [&](const float (&e)[SPL]) {
        float m = e[0];
#pragma unroll
        for (int q = 1; q < SPL; ++q) m = fmaxf(m, e[q]);
        m = redux_max(m);
        return (m == PC_NEG_INF) ? 0.f : m;
    };
    // shifted log2 emissions: b*log2e - g2; the entry state emits log 1 = 0
    auto shift_row = [&](const float (&e)[SPL], float g2, float (&es)[SPL]) {
#pragma unroll
        for (int q = 0; q < SPL; ++q)
            es[q] = (kind[q] == 1) ? fmaf(e[q], kLog2e, -g2) : (kind[q] == 0 ? -g2 : PC_NEG_INF);
    };

    // ------------------------------------------------------------------ backward (LHMM.py:353-366)
    float bh[SPL];
    float Cb = 0.f;  // log2 beta_t = bh + Cb; an integer (exact in fp32)
#pragma unroll
    for (int q = 0; q < SPL; ++q) bh[q] = (kind[q] != 2) ? 0.f : PC_NEG_INF;
    {
        // step tau (= T-1 .. 1) consumes emission row tau, stores nb_tau and produces beta_hat_{tau-1}
        int tau_hi = T - 1;
        float *np_[SPL];  // -> (frame tau, this state) of the nb rows, walking down
        float4 *fp_ = fs + tau_hi;
#pragma unroll
        for (int q = 0; q < SPL; ++q) np_[q] = nq[q] + (unsigned)tau_hi * spu;
        // rows t_first, t_first - 1, .. of b -> stage `st` (clamped at row 0: the tail chunk ignores the extras)
        auto prefetch_b = [&](int st, int t_first) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const unsigned off = (unsigned)max(t_first - k, 0) * spu;
#pragma unroll
                for (int q = 0; q < SPL; ++q)
                    if (emit[q]) cp_async4(&sm_b[st][k][lane * SPL + q], bq[q] + off);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int d = 0; d < NST - 1; ++d) prefetch_b(d, tau_hi - d * CH);
        int stage = 0;
        auto chunk = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            float es[CH][SPL], g2[CH];
            cp_async_wait<NST - 2>();
#pragma unroll
            for (int k = 0; k < CH; ++k) {  // off the dependency chain
                float e[SPL];
#pragma unroll
                for (int q = 0; q < SPL; ++q) e[q] = emit[q] ? sm_b[stage][k][lane * SPL + q] : PC_NEG_INF;
                g2[k] = ceilf(frame_max(e) * kLog2e);
                shift_row(e, g2[k], es[k]);
            }
            prefetch_b((stage + NST - 1) % NST, tau_hi - (NST - 1) * CH);
            stage = (stage + 1) % NST;
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int tau = tau_hi - k;
                if (FULL || tau >= 1) {
                    float nb[SPL], raw[SPL];
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        nb[q] = bh[q] + es[k][q];
                        if (emit[q]) *np_[q] = nb[q];
                        np_[q] -= spu;
                    }
                    // lane 0 holds the entry state in slot 0; log2(b_tau beta_tau) = nb_tau + (Cb + g2)
                    if (lane == 0) *fp_ = make_float4(Cb + g2[k], g2[k], nb[0], 0.f);
                    --fp_;
                    float up = __shfl_down_sync(0xffffffffu, nb[0], 1);
                    if (lane == 31) up = PC_NEG_INF;
#pragma unroll
                    for (int q = 0; q < SPL; ++q) {
                        const float nxt = (q + 1 < SPL) ? nb[(q + 1) % SPL] : up;
                        raw[q] = logadd2(ls[q] + nb[q], ln[q] + nxt);
                    }
                    float r = 0.f;
                    if (FULL && k == CH - 1) {
                        float m = raw[0];
#pragma unroll
                        for (int q = 1; q < SPL; ++q) m = fmaxf(m, raw[q]);
                        m = redux_max(m);
                        r = (m == PC_NEG_INF) ? 0.f : ceilf(m);
                    }
                    Cb += g2[k] + r;
#pragma unroll
                    for (int q = 0; q < SPL; ++q) bh[q] = raw[q] - r;
                }
            }
            tau_hi -= CH;
        };
        while (tau_hi >= CH) chunk(std::true_type{});
        if (tau_hi >= 1) chunk(std::false_type{});
        cp_async_wait_all();  // the look-ahead requests past row 0 still target the ring the forward pass reuses
    }
    if (trace) g_fw_dbg[1] = clock64();
    float e0[SPL];
    load_row(0, e0);
    const float g20 = ceilf(frame_max(e0) * kLog2e);

    // ----------------------------------------- pi iteration (LHMM.py:447-452,526-544; A.3)
    // natural-log fp64 on w = B[:,0] + beta_0, relative to Cb (added back for log P)
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
    double q_prev = (double)PC_NEG_INF, qn = (double)PC_NEG_INF;
    for (int guard = 0; guard < 100000; ++guard) {
        double m = (double)PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) m = fmax(m, lp[q] + w[q]);
        m = warp_max_d(m);
        qn = m;
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
    // log gamma_0 = log pi + w - q (LHMM.py:486-500 at t = 0); log P(O) = q + backward shifts
    float tmx[SPL];  // running maximum of log gamma per state over the current 128-frame tile (K3's flags)
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        const double lg0 = lp_used[q] + w[q] - qn;  // NaN when the utterance has no path at all
        tmx[q] = (kind[q] == 1) ? (float)lg0 : PC_NEG_INF;
        if (kind[q] == 1) gu[col[q]] = (float)lg0;
    }
    if (lane == 0) {
        utt_logp[u] = qn + (double)Cb * 0.6931471805599453;
        utt_iters[u] = iters;
    }
    if (trace) g_fw_dbg[2] = clock64();

    // K3's activity flags: every (tile, position) flag is written exactly once per run by the lane that holds
    // the position's first state (the maxima of the position's other two states arrive by shuffle): bit k =
    // the 32-frame block k of the tile carries posterior mass (some log gamma above PC_ACTIVE_MIN_LGAM)
    int32_t *flag0[SPL];  // -> flag of (tile 0, this state's position); the pair's tiles are consecutive
#pragma unroll
    for (int q = 0; q < SPL; ++q)
        flag0[q] = v.tile_active + ((kind[q] == 1) ? v.pair_tile0[p0 + col[q] / PC_EMIT] : 0);
    auto next_state = [&](const float (&x)[SPL], float (&y)[SPL]) {  // y[state s] = x[state s + 1]
        const float up = __shfl_down_sync(0xffffffffu, x[0], 1);
#pragma unroll
        for (int q = 0; q < SPL; ++q) y[q] = (q + 1 < SPL) ? x[(q + 1) % SPL] : up;
    };
    int tmask[SPL];
#pragma unroll
    for (int q = 0; q < SPL; ++q) tmask[q] = 0;
    // end of the 32-frame block that holds frame t (t = 31 mod 32, or the utterance's last frame)
    auto flag_block = [&](int t) {
        float m1[SPL], m2[SPL];
        next_state(tmx, m1);
        next_state(m1, m2);
        const int bit = 1 << ((t / PC_BLOCK_ROWS) & (PC_TILE_ROWS / PC_BLOCK_ROWS - 1));
        const bool tile_end = (t & (PC_TILE_ROWS - 1)) == PC_TILE_ROWS - 1 || t == T - 1;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            if (fmaxf(fmaxf(tmx[q], m1[q]), m2[q]) > PC_ACTIVE_MIN_LGAM) tmask[q] |= bit;
            tmx[q] = PC_NEG_INF;
            if (tile_end) {
                if (kind[q] == 1 && col[q] % PC_EMIT == 0) flag0[q][t / PC_TILE_ROWS] = tmask[q];
                tmask[q] = 0;
            }
        }
    };
    if (T == 1) flag_block(0);
    __syncwarp();  // lane 0's per-frame records are read by every lane below

    // ------------------------------------------------------------------ forward (LHMM.py:335-351)
    // + log gamma (LHMM.py:486-500) + expected transition counts (LHMM.py:431-445)
    float ah[SPL];
    float Sa;      // (forward shifts up to the previous frame) - floor(log2 P): an exact integer
    float fPc;     // frac(log2 P) + the measured drift of the normaliser
    {
        float m = PC_NEG_INF;
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            const float es0 = (kind[q] == 1) ? fmaf(e0[q], kLog2e, -g20) : (kind[q] == 0 ? -g20 : PC_NEG_INF);
            ah[q] = (kind[q] == 2) ? PC_NEG_INF : (float)(lp_used[q] * 1.4426950408889634) + es0;
            m = fmaxf(m, ah[q]);
        }
        m = redux_max(m);
        m = (m == PC_NEG_INF) ? 0.f : ceilf(m);
#pragma unroll
        for (int q = 0; q < SPL; ++q) ah[q] -= m;
        const double LP = qn * 1.4426950408889634 + (double)Cb;  // log2 P(O)
        const double IP = floor(LP);
        Sa = (float)((double)g20 + (double)m - IP);
        fPc = (float)(LP - IP);
        if (!(fabs(LP) < 1e30)) { Sa = 0.f; fPc = 0.f; }  // no path at all: the rows come out -inf / NaN either way
    }
    float ms[SPL], cs[SPL], mn[SPL], cn[SPL];  // running log-sum-exp of the stay / move counts
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        ms[q] = mn[q] = -1e30f;  // finite floor: 2^(x - floor) = 0 for every real x below it
        cs[q] = cn[q] = 0.f;
    }
    {
        int tau_lo = 1;
        float *gp_[SPL];  // -> (frame t, this state) of the log gamma rows, walking up
#pragma unroll
        for (int q = 0; q < SPL; ++q) gp_[q] = gq[q] + spu;
        // rows t0 .. t0 + CH - 1 of b and nb and the per-frame records -> stage `st` (clamped at row T - 1)
        auto prefetch_f = [&](int st, int t0) {
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int t = min(t0 + k, T - 1);
                const unsigned off = (unsigned)t * spu;
#pragma unroll
                for (int q = 0; q < SPL; ++q)
                    if (emit[q]) {
                        cp_async4(&sm_b[st][k][lane * SPL + q], bq[q] + off);
                        cp_async4(&sm_n[st][k][lane * SPL + q], nq[q] + off);
                    }
                if (lane == k) cp_async16(&sm_f[st][k], fs + t);
            }
            cp_async_commit();
        };
#pragma unroll
        for (int d = 0; d < NST - 1; ++d) prefetch_f(d, tau_lo + d * CH);
        int stage = 0;
        auto chunk = [&](auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            float es[CH][SPL], nbc[CH][SPL], g2[CH], db[CH];
            cp_async_wait<NST - 2>();
            __syncwarp();  // the per-frame records were fetched by lanes 0 .. CH-1
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const float4 f = sm_f[stage][k];
                db[k] = f.x;
                g2[k] = f.y;
                float e[SPL];
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    e[q] = emit[q] ? sm_b[stage][k][lane * SPL + q] : PC_NEG_INF;
                    nbc[k][q] = emit[q] ? sm_n[stage][k][lane * SPL + q] : (kind[q] == 0 ? f.z : PC_NEG_INF);
                }
                shift_row(e, g2[k], es[k]);
            }
            __syncwarp();  // every lane has read the records before the other stage's refill can land on a reused slot
            prefetch_f((stage + NST - 1) % NST, tau_lo + (NST - 1) * CH);
            stage = (stage + 1) % NST;
            float xs[CH][SPL], xn[CH][SPL], lg2_last[SPL];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int t = tau_lo + k;
                const bool valid = FULL || t <= T - 1;
                // scale constant of frame t: (exact integer) - (fraction of log2 P + drift)
                const float cx = (Sa + db[k]) - fPc;
                float left = __shfl_up_sync(0xffffffffu, ah[SPL - 1] + ln[SPL - 1], 1);
                if (lane == 0) left = PC_NEG_INF;
                float raw[SPL];
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    const float stay = ah[q] + ls[q];
                    const float move = (q > 0) ? ah[(q + SPL - 1) % SPL] + ln[(q + SPL - 1) % SPL] : left;
                    raw[q] = logadd2(stay, move);
                    const float nc = nbc[k][q] + cx;
                    xs[k][q] = valid ? stay + nc : PC_NEG_INF;
                    xn[k][q] = valid ? move + nc : PC_NEG_INF;
                    const float l2 = raw[q] + nc;
                    if (k == CH - 1) lg2_last[q] = l2;
                    const float lg = l2 * kLn2;
                    if (valid && emit[q]) *gp_[q] = lg;
                    gp_[q] += spu;
                    if (valid) tmx[q] = fmaxf(tmx[q], lg);  // (only emitting states' maxima are ever read)
                }
                if (valid) {
                    float r = 0.f;
                    if (FULL && k == CH - 1) {
                        float m = raw[0] + es[k][0];
#pragma unroll
                        for (int q = 1; q < SPL; ++q) m = fmaxf(m, raw[q] + es[k][q]);
                        m = redux_max(m);
                        r = (m == PC_NEG_INF) ? 0.f : ceilf(m);
                    }
                    Sa += g2[k] + r;
#pragma unroll
                    for (int q = 0; q < SPL; ++q) ah[q] = raw[q] + es[k][q] - r;
                }
                // a block's last frame, t = 31 mod 32, sits at k = CH - 2 of its chunk (chunks start at t = 1 mod CH)
                if (k == CH - 2 && valid && (t & (PC_BLOCK_ROWS - 1)) == PC_BLOCK_ROWS - 1) flag_block(t);
            }
            // expected counts: log-sum-exp over the chunk with one shared reference per state (branch-free)
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                float gs = xs[0][q], gn = xn[0][q];
#pragma unroll
                for (int k = 1; k < CH; ++k) { gs = fmaxf(gs, xs[k][q]); gn = fmaxf(gn, xn[k][q]); }
                const float ns = fmaxf(ms[q], gs), nn = fmaxf(mn[q], gn);
                float as = cs[q] * ex2f(ms[q] - ns), an = cn[q] * ex2f(mn[q] - nn);
#pragma unroll
                for (int k = 0; k < CH; ++k) { as += ex2f(xs[k][q] - ns); an += ex2f(xn[k][q] - nn); }
                ms[q] = ns; cs[q] = as; mn[q] = nn; cn[q] = an;
            }
            if (FULL) {
                // drift of the normaliser at the chunk's last frame: log2 sum_j gamma(j) should be 0
                float m = lg2_last[0];
#pragma unroll
                for (int q = 1; q < SPL; ++q) m = fmaxf(m, lg2_last[q]);
                m = redux_max(m);
                float z = 0.f;
#pragma unroll
                for (int q = 0; q < SPL; ++q) z += ex2f(lg2_last[q] - m);
                z = warp_sum(z);
                const float drift = m + lg2f(z);
                if (fabsf(drift) < 1e-2f) fPc += drift;  // (an utterance without any path has no normaliser)
            }
            tau_lo += CH;
        };
        while (tau_lo + CH - 1 <= T - 1) chunk(std::true_type{});
        if (tau_lo <= T - 1) chunk(std::false_type{});
        cp_async_wait_all();
        if (T > 1 && ((T - 1) & (PC_BLOCK_ROWS - 1)) != PC_BLOCK_ROWS - 1) flag_block(T - 1);  // the last, partial block
    }
    // the move counts were collected at the destination state: state s's "next" count sits with state s + 1
    {
        float mn1[SPL], cn1[SPL];
        next_state(mn, mn1);
        next_state(cn, cn1);
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            if (kind[q] == 1) {
                const int s = lane * SPL + q;
                float *o = pair_trans + (size_t)(p0 + (s - 1) / PC_EMIT) * PC_TRANS_SLOTS + ((s - 1) % PC_EMIT) * 3;
                const float ks = (cs[q] > 0.f) ? ms[q] + lg2f(cs[q]) : PC_NEG_INF;
                const float kn = (s < NE && cn1[q] > 0.f) ? mn1[q] + lg2f(cn1[q]) : PC_NEG_INF;
                o[0] = ks * kLn2;
                o[1] = kn * kLn2;
                o[2] = logadd2(ks, kn) * kLn2;  // occupancy over t < T-1 = self + next
            }
        }
    }
    if (trace) g_fw_dbg[3] = clock64();
}

template <int SPL>
int launch_fw(pc_handle h, const CorpusView &v, const float *b, const double *log_self, const double *log_next,
              float *lgam, float *scratch0, double *utt_logp, int32_t *utt_iters, float *pair_trans,
              cudaStream_t st) {
    const int blocks = (v.n_utt + FW_WPB - 1) / FW_WPB;
    fwdbwd_warp_kernel<SPL><<<blocks, FW_WPB * 32, 0, st>>>(v, b, log_self, log_next, lgam,
                                                            reinterpret_cast<float4 *>(scratch0), utt_logp, utt_iters,
                                                            pair_trans);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

// block 0's phase clocks: [0] start, [1] backward done, [2] pi iteration done, [3] forward done
extern "C" int pc_debug_read_fw(long long *host_out) {
    return cudaMemcpyFromSymbol(host_out, g_fw_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -2;
}

int launch_forward_backward_warp(pc_handle h, const CorpusView &v, const float *b, const double *log_self,
                                 const double *log_next, float *lgam, float *scratch0, double *utt_logp,
                                 int32_t *utt_iters, float *pair_trans, cudaStream_t st) {
    if (v.n_utt == 0) return PC_OK;
    const int states = PC_EMIT * v.max_labels + 1;
    if (states <= 32) return launch_fw<1>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 64) return launch_fw<2>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 128) return launch_fw<4>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 256) return launch_fw<8>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    pc_set_error("pc_forward_backward: %d labels per utterance exceeds the limit of 85", v.max_labels);
    return PC_ERR_UNSUPPORTED;
}
