// K2: log-space forward-backward over the banded sentence HMM.  Three warps per utterance: a CHAIN
// warp that runs nothing but the two T-step recurrences, and two HELPER warps that turn alpha /
// beta rows into posteriors (A) and expected transition counts (B) while the forward chain is
// still running.
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
//     Utterances are weighted by their likelihood there, so an expected count of e^-100 in a likely
//     utterance outweighs a count of 1 in an unlikely one: the counts need the full log range,
//     which is why the recurrences stay in the log domain (a scaled linear-domain kernel was
//     measured 8x faster and rejected for flushing those tails, profiles/experiments/).
//
// Arithmetic: fp32 log2-domain.  Every frame's emissions are shifted by g_t = max_j b_t(j) (one
// CREDUX.MAX.F32, computed ahead of the recurrence from the prefetched row: off the dependency
// chain) and every FB_RENORM frames the state vector is renormalised exactly (one CREDUX on the
// chain), so |alpha_hat|, |beta_hat| stay O(100) and the fp32 ulp stays ~1e-5 or below (SURVEY §7
// hard part 2b).  Only the backward pass accumulates its shifts (fp64): log P(O) is the pi
// iteration's last q plus that sum.  Nothing else needs absolute scales:
//   * log gamma_t(j) = alpha_hat_t(j) + beta_hat_t(j) - LSE_j(...) per frame (LHMM.py:486-500),
//   * xi_t(i,i) / P = gamma_{t-1}(i) * [stay term / (stay + move)] of the backward recurrence at t,
//     whose frame shift cancels in the ratio (LHMM.py:431-445).
// Chain warp, per frame and pass: one shuffle (the j+1 / j-1 neighbour), one log-add, one store.
// The forward pass hands alpha_hat rows to helper A through a shared-memory ring (mbarrier full /
// empty per group of FB_GRP frames), helper A hands log gamma rows to helper B the same way;
// beta_hat rows travel through a scratch buffer of the corpus (same layout as b).  Lane l owns states [l*SPL, (l+1)*SPL); emissions
// are time-major (b[t][s]), so a warp reads / writes one coalesced row per frame.
#include <type_traits>

#include "tc_common.cuh"

#define FB_RENORM 8
#define FB_GRP 8    // frames per ring group (producer / consumer hand-over granularity)
#define FB_NGRP 4   // ring depth in groups
// utterances (warp triples) per block: 4 with the warps in role-major order when registers allow, so
// that every SM sub-partition (warp index mod 4) hosts one chain warp and one helper of each kind per
// block; with 2 (utterance-major order) one chain warp shared its sub-partition with a helper and
// ran the forward pass at 272 clk per frame while the other ran alone
template <int SPL>
struct FbCfg {
    static constexpr int UPB = (SPL <= 2) ? 4 : 2;
};

__device__ long long g_fb_dbg[16];

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
// warp-wide maximum in one instruction (sm_100a CREDUX); NaN operands are ignored
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

template <int SPL>
struct PairSmem {
    float ring[FB_NGRP * FB_GRP][32 * SPL];  // alpha_hat rows of frames 1.. (slot (t-1) % depth)
    float lgr[FB_NGRP * FB_GRP][32 * SPL];   // log2 gamma rows of the same frames (helper A -> helper B)
    float row0[32 * SPL];                    // log2 gamma_0 (all materialised states)
    float tmx_row[32 * SPL + 2];             // per-state maxima of log gamma over a tile (helper A, at tile ends)
    uint64_t full[FB_NGRP], empty[FB_NGRP], full2[FB_NGRP], empty2[FB_NGRP], start;
};

template <int SPL, int CH>
__global__ void __launch_bounds__(FbCfg<SPL>::UPB * 96)
fwdbwd_kernel(CorpusView v, const float *__restrict__ b, const double *__restrict__ log_self,
              const double *__restrict__ log_next, float *lgam, float *scratch,
              double *__restrict__ utt_logp, int32_t *__restrict__ utt_iters,
              float *__restrict__ pair_trans) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int FB_UPB = FbCfg<SPL>::UPB;
    // role: 0 chain, 1 helper A (log gamma), 2 helper B (transition counts); slot: utterance inside the block
    const int slot = (FB_UPB == 4) ? (warp & 3) : warp / 3;
    const int role = (FB_UPB == 4) ? (warp >> 2) : warp - 3 * slot;
    const bool helper = role != 0;
    const int idx = blockIdx.x * FB_UPB + slot;
    if (idx >= v.n_utt) return;      // the three warps of an utterance leave together
    PairSmem<SPL> *sm = reinterpret_cast<PairSmem<SPL> *>(smem_raw) + slot;
    if (role == 0 && lane == 0) {
        for (int i = 0; i < FB_NGRP; ++i) {
            tc::mbar_init(&sm->full[i], 1); tc::mbar_init(&sm->empty[i], 1);
            tc::mbar_init(&sm->full2[i], 1); tc::mbar_init(&sm->empty2[i], 1);
        }
        tc::mbar_init(&sm->start, 1);
        tc::mbar_fence_init();
    }
    asm volatile("bar.sync %0, 96;" ::"r"(slot + 1) : "memory");  // this utterance's warps only

    const int u = v.fb_order[idx];
    const int64_t f0 = v.frame_off[u];
    const int T = (int)(v.frame_off[u + 1] - f0);
    const int64_t p0 = v.pair_off[u];
    const int L = (int)(v.pair_off[u + 1] - p0);
    const int NE = PC_EMIT * L;  // states 0..NE are materialised; N = NE + 2
    const int sp = pc_spad(L);
    const float *bu = b + v.emis_off[u];
    float *gu = lgam + v.emis_off[u];
    float *bsu = v.scratch1 + v.emis_off[u];  // beta_hat rows of the emitting states
    float *eb = scratch + f0;                 // beta_hat of the entry state, one float per frame

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
        float m = e[0];
#pragma unroll
        for (int q = 1; q < SPL; ++q) m = fmaxf(m, e[q]);
        m = redux_max(m);
        return (m == PC_NEG_INF) ? 0.f : m;
    };
    // shifted log2 emissions: (b - g)*log2e; the entry state emits log 1 = 0
    auto shift_row = [&](const float (&e)[SPL], float g, float (&es)[SPL]) {
#pragma unroll
        for (int q = 0; q < SPL; ++q)
            es[q] = (kind[q] == 1) ? (e[q] - g) * kLog2e : (kind[q] == 0 ? -g * kLog2e : PC_NEG_INF);
    };

    const bool trace = (blockIdx.x == 0 && slot == 0 && lane == 0);
    if (trace) g_fb_dbg[role * 4] = clock64();
    // per-lane row pointers: an emitting state lives in column col of the [T][SP] blocks, the entry
    // state's beta_hat in eb (one float per frame); one predicated store serves both
    bool live[SPL];
    int stride[SPL];
    float *grow[SPL];         // -> (frame 0, this state) in lgam (emitting states)
    float *hrow[SPL];         // -> (frame 0, this state) in the beta_hat rows / eb
    const float *brow[SPL];   // -> (frame 0, this state) in b (emitting states only)
#pragma unroll
    for (int q = 0; q < SPL; ++q) {
        live[q] = kind[q] != 2;
        stride[q] = (kind[q] == 1) ? sp : 1;
        grow[q] = gu + col[q];
        hrow[q] = (kind[q] == 1) ? bsu + col[q] : eb;
        brow[q] = bu + col[q];
    }
    if (!helper) {
        // ============================================================ CHAIN WARP
        // ---------------------------------------------------------- backward (LHMM.py:353-366)
        float bh[SPL];
        double Cb = 0.0;  // log2 beta_t = bh + Cb
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            bh[q] = live[q] ? 0.f : PC_NEG_INF;
            if (live[q]) hrow[q][(size_t)(T - 1) * stride[q]] = 0.f;
        }
        {
            // step tau (= T-1 .. 1) consumes emission row tau and produces beta_hat_{tau-1}
            float e_nxt[CH][SPL];
            int tau_hi = T - 1;
#pragma unroll
            for (int k = 0; k < CH; ++k) load_row(max(tau_hi - k, 0), e_nxt[k]);
            // FULL: all CH frames of the chunk exist (no per-frame test, renormalisation at a fixed
            // position); otherwise the tail of < CH frames
            auto chunk = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                float es[CH][SPL], g[CH];
                double gs = 0.0;
#pragma unroll
                for (int k = 0; k < CH; ++k) {  // off the dependency chain
                    g[k] = frame_max(e_nxt[k]);
                    shift_row(e_nxt[k], g[k], es[k]);
                    if (FULL || tau_hi - k >= 1) gs += (double)g[k];
                }
                Cb += gs * 1.4426950408889634;
#pragma unroll
                for (int k = 0; k < CH; ++k) load_row(max(tau_hi - CH - k, 0), e_nxt[k]);  // prefetch
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const int tau = tau_hi - k;
                    if (FULL || tau >= 1) {
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
                        if (FULL && k == CH - 1) {  // exact renormalisation once per full chunk
                            float m = raw[0];
#pragma unroll
                            for (int q = 1; q < SPL; ++q) m = fmaxf(m, raw[q]);
                            m = redux_max(m);
                            r = (m == PC_NEG_INF) ? 0.f : m;
                            Cb += (double)r;
                        }
#pragma unroll
                        for (int q = 0; q < SPL; ++q) {
                            bh[q] = raw[q] - r;
                            if (live[q]) hrow[q][(size_t)(tau - 1) * stride[q]] = bh[q];
                        }
                    }
                }
                tau_hi -= CH;
            };
            while (tau_hi >= CH) chunk(std::true_type{});
            if (tau_hi >= 1) chunk(std::false_type{});
        }
        if (trace) g_fb_dbg[1] = clock64();
        float e0[SPL];
        load_row(0, e0);
        const float g0 = frame_max(e0);

        // ------------------------------------- pi iteration (LHMM.py:447-452,526-544; A.3)
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
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            const double lg0 = lp_used[q] + w[q] - qn;  // NaN when the utterance has no path at all
            sm->row0[lane * SPL + q] = (kind[q] == 2) ? PC_NEG_INF : (float)(lg0 * 1.4426950408889634);
            if (kind[q] == 1) gu[col[q]] = (float)lg0;
        }
        if (lane == 0) {
            utt_logp[u] = qn + Cb * 0.6931471805599453;
            utt_iters[u] = iters;
        }
        __threadfence_block();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&sm->start);  // beta_hat rows and row0 are in place
        if (trace) g_fb_dbg[2] = clock64();

        // ---------------------------------------------------------- forward (LHMM.py:335-351)
        float ah[SPL];
        {
            float es0[SPL];
            shift_row(e0, g0, es0);
            float m = PC_NEG_INF;
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                ah[q] = (kind[q] == 2) ? PC_NEG_INF : (float)(lp_used[q] * 1.4426950408889634) + es0[q];
                m = fmaxf(m, ah[q]);
            }
            m = redux_max(m);
            if (m == PC_NEG_INF) m = 0.f;
#pragma unroll
            for (int q = 0; q < SPL; ++q) ah[q] -= m;
        }
        {
            float e_nxt[CH][SPL];
            int tau_lo = 1;
#pragma unroll
            for (int k = 0; k < CH; ++k) load_row(min(tau_lo + k, T - 1), e_nxt[k]);
            static_assert(CH == FB_GRP, "one prefetch chunk = one ring group");
            uint32_t grp = 0;  // running group counter
            auto chunk = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                float es[CH][SPL];
#pragma unroll
                for (int k = 0; k < CH; ++k) shift_row(e_nxt[k], frame_max(e_nxt[k]), es[k]);
#pragma unroll
                for (int k = 0; k < CH; ++k) load_row(min(tau_lo + CH + k, T - 1), e_nxt[k]);  // prefetch
                const int gs = grp % FB_NGRP;
                tc::mbar_wait(&sm->empty[gs], ((grp / FB_NGRP) & 1) ^ 1);  // the helper is done with this slot
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    if (FULL || tau_lo + k <= T - 1) {
                        float left = __shfl_up_sync(0xffffffffu, ah[SPL - 1] + ln[SPL - 1], 1);
                        if (lane == 0) left = PC_NEG_INF;
                        float raw[SPL];
#pragma unroll
                        for (int q = 0; q < SPL; ++q) {
                            const float from_left = (q > 0) ? ah[(q + SPL - 1) % SPL] + ln[(q + SPL - 1) % SPL] : left;
                            raw[q] = logadd2(ah[q] + ls[q], from_left) + es[k][q];
                        }
                        float r = 0.f;
                        if (FULL && k == CH - 1) {
                            float m = raw[0];
#pragma unroll
                            for (int q = 1; q < SPL; ++q) m = fmaxf(m, raw[q]);
                            m = redux_max(m);
                            r = (m == PC_NEG_INF) ? 0.f : m;
                        }
#pragma unroll
                        for (int q = 0; q < SPL; ++q) {
                            ah[q] = raw[q] - r;
                            sm->ring[gs * FB_GRP + k][lane * SPL + q] = ah[q];
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&sm->full[gs]);
                tau_lo += CH;
                ++grp;
            };
            while (tau_lo + CH - 1 <= T - 1) chunk(std::true_type{});
            if (tau_lo <= T - 1) chunk(std::false_type{});
        }
        if (trace) g_fb_dbg[3] = clock64();
    } else if (role == 1) {
        // ============================================================ HELPER A: log gamma
        // per frame t >= 1: log gamma_t(j) = alpha_hat + beta_hat - LSE_j (LHMM.py:486-500).  A group
        // of FB_GRP frames goes stage by stage, branch-free, so that the frames' independent
        // reductions overlap.  Rows go to lgam (natural log) and to helper B (log2).
        tc::mbar_wait(&sm->start, 0);
        float bt_nxt[FB_GRP][SPL];
        int tau_lo = 1;
        auto load_beta = [&](int t0, float (&x)[FB_GRP][SPL]) {
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                const int t = min(t0 + k, T - 1);
#pragma unroll
                for (int q = 0; q < SPL; ++q) x[k][q] = live[q] ? __ldcg(hrow[q] + t * stride[q]) : PC_NEG_INF;
            }
        };
        load_beta(tau_lo, bt_nxt);
        // K3's activity flags, set while the rows are at hand: running maximum of log gamma per state over
        // the current 128-frame tile (frame 0 comes from the chain warp's row0), flag of the (tile, position)
        // pair at the tile's last frame - the same values and threshold as K3's own pre-pass reads back
        float tmx[SPL];
        int32_t *flag0[SPL];  // -> flag of (tile 0, this state's position); the pair's tiles are consecutive
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            tmx[q] = (kind[q] == 1) ? sm->row0[lane * SPL + q] * kLn2 : PC_NEG_INF;
            flag0[q] = v.tile_active + ((kind[q] == 1) ? v.pair_tile0[p0 + col[q] / PC_EMIT] : 0);
        }
        // every (tile, position) flag is written exactly once per run, 0 or 1, by the lane that holds the
        // position's first state (the three states' maxima meet in shared memory): no memset, no atomics
        auto flag_tile = [&](int k_tile) {
#pragma unroll
            for (int q = 0; q < SPL; ++q) sm->tmx_row[lane * SPL + q] = tmx[q];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                if (kind[q] == 1 && col[q] % PC_EMIT == 0) {
                    const int s = lane * SPL + q;
                    const float m = fmaxf(fmaxf(sm->tmx_row[s], sm->tmx_row[s + 1]), sm->tmx_row[s + 2]);
                    flag0[q][k_tile] = m > PC_ACTIVE_MIN_LGAM ? 0xF : 0;  // (per-tile resolution: every 32-frame block)
                }
                tmx[q] = PC_NEG_INF;
            }
            __syncwarp();
        };
        uint32_t grp = 0;
        while (tau_lo <= T - 1) {
            float bt[FB_GRP][SPL];
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
#pragma unroll
                for (int q = 0; q < SPL; ++q) bt[k][q] = bt_nxt[k][q];
            }
            load_beta(tau_lo + FB_GRP, bt_nxt);  // prefetch the next group
            const int gs = grp % FB_NGRP;
            tc::mbar_wait(&sm->full[gs], (grp / FB_NGRP) & 1);
            float ab[FB_GRP][SPL], mm[FB_GRP], ssum[FB_GRP];
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                float m = PC_NEG_INF;
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    ab[k][q] = sm->ring[gs * FB_GRP + k][lane * SPL + q] + bt[k][q];
                    m = fmaxf(m, ab[k][q]);
                }
                mm[k] = m;
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->empty[gs]);  // the chain warp may refill this slot
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                const float m = redux_max(mm[k]);
                mm[k] = (m == PC_NEG_INF) ? 0.f : m;
            }
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                float sacc = 0.f;
#pragma unroll
                for (int q = 0; q < SPL; ++q) sacc += ex2f(ab[k][q] - mm[k]);
                ssum[k] = sacc;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < FB_GRP; ++k) ssum[k] += __shfl_xor_sync(0xffffffffu, ssum[k], o);
            }
            tc::mbar_wait(&sm->empty2[gs], ((grp / FB_NGRP) & 1) ^ 1);  // helper B is done with this slot
            const int kb = (tau_lo | (PC_TILE_ROWS - 1)) - tau_lo;  // group index of the tile's last frame (>= FB_GRP: not here)
            float tnx[SPL];
#pragma unroll
            for (int q = 0; q < SPL; ++q) tnx[q] = PC_NEG_INF;
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                const float Z = mm[k] + lg2f(ssum[k]);  // -inf when the frame carries no mass at all
                const bool valid = tau_lo + k <= T - 1;
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    const float lg = ab[k][q] - Z;
                    sm->lgr[gs * FB_GRP + k][lane * SPL + q] = lg;
                    if (valid && kind[q] == 1) {
                        grow[q][(tau_lo + k) * sp] = lg * kLn2;
                        // frames up to the tile's last one (index kb inside the group) / frames of the next tile
                        if (k <= kb) tmx[q] = fmaxf(tmx[q], lg * kLn2);
                        else tnx[q] = fmaxf(tnx[q], lg * kLn2);
                    }
                }
            }
            if (kb < FB_GRP && tau_lo + kb <= T - 1) {  // a tile's last frame lies in this group: one uniform branch per group
                flag_tile(tau_lo / PC_TILE_ROWS);
#pragma unroll
                for (int q = 0; q < SPL; ++q) tmx[q] = tnx[q];
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->full2[gs]);
            tau_lo += FB_GRP;
            ++grp;
        }
        if (((T - 1) & (PC_TILE_ROWS - 1)) != PC_TILE_ROWS - 1) flag_tile((T - 1) / PC_TILE_ROWS);  // the last, partial tile
        if (trace) g_fb_dbg[5] = clock64();
    } else {
        // ============================================================ HELPER B: transition counts
        // gamma_{t-1}(i) split into stay / move by the backward recurrence at t (LHMM.py:431-445;
        // any common shift of frame t cancels in the ratio), log-sum-exp over t per state
        float ms[SPL], cs[SPL], mn[SPL], cn[SPL], lg_prev[SPL];
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            ms[q] = mn[q] = -1e30f;  // finite floor: 2^(x - floor) = 0 for every real x below it
            cs[q] = cn[q] = 0.f;
        }
        // nb_t(j) = log2 emission + beta_hat: GMM states from b, the entry state emits log 1
        auto load_nb = [&](int t0, float (&x)[FB_GRP][SPL]) {
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                const int t = min(t0 + k, T - 1);
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    const float e = (kind[q] == 1) ? __ldg(brow[q] + t * sp) * kLog2e : 0.f;
                    x[k][q] = live[q] ? __ldcg(hrow[q] + t * stride[q]) + e : PC_NEG_INF;
                }
            }
        };
        tc::mbar_wait(&sm->start, 0);
#pragma unroll
        for (int q = 0; q < SPL; ++q) lg_prev[q] = sm->row0[lane * SPL + q];
        float nb_nxt[FB_GRP][SPL];
        int tau_lo = 1;
        load_nb(tau_lo, nb_nxt);
        uint32_t grp = 0;
        while (tau_lo <= T - 1) {
            float nb[FB_GRP][SPL];
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
#pragma unroll
                for (int q = 0; q < SPL; ++q) nb[k][q] = nb_nxt[k][q];
            }
            load_nb(tau_lo + FB_GRP, nb_nxt);  // prefetch the next group
            float xs[FB_GRP][SPL], xn[FB_GRP][SPL], fs[FB_GRP][SPL], fn[FB_GRP][SPL];
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                float up = __shfl_down_sync(0xffffffffu, nb[k][0], 1);
                if (lane == 31) up = PC_NEG_INF;
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    const float nxt = (q + 1 < SPL) ? nb[k][(q + 1) % SPL] : up;
                    const float sv = ls[q] + nb[k][q], mv = ln[q] + nxt;
                    float split = logadd2(sv, mv);
                    split = (split == PC_NEG_INF) ? 0.f : split;
                    fs[k][q] = sv - split;
                    fn[k][q] = mv - split;
                }
            }
            const int gs = grp % FB_NGRP;
            tc::mbar_wait(&sm->full2[gs], (grp / FB_NGRP) & 1);
#pragma unroll
            for (int k = 0; k < FB_GRP; ++k) {
                const bool valid = tau_lo + k <= T - 1;
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    const float lgp = (k == 0) ? lg_prev[q] : sm->lgr[gs * FB_GRP + k - 1][lane * SPL + q];
                    xs[k][q] = valid ? lgp + fs[k][q] : PC_NEG_INF;
                    xn[k][q] = valid ? lgp + fn[k][q] : PC_NEG_INF;
                }
            }
            {
                const int last = min(FB_GRP - 1, T - 1 - tau_lo);
#pragma unroll
                for (int q = 0; q < SPL; ++q) lg_prev[q] = sm->lgr[gs * FB_GRP + last][lane * SPL + q];
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&sm->empty2[gs]);
#pragma unroll
            for (int q = 0; q < SPL; ++q) {
                // running log-sum-exp with a group-wide reference: one rescale per group
                float gm_s = xs[0][q], gm_n = xn[0][q];
#pragma unroll
                for (int k = 1; k < FB_GRP; ++k) { gm_s = fmaxf(gm_s, xs[k][q]); gm_n = fmaxf(gm_n, xn[k][q]); }
                const float ns = fmaxf(ms[q], gm_s), nn = fmaxf(mn[q], gm_n);
                float as = cs[q] * ex2f(ms[q] - ns), an = cn[q] * ex2f(mn[q] - nn);
#pragma unroll
                for (int k = 0; k < FB_GRP; ++k) { as += ex2f(xs[k][q] - ns); an += ex2f(xn[k][q] - nn); }
                ms[q] = ns; cs[q] = as; mn[q] = nn; cn[q] = an;
            }
            tau_lo += FB_GRP;
            ++grp;
        }
#pragma unroll
        for (int q = 0; q < SPL; ++q) {
            if (kind[q] == 1) {
                const int s = lane * SPL + q;
                float *o = pair_trans + (size_t)(p0 + (s - 1) / PC_EMIT) * PC_TRANS_SLOTS +
                           ((s - 1) % PC_EMIT) * 3;
                const float ks = (cs[q] > 0.f) ? ms[q] + lg2f(cs[q]) : PC_NEG_INF;
                const float kn = (cn[q] > 0.f) ? mn[q] + lg2f(cn[q]) : PC_NEG_INF;
                o[0] = ks * kLn2;
                o[1] = kn * kLn2;
                o[2] = logadd2(ks, kn) * kLn2;  // occupancy over t < T-1 = self + next
            }
        }
        if (trace) g_fb_dbg[9] = clock64();
    }
}

template <int SPL>
int launch_fb(pc_handle h, const CorpusView &v, const float *b, const double *log_self,
              const double *log_next, float *lgam, float *scratch0, double *utt_logp,
              int32_t *utt_iters, float *pair_trans, cudaStream_t st) {
    constexpr int FB_UPB = FbCfg<SPL>::UPB;
    const int blocks = (v.n_utt + FB_UPB - 1) / FB_UPB;
    const size_t smem = FB_UPB * sizeof(PairSmem<SPL>);
    auto kern = fwdbwd_kernel<SPL, FB_GRP>;
    if (smem > 48 * 1024)
        PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<blocks, FB_UPB * 96, smem, st>>>(v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters,
                                           pair_trans);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

// block 0's phase clocks: [0] start, [1] backward done, [2] pi iteration done, [3] forward done,
// [8] helper start, [9] helper done (tuning aid, profiles/exp_k2.py)
extern "C" int pc_debug_read_fb(long long *host_out) {
    return cudaMemcpyFromSymbol(host_out, g_fb_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -2;
}

int launch_forward_backward(pc_handle h, const CorpusView &v, const float *b,
                            const double *log_self, const double *log_next, float *lgam,
                            float *scratch0, double *utt_logp, int32_t *utt_iters,
                            float *pair_trans, cudaStream_t st) {
    if (v.n_utt == 0) return PC_OK;
    if (h->k2_kernel)  // one warp per utterance (fwdbwd_warp.cu); this file's kernel stays as the cross-check
        return launch_forward_backward_warp(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters,
                                            pair_trans, st);
    const int states = PC_EMIT * v.max_labels + 1;
    if (states <= 32) return launch_fb<1>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 64) return launch_fb<2>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 128) return launch_fb<4>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    if (states <= 256) return launch_fb<8>(h, v, b, log_self, log_next, lgam, scratch0, utt_logp, utt_iters, pair_trans, st);
    pc_set_error("pc_forward_backward: %d labels per utterance exceeds the limit of 85", v.max_labels);
    return PC_ERR_UNSUPPORTED;
}
