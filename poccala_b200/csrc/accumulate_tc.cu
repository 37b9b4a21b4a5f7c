// K3 (tensor-core generation): Baum-Welch sufficient statistics as two chained tcgen05
// contractions per 128-frame tile (FlashAttention-shaped, posteriors never leave the SM):
//
//   S[t, g]   = <[x_t | x_t^2], W_g>                      MMA1: frames x Gaussians, K = 80
//   P[t, g]   = exp(S[t, g] - b_t(state(g)) + lgam_t(state(g)))      (gamma_t(j, m), A.3)
//   D2[f, g] += sum_t [x_t | x_t^2][f] * P[t, g]          MMA2: features x Gaussians, K = frames
//
// Restates LHMM.update_acc -> Clustering.GMM.update_acc (LHMM.py:473-507, Clustering.py:653-680)
// in the linear-equivalent form of SURVEY A.4 (see accumulate_simt.cu).  All operands are fp16
// (hi, lo) pairs and every contraction is the 3-product error-compensated sum, fp32 accumulation
// in TMEM.  The frame tile written once by the converter warps serves BOTH contractions: as the
// K-major A operand of MMA1 (K = feature) and, through an MN-major descriptor over the same
// bytes, as the A operand of MMA2 (M = feature, K = frame).  P is written by the softmax warps as
// the MN-major B operand of MMA2.  D2 stays in TMEM for a whole work item (<= 16 tiles of one
// unit) and is flushed with fp64 atomics.
//
//   warp 8      TMA producer : raw X rows -> 2-stage ring (cp.async.bulk)
//   warps 4-7   converters   : raw rows -> operand tile; once per item the unit's W rows -> B
//   warp 9      MMA issuer   : MMA1(i), then MMA2(i-1) (software pipelined)
//   warps 0-3   softmax      : tcgen05.ld S, exp2, hi/lo split, P tile; at item end the flush
#include "tc_common.cuh"

namespace {

using tc::T_KCH;
using tc::T_PIECE;
using tc::T_ROWS;
constexpr int RAW_STAGES = 2;
constexpr int A_STAGES = 2;
constexpr int RAW_BYTES = T_ROWS * PC_XS * 4;
constexpr int NTHREADS = 320;
constexpr float LOG2E = 1.4426950408889634f;
// posteriors are stored as P * 2^15 so that the fp16 window [6e-8, 65504] covers [1.8e-12, 2]
constexpr float P_SHIFT = 15.f;
constexpr float P_UNSHIFT = 1.f / 32768.f;

// NC = Gaussians handled per work item (a unit's 3*MIX Gaussians, or a slice of them)
template <int MIX>
struct Cfg {
    static constexpr int N_UNIT = PC_EMIT * MIX;
    static constexpr int NC = N_UNIT <= 96 ? N_UNIT : 96;
    static constexpr int N_SLICES = (N_UNIT + NC - 1) / NC;
    static constexpr int NPAD = (NC + 15) & ~15;
    static constexpr int B_PIECE = T_KCH * NPAD * 16;
    static constexpr int P_PIECE = (NPAD / 8) * T_ROWS * 16;
    static constexpr int P_STAGES = NPAD <= 48 ? 2 : 1;
    static constexpr int S_STRIDE = NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : 128);
    static constexpr int S_BUFS = 2;
    static constexpr int D2_COL = S_STRIDE * S_BUFS;
    static constexpr int TM_COLS = 512;
    static constexpr int SMEM = 1024 + RAW_STAGES * RAW_BYTES + A_STAGES * 2 * T_PIECE + 2 * B_PIECE +
                                P_STAGES * 2 * P_PIECE + 3 * NPAD * 4 + 256;
    static_assert(N_UNIT % NC == 0, "slices must tile the unit");
    static_assert(D2_COL + NPAD <= 512, "TMEM budget");
};

struct Bars {
    uint64_t raw_full[RAW_STAGES], raw_empty[RAW_STAGES];
    uint64_t a_full[A_STAGES], a_empty[A_STAGES];
    uint64_t s_full[2], s_empty[2];
    uint64_t p_full[2], p_empty[2];
    uint64_t d2_full;
    uint32_t tmem_base;
};

// S (TMEM, one frame per thread) -> P = exp2(S * scale * log2e + d) -> fp16 hi / lo rows of the
// MN-major P tile.  G0 = first Gaussian of the slice inside the unit (selects the state of a column
// at compile time).
template <int MIX, int G0>
__device__ __forceinline__ void softmax_tile(uint32_t taddr, const float (&dl)[PC_EMIT],
                                             const float *__restrict__ scale_s, uint8_t *ph,
                                             uint8_t *pl, int r) {
    using C = Cfg<MIX>;
#pragma unroll
    for (int j = 0; j < C::NPAD / 16; ++j) {
        float t16[16];
        tc::tmem_ld16(taddr + j * 16, t16);
        tc::tmem_ld_wait();
        uint32_t h[8], l[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c0 = j * 16 + 2 * e, c1 = c0 + 1;
            constexpr int SMAX = PC_EMIT - 1;
            const int s0 = (G0 + c0) / MIX < SMAX ? (G0 + c0) / MIX : SMAX;
            const int s1 = (G0 + c1) / MIX < SMAX ? (G0 + c1) / MIX : SMAX;
            float p0 = exp2f(fmaf(t16[2 * e], scale_s[c0] * LOG2E, dl[s0]));
            float p1 = exp2f(fmaf(t16[2 * e + 1], scale_s[c1] * LOG2E, dl[s1]));
            if (c0 >= C::NC) p0 = 0.f;
            if (c1 >= C::NC) p1 = 0.f;
            tc::split2(p0, p1, h[e], l[e]);
        }
        // columns 16j..16j+7 -> Gaussian block 2j, 16j+8..16j+15 -> block 2j+1
        *reinterpret_cast<uint4 *>(ph + (2 * j) * T_ROWS * 16 + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(pl + (2 * j) * T_ROWS * 16 + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
        *reinterpret_cast<uint4 *>(ph + (2 * j + 1) * T_ROWS * 16 + r * 16) = make_uint4(h[4], h[5], h[6], h[7]);
        *reinterpret_cast<uint4 *>(pl + (2 * j + 1) * T_ROWS * 16 + r * 16) = make_uint4(l[4], l[5], l[6], l[7]);
    }
}

template <int MIX>
__global__ void __launch_bounds__(NTHREADS, 1)
accumulate_tc_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W,
                     const float *__restrict__ b, const float *__restrict__ lgam,
                     double *__restrict__ acc) {
    using C = Cfg<MIX>;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars *bars = reinterpret_cast<Bars *>(smem);
    uint8_t *raw_s = smem + 1024;
    uint8_t *a_s = raw_s + RAW_STAGES * RAW_BYTES;
    uint8_t *b_s = a_s + A_STAGES * 2 * T_PIECE;
    uint8_t *p_s = b_s + 2 * C::B_PIECE;
    float *scale_s = reinterpret_cast<float *>(p_s + C::P_STAGES * 2 * C::P_PIECE);
    float *bias_s = scale_s + C::NPAD;
    uint32_t *rowmax_s = reinterpret_cast<uint32_t *>(bias_s + C::NPAD);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < RAW_STAGES; ++i) { tc::mbar_init(&bars->raw_full[i], 1); tc::mbar_init(&bars->raw_empty[i], 4); }
        for (int i = 0; i < A_STAGES; ++i) { tc::mbar_init(&bars->a_full[i], 4); tc::mbar_init(&bars->a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&bars->s_full[i], 1); tc::mbar_init(&bars->s_empty[i], 4);
            tc::mbar_init(&bars->p_full[i], 4); tc::mbar_init(&bars->p_empty[i], 1);
        }
        tc::mbar_init(&bars->d2_full, 1);
        tc::mbar_fence_init();
    }
    if (warp == 9) tc::tmem_alloc(&bars->tmem_base, C::TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    uint32_t n_raw = 0, n_a = 0, n_s = 0, n_p = 0, n_item = 0;
    const int n_work = v.n_items * C::N_SLICES;
    for (int work = blockIdx.x; work < n_work; work += gridDim.x, ++n_item) {
        const int item = work / C::N_SLICES, slice = work - item * C::N_SLICES;
        const int unit = v.item_unit[item];
        const int g0 = slice * C::NC;  // first Gaussian of the slice inside the unit
        const int64_t lo = v.item_tile_lo[item], hi = v.item_tile_lo[item + 1];
        const int n_tiles = (int)(hi - lo);
        if (warp >= 4 && warp < 8) {
            tc::load_gauss_operand<C::NPAD>(W + ((size_t)unit * C::N_UNIT + g0) * PC_KA, C::NC,
                                            threadIdx.x - 128, b_s, b_s + C::B_PIECE, scale_s, bias_s,
                                            rowmax_s);
        }
        __syncthreads();

        if (warp == 8) {
            // ------------------------------------------------------------ TMA producer
            for (int i = 0; i < n_tiles; ++i, ++n_raw) {
                const int s = n_raw % RAW_STAGES;
                tc::mbar_wait(&bars->raw_empty[s], ((n_raw / RAW_STAGES) & 1) ^ 1);
                if (lane == 0) {
                    const int64_t tile = lo + i;
                    const uint32_t bytes = (uint32_t)v.tile_rows[tile] * PC_XS * 4;
                    tc::mbar_expect_tx(&bars->raw_full[s], bytes);
                    tc::tma_load_1d(raw_s + s * RAW_BYTES, X + (size_t)v.tile_xrow[tile] * PC_XS, bytes,
                                    &bars->raw_full[s]);
                }
                __syncwarp();
            }
        } else if (warp >= 4 && warp < 8) {
            // ------------------------------------------------------------ converters
            const int r = threadIdx.x - 128;
            for (int i = 0; i < n_tiles; ++i, ++n_raw, ++n_a) {
                const int rs = n_raw % RAW_STAGES, as = n_a % A_STAGES;
                const int rows = v.tile_rows[lo + i];
                tc::mbar_wait(&bars->raw_full[rs], (n_raw / RAW_STAGES) & 1);
                tc::mbar_wait(&bars->a_empty[as], ((n_a / A_STAGES) & 1) ^ 1);
                uint8_t *ah = a_s + as * 2 * T_PIECE;
                tc::convert_frame_row(raw_s + rs * RAW_BYTES, r, r < rows, ah, ah + T_PIECE);
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&bars->raw_empty[rs]);
                    tc::mbar_arrive(&bars->a_full[as]);
                }
            }
        } else if (warp == 9) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc1 = tc::umma_idesc_f16(T_ROWS, C::NPAD, 0, 0);
            constexpr uint32_t idesc2 = tc::umma_idesc_f16(128, C::NPAD, 1, 1);
            const uint32_t a_base = tc::smem_u32(a_s), b_base = tc::smem_u32(b_s), p_base = tc::smem_u32(p_s);
            const uint32_t d2 = tmem_base + C::D2_COL;
            uint32_t n_a2 = n_a, n_p2 = n_p;  // counters of the lagging MMA2 stream
            for (int i = 0; i <= n_tiles; ++i) {
                if (i < n_tiles) {
                    const int as = n_a % A_STAGES, sb = n_s % C::S_BUFS;
                    tc::mbar_wait(&bars->a_full[as], (n_a / A_STAGES) & 1);
                    tc::mbar_wait(&bars->s_empty[sb], ((n_s / C::S_BUFS) & 1) ^ 1);
                    tc::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t d = tmem_base + sb * C::S_STRIDE;
                        const uint32_t ah = a_base + as * 2 * T_PIECE, al = ah + T_PIECE;
                        const uint32_t bh = b_base, bl = b_base + C::B_PIECE;
                        uint32_t accum = 0;
#pragma unroll
                        for (int p = 0; p < 3; ++p) {
                            const uint32_t ap = (p == 2) ? al : ah;
                            const uint32_t bp = (p == 1) ? bl : bh;
#pragma unroll
                            for (int k = 0; k < T_KCH / 2; ++k) {
                                const uint64_t ad = tc::umma_desc(ap + 2 * k * T_ROWS * 16, T_ROWS * 16, 128);
                                const uint64_t bd = tc::umma_desc(bp + 2 * k * C::NPAD * 16, C::NPAD * 16, 128);
                                tc::mma_f16_ss(d, ad, bd, idesc1, accum);
                                accum = 1;
                            }
                        }
                        tc::tc_commit(&bars->s_full[sb]);
                    }
                    __syncwarp();
                    ++n_a;
                    ++n_s;
                }
                if (i >= 1) {
                    // MMA2 of tile i-1: D2[f, g] += sum_t A[t, f] * P[t, g]
                    const int as = n_a2 % A_STAGES, ps = n_p2 % C::P_STAGES;
                    tc::mbar_wait(&bars->p_full[ps], (n_p2 / C::P_STAGES) & 1);
                    tc::tc_fence_after();
                    if (lane == 0) {
                        const uint32_t ah = a_base + as * 2 * T_PIECE, al = ah + T_PIECE;
                        const uint32_t ph = p_base + ps * 2 * C::P_PIECE, pl = ph + C::P_PIECE;
                        uint32_t accum = (i == 1) ? 0u : 1u;  // first tile of the item resets D2
#pragma unroll
                        for (int p = 0; p < 3; ++p) {
                            const uint32_t ap = (p == 1) ? al : ah;
                            const uint32_t pp = (p == 2) ? pl : ph;
#pragma unroll
                            for (int k = 0; k < T_ROWS / 16; ++k) {
                                // MN-major views: 8 frames x 16 B core matrices; LBO = 128 B between
                                // 8-frame groups, SBO = 2048 B between 8-feature / 8-Gaussian blocks
                                const uint64_t ad = tc::umma_desc(ap + k * 256, 128, T_ROWS * 16);
                                const uint64_t bd = tc::umma_desc(pp + k * 256, 128, T_ROWS * 16);
                                tc::mma_f16_ss(d2, ad, bd, idesc2, accum);
                                accum = 1;
                            }
                        }
                        tc::tc_commit(&bars->a_empty[as]);
                        tc::tc_commit(&bars->p_empty[ps]);
                        if (i == n_tiles) tc::tc_commit(&bars->d2_full);
                    }
                    __syncwarp();
                    ++n_a2;
                    ++n_p2;
                }
            }
            n_p = n_p2;
        } else {
            // ------------------------------------------------------------ softmax warps 0-3
            const int r = threadIdx.x;
            for (int i = 0; i < n_tiles; ++i, ++n_s, ++n_p) {
                const int sb = n_s % C::S_BUFS, ps = n_p % C::P_STAGES;
                const int64_t tile = lo + i;
                const int rows = v.tile_rows[tile];
                const int tp = v.tile_tp[tile];
                // d = (lgam - b) * log2(e) for the states this slice touches; -inf kills the row
                float dl[PC_EMIT];
                {
                    const size_t o = (size_t)v.tile_boff[tile] + r;
#pragma unroll
                    for (int s = 0; s < PC_EMIT; ++s) {
                        float d = PC_NEG_INF;
                        if (r < rows) {
                            const float lg = __ldg(lgam + o + (size_t)s * tp), bb = __ldg(b + o + (size_t)s * tp);
                            d = (lg == PC_NEG_INF) ? PC_NEG_INF : fmaf(lg - bb, LOG2E, P_SHIFT);
                        }
                        dl[s] = d;
                    }
                }
                tc::mbar_wait(&bars->s_full[sb], (n_s / C::S_BUFS) & 1);
                tc::tc_fence_after();
                tc::mbar_wait(&bars->p_empty[ps], ((n_p / C::P_STAGES) & 1) ^ 1);
                const uint32_t taddr = tmem_base + sb * C::S_STRIDE + ((uint32_t)(warp * 32) << 16);
                uint8_t *ph = p_s + ps * 2 * C::P_PIECE, *pl = ph + C::P_PIECE;
                if (C::N_SLICES == 1 || slice == 0)
                    softmax_tile<MIX, 0>(taddr, dl, scale_s, ph, pl, r);
                else
                    softmax_tile<MIX, C::NC>(taddr, dl, scale_s, ph, pl, r);
                tc::tc_fence_before();
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tc::mbar_arrive(&bars->s_empty[sb]);
                    tc::mbar_arrive(&bars->p_full[ps]);
                }
            }
            // ---------------------------------------------------- flush D2 (lane = feature)
            tc::mbar_wait(&bars->d2_full, n_item & 1);
            tc::tc_fence_after();
            if (warp < 3) {
                const int f = threadIdx.x;
                const uint32_t taddr = tmem_base + C::D2_COL + ((uint32_t)(warp * 32) << 16);
                double *dst = acc + ((size_t)unit * C::N_UNIT + g0) * PC_KA + f;
#pragma unroll
                for (int j = 0; j < C::NPAD / 16; ++j) {
                    float t16[16];
                    tc::tmem_ld16(taddr + j * 16, t16);
                    tc::tmem_ld_wait();
                    if (f < PC_KA) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const int g = j * 16 + e;
                            if (g < C::NC && bias_s[g] == 0.f) atomicAdd(dst + (size_t)g * PC_KA, (double)(t16[e] * P_UNSHIFT));
                        }
                    }
                }
            }
            tc::tc_fence_before();
        }
        __syncthreads();
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 9) tc::tmem_dealloc(tmem_base, C::TM_COLS);
}

template <int MIX>
int launch_mix(pc_handle h, const CorpusView &v, const float *X, const float *W, const float *b,
               const float *lgam, double *acc, cudaStream_t st) {
    auto kern = accumulate_tc_kernel<MIX>;
    PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MIX>::SMEM));
    const int n_work = v.n_items * Cfg<MIX>::N_SLICES;
    const int grid = n_work < h->sm_count ? n_work : h->sm_count;
    kern<<<grid, NTHREADS, Cfg<MIX>::SMEM, st>>>(v, X, W, b, lgam, acc);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

bool accumulate_tc_supported(int mix) { return mix == 4 || mix == 8 || mix == 16 || mix == 32 || mix == 64; }

int launch_accumulate_tc(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                         const float *b, const float *lgam, double *acc, cudaStream_t st) {
    if (v.n_items == 0) return PC_OK;
    switch (mix) {
        case 4: return launch_mix<4>(h, v, X, W, b, lgam, acc, st);
        case 8: return launch_mix<8>(h, v, X, W, b, lgam, acc, st);
        case 16: return launch_mix<16>(h, v, X, W, b, lgam, acc, st);
        case 32: return launch_mix<32>(h, v, X, W, b, lgam, acc, st);
        case 64: return launch_mix<64>(h, v, X, W, b, lgam, acc, st);
    }
    pc_set_error("launch_accumulate_tc: mix=%d not covered", mix);
    return PC_ERR_UNSUPPORTED;
}
