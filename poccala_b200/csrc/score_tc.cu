// K1 (tensor-core generation): GMM scoring as a tcgen05 contraction with a fused log-sum-exp.
//
// Same arithmetic as score_simt.cu (LHMM.cal_observation_pro -> GMM.point -> gaussian_function,
// LHMM.py:163-187, Clustering.py:740-767, util.py:20-36,54-77):
//     c[t, g] = <[x_t (39), 1 | x_t^2 (39), 1], W_g>,     b[t, s] = logsumexp_{g in s} c[t, g]
// mapped on the 5th-generation tensor cores.  Both operands are split into fp16 (hi, lo) pairs
// (x = hi + lo keeps 22 significant bits) and the contraction is the error-compensated sum
//     A_hi*B_hi + A_hi*B_lo + A_lo*B_hi      (fp32 accumulation in TMEM)
// which reproduces the fp32 result (a single fp16/bf16/tf32 pass does not meet the 1e-4 parity
// bound).  Rows of W whose entries exceed the fp16 range are scaled by a power of two
// and the accumulator is scaled back in the epilogue.
//
// One persistent CTA per SM walks work items (runs of 128-frame tiles of one unit):
//   warp 8      TMA producer : cp.async.bulk of raw X rows (128 x 160 B) into a 3-stage ring
//   warps 4-7   converters   : raw fp32 rows -> [x | x^2] fp16 hi/lo operand tiles (2 stages),
//                              and once per item the unit's W rows -> resident B_hi / B_lo
//   warp 9      MMA issuer   : 15 tcgen05.mma (M=128, N=NPAD, K=16) per tile, tcgen05.commit
//   warps 0-3   epilogue     : tcgen05.ld (one frame per thread, all Gaussians in registers),
//                              log-sum-exp per state, coalesced store of b
// Shared-memory operand layout (no swizzle): 16-byte chunk c of row r at c*rows*16 + r*16, i.e.
// 8x16 B core matrices, SBO = 128 B between 8-row groups, LBO = rows*16 B between K chunks.
#include "tc_common.cuh"

namespace {

constexpr int ROWS = PC_TILE_ROWS;        // 128 frames per tile (UMMA M)
constexpr int KCH = PC_KA / 8;            // 10 sixteen-byte chunks of 8 halves along K
constexpr int RAW_STAGES = 3;
constexpr int A_STAGES = 2;
constexpr int RAW_BYTES = ROWS * PC_XS * 4;   // 20480
constexpr int A_PIECE = KCH * ROWS * 16;      // 20480 per hi / lo piece
constexpr int NTHREADS = 320;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

template <int MIX>
struct Cfg {
    static constexpr int N_REAL = PC_EMIT * MIX;
    static constexpr int NPAD = (N_REAL + 15) & ~15;
    static constexpr int B_PIECE = KCH * NPAD * 16;
    static constexpr int TM_STRIDE = NPAD <= 32 ? 32 : (NPAD <= 64 ? 64 : (NPAD <= 128 ? 128 : 256));
    static constexpr int TM_BUFS = TM_STRIDE >= 256 ? 2 : 4;
    static constexpr int TM_COLS = TM_STRIDE * TM_BUFS;
    static constexpr int SMEM = 1024 + RAW_STAGES * RAW_BYTES + A_STAGES * 2 * A_PIECE + 2 * B_PIECE +
                                3 * NPAD * 4 + 256;
};

struct Bars {
    uint64_t raw_full[RAW_STAGES], raw_empty[RAW_STAGES];
    uint64_t a_full[A_STAGES], a_empty[A_STAGES];
    uint64_t tm_full[4], tm_empty[4];
    uint32_t tmem_base;
};

template <int MIX>
__global__ void __launch_bounds__(NTHREADS, 1)
score_tc_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W,
                float *__restrict__ b) {
    using C = Cfg<MIX>;
    extern __shared__ __align__(1024) uint8_t smem[];
    Bars *bars = reinterpret_cast<Bars *>(smem);
    uint8_t *raw_s = smem + 1024;
    uint8_t *a_s = raw_s + RAW_STAGES * RAW_BYTES;
    uint8_t *b_s = a_s + A_STAGES * 2 * A_PIECE;
    float *scale_s = reinterpret_cast<float *>(b_s + 2 * C::B_PIECE);
    float *bias_s = scale_s + C::NPAD;
    uint32_t *rowmax_s = reinterpret_cast<uint32_t *>(bias_s + C::NPAD);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < RAW_STAGES; ++i) { tc::mbar_init(&bars->raw_full[i], 1); tc::mbar_init(&bars->raw_empty[i], 4); }
        for (int i = 0; i < A_STAGES; ++i) { tc::mbar_init(&bars->a_full[i], 4); tc::mbar_init(&bars->a_empty[i], 1); }
        for (int i = 0; i < 4; ++i) { tc::mbar_init(&bars->tm_full[i], 1); tc::mbar_init(&bars->tm_empty[i], 4); }
        tc::mbar_fence_init();
    }
    if (warp == 9) tc::tmem_alloc(&bars->tmem_base, C::TM_COLS);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;

    uint32_t n_raw = 0, n_a = 0, n_tm = 0;  // per-role running tile counters (stage / parity)
    for (int item = blockIdx.x; item < v.n_items; item += gridDim.x) {
        const int64_t lo = v.item_tile_lo[item], hi = v.item_tile_lo[item + 1];
        const int n_tiles = (int)(hi - lo);
        // ------------------------------------------------ resident B operand for this item's unit
        if (warp >= 4 && warp < 8) {
            const int tid = threadIdx.x - 128;
            const float *wu = W + (size_t)v.item_unit[item] * C::N_REAL * PC_KA;
            for (int n = tid; n < C::NPAD; n += 128) rowmax_s[n] = 0u;
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // pass 1: row maxima (the bit pattern of |w| orders like the value); a row whose
            // constant is not finite (alpha = 0 -> log 0) is marked dead with 0xffffffff
            for (int task = tid; task < C::N_REAL * 5; task += 128) {
                const int n = task / 5, c = task - n * 5;
                const float4 *src = reinterpret_cast<const float4 *>(wu + (size_t)n * PC_KA);
                float m = 0.f;
                bool dead = false;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float4 q0 = __ldg(src + (c + 5 * half) * 2), q1 = __ldg(src + (c + 5 * half) * 2 + 1);
                    float vals[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float a = fabsf(vals[e]);
                        if (a <= 3.0e38f) m = fmaxf(m, a);
                    }
                    if (c == 4 && half == 0 && !(fabsf(vals[7]) <= 3.0e38f)) dead = true;
                }
                atomicMax(&rowmax_s[n], dead ? 0xffffffffu : __float_as_uint(m));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            // pass 2: scale by 2^-e so that the row fits fp16, split into hi / lo, store chunks.
            // Task (n, c) converts chunks c and c+5 of row n (the constant's two columns 39 / 79
            // are both in task c = 4, so the residual of the first pair can be folded into the second).
            for (int task = tid; task < C::NPAD * 5; task += 128) {
                const int n = task / 5, c = task - n * 5;
                uint32_t hi8[2][4], lo8[2][4];
#pragma unroll
                for (int half = 0; half < 2; ++half)
#pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2) hi8[half][e2] = lo8[half][e2] = 0u;
                const uint32_t mbits = (n < C::N_REAL) ? rowmax_s[n] : 0xffffffffu;
                if (mbits != 0xffffffffu) {
                    const float mx = __uint_as_float(mbits);
                    int e = 0;
                    if (mx > 16384.f) e = (int)((mbits >> 23) & 0xff) - 127 - 13;
                    const float inv = __uint_as_float((uint32_t)(127 - e) << 23);
                    const float4 *src = reinterpret_cast<const float4 *>(wu + (size_t)n * PC_KA);
                    float kres = 0.f;  // what the first (hi, lo) pair of the constant leaves over
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float4 q0 = __ldg(src + (c + 5 * half) * 2), q1 = __ldg(src + (c + 5 * half) * 2 + 1);
                        float vals[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                        for (int e2 = 0; e2 < 8; ++e2) vals[e2] *= inv;
                        if (c == 4) {
                            if (half == 0) {
                                const __half kh = __float2half_rn(vals[7]);
                                const float r1 = vals[7] - __half2float(kh);
                                kres = r1 - __half2float(__float2half_rn(r1));
                            } else {
                                vals[7] += kres;
                            }
                        }
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2)
                            tc::split2(vals[2 * e2], vals[2 * e2 + 1], hi8[half][e2], lo8[half][e2]);
                    }
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int chunk = c + 5 * half;
                    uint4 *dh = reinterpret_cast<uint4 *>(b_s + chunk * C::NPAD * 16 + n * 16);
                    uint4 *dl = reinterpret_cast<uint4 *>(b_s + C::B_PIECE + chunk * C::NPAD * 16 + n * 16);
                    *dh = make_uint4(hi8[half][0], hi8[half][1], hi8[half][2], hi8[half][3]);
                    *dl = make_uint4(lo8[half][0], lo8[half][1], lo8[half][2], lo8[half][3]);
                }
            }
            for (int n = tid; n < C::NPAD; n += 128) {
                const uint32_t mbits = (n < C::N_REAL) ? rowmax_s[n] : 0xffffffffu;
                float sc = 1.f, bias = 0.f;
                if (mbits == 0xffffffffu) {
                    bias = PC_NEG_INF;  // dead row: weights are zero, score = log 0
                } else if (__uint_as_float(mbits) > 16384.f) {
                    const int e = (int)((mbits >> 23) & 0xff) - 127 - 13;
                    sc = __uint_as_float((uint32_t)(127 + e) << 23);
                }
                scale_s[n] = sc;
                bias_s[n] = bias;
            }
            tc::fence_proxy_async();
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
            const int r = threadIdx.x - 128;  // row of the tile
            for (int i = 0; i < n_tiles; ++i, ++n_raw, ++n_a) {
                const int rs = n_raw % RAW_STAGES, as = n_a % A_STAGES;
                const int rows = v.tile_rows[lo + i];
                tc::mbar_wait(&bars->raw_full[rs], (n_raw / RAW_STAGES) & 1);
                float x[PC_XS];
                if (r < rows) {
                    const float4 *src = reinterpret_cast<const float4 *>(raw_s + rs * RAW_BYTES + r * PC_XS * 4);
#pragma unroll
                    for (int q = 0; q < PC_XS / 4; ++q) {
                        float4 t4 = src[q];
                        x[4 * q] = t4.x; x[4 * q + 1] = t4.y; x[4 * q + 2] = t4.z; x[4 * q + 3] = t4.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < PC_XS; ++q) x[q] = 0.f;
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars->raw_empty[rs]);
                tc::mbar_wait(&bars->a_empty[as], ((n_a / A_STAGES) & 1) ^ 1);
                uint8_t *ah = a_s + as * 2 * A_PIECE, *al = ah + A_PIECE;
#pragma unroll
                for (int c = 0; c < 5; ++c) {
                    uint32_t h[4], l[4], h2[4], l2[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float p = x[8 * c + 2 * e], q = x[8 * c + 2 * e + 1];
                        tc::split2(p, q, h[e], l[e]);
                        tc::split2(p * p, q * q, h2[e], l2[e]);
                    }
                    *reinterpret_cast<uint4 *>(ah + c * ROWS * 16 + r * 16) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4 *>(al + c * ROWS * 16 + r * 16) = make_uint4(l[0], l[1], l[2], l[3]);
                    *reinterpret_cast<uint4 *>(ah + (c + 5) * ROWS * 16 + r * 16) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
                    *reinterpret_cast<uint4 *>(al + (c + 5) * ROWS * 16 + r * 16) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
                }
                tc::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars->a_full[as]);
            }
        } else if (warp == 9) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc = tc::umma_idesc_f16(ROWS, C::NPAD, 0, 0);
            const uint32_t a_base = tc::smem_u32(a_s), b_base = tc::smem_u32(b_s);
            for (int i = 0; i < n_tiles; ++i, ++n_a, ++n_tm) {
                const int as = n_a % A_STAGES, tb = n_tm % C::TM_BUFS;
                tc::mbar_wait(&bars->a_full[as], (n_a / A_STAGES) & 1);
                tc::mbar_wait(&bars->tm_empty[tb], ((n_tm / C::TM_BUFS) & 1) ^ 1);
                tc::tc_fence_after();
                if (lane == 0) {
                    const uint32_t d = tmem_base + tb * C::TM_STRIDE;
                    const uint32_t ah = a_base + as * 2 * A_PIECE, al = ah + A_PIECE;
                    const uint32_t bh = b_base, bl = b_base + C::B_PIECE;
                    uint32_t accum = 0;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {
                        const uint32_t ap = (p == 2) ? al : ah;
                        const uint32_t bp = (p == 1) ? bl : bh;
#pragma unroll
                        for (int k = 0; k < KCH / 2; ++k) {
                            const uint64_t ad = tc::umma_desc(ap + 2 * k * ROWS * 16, ROWS * 16, 128);
                            const uint64_t bd = tc::umma_desc(bp + 2 * k * C::NPAD * 16, C::NPAD * 16, 128);
                            tc::mma_f16_ss(d, ad, bd, idesc, accum);
                            accum = 1;
                        }
                    }
                    tc::tc_commit(&bars->a_empty[as]);
                    tc::tc_commit(&bars->tm_full[tb]);
                }
                __syncwarp();
            }
        } else {
            // ------------------------------------------------------------ epilogue (warps 0-3)
            const int r = threadIdx.x;  // row of the tile == TMEM lane
            for (int i = 0; i < n_tiles; ++i, ++n_tm) {
                const int tb = n_tm % C::TM_BUFS;
                const int64_t tile = lo + i;
                const int rows = v.tile_rows[tile];
                const int tp = v.tile_tp[tile];
                float *out = b + v.tile_boff[tile] + r;
                tc::mbar_wait(&bars->tm_full[tb], (n_tm / C::TM_BUFS) & 1);
                tc::tc_fence_after();
                const uint32_t taddr = tmem_base + tb * C::TM_STRIDE + ((uint32_t)(warp * 32) << 16);
                float res[PC_EMIT];
                constexpr int LD_PER_STATE = MIX >= 16 ? MIX / 16 : 0;
                float small[MIX >= 16 ? 1 : C::NPAD];
                if constexpr (MIX < 16) {
#pragma unroll
                    for (int j = 0; j < C::NPAD / 16; ++j) {
                        float t16[16];
                        tc::tmem_ld16(taddr + j * 16, t16);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 16; ++e) small[j * 16 + e] = t16[e];
                    }
                }
#pragma unroll
                for (int s = 0; s < PC_EMIT; ++s) {
                    float mx = PC_NEG_INF, sum = 0.f;
                    float vbuf[MIX];
                    if constexpr (MIX >= 16) {
#pragma unroll
                        for (int j = 0; j < LD_PER_STATE; ++j) {
                            float t16[16];
                            tc::tmem_ld16(taddr + s * MIX + j * 16, t16);
                            tc::tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e) vbuf[j * 16 + e] = t16[e];
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < MIX; ++e) vbuf[e] = small[s * MIX + e];
                    }
#pragma unroll
                    for (int e = 0; e < MIX; ++e) {
                        vbuf[e] = fmaf(vbuf[e], scale_s[s * MIX + e], bias_s[s * MIX + e]);
                        mx = fmaxf(mx, vbuf[e]);
                    }
                    if (mx == PC_NEG_INF) {
                        res[s] = PC_NEG_INF;
                    } else {
                        const float ms = mx * LOG2E;
#pragma unroll
                        for (int e = 0; e < MIX; ++e) sum += exp2f(fmaf(vbuf[e], LOG2E, -ms));
                        res[s] = mx + LN2 * log2f(sum);
                    }
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&bars->tm_empty[tb]);
                if (r < rows) {
#pragma unroll
                    for (int s = 0; s < PC_EMIT; ++s) out[(size_t)s * tp] = res[s];
                }
            }
        }
        __syncthreads();  // the item's MMAs are complete (epilogue consumed the last tile): B may change
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 9) tc::tmem_dealloc(tmem_base, C::TM_COLS);
}

template <int MIX>
int launch_mix(pc_handle h, const CorpusView &v, const float *X, const float *W, float *b,
               cudaStream_t st) {
    auto kern = score_tc_kernel<MIX>;
    PC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<MIX>::SMEM));
    int grid = v.n_items < h->sm_count ? v.n_items : h->sm_count;
    kern<<<grid, NTHREADS, Cfg<MIX>::SMEM, st>>>(v, X, W, b);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // namespace

bool score_tc_supported(int mix) { return mix == 4 || mix == 8 || mix == 16 || mix == 32 || mix == 64; }

int launch_score_tc(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                    float *b, cudaStream_t st) {
    if (v.n_items == 0) return PC_OK;
    switch (mix) {
        case 4: return launch_mix<4>(h, v, X, W, b, st);
        case 8: return launch_mix<8>(h, v, X, W, b, st);
        case 16: return launch_mix<16>(h, v, X, W, b, st);
        case 32: return launch_mix<32>(h, v, X, W, b, st);
        case 64: return launch_mix<64>(h, v, X, W, b, st);
    }
    pc_set_error("launch_score_tc: mix=%d not covered", mix);
    return PC_ERR_UNSUPPORTED;
}
