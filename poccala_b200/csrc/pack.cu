// Model packing and frame preparation: the HBM-resident operand formats.
//
// pack_gmm: util.gaussian_function's log branch (util.py:20-36) plus the log(alpha) term of
// Clustering.GMM.point (Clustering.py:753-757) folded into one row per Gaussian:
//     score(x, g) = <[x' (39), 1 | x'^2 (39), 1], W_g>,   x' = (x - shift) * inv_scale
//     W_g = [ mu'/var' (39), k_hi | -1/(2 var') (39), k_lo ]
//     k   = log alpha - D/2 log 2pi - 1/2 sum_d var_d (Q1: ORIGINAL variances, not log-det)
//           - 1/2 sum_d mu'^2/var'
// All arithmetic in fp64.  W buffer layout (G Gaussians, U = G / (3*mix) units):
//     [0, 320 G)              fp32 rows [G][80]                         (CUDA-core kernels)
//     [320 G, 324 G)          float scale[G]
//     [324 G, 328 G)          int32 flags[G]   (flags[0] != 0 <=> some row has scale != 1)
//     [pc_w16_offset(G), ..)  per unit one tcgen05 operand image: fp16 [NPAD/8 row groups][2 (hi, lo)]
//                             [10 chunks][8 rows][8 halves], NPAD = 3*mix rounded up to 16 (padding
//                             rows are zero).  Every [8 rows][16 B] block is one UMMA core matrix
//                             (LBO = 128 B between K chunks, SBO = 2560 B between row groups), units
//                             and slices of units concatenate along N, and a unit is ONE cp.async.bulk.
//     [pc_w16s_offset(G, 3*mix), ..)  the same images with hi and lo as two contiguous pieces per unit
//                             ([NPAD/8 row groups][10 chunks][8 rows][8 halves] each, SBO = 1280 B)
// A row is scaled by 2^-e so that it fits fp16 (scale = 2^e undoes it in the epilogue); a row whose
// constant is not finite (alpha = 0 -> log 0) is stored as zero weights with the constant -60000,
// which underflows to probability 0 against any live component (the fp32 row keeps -inf).
//
// prepare_frames: X buffer = fp32 rows [F][40] (cols [0,D) standardised data, col 39 = 1), then at
// pc_x16_offset(F) one operand image per 128-frame tile of every utterance:
// fp16 [2 (hi, lo)][10 chunks][128 rows][8] of the augmented row [x | x^2] (chunks 0-4 = x, 5-9 = x^2),
// rows past the end of the utterance zero: a tile is ONE 40 KiB cp.async.bulk.  Behind those, at
// pc_x32_offset(F, tiles), the same tiles as fp32, float [4 blocks][10 quads][32 rows][4] (20 KiB, padding rows
// zero): what the scoring kernel for narrow units and the accumulation kernel stream and convert in shared
// memory (conflict-free 16-byte reads, one row per thread); a 32-row block is one contiguous 5 KiB piece, the
// unit in which the accumulation kernel gathers the frames that carry posterior mass.
#include <cuda_fp16.h>

#include "common.cuh"

__device__ __forceinline__ void split_h(float a, __half &hi, __half &lo) {
    hi = __float2half_rn(a);
    lo = __float2half_rn(a - __half2float(hi));
}

// One warp per (unit, padded row); lane l owns elements l, l+32, l+64 of the 80-element row.
__global__ void pack_gmm_kernel(const double *__restrict__ mean, const double *__restrict__ var,
                                const double *__restrict__ alpha, const double *__restrict__ shift,
                                const double *__restrict__ inv_scale, int n_gauss, int dim, int n_unit,
                                float *__restrict__ W) {
    const int npad = (n_unit + 15) & ~15;
    const int n_units = n_gauss / n_unit;
    const int lane = threadIdx.x & 31;
    const int slot = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);  // (unit, padded row)
    if (slot >= n_units * npad) return;
    const int unit = slot / npad, row = slot - unit * npad;
    uint8_t *img = reinterpret_cast<uint8_t *>(W) + pc_w16_offset(n_gauss) + (size_t)unit * 2 * (PC_KA / 8) * npad * 16;
    uint8_t *grp = img + (size_t)(row >> 3) * PC_WGROUP_BYTES + (row & 7) * 16;
    // the split image set: hi piece, then lo piece, each [npad / 8 row groups][10 chunks][8 rows][8 halves]
    const size_t piece = (size_t)(PC_KA / 8) * npad * 16;
    uint8_t *grp_s = reinterpret_cast<uint8_t *>(W) + pc_w16s_offset(n_gauss, n_unit) + (size_t)unit * 2 * piece +
                     (size_t)(row >> 3) * (PC_WGROUP_BYTES / 2) + (row & 7) * 16;
    if (row >= n_unit) {  // padding row of the operand images
        if (lane < 2 * (PC_KA / 8)) {
            *reinterpret_cast<uint4 *>(grp + (lane / (PC_KA / 8)) * (PC_WGROUP_BYTES / 2) + (lane % (PC_KA / 8)) * 128) =
                make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(grp_s + (lane / (PC_KA / 8)) * piece + (lane % (PC_KA / 8)) * 128) =
                make_uint4(0u, 0u, 0u, 0u);
        }
        return;
    }
    const int g = unit * n_unit + row;
    const double LOG_2PI = 1.8378770664093453;  // np.log(2*pi), util.py:14
    // per-dimension terms: lane handles d = lane and d = lane + 32
    double wx[2] = {0.0, 0.0}, wq[2] = {0.0, 0.0}, sum_var = 0.0, quad = 0.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int d = lane + 32 * h;
        if (d < dim) {
            const double mu = mean[(size_t)g * dim + d];
            const double v = var[(size_t)g * dim + d];
            const double sh = shift ? shift[d] : 0.0;
            const double is = inv_scale ? inv_scale[d] : 1.0;
            const double mu_s = (mu - sh) * is;
            const double v_s = v * is * is;
            sum_var += v;
            quad += mu_s * mu_s / v_s;
            wx[h] = (double)(float)(mu_s / v_s);
            wq[h] = (double)(float)(-0.5 / v_s);
        }
    }
    sum_var = warp_sum_d(sum_var);
    quad = warp_sum_d(quad);
    const double k = log(alpha[g]) - 0.5 * dim * LOG_2PI - 0.5 * sum_var - 0.5 * quad;
    const float k_hi = (float)k;
    const float k_lo = isfinite(k) ? (float)(k - (double)k_hi) : 0.f;
    const bool dead = !isfinite(k);
    // the row's largest finite magnitude decides the power-of-two scale that fits it into fp16
    float mx = 0.f;
    {
        const float cand[6] = {(float)wx[0], (float)wx[1], (float)wq[0], (float)wq[1], k_hi, k_lo};
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const float a = fabsf(cand[i]);
            if (a <= 3.0e38f) mx = fmaxf(mx, a);
        }
        mx = warp_max(mx);
    }
    int e = 0;
    if (mx > 16384.f) e = (int)((__float_as_uint(mx) >> 23) & 0xff) - 127 - 13;
    const double inv = ldexp(1.0, -e);
    float *scale = W + (size_t)n_gauss * PC_KA;
    int *flags = reinterpret_cast<int *>(scale + n_gauss);
    if (lane == 0) {
        scale[g] = dead ? 1.f : (float)ldexp(1.0, e);
        if (!dead && e != 0) atomicOr(flags, 1);
    }
    // the constant keeps 4 fp16 pieces: (hi, lo) of k in column 39, (hi, lo) of the rest in column 79
    __half k1h = __float2half_rn(-60000.f), k1l = __float2half_rn(0.f), k2h = k1l, k2l = k1l;
    if (!dead) {
        const double ks = k * inv;
        k1h = __float2half_rn((float)ks);
        double r = ks - (double)__half2float(k1h);
        k1l = __float2half_rn((float)r);
        r -= (double)__half2float(k1l);
        k2h = __float2half_rn((float)r);
        r -= (double)__half2float(k2h);
        k2l = __float2half_rn((float)r);
    }
    float *wrow = W + (size_t)g * PC_KA;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
        const int i = lane + 32 * t;  // element of the 80-element row
        const bool valid = i < PC_KA;
        const int ii = valid ? i : 0;
        const int d = (ii < PC_XS) ? ii : ii - PC_XS;  // 39 for the two constant columns: a zero lane
        // element d lives in lane d % 32, half d / 32 (all lanes take part in the shuffles)
        const double vx0 = __shfl_sync(0xffffffffu, wx[0], d & 31), vx1 = __shfl_sync(0xffffffffu, wx[1], d & 31);
        const double vq0 = __shfl_sync(0xffffffffu, wq[0], d & 31), vq1 = __shfl_sync(0xffffffffu, wq[1], d & 31);
        const double vx = (d < 32) ? vx0 : vx1, vq = (d < 32) ? vq0 : vq1;
        float wf = (d < dim) ? (float)((ii < PC_XS) ? vx : vq) : 0.f;  // fp32 row (CUDA-core kernels)
        __half hi, lo;
        if (dead) hi = lo = __float2half_rn(0.f);
        else split_h((float)((double)wf * inv), hi, lo);
        if (ii == PC_XS - 1) { wf = k_hi; hi = k1h; lo = k1l; }
        if (ii == PC_KA - 1) { wf = k_lo; hi = k2h; lo = k2l; }
        if (valid) {
            wrow[i] = wf;
            const int c = i >> 3, j = i & 7;
            reinterpret_cast<__half *>(grp + c * 128)[j] = hi;
            reinterpret_cast<__half *>(grp + PC_WGROUP_BYTES / 2 + c * 128)[j] = lo;
            reinterpret_cast<__half *>(grp_s + c * 128)[j] = hi;
            reinterpret_cast<__half *>(grp_s + piece + c * 128)[j] = lo;
        }
    }
}

// fp32 rows only (dense scoring sweep; no corpus): one thread per (frame, 8-feature chunk)
template <typename T>
__global__ void prepare_rows_kernel(const T *__restrict__ x, int64_t n, int dim,
                                    const double *__restrict__ shift,
                                    const double *__restrict__ inv_scale, float *__restrict__ X,
                                    int *__restrict__ clamped) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 5) return;
    int n_clamped = 0;
    const int64_t f = i / 5;
    const int c = (int)(i - f * 5);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = 8 * c + j;
        float out;
        if (d < dim) {
            double val = (double)x[f * dim + d];
            out = (float)((val - (shift ? shift[d] : 0.0)) * (inv_scale ? inv_scale[d] : 1.0));
            if (!(fabsf(out) <= 240.f)) ++n_clamped;  // reported through option "clamped"
            out = fminf(fmaxf(out, -240.f), 240.f);
        } else {
            out = (d == PC_XS - 1) ? 1.f : 0.f;
        }
        v[j] = out;
    }
    float4 *row = reinterpret_cast<float4 *>(X + (size_t)f * PC_XS + 8 * c);
    row[0] = make_float4(v[0], v[1], v[2], v[3]);
    row[1] = make_float4(v[4], v[5], v[6], v[7]);
    if (n_clamped) atomicAdd(clamped, n_clamped);
}

// corpus frames: one thread per (tile image, row, 8-feature chunk)
template <typename T>
__global__ void prepare_frames_kernel(CorpusView cv, const T *__restrict__ x, int dim,
                                      const double *__restrict__ shift,
                                      const double *__restrict__ inv_scale, float *__restrict__ X,
                                      int64_t xtile_lo, int64_t xtile_hi, int *__restrict__ clamped) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (xtile_hi - xtile_lo) * PC_TILE_ROWS * 5) return;
    const int64_t blk = xtile_lo + i / (PC_TILE_ROWS * 5);
    const int rem = (int)(i % (PC_TILE_ROWS * 5));
    const int c = rem / PC_TILE_ROWS, r = rem - c * PC_TILE_ROWS;  // consecutive threads -> consecutive rows
    const int u = cv.xtile_utt[blk];
    const int64_t f0 = cv.frame_off[u];
    const int T_u = (int)(cv.frame_off[u + 1] - f0);
    const int t = cv.xtile_t0[blk] + r;
    const bool valid = t < T_u;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = 8 * c + j;
        float out = 0.f;
        if (valid) {
            if (d < dim) {
                double val = (double)x[(f0 + t) * dim + d];
                out = (float)((val - (shift ? shift[d] : 0.0)) * (inv_scale ? inv_scale[d] : 1.0));
                // keep x^2 inside the fp16 range of the tensor-core operands (|x'| is in standard
                // deviations when the engine standardises, so this never triggers on sane data; when it
                // does the caller hears about it: option "clamped", PC_ERR_INVALID from the host entry point)
                if (!(fabsf(out) <= 240.f)) atomicAdd(clamped, 1);
                out = fminf(fmaxf(out, -240.f), 240.f);
            } else {
                out = (d == PC_XS - 1) ? 1.f : 0.f;
            }
        }
        v[j] = out;
    }
    if (valid) {
        float4 *row = reinterpret_cast<float4 *>(X + (size_t)(f0 + t) * PC_XS + 8 * c);
        row[0] = make_float4(v[0], v[1], v[2], v[3]);
        row[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    {
        float4 *q32 = reinterpret_cast<float4 *>(reinterpret_cast<uint8_t *>(X) + pc_x32_offset(cv.total_frames, cv.n_xtiles) +
                                                 (size_t)blk * PC_X32TILE_BYTES);
        // [4 blocks of 32 rows][10 quads][32 rows][4 floats]: a 32-row block is one contiguous 5 KB piece
        float4 *blk32 = q32 + (r / PC_BLOCK_ROWS) * (PC_BLOCK_ROWS * PC_XS / 4) + (r % PC_BLOCK_ROWS);
        blk32[(2 * c) * PC_BLOCK_ROWS] = make_float4(v[0], v[1], v[2], v[3]);
        blk32[(2 * c + 1) * PC_BLOCK_ROWS] = make_float4(v[4], v[5], v[6], v[7]);
    }
    uint8_t *img = reinterpret_cast<uint8_t *>(X) + pc_x16_offset(cv.total_frames) + (size_t)blk * PC_XTILE_BYTES;
    __half xh[8], xl[8], qh[8], ql[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        split_h(v[j], xh[j], xl[j]);
        split_h(v[j] * v[j], qh[j], ql[j]);
    }
    constexpr int CH = PC_TILE_ROWS * 16;   // bytes per chunk
    constexpr int PIECE = (PC_KA / 8) * CH;  // bytes per piece
    *reinterpret_cast<uint4 *>(img + (size_t)c * CH + r * 16) = *reinterpret_cast<uint4 *>(xh);
    *reinterpret_cast<uint4 *>(img + (size_t)(c + 5) * CH + r * 16) = *reinterpret_cast<uint4 *>(qh);
    *reinterpret_cast<uint4 *>(img + PIECE + (size_t)c * CH + r * 16) = *reinterpret_cast<uint4 *>(xl);
    *reinterpret_cast<uint4 *>(img + PIECE + (size_t)(c + 5) * CH + r * 16) = *reinterpret_cast<uint4 *>(ql);
}

int launch_pack_gmm(pc_handle h, const double *mean, const double *var, const double *alpha,
                    const double *shift, const double *inv_scale, int n_gauss, int dim, int mix,
                    float *W, cudaStream_t st) {
    if (n_gauss == 0) return PC_OK;
    const int n_unit = mix > 0 ? PC_EMIT * mix : n_gauss;  // mix = 0: flat list = one pseudo unit
    const int npad = (n_unit + 15) & ~15;
    const int slots = (n_gauss / n_unit) * npad;
    const int threads = 128;  // one warp per slot
    const int blocks = (int)(((int64_t)slots * 32 + threads - 1) / threads);
    PC_CUDA_TRY(cudaMemsetAsync(reinterpret_cast<char *>(W) + (size_t)n_gauss * 324, 0, sizeof(int), st));
    pack_gmm_kernel<<<blocks, threads, 0, st>>>(mean, var, alpha, shift, inv_scale, n_gauss, dim, n_unit, W);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int launch_prepare_rows(pc_handle h, const void *x, int is_f64, int64_t n, int dim,
                        const double *shift, const double *inv_scale, float *X, cudaStream_t st) {
    if (n == 0) return PC_OK;
    const int threads = 256;
    const int64_t blocks = (n * 5 + threads - 1) / threads;
    if (blocks > 2147483647LL) {
        pc_set_error("prepare_rows: too many frames for one launch");
        return PC_ERR_UNSUPPORTED;
    }
    if (is_f64)
        prepare_rows_kernel<double><<<(unsigned)blocks, threads, 0, st>>>((const double *)x, n, dim, shift, inv_scale, X, h->dev_counters + PC_CNT_CLAMPED);
    else
        prepare_rows_kernel<float><<<(unsigned)blocks, threads, 0, st>>>((const float *)x, n, dim, shift, inv_scale, X, h->dev_counters + PC_CNT_CLAMPED);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int launch_prepare_frames(pc_handle h, const CorpusView &cv, const void *x, int is_f64, int dim,
                          const double *shift, const double *inv_scale, float *X, int64_t xtile_lo,
                          int64_t xtile_hi, cudaStream_t st) {
    if (xtile_hi <= xtile_lo) return PC_OK;
    const int threads = 256;
    const int64_t blocks = ((xtile_hi - xtile_lo) * PC_TILE_ROWS * 5 + threads - 1) / threads;
    if (blocks > 2147483647LL) {
        pc_set_error("prepare_frames: too many frames for one launch");
        return PC_ERR_UNSUPPORTED;
    }
    if (is_f64)
        prepare_frames_kernel<double><<<(unsigned)blocks, threads, 0, st>>>(cv, (const double *)x, dim, shift, inv_scale, X, xtile_lo, xtile_hi, h->dev_counters + PC_CNT_CLAMPED);
    else
        prepare_frames_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(cv, (const float *)x, dim, shift, inv_scale, X, xtile_lo, xtile_hi, h->dev_counters + PC_CNT_CLAMPED);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}
