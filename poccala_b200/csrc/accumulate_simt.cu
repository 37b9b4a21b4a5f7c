// K3 (CUDA-core generation): Baum-Welch sufficient statistics.
//
// Restates LHMM.update_acc -> Clustering.GMM.update_acc (LHMM.py:473-507, Clustering.py:653-680)
// in the linear-equivalent form of SURVEY A.4:
//     gamma_t(j,m) = exp(lgam_t(j) + c_t(j,m) - b_t(j)),   acc[g] += gamma_t(j,m) * [x, x^2, 1, 1]
// (the reference keeps log-domain sums with a +100 bias on x and centres the variance on the old
// mean; both are algebraic rewrites of these three moments, applied in the M-step).
// The component scores c are recomputed from X and W instead of being stored.
// One CTA per work item; thread (g, fl) owns Gaussian g of the item's unit and every FL-th frame;
// W_g and the 80 partial sums live in registers, the augmented frame tile in shared memory.
// Partial sums are flushed with fp64 atomics once per item.
#include "common.cuh"

#define ACC_MAX_THREADS 256

__global__ void __launch_bounds__(ACC_MAX_THREADS)
accumulate_simt_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W,
                       int mix, int frame_lanes, const float *__restrict__ b,
                       const float *__restrict__ lgam, double *__restrict__ acc) {
    __shared__ __align__(16) float xa_s[PC_TILE_ROWS][PC_KA];
    __shared__ float d_s[PC_EMIT][PC_TILE_ROWS];  // lgam - b per (state, frame)
    const int item = blockIdx.x;
    const int unit = v.item_unit[item];
    const int n_g = PC_EMIT * mix;
    const int g = threadIdx.x % n_g;
    const int fl = threadIdx.x / n_g;
    const int r = g / mix;
    float w[PC_KA], a[PC_KA];
    {
        const float4 *src =
            reinterpret_cast<const float4 *>(W + ((size_t)unit * n_g + g) * PC_KA);
#pragma unroll
        for (int i = 0; i < PC_KA / 4; ++i) {
            float4 q = __ldg(src + i);
            w[4 * i] = q.x; w[4 * i + 1] = q.y; w[4 * i + 2] = q.z; w[4 * i + 3] = q.w;
        }
    }
#pragma unroll
    for (int k = 0; k < PC_KA; ++k) a[k] = 0.f;

    const int64_t lo = v.item_tile_lo[item], hi = v.item_tile_lo[item + 1];
    for (int64_t tile = lo; tile < hi; ++tile) {
        const int64_t pair = v.tile_pair[tile];
        const int u = v.pair_utt[pair];
        const int pos = (int)(pair - v.pair_off[u]);
        const int T = (int)(v.frame_off[u + 1] - v.frame_off[u]);
        const int t0 = v.tile_t0[tile];
        const int rows = min(PC_TILE_ROWS, T - t0);
        const int sp = pc_spad((int)(v.pair_off[u + 1] - v.pair_off[u]));
        __syncthreads();  // previous tile fully consumed
        for (int i = threadIdx.x; i < rows * PC_XS; i += blockDim.x) {
            int f = i / PC_XS, d = i - f * PC_XS;
            float x = __ldg(X + (size_t)(v.frame_off[u] + t0 + f) * PC_XS + d);
            xa_s[f][d] = x;  // [x (39), 1 | x^2 (39), 1]
            xa_s[f][PC_XS + d] = x * x;
        }
        for (int i = threadIdx.x; i < PC_EMIT * rows; i += blockDim.x) {
            int rr = i / rows, f = i - rr * rows;
            size_t o = v.emis_off[u] + (size_t)(t0 + f) * sp + PC_EMIT * pos + rr;
            float lg = __ldg(lgam + o), bb = __ldg(b + o);
            d_s[rr][f] = (lg == PC_NEG_INF) ? PC_NEG_INF : lg - bb;
        }
        __syncthreads();
        for (int f = fl; f < rows; f += frame_lanes) {
            const float4 *x4 = reinterpret_cast<const float4 *>(&xa_s[f][0]);
            float c0 = 0.f, c1 = 0.f;
#pragma unroll
            for (int i = 0; i < PC_KA / 4; ++i) {
                float4 q = x4[i];
                c0 = fmaf(q.x, w[4 * i], c0);
                c1 = fmaf(q.y, w[4 * i + 1], c1);
                c0 = fmaf(q.z, w[4 * i + 2], c0);
                c1 = fmaf(q.w, w[4 * i + 3], c1);
            }
            float d = d_s[r][f];
            float p = (d == PC_NEG_INF) ? 0.f : __expf(c0 + c1 + d);
#pragma unroll
            for (int i = 0; i < PC_KA / 4; ++i) {
                float4 q = x4[i];
                a[4 * i] = fmaf(p, q.x, a[4 * i]);
                a[4 * i + 1] = fmaf(p, q.y, a[4 * i + 1]);
                a[4 * i + 2] = fmaf(p, q.z, a[4 * i + 2]);
                a[4 * i + 3] = fmaf(p, q.w, a[4 * i + 3]);
            }
        }
    }
    if (fl < frame_lanes) {
        double *dst = acc + ((size_t)unit * n_g + g) * PC_KA;
#pragma unroll
        for (int k = 0; k < PC_KA; ++k) atomicAdd(dst + k, (double)a[k]);
    }
}

int launch_accumulate_simt(pc_handle h, const CorpusView &v, const float *X, const float *W,
                           int mix, const float *b, const float *lgam, double *acc,
                           cudaStream_t st) {
    if (v.n_items == 0) return PC_OK;
    int n_g = PC_EMIT * mix;
    if (n_g > ACC_MAX_THREADS) {
        pc_set_error("pc_accumulate: mix=%d exceeds the CUDA-core kernel's limit (%d)", mix,
                     ACC_MAX_THREADS / PC_EMIT);
        return PC_ERR_UNSUPPORTED;
    }
    int fl = ACC_MAX_THREADS / n_g;
    if (fl > 16) fl = 16;
    accumulate_simt_kernel<<<v.n_items, n_g * fl, 0, st>>>(v, X, W, mix, fl, b, lgam, acc);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}
