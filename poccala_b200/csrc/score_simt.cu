// K1 (CUDA-core generation): GMM scoring + log-sum-exp over mixtures.
//
// Restates LHMM.cal_observation_pro -> GMM.point -> util.gaussian_function (LHMM.py:163-187,
// Clustering.py:740-767, util.py:20-36) as a dot product of the augmented frame [x, x^2, 1, 1]
// with the packed Gaussian row (pack.cu), followed by an online log-sum-exp over each state's
// mixture components (util.py:54-77).  One CTA per work item (a run of 128-frame tiles that all
// belong to one unit): the unit's 3*mix rows of W stay in shared memory, one thread owns one frame.
#include "common.cuh"

__device__ __forceinline__ void load_aug_row(const float *__restrict__ xrow, float (&xa)[PC_KA]) {
    const float4 *p = reinterpret_cast<const float4 *>(xrow);
    float x[PC_XS];
#pragma unroll
    for (int i = 0; i < PC_XS / 4; ++i) {
        float4 v = __ldg(p + i);
        x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int d = 0; d < PC_XS; ++d) {  // [x (39), 1 | x^2 (39), 1]
        xa[d] = x[d];
        xa[PC_XS + d] = x[d] * x[d];
    }
}

__device__ __forceinline__ float dot_aug(const float (&xa)[PC_KA], const float *__restrict__ w) {
    const float4 *w4 = reinterpret_cast<const float4 *>(w);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int i = 0; i < PC_KA / 4; ++i) {
        float4 v = w4[i];
        a0 = fmaf(xa[4 * i], v.x, a0);
        a1 = fmaf(xa[4 * i + 1], v.y, a1);
        a0 = fmaf(xa[4 * i + 2], v.z, a0);
        a1 = fmaf(xa[4 * i + 3], v.w, a1);
    }
    return a0 + a1;
}

// Online log-sum-exp state: (mx, sum) with the reference's +-inf conventions.
struct Lse {
    float mx, sum;
    __device__ __forceinline__ void init() { mx = PC_NEG_INF; sum = 0.f; }
    __device__ __forceinline__ void add(float c) {
        if (c > mx) {
            sum = sum * __expf(mx - c) + 1.f;
            mx = c;
        } else if (c > PC_NEG_INF) {
            sum += __expf(c - mx);
        }
    }
    __device__ __forceinline__ float value() const {
        return (mx == PC_NEG_INF) ? PC_NEG_INF : mx + __logf(sum);
    }
};

__global__ void __launch_bounds__(PC_TILE_ROWS)
score_simt_kernel(CorpusView v, const float *__restrict__ X, const float *__restrict__ W, int mix,
                  float *__restrict__ b) {
    extern __shared__ __align__(16) float w_s[];  // [3*mix][PC_KA]
    const int item = blockIdx.x;
    const int unit = v.item_unit[item];
    const int n_rows = PC_EMIT * mix;
    {
        const float4 *src = reinterpret_cast<const float4 *>(W + (size_t)unit * n_rows * PC_KA);
        float4 *dst = reinterpret_cast<float4 *>(w_s);
        for (int i = threadIdx.x; i < n_rows * PC_KA / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int64_t lo = v.item_tile_lo[item], hi = v.item_tile_lo[item + 1];
    for (int64_t tile = lo; tile < hi; ++tile) {
        const int64_t pair = v.tile_pair[tile];
        const int u = v.pair_utt[pair];
        const int pos = (int)(pair - v.pair_off[u]);
        const int T = (int)(v.frame_off[u + 1] - v.frame_off[u]);
        const int t = v.tile_t0[tile] + threadIdx.x;
        if (t >= T) continue;
        float xa[PC_KA];
        load_aug_row(X + (size_t)(v.frame_off[u] + t) * PC_XS, xa);
        const int sp = pc_spad((int)(v.pair_off[u + 1] - v.pair_off[u]));
        float *out = b + v.emis_off[u] + (size_t)t * sp + PC_EMIT * pos;
        for (int r = 0; r < PC_EMIT; ++r) {
            Lse l;
            l.init();
            const float *wr = w_s + (size_t)r * mix * PC_KA;
            for (int m = 0; m < mix; ++m) l.add(dot_aug(xa, wr + (size_t)m * PC_KA));
            out[r] = l.value();
        }
    }
}

// Dense sweep (BASELINE config 3): every frame against a chunk of states per CTA.
__global__ void __launch_bounds__(PC_TILE_ROWS)
score_dense_simt_kernel(const float *__restrict__ X, int64_t n, const float *__restrict__ W,
                        int n_states, int mix, int states_per_cta, float *__restrict__ out) {
    extern __shared__ __align__(16) float w_s[];
    const int s0 = blockIdx.y * states_per_cta;
    const int ns = min(states_per_cta, n_states - s0);
    {
        const float4 *src = reinterpret_cast<const float4 *>(W + (size_t)s0 * mix * PC_KA);
        float4 *dst = reinterpret_cast<float4 *>(w_s);
        for (int i = threadIdx.x; i < ns * mix * PC_KA / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    for (int64_t f = (int64_t)blockIdx.x * PC_TILE_ROWS + threadIdx.x; f < n;
         f += (int64_t)gridDim.x * PC_TILE_ROWS) {
        float xa[PC_KA];
        load_aug_row(X + (size_t)f * PC_XS, xa);
        for (int s = 0; s < ns; ++s) {
            Lse l;
            l.init();
            const float *wr = w_s + (size_t)s * mix * PC_KA;
            for (int m = 0; m < mix; ++m) l.add(dot_aug(xa, wr + (size_t)m * PC_KA));
            out[(size_t)f * n_states + s0 + s] = l.value();
        }
    }
}

int launch_score_simt(pc_handle h, const CorpusView &v, const float *X, const float *W, int mix,
                      float *b, cudaStream_t st) {
    if (v.n_items == 0) return PC_OK;
    size_t smem = (size_t)PC_EMIT * mix * PC_KA * sizeof(float);
    if (smem > 200 * 1024) {
        pc_set_error("pc_gmm_score: mix=%d needs %zu B of shared memory (limit 200 KiB)", mix, smem);
        return PC_ERR_UNSUPPORTED;
    }
    PC_CUDA_TRY(cudaFuncSetAttribute(score_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem));
    score_simt_kernel<<<v.n_items, PC_TILE_ROWS, smem, st>>>(v, X, W, mix, b);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int launch_score_dense_simt(pc_handle h, const float *X, int64_t n, const float *W, int n_states,
                            int mix, float *out, cudaStream_t st) {
    if (n == 0 || n_states == 0) return PC_OK;
    int per = (int)((96 * 1024) / ((size_t)mix * PC_KA * sizeof(float)));
    if (per < 1) {
        pc_set_error("pc_gmm_score_dense: mix=%d too large", mix);
        return PC_ERR_UNSUPPORTED;
    }
    if (per > n_states) per = n_states;
    size_t smem = (size_t)per * mix * PC_KA * sizeof(float);
    PC_CUDA_TRY(cudaFuncSetAttribute(score_dense_simt_kernel,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t ftiles = (n + PC_TILE_ROWS - 1) / PC_TILE_ROWS;
    int gx = (int)(ftiles < (int64_t)h->sm_count * 4 ? ftiles : (int64_t)h->sm_count * 4);
    int gy = (n_states + per - 1) / per;
    if (gy > 65535) {
        pc_set_error("pc_gmm_score_dense: too many state chunks (%d)", gy);
        return PC_ERR_UNSUPPORTED;
    }
    score_dense_simt_kernel<<<dim3(gx, gy), PC_TILE_ROWS, smem, st>>>(X, n, W, n_states, mix, per,
                                                                      out);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}
