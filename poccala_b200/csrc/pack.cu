// Model packing and frame preparation.
//
// pack_gmm: util.gaussian_function's log branch (util.py:20-36) plus the log(alpha) term of
// Clustering.GMM.point (Clustering.py:753-757) folded into one fp32 row per Gaussian:
//     score(x, g) = <[x', x'^2, 1, 1], W_g>,   x' = (x - shift) * inv_scale
//     W_g = [ mu'/var' (39), k_hi | -1/(2 var') (39), k_lo ]   (matches [x (39), 1 | x^2 (39), 1])
//     k   = log alpha - D/2 log 2pi - 1/2 sum_d var_d (Q1: ORIGINAL variances, not log-det)
//           - 1/2 sum_d mu'^2/var'
// All arithmetic in fp64; k is stored as an fp32 pair so the constant keeps ~48 bits.
#include "common.cuh"

__global__ void pack_gmm_kernel(const double *__restrict__ mean, const double *__restrict__ var,
                                const double *__restrict__ alpha, const double *__restrict__ shift,
                                const double *__restrict__ inv_scale, int n_gauss, int dim,
                                float *__restrict__ W) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gauss) return;
    const double LOG_2PI = 1.8378770664093453;  // np.log(2*pi), util.py:14
    float *w = W + (size_t)g * PC_KA;
    double sum_var = 0.0, quad = 0.0;
    for (int d = 0; d < PC_DIM_MAX; ++d) {
        if (d < dim) {
            double mu = mean[(size_t)g * dim + d];
            double v = var[(size_t)g * dim + d];
            double sh = shift ? shift[d] : 0.0;
            double is = inv_scale ? inv_scale[d] : 1.0;
            double mu_s = (mu - sh) * is;
            double v_s = v * is * is;
            sum_var += v;
            quad += mu_s * mu_s / v_s;
            w[d] = (float)(mu_s / v_s);
            w[PC_XS + d] = (float)(-0.5 / v_s);
        } else {
            w[d] = 0.f;
            w[PC_XS + d] = 0.f;
        }
    }
    double k = log(alpha[g]) - 0.5 * dim * LOG_2PI - 0.5 * sum_var - 0.5 * quad;
    float k_hi = (float)k;
    float k_lo = isfinite(k) ? (float)(k - (double)k_hi) : 0.f;
    w[PC_XS - 1] = k_hi;
    w[PC_KA - 1] = k_lo;
}

template <typename T>
__global__ void prepare_frames_kernel(const T *__restrict__ x, int64_t n, int dim,
                                      const double *__restrict__ shift,
                                      const double *__restrict__ inv_scale, float *__restrict__ X) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per output float
    if (i >= n * PC_XS) return;
    int64_t f = i / PC_XS;
    int d = (int)(i - f * PC_XS);
    float out;
    if (d < dim) {
        double v = (double)x[f * dim + d];
        double sh = shift ? shift[d] : 0.0;
        double is = inv_scale ? inv_scale[d] : 1.0;
        out = (float)((v - sh) * is);
    } else {
        out = (d == PC_XS - 1) ? 1.f : 0.f;
    }
    X[i] = out;
}

int launch_pack_gmm(pc_handle h, const double *mean, const double *var, const double *alpha,
                    const double *shift, const double *inv_scale, int n_gauss, int dim, float *W,
                    cudaStream_t st) {
    if (n_gauss == 0) return PC_OK;
    int threads = 128;
    int blocks = (n_gauss + threads - 1) / threads;
    pack_gmm_kernel<<<blocks, threads, 0, st>>>(mean, var, alpha, shift, inv_scale, n_gauss, dim, W);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

int launch_prepare_frames(pc_handle h, const void *x, int is_f64, int64_t n, int dim,
                          const double *shift, const double *inv_scale, float *X, cudaStream_t st) {
    if (n == 0) return PC_OK;
    int threads = 256;
    int64_t total = n * PC_XS;
    int64_t blocks = (total + threads - 1) / threads;
    if (blocks > 2147483647LL) {
        pc_set_error("prepare_frames: too many frames for one launch");
        return PC_ERR_UNSUPPORTED;
    }
    if (is_f64)
        prepare_frames_kernel<double><<<(unsigned)blocks, threads, 0, st>>>(
            (const double *)x, n, dim, shift, inv_scale, X);
    else
        prepare_frames_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(
            (const float *)x, n, dim, shift, inv_scale, X);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}
