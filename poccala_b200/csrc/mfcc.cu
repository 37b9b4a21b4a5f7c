// Front end: MFCC features and the cepstral-distance voice activity detector, the step in front
// of the E-step (AcousticModel.__load_audio, AcousticModel.py:463-477 -> AudioProcessing.MFCC.mfcc,
// AudioProcessing.py:416-448, and AudioProcessing.VAD, :450-542).  fp64 like the reference; the arithmetic is kept
// as the reference writes it, including what a textbook front end does differently:
//   * pre-emphasis y[n] = s[n+1] - 0.98 s[n], a zero appended (:195-198);
//   * the Hamming "window" scales frame f as a whole by 0.54 - 0.46 cos(2 pi f / (F - 1)) - it runs over the
//     frame index, not over the samples of a frame (:243-246);
//   * the spectrum is the MAGNITUDE of rfft(frame, nfft) (frames longer than nfft are cut, :262-263), the frame
//     energy the sum of the magnitudes (:338), the filter bank the caller's response matrix (the mirror builds it
//     with the reference's own expression, rising flanks on both sides, :318-326);
//   * DCT with the (2k - 1) argument and the 2 / sqrt(n_filters) factor on every coefficient (:359-367);
//   * deltas over +-2 frames with edge padding, denominator 10 (:405-412).
// One block per frame: bit-reversed load, radix-2 FFT in shared memory (9 stages at nfft = 512), magnitudes,
// block reductions for the energy and the filters, DCT by the first warps.
#include "common.cuh"

namespace {

constexpr int MF_THREADS = 256;
constexpr int MF_MAX_NFFT = 2048;
constexpr int MF_MAX_FILTERS = 64;

__global__ void __launch_bounds__(MF_THREADS)
mfcc_frame_kernel(const double *__restrict__ signal, int64_t n_samples, int framesize, int step, int n_frames, int nfft,
                  int log2n, const double *__restrict__ fbank, int n_filters, int n_ceps, int cal_energy, int out_dim,
                  double *__restrict__ out) {
    extern __shared__ double sh[];
    double *re = sh, *im = sh + nfft, *mag = sh + 2 * nfft, *logfb = mag + (nfft / 2 + 1);
    __shared__ double red[MF_THREADS / 32];
    const int f = blockIdx.x, tid = threadIdx.x;
    const double PI = 3.141592653589793;
    // math.cos(2 * math.pi * i / (length - 1)); a single frame divides by zero in the reference (the host side refuses it)
    const double win = 0.54 - 0.46 * cos(2.0 * PI * (double)f / (double)(n_frames - 1));
    const int64_t base = (int64_t)f * step;
    for (int i = tid; i < nfft; i += MF_THREADS) {
        double v = 0.0;
        if (i < framesize) {
            const int64_t n = base + i;
            if (n < n_samples - 1) v = signal[n + 1] - 0.98 * signal[n];  // the appended zero and the padding stay 0
            v *= win;
        }
        const int j = (int)(__brev((unsigned)i) >> (32 - log2n));
        re[j] = v;
        im[j] = 0.0;
    }
    __syncthreads();
    for (int s = 1; s <= log2n; ++s) {
        const int half = 1 << (s - 1);
        for (int b = tid; b < nfft / 2; b += MF_THREADS) {
            const int grp = b / half, k = b - grp * half;
            const int i0 = grp * 2 * half + k, i1 = i0 + half;
            double sn, cs;
            sincospi(-(double)k / (double)half, &sn, &cs);
            const double tr = re[i1] * cs - im[i1] * sn, ti = re[i1] * sn + im[i1] * cs;
            const double ur = re[i0], ui = im[i0];
            re[i0] = ur + tr; im[i0] = ui + ti;
            re[i1] = ur - tr; im[i1] = ui - ti;
        }
        __syncthreads();
    }
    const int n_bins = nfft / 2 + 1;
    double e = 0.0;
    for (int k = tid; k < n_bins; k += MF_THREADS) {
        const double m = hypot(re[k], im[k]);
        mag[k] = m;
        e += m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    if ((tid & 31) == 0) red[tid >> 5] = e;
    __syncthreads();
    double energy = 0.0;
    for (int w = 0; w < MF_THREADS / 32; ++w) energy += red[w];
    // filter bank: one warp per filter, then the logarithm (np.log: -inf for an empty filter, as in the reference)
    for (int m = tid >> 5; m < n_filters; m += MF_THREADS / 32) {
        double s = 0.0;
        for (int k = tid & 31; k < n_bins; k += 32) s += mag[k] * fbank[(size_t)m * n_bins + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if ((tid & 31) == 0) logfb[m] = log(s);
    }
    __syncthreads();
    if (tid < n_ceps) {
        const double coef = 2.0 / sqrt((double)n_filters);
        double s = 0.0;
        for (int k = 0; k < n_filters; ++k)
            s += coef * logfb[k] * cos(PI * (double)(2 * k - 1) * (double)tid / (double)(2 * n_filters));
        if (tid == 0 && cal_energy) s = log(energy);
        out[(size_t)f * out_dim + tid] = s;
    }
}

// out[t][dst .. dst + n) = delta of out[t][src .. src + n): sum_{i = -2..2} i * x[clamp(t + i)] / 10
__global__ void delta_kernel(double *__restrict__ out, int n_frames, int out_dim, int src, int dst, int n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n_frames * n) return;
    const int t = (int)(i / n), j = (int)(i - (int64_t)t * n);
    double s = 0.0;
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
        const int tt = min(max(t + d, 0), n_frames - 1);
        s += (double)d * out[(size_t)tt * out_dim + src + j];
    }
    out[(size_t)t * out_dim + dst + j] = s / 10.0;
}

// VAD.mel_distance + VAD.osf (AudioProcessing.py:462-506): noise estimate from the first `sample` frames (mean, then
// `sample` steps of noise = alpha noise + (1 - alpha) mfcc[i]), Euclidean distance of every frame to it, and the order
// statistics filter: for sample <= t < T - sample the window dist[t - sample, t + sample) sorted ascending, the values at
// h = int(beta (2 sample + 1)) and h + 1 blended as (1 - beta) d + beta d'.  One block; frames across threads.
__global__ void __launch_bounds__(256)
vad_distance_kernel(const double *__restrict__ mfcc, int n_frames, int dim, int sample, double alpha, double beta,
                    double *__restrict__ dist, double *__restrict__ dist_osf) {
    __shared__ double noise[64];
    if ((int)threadIdx.x < dim) {
        const int d = threadIdx.x;
        double s = 0.0;
        for (int i = 0; i < sample; ++i) s += mfcc[(size_t)i * dim + d];  // simple.sum(axis=0)
        double nz = 1.0 / (double)sample * s;
        for (int i = 0; i < sample; ++i) nz = alpha * nz + (1.0 - alpha) * mfcc[(size_t)i * dim + d];
        noise[d] = nz;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n_frames; t += blockDim.x) {
        double s = 0.0;
        for (int d = 0; d < dim; ++d) {
            const double v = noise[d] - mfcc[(size_t)t * dim + d];
            s += v * v;
        }
        dist[t] = sqrt(s);
    }
    __syncthreads();
    __threadfence_block();
    const int h = (int)(beta * (double)(2 * sample + 1));
    for (int t = threadIdx.x; t < n_frames; t += blockDim.x) {
        double v = dist[t];
        if (t >= sample && t < n_frames - sample) {
            double w[64];
            const int n = 2 * sample;
            for (int i = 0; i < n; ++i) {  // insertion sort of the window
                const double x = dist[t - sample + i];
                int j = i;
                while (j > 0 && w[j - 1] > x) { w[j] = w[j - 1]; --j; }
                w[j] = x;
            }
            v = (1.0 - beta) * w[h] + beta * w[h + 1];
        }
        dist_osf[t] = v;
    }
}

}  // namespace

extern "C" int pc_mfcc(pc_handle h, const double *dev_signal, int64_t n_samples, int32_t framesize, int32_t step,
                       int32_t n_frames, int32_t nfft, const double *dev_fbank, int32_t n_filters, int32_t n_ceps,
                       int32_t cal_energy, int32_t n_delta, double *dev_out, void *stream) {
    if (!h || !dev_signal || !dev_fbank || !dev_out) {
        pc_set_error("pc_mfcc: NULL argument");
        return PC_ERR_INVALID;
    }
    int log2n = 0;
    while ((1 << log2n) < nfft) ++log2n;
    if (nfft < 64 || nfft > MF_MAX_NFFT || (1 << log2n) != nfft || n_filters < 1 || n_filters > MF_MAX_FILTERS || n_ceps < 1 ||
        n_ceps > 32 || n_delta < 0 || n_delta > 2 || framesize < 1 || step < 1 || n_frames < 2 || n_samples < 1) {
        pc_set_error("pc_mfcc: nfft=%d (power of two, 64..%d), %d filters (<= %d), %d coefficients (<= 32), %d deltas "
                     "(0..2), frame size %d, step %d, %d frames (>= 2)", nfft, MF_MAX_NFFT, n_filters, MF_MAX_FILTERS,
                     n_ceps, n_delta, framesize, step, n_frames);
        return PC_ERR_INVALID;
    }
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int out_dim = n_ceps * (1 + n_delta);
    const size_t smem = (size_t)(2 * nfft + nfft / 2 + 1 + n_filters) * sizeof(double);
    PC_CUDA_TRY(cudaFuncSetAttribute(mfcc_frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mfcc_frame_kernel<<<n_frames, MF_THREADS, smem, st>>>(dev_signal, n_samples, framesize, step, n_frames, nfft, log2n,
                                                         dev_fbank, n_filters, n_ceps, cal_energy, out_dim, dev_out);
    PC_LAUNCH_CHECK();
    h->launches++;
    for (int k = 0; k < n_delta; ++k) {
        const int64_t n = (int64_t)n_frames * n_ceps;
        delta_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dev_out, n_frames, out_dim, k * n_ceps, (k + 1) * n_ceps, n_ceps);
        PC_LAUNCH_CHECK();
        h->launches++;
    }
    return PC_OK;
}

extern "C" int pc_vad_distance(pc_handle h, const double *dev_mfcc, int32_t n_frames, int32_t dim, int32_t sample_size,
                               double alpha, double beta, double *dev_dist, double *dev_dist_osf, void *stream) {
    if (!h || !dev_mfcc || !dev_dist || !dev_dist_osf) {
        pc_set_error("pc_vad_distance: NULL argument");
        return PC_ERR_INVALID;
    }
    if (dim < 1 || dim > 64 || sample_size < 1 || sample_size > 32 || n_frames < sample_size) {
        pc_set_error("pc_vad_distance: dim=%d (<= 64), sample size %d (<= 32), %d frames (>= sample size)", dim, sample_size,
                     n_frames);
        return PC_ERR_INVALID;
    }
    cudaSetDevice(h->device);
    vad_distance_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dev_mfcc, n_frames, dim, sample_size, alpha, beta, dev_dist,
                                                            dev_dist_osf);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}
