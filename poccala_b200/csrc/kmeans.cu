// K5: greedy k-means passes of ClusterInitialization.kmeans(algorithm=1) (Clustering.py:894-940)
// with the dimension-0 metric of cal_distance (Clustering.py:796-801), plus the
// statistics the call returns (Clustering.py:941-961, cal_variance :807-832).
//
// What the reference computes, restated for the device:
//   * centres are fixed within a pass; for k = 0..K-1 IN ORDER the cluster takes the FIRST point
//     with the strictly smallest distance to its centre among the points it does not own that are
//     either unowned or strictly closer to centre k than to their owner's centre; a cluster that
//     finds nothing aborts the pass for the clusters after it (Q10);
//   * a moved point leaves its previous cluster only if it holds a non-seed key (Q11): a seed
//     point stays in its seed cluster for ever and is counted twice once another cluster takes it;
//   * after the K moves every centre becomes the mean of its members, summed SEQUENTIALLY in
//     insertion order in fp64 (Clustering.py:880-891) - memberships depend on the centres through
//     < / <= comparisons, so the metric coordinate is reproduced in exactly that order;
//   * stop when a pass moves nothing.
// The K moves of a pass are dependent (each changes the ownership the next one sees), so the
// data-parallel part is the masked arg-min over the points: one CTA per problem (one problem per
// HMM state, AcousticModel.py:553-554), the metric coordinate packed contiguously (8 B per point
// and arg-min; staged in shared memory when the problem fits), ownership as one byte per point.
// Member lists are append-only arrays with tombstones, compacted (and their running sum rebuilt in
// insertion order) only for clusters that lost a point in the pass; clusters that only gained
// points extend their running sum by one addition, which IS the sequential sum.
#include <algorithm>
#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int KM_MAX_K = 127;
constexpr int KM_SMEM_POINTS = 20480;  // 8 B coordinate + 1 B owner per point -> 180 KiB

struct KmProblem {
    int64_t point0;   // first point (row of x) of the problem
    int64_t member0;  // first int of the problem's member arrays in the workspace
    int64_t out0;     // first int of the problem's region of member_list (= point0 + p*k)
    int32_t n;
    int32_t cap;      // capacity of one cluster's member array
};

// (|a-b| ** 2) ** 0.5 as the reference evaluates it.  In IEEE arithmetic sqrt(d*d) == |d| exactly
// unless d*d under- or overflows, so only those ranges take the slow path.
__device__ __forceinline__ double km_dist(double a, double b) {
    const double d = fabs(a - b);
    if (d > 1e150 || (d < 1e-150 && d != 0.0)) return sqrt(d * d);
    return d;
}

__global__ void km_gather_kernel(const double *__restrict__ x, int dim, int64_t n,
                                 double *__restrict__ x0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x0[i] = x[i * dim];
}

__global__ void __launch_bounds__(1024)
kmeans_kernel(const KmProblem *__restrict__ problems, const double *__restrict__ x0_all, int k,
              const int32_t *__restrict__ seeds, int8_t *own_all, int32_t *pos_all,
              int32_t *members_all, int32_t *__restrict__ owner_out,
              int32_t *__restrict__ member_list, int32_t *__restrict__ member_count,
              int32_t *__restrict__ passes_out, int64_t *__restrict__ moves_out, int64_t max_passes,
              int smem_points) {
    extern __shared__ __align__(16) unsigned char km_smem[];
    __shared__ double centre[KM_MAX_K + 1], rsum[KM_MAX_K + 1];
    __shared__ int32_t live[KM_MAX_K + 1], tail[KM_MAX_K + 1], dirty[KM_MAX_K + 1];
    __shared__ double red_d[32];
    __shared__ int32_t red_i[32];
    __shared__ int32_t s_pick;
    __shared__ int32_t s_moved;

    const KmProblem pr = problems[blockIdx.x];
    const int n = pr.n, cap = pr.cap;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const double *xg = x0_all + pr.point0;
    int8_t *own_g = own_all + pr.point0;
    int32_t *pos = pos_all + pr.point0;
    int32_t *members = members_all + pr.member0;
    const int32_t *seed = seeds + (size_t)blockIdx.x * k;

    // stage the metric coordinate and the ownership bytes in shared memory when they fit
    const bool in_smem = n <= smem_points;
    const double *xs = xg;
    int8_t *own = own_g;
    if (in_smem) {
        double *sx = reinterpret_cast<double *>(km_smem);
        int8_t *so = reinterpret_cast<int8_t *>(km_smem + (size_t)smem_points * 8);
        for (int i = tid; i < n; i += nthr) sx[i] = xg[i];
        xs = sx;
        own = so;
    }
    for (int i = tid; i < n; i += nthr) {
        own[i] = -1;
        pos[i] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        // seed kk owns its point (a later duplicate overrides, Clustering.py:1016-1017) and is
        // entry 0 of the cluster's member array: the key -1 entry that is never deleted
        for (int kk = 0; kk < k; ++kk) {
            const int s = seed[kk];
            own[s] = (int8_t)kk;
            centre[kk] = xs[s];
            rsum[kk] = xs[s];  // 0.0 + x == x
            members[(size_t)kk * cap] = s;
            tail[kk] = 1;
            live[kk] = 1;
            dirty[kk] = 0;
        }
    }
    __syncthreads();

    int64_t passes = 0, moves = 0;
    while (passes < max_passes) {
        ++passes;
        if (tid == 0) s_moved = 0;
        for (int kk = 0; kk < k; ++kk) {
            const double ck = centre[kk];
            // `dist < min_dist` with min_dist = sys.maxsize (Clustering.py:896,913)
            double bd = 9223372036854775807.0;
            int bi = 0x7fffffff;
            for (int i = tid; i < n; i += nthr) {
                const int o = own[i];
                if (o == kk) continue;
                const double xi = xs[i];
                const double d = km_dist(ck, xi);
                if (o >= 0 && km_dist(centre[o], xi) <= d) continue;
                if (d < bd) {  // i ascends within a thread: the first strictly smallest wins
                    bd = d;
                    bi = i;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (oi != 0x7fffffff && (bi == 0x7fffffff || od < bd || (od == bd && oi < bi))) {
                    bd = od;
                    bi = oi;
                }
            }
            if (lane == 0) {
                red_d[warp] = bd;
                red_i[warp] = bi;
            }
            __syncthreads();
            if (warp == 0) {
                bd = (lane < nwarp) ? red_d[lane] : 0.0;
                bi = (lane < nwarp) ? red_i[lane] : 0x7fffffff;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, bd, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || od < bd || (od == bd && oi < bi))) {
                        bd = od;
                        bi = oi;
                    }
                }
                if (lane == 0) {
                    s_pick = bi;
                    if (bi != 0x7fffffff) {
                        // the move (Clustering.py:920-934)
                        const int o = own[bi];
                        const int ps = pos[bi];
                        if (o >= 0 && ps >= 0) {  // keyed entry: leaves its old cluster
                            members[(size_t)o * cap + ps] = -1;
                            live[o]--;
                            dirty[o] = 1;
                        }
                        own[bi] = (int8_t)kk;
                        pos[bi] = tail[kk];
                        members[(size_t)kk * cap + tail[kk]] = bi;
                        tail[kk]++;
                        live[kk]++;
                        if (!dirty[kk]) rsum[kk] += xs[bi];
                        s_moved = 1;
                    }
                }
            }
            __syncthreads();
            if (s_pick == 0x7fffffff) break;  // Q10: the remaining clusters of the pass are skipped
            ++moves;
        }
        // new centres (Clustering.py:880-891,938-939): one warp per cluster; a cluster that lost a
        // point compacts its array and re-sums in insertion order
        for (int kk = warp; kk < k; kk += nwarp) {
            if (dirty[kk]) {
                int32_t *m = members + (size_t)kk * cap;
                const int tl = tail[kk];
                double s = 0.0;
                int w = 0;
                for (int base = 0; base < tl; base += 32) {
                    const int j = base + lane;
                    const int idx = (j < tl) ? m[j] : -1;
                    const double v = (idx >= 0) ? xs[idx] : 0.0;
                    const unsigned mask = __ballot_sync(0xffffffffu, idx >= 0);
                    __syncwarp();
                    if (idx >= 0) {
                        const int dst = w + __popc(mask & ((1u << lane) - 1u));
                        m[dst] = idx;
                        if (dst > 0) pos[idx] = dst;  // entry 0 is the seed entry (no key)
                    }
                    w += __popc(mask);
                    for (int q = 0; q < 32; ++q) {  // sequential fp64 sum in insertion order
                        const double vq = __shfl_sync(0xffffffffu, v, q);
                        if ((mask >> q) & 1u) s += vq;
                    }
                }
                if (lane == 0) {
                    tail[kk] = w;
                    rsum[kk] = s;
                    dirty[kk] = 0;
                }
            }
        }
        __syncthreads();
        for (int kk = tid; kk < k; kk += nthr) centre[kk] = rsum[kk] / (double)live[kk];
        const int any_move = s_moved;
        __syncthreads();  // every thread has read s_moved before thread 0 clears it again
        if (!any_move) break;
    }

    // outputs: ownership, ordered member lists (compacted), counts
    for (int i = tid; i < n; i += nthr) owner_out[pr.point0 + i] = own[i];
    __shared__ int32_t out_off[KM_MAX_K + 2];
    if (tid == 0) {
        int o = 0;
        for (int kk = 0; kk < k; ++kk) {
            out_off[kk] = o;
            o += live[kk];
            member_count[(size_t)blockIdx.x * k + kk] = live[kk];
        }
        passes_out[blockIdx.x] = (int32_t)passes;
        moves_out[blockIdx.x] = moves;
    }
    __syncthreads();
    for (int kk = warp; kk < k; kk += nwarp) {
        const int32_t *m = members + (size_t)kk * cap;
        int32_t *dst = member_list + pr.out0 + out_off[kk];
        const int tl = tail[kk];
        int w = 0;
        for (int base = 0; base < tl; base += 32) {
            const int j = base + lane;
            const int idx = (j < tl) ? m[j] : -1;
            const unsigned mask = __ballot_sync(0xffffffffu, idx >= 0);
            if (idx >= 0) dst[w + __popc(mask & ((1u << lane) - 1u))] = idx;
            w += __popc(mask);
        }
    }
}

// The same passes for ONE LARGE problem per thread-block CLUSTER (configs[4] at full size: ~175 000 points per state,
// 13 200 passes x 64 arg-mins, 74 s on a single CTA): the points are cut into contiguous slices, one per CTA of the
// cluster, coordinate and ownership of a slice live in that CTA's shared memory, and every masked arg-min is a local
// scan + one candidate per CTA written into CTA 0's shared memory over DSMEM.  CTA 0 picks (smallest distance, then
// smallest index: the slices are contiguous, so this is the single-CTA order), performs the move - the ownership byte
// of the moved point sits in its slice's CTA, reached through DSMEM - and broadcasts the pick.  Centres are replicated;
// member arrays, positions and the insertion-order sums stay with CTA 0 (global memory), exactly as in kmeans_kernel.
// Two cluster barriers per arg-min.
constexpr int KM_MAX_CLUSTER = 16;

__global__ void __launch_bounds__(1024)
kmeans_cluster_kernel(const KmProblem *__restrict__ problems, const double *__restrict__ x0_all, int k,
                      const int32_t *__restrict__ seeds, int32_t *pos_all, int32_t *members_all,
                      int32_t *__restrict__ owner_out, int32_t *__restrict__ member_list,
                      int32_t *__restrict__ member_count, int32_t *__restrict__ passes_out,
                      int64_t *__restrict__ moves_out, int64_t max_passes, int slice_cap) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int problem = blockIdx.x / CL;
    extern __shared__ __align__(16) unsigned char km_smem[];
    double *sx = reinterpret_cast<double *>(km_smem);
    int8_t *so = reinterpret_cast<int8_t *>(km_smem + (size_t)slice_cap * 8);
    __shared__ double centre[KM_MAX_K + 1], rsum[KM_MAX_K + 1];
    __shared__ int32_t live[KM_MAX_K + 1], tail[KM_MAX_K + 1], dirty[KM_MAX_K + 1];
    __shared__ double red_d[32];
    __shared__ int32_t red_i[32];
    __shared__ double cand_d[KM_MAX_CLUSTER];   // used in CTA 0: one candidate per CTA
    __shared__ int32_t cand_i[KM_MAX_CLUSTER];
    __shared__ int32_t s_pick, s_moved;

    const KmProblem pr = problems[problem];
    const int n = pr.n, cap = pr.cap;
    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    const double *xg = x0_all + pr.point0;
    int32_t *pos = pos_all + pr.point0;
    int32_t *members = members_all + pr.member0;
    const int32_t *seed = seeds + (size_t)problem * k;
    const int S = (n + CL - 1) / CL;                    // slice length (<= slice_cap)
    const int lo = min(n, rank * S), hi = min(n, lo + S);
    const int m = hi - lo;

    for (int i = tid; i < m; i += nthr) {
        sx[i] = xg[lo + i];
        so[i] = -1;
    }
    if (rank == 0)
        for (int i = tid; i < n; i += nthr) pos[i] = -1;
    __syncthreads();
    // seeds (Clustering.py:1016-1017): every CTA marks the seeds of its slice (a later duplicate overrides) and takes
    // the centres; CTA 0 starts the member arrays
    if (tid == 0) {
        for (int kk = 0; kk < k; ++kk) {
            const int sd = seed[kk];
            if (sd >= lo && sd < hi) so[sd - lo] = (int8_t)kk;
            centre[kk] = xg[sd];
            if (rank == 0) {
                rsum[kk] = xg[sd];
                members[(size_t)kk * cap] = sd;
                tail[kk] = 1;
                live[kk] = 1;
                dirty[kk] = 0;
            }
        }
    }
    cluster.sync();

    int64_t passes = 0, moves = 0;
    while (passes < max_passes) {
        ++passes;
        if (tid == 0) s_moved = 0;
        for (int kk = 0; kk < k; ++kk) {
            const double ck = centre[kk];
            double bd = 9223372036854775807.0;
            int bi = 0x7fffffff;
            for (int i = tid; i < m; i += nthr) {
                const int o = so[i];
                if (o == kk) continue;
                const double xi = sx[i];
                const double d = km_dist(ck, xi);
                if (o >= 0 && km_dist(centre[o], xi) <= d) continue;
                if (d < bd) {
                    bd = d;
                    bi = lo + i;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double od = __shfl_xor_sync(0xffffffffu, bd, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (oi != 0x7fffffff && (bi == 0x7fffffff || od < bd || (od == bd && oi < bi))) {
                    bd = od;
                    bi = oi;
                }
            }
            if (lane == 0) {
                red_d[warp] = bd;
                red_i[warp] = bi;
            }
            __syncthreads();
            if (warp == 0) {
                bd = (lane < nwarp) ? red_d[lane] : 0.0;
                bi = (lane < nwarp) ? red_i[lane] : 0x7fffffff;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, bd, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || od < bd || (od == bd && oi < bi))) {
                        bd = od;
                        bi = oi;
                    }
                }
                if (lane == 0) {  // this CTA's candidate -> CTA 0
                    *cluster.map_shared_rank(&cand_d[rank], 0) = bd;
                    *cluster.map_shared_rank(&cand_i[rank], 0) = bi;
                }
            }
            cluster.sync();
            if (rank == 0 && warp == 0) {
                bd = (lane < CL) ? cand_d[lane] : 0.0;
                bi = (lane < CL) ? cand_i[lane] : 0x7fffffff;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const double od = __shfl_xor_sync(0xffffffffu, bd, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (oi != 0x7fffffff && (bi == 0x7fffffff || od < bd || (od == bd && oi < bi))) {
                        bd = od;
                        bi = oi;
                    }
                }
                if (lane == 0 && bi != 0x7fffffff) {
                    // the move (Clustering.py:920-934); the point's ownership byte lives in its slice's CTA
                    int8_t *own_p = cluster.map_shared_rank(&so[bi % S], bi / S);
                    const int o = *own_p;
                    const int ps = pos[bi];
                    if (o >= 0 && ps >= 0) {
                        members[(size_t)o * cap + ps] = -1;
                        live[o]--;
                        dirty[o] = 1;
                    }
                    *own_p = (int8_t)kk;
                    pos[bi] = tail[kk];
                    members[(size_t)kk * cap + tail[kk]] = bi;
                    tail[kk]++;
                    live[kk]++;
                    if (!dirty[kk]) rsum[kk] += xg[bi];
                    s_moved = 1;
                }
                const int pick = __shfl_sync(0xffffffffu, bi, 0);
                if (lane < CL) *cluster.map_shared_rank(&s_pick, lane) = pick;
            }
            cluster.sync();
            if (s_pick == 0x7fffffff) break;  // Q10
            ++moves;
        }
        // new centres by CTA 0 (member arrays and the running sums are its), then to every CTA
        if (rank == 0) {
            for (int kk = warp; kk < k; kk += nwarp) {
                if (dirty[kk]) {
                    int32_t *mm = members + (size_t)kk * cap;
                    const int tl = tail[kk];
                    double s = 0.0;
                    int w = 0;
                    for (int base = 0; base < tl; base += 32) {
                        const int j = base + lane;
                        const int idx = (j < tl) ? mm[j] : -1;
                        const double v = (idx >= 0) ? xg[idx] : 0.0;
                        const unsigned mask = __ballot_sync(0xffffffffu, idx >= 0);
                        __syncwarp();
                        if (idx >= 0) {
                            const int dst = w + __popc(mask & ((1u << lane) - 1u));
                            mm[dst] = idx;
                            if (dst > 0) pos[idx] = dst;
                        }
                        w += __popc(mask);
                        for (int q = 0; q < 32; ++q) {
                            const double vq = __shfl_sync(0xffffffffu, v, q);
                            if ((mask >> q) & 1u) s += vq;
                        }
                    }
                    if (lane == 0) {
                        tail[kk] = w;
                        rsum[kk] = s;
                        dirty[kk] = 0;
                    }
                }
            }
            __syncthreads();
            const int any = s_moved;
            for (int j = tid; j < k * CL; j += nthr) {
                const int kk = j % k, c = j / k;
                *cluster.map_shared_rank(&centre[kk], c) = rsum[kk] / (double)live[kk];
            }
            if (tid < CL) *cluster.map_shared_rank(&s_moved, tid) = any;
        }
        cluster.sync();
        const int any_move = s_moved;
        cluster.sync();  // everybody has read the flag before CTA 0 clears it in the next pass
        if (!any_move) break;
    }

    for (int i = tid; i < m; i += nthr) owner_out[pr.point0 + lo + i] = so[i];
    if (rank == 0) {
        __shared__ int32_t out_off[KM_MAX_K + 2];
        if (tid == 0) {
            int o = 0;
            for (int kk = 0; kk < k; ++kk) {
                out_off[kk] = o;
                o += live[kk];
                member_count[(size_t)problem * k + kk] = live[kk];
            }
            passes_out[problem] = (int32_t)passes;
            moves_out[problem] = moves;
        }
        __syncthreads();
        for (int kk = warp; kk < k; kk += nwarp) {
            const int32_t *mm = members + (size_t)kk * cap;
            int32_t *dst = member_list + pr.out0 + out_off[kk];
            const int tl = tail[kk];
            int w = 0;
            for (int base = 0; base < tl; base += 32) {
                const int j = base + lane;
                const int idx = (j < tl) ? mm[j] : -1;
                const unsigned mask = __ballot_sync(0xffffffffu, idx >= 0);
                if (idx >= 0) dst[w + __popc(mask & ((1u << lane) - 1u))] = idx;
                w += __popc(mask);
            }
        }
    }
    cluster.sync();  // no CTA leaves while its shared memory may still be addressed
}

// mean / variance / weight of every cluster (Clustering.py:880-891 cal_center, :807-832
// cal_variance(algorithm='kmeans'), :947 alpha): one thread per (cluster, dimension), sequential
// fp64 sums in insertion order; variance floor 1e-4, returned as (sqrt(v))^2 like np.diag(std**2).
__global__ void kmeans_finish_kernel(const KmProblem *__restrict__ problems,
                                     const double *__restrict__ x, int dim, int k,
                                     const int32_t *__restrict__ member_list,
                                     const int32_t *__restrict__ member_count,
                                     double *__restrict__ mean, double *__restrict__ var,
                                     double *__restrict__ alpha) {
    const int p = blockIdx.x, kk = blockIdx.y;
    const KmProblem pr = problems[p];
    const int32_t *cnt = member_count + (size_t)p * k;
    int64_t off = pr.out0;
    for (int q = 0; q < kk; ++q) off += cnt[q];
    const int len = cnt[kk];
    const int32_t *m = member_list + off;
    const double *xp = x + pr.point0 * dim;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        double s = 0.0;
        for (int j = 0; j < len; ++j) s += xp[(size_t)m[j] * dim + d];
        const double c = s / (double)len;
        double v = 0.0;
        for (int j = 0; j < len; ++j) {
            const double e = c - xp[(size_t)m[j] * dim + d];
            v += e * e;
        }
        v /= (double)len;
        if (v < 1e-4) v = 1e-4;
        const double sd = sqrt(v);
        const size_t o = ((size_t)p * k + kk) * dim + d;
        mean[o] = c;
        var[o] = sd * sd;
    }
    if (threadIdx.x == 0) alpha[(size_t)p * k + kk] = (double)len / (double)pr.n;
}

struct KmLayout {
    size_t o_table, o_x0, o_own, o_pos, o_members, total;
    std::vector<KmProblem> table;
};

}  // namespace

static int km_layout(int32_t n_problems, const int64_t *point_off, int32_t k, KmLayout &L) {
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    L.table.resize(n_problems);
    int64_t member = 0;
    for (int p = 0; p < n_problems; ++p) {
        const int64_t n = point_off[p + 1] - point_off[p];
        if (n < 1 || n > 0x7ffffff0) return -1;
        KmProblem &q = L.table[p];
        q.point0 = point_off[p];
        q.n = (int32_t)n;
        q.cap = (int32_t)n + k + 2;
        q.member0 = member;
        q.out0 = point_off[p] + (int64_t)p * k;
        member += (int64_t)k * q.cap;
    }
    const int64_t total_points = n_problems ? point_off[n_problems] : 0;
    size_t off = 0;
    L.o_table = off; off = al(off + sizeof(KmProblem) * (size_t)n_problems);
    L.o_x0 = off;    off = al(off + 8 * (size_t)total_points);
    L.o_own = off;   off = al(off + (size_t)total_points);
    L.o_pos = off;   off = al(off + 4 * (size_t)total_points);
    L.o_members = off; off = al(off + 4 * (size_t)member);
    L.total = off;
    return 0;
}

extern "C" {

int64_t pc_kmeans_workspace_bytes(int32_t n_problems, const int64_t *host_point_off, int32_t k) {
    if (n_problems < 0 || k < 1 || k > KM_MAX_K || (n_problems > 0 && !host_point_off)) return -1;
    KmLayout L;
    if (km_layout(n_problems, host_point_off, k, L)) return -1;
    return (int64_t)L.total;
}

int pc_kmeans_run(pc_handle h, int32_t n_problems, const int64_t *host_point_off,
                  const double *dev_x, int32_t dim, int32_t k, const int32_t *dev_seed_points,
                  void *dev_workspace, int32_t *dev_owner, int32_t *dev_member_list,
                  int32_t *dev_member_count, int32_t *dev_passes, int64_t *dev_moves,
                  int64_t max_passes, void *stream) {
    PC_REQUIRE(h != nullptr, "pc_kmeans_run: NULL handle");
    PC_CUDA_TRY(cudaSetDevice(h->device));
    PC_REQUIRE(n_problems >= 0, "pc_kmeans_run: n_problems=%d", n_problems);
    PC_REQUIRE(k >= 1 && k <= KM_MAX_K, "pc_kmeans_run: k=%d outside [1,%d]", k, KM_MAX_K);
    PC_REQUIRE(dim >= 1, "pc_kmeans_run: dim=%d", dim);
    if (n_problems == 0) return PC_OK;
    PC_REQUIRE(host_point_off && dev_x && dev_seed_points && dev_workspace && dev_owner &&
                   dev_member_list && dev_member_count && dev_passes && dev_moves,
               "pc_kmeans_run: NULL argument");
    PC_REQUIRE(max_passes >= 1, "pc_kmeans_run: max_passes=%lld", (long long)max_passes);
    KmLayout L;
    PC_REQUIRE(km_layout(n_problems, host_point_off, k, L) == 0,
               "pc_kmeans_run: every problem needs between 1 and 2^31-16 points");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)dev_workspace;
    PC_CUDA_TRY(cudaMemcpyAsync(ws + L.o_table, L.table.data(), sizeof(KmProblem) * (size_t)n_problems,
                                cudaMemcpyHostToDevice, st));
    const int64_t total_points = host_point_off[n_problems];
    int max_n = 0;
    for (auto &q : L.table) max_n = std::max(max_n, q.n);
    km_gather_kernel<<<(unsigned)((total_points + 255) / 256), 256, 0, st>>>(
        dev_x, dim, total_points, (double *)(ws + L.o_x0));
    PC_LAUNCH_CHECK();
    // One thread-block cluster per problem when that shortens the run: the cluster answers an arg-min ~6x faster than
    // one CTA at 175 000 points but occupies up to 16 SMs, so it pays when the SMs would otherwise idle (few problems:
    // the states of one rank of a multi-GPU job, a single large data set) - with more than sm_count / 2 problems one
    // CTA each has the better throughput.  Option "kmeans_cluster": 0 = never, 1 = this rule, 2 = always (tests).
    int CL = 1;
    while (CL < KM_MAX_CLUSTER && (int64_t)n_problems * (CL * 2) <= h->sm_count) CL *= 2;
    if (h->kmeans_cluster == 2) CL = KM_MAX_CLUSTER;
    int need = 1;
    while ((int64_t)need * KM_SMEM_POINTS < max_n) need *= 2;
    const bool few = (int64_t)n_problems * 2 <= h->sm_count;
    if (h->kmeans_cluster && need <= KM_MAX_CLUSTER && ((max_n > 8192 && CL > 1 && few) || h->kmeans_cluster == 2)) {
        CL = std::max(CL, need);
        const int slice_cap = std::max(1, (max_n + CL - 1) / CL);
        const size_t smem = (size_t)slice_cap * 9 + 16;
        PC_CUDA_TRY(cudaFuncSetAttribute(kmeans_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (CL > 8) PC_CUDA_TRY(cudaFuncSetAttribute(kmeans_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)n_problems * CL);
        cfg.blockDim = dim3(1024);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kmeans_cluster_kernel, (const KmProblem *)(ws + L.o_table),
                                       (const double *)(ws + L.o_x0), (int)k, dev_seed_points, (int32_t *)(ws + L.o_pos),
                                       (int32_t *)(ws + L.o_members), dev_owner, dev_member_list, dev_member_count,
                                       dev_passes, dev_moves, (int64_t)max_passes, slice_cap));
        h->launches += 2;
        return PC_OK;
    }
    const int smem_points = std::min(max_n, KM_SMEM_POINTS);
    const size_t smem = (size_t)smem_points * 9 + 16;
    PC_CUDA_TRY(cudaFuncSetAttribute(kmeans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = max_n <= 2048 ? 256 : (max_n <= 8192 ? 512 : 1024);
    kmeans_kernel<<<n_problems, threads, smem, st>>>(
        (const KmProblem *)(ws + L.o_table), (const double *)(ws + L.o_x0), k, dev_seed_points,
        (int8_t *)(ws + L.o_own), (int32_t *)(ws + L.o_pos), (int32_t *)(ws + L.o_members), dev_owner,
        dev_member_list, dev_member_count, dev_passes, dev_moves, max_passes, smem_points);
    PC_LAUNCH_CHECK();
    h->launches += 2;
    return PC_OK;
}

int pc_kmeans_finish(pc_handle h, int32_t n_problems, const int64_t *host_point_off,
                     const double *dev_x, int32_t dim, int32_t k, const void *dev_workspace,
                     const int32_t *dev_member_list, const int32_t *dev_member_count,
                     double *dev_mean, double *dev_var, double *dev_alpha, void *stream) {
    PC_REQUIRE(h != nullptr, "pc_kmeans_finish: NULL handle");
    PC_CUDA_TRY(cudaSetDevice(h->device));
    PC_REQUIRE(n_problems >= 0 && k >= 1 && k <= KM_MAX_K && dim >= 1, "pc_kmeans_finish: bad sizes");
    if (n_problems == 0) return PC_OK;
    PC_REQUIRE(host_point_off && dev_x && dev_workspace && dev_member_list && dev_member_count &&
                   dev_mean && dev_var && dev_alpha,
               "pc_kmeans_finish: NULL argument");
    // the problem table written by pc_kmeans_run sits at the start of the workspace
    kmeans_finish_kernel<<<dim3(n_problems, k), 64, 0, (cudaStream_t)stream>>>(
        (const KmProblem *)dev_workspace, dev_x, dim, k, dev_member_list, dev_member_count, dev_mean,
        dev_var, dev_alpha);
    PC_LAUNCH_CHECK();
    h->launches++;
    return PC_OK;
}

}  // extern "C"
