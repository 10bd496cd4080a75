// cs_stats.cu — cell-sharded sufficient statistics for the multi-GPU M-step.
//
// Each rank holds a contiguous shard of cells with fixed labels; the per-cluster sums,
// non-NA counts and cluster sizes (R/BayesNonparCluster.R:221-224) and the per-segment
// squared residuals (:234) of the shard are produced here, and the caller all-reduces
// the (K*R*2 + K + 2R) doubles across ranks (NCCL over NVLink; latency-bound, ~100 KB).
#include "cs_common.cuh"

#include <math_constants.h>

namespace {

constexpr int ST_CHUNK = 512;
constexpr int ST_UNROLL = 8;
// a thread per segment: one warp-multiple that covers R (no thread with two segments while its neighbours idle)
inline int st_threads(int R) { return R >= 1024 ? 1024 : (R < 64 ? 64 : ((R + 31) / 32) * 32); }

// grid: cell chunks.  Shared accumulators when K*R fits, global atomics otherwise.
__global__ void __launch_bounds__(1024)
cluster_stats_kernel(int N, int R, int K, const double *__restrict__ X, const int *__restrict__ Z,
                     double *__restrict__ sum_x, double *__restrict__ cnt_x, double *__restrict__ nk, int use_smem)
{
    extern __shared__ double s_acc[]; // K*R sums, K*R counts
    const int i0 = blockIdx.x * ST_CHUNK, i1 = min(N, i0 + ST_CHUNK);
    if (use_smem) {
        for (int e = threadIdx.x; e < 2 * K * R; e += blockDim.x) s_acc[e] = 0.0;
        __syncthreads();
    }
    // a thread owns segment r; ST_UNROLL cells' values are in flight at a time (cells still in order)
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        for (int i = i0; i < i1; i += ST_UNROLL) {
            int k[ST_UNROLL];
            double v[ST_UNROLL];
#pragma unroll
            for (int u = 0; u < ST_UNROLL; u++) {
                const bool in = i + u < i1;
                k[u] = in ? Z[i + u] - 1 : -1;
                v[u] = in ? X[(size_t)(i + u) * R + r] : CUDART_NAN;
            }
#pragma unroll
            for (int u = 0; u < ST_UNROLL; u++) {
                if ((unsigned)k[u] >= (unsigned)K || cs_isna(v[u])) continue;
                if (use_smem) {
                    s_acc[k[u] * R + r] += v[u]; // thread owns column r: no race
                    s_acc[K * R + k[u] * R + r] += 1.0;
                } else {
                    atomicAdd(sum_x + (size_t)k[u] * R + r, v[u]);
                    atomicAdd(cnt_x + (size_t)k[u] * R + r, 1.0);
                }
            }
        }
    }
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) { // cluster sizes
        const int k = Z[i] - 1;
        if (k >= 0 && k < K) atomicAdd(nk + k, 1.0);
    }
    if (use_smem) {
        __syncthreads();
        for (int e = threadIdx.x; e < K * R; e += blockDim.x) {
            if (s_acc[K * R + e] != 0.0) {
                atomicAdd(sum_x + e, s_acc[e]);
                atomicAdd(cnt_x + e, s_acc[K * R + e]);
            }
        }
    }
}

__global__ void __launch_bounds__(1024)
sigma_stats_kernel(int N, int R, int K, const double *__restrict__ X, const int *__restrict__ Z,
                   const double *__restrict__ U, double *__restrict__ sq, double *__restrict__ cnt)
{
    const int i0 = blockIdx.x * ST_CHUNK, i1 = min(N, i0 + ST_CHUNK);
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        double s = 0.0, c = 0.0;
        for (int i = i0; i < i1; i += ST_UNROLL) { // ST_UNROLL (x, u) pairs in flight, cells in order
            double x[ST_UNROLL], u[ST_UNROLL];
            bool lab[ST_UNROLL];
#pragma unroll
            for (int q = 0; q < ST_UNROLL; q++) {
                const bool in = i + q < i1;
                const int k = in ? Z[i + q] - 1 : -1;
                lab[q] = (unsigned)k < (unsigned)K;
                x[q] = in ? X[(size_t)(i + q) * R + r] : CUDART_NAN;
                u[q] = lab[q] ? U[(size_t)k * R + r] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < ST_UNROLL; q++) {
                if (cs_isna(x[q])) continue;
                c += 1.0;
                if (!lab[q]) continue;
                const double d = x[q] - u[q];
                if (!cs_isna(d)) s += d * d;
            }
        }
        atomicAdd(sq + r, s);
        atomicAdd(cnt + r, c);
    }
}

// per-gene sums and sums of squares over the cells of a CSC (gene x cell) matrix — the inputs
// of the variance filter of EstRegionCov (R/EstRegionCov.R:42-48) without densifying the matrix.
// Integer counts give exact, order-free FP64 sums (below 2^53).
__global__ void gene_moments_kernel(int64_t nnz, const int32_t *__restrict__ rowidx, const double *__restrict__ x,
                                    double *__restrict__ sum, double *__restrict__ sumsq)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[e];
        const int g = rowidx[e];
        atomicAdd(sum + g, v);
        atomicAdd(sumsq + g, v * v);
    }
}

} // namespace

// host (R layout) form: the dgCMatrix slots @i @x of a G x N matrix with nnz entries
extern "C" int cs_gene_moments(int G, int64_t nnz, const int32_t *rowidx, const double *x, double *sum, double *sumsq)
{
    CS_REQUIRE(G >= 0 && nnz >= 0 && sum && sumsq && (nnz == 0 || (rowidx && x)), "cs_gene_moments: bad arguments");
    for (int64_t e = 0; e < nnz; e++)
        CS_REQUIRE(rowidx[e] >= 0 && rowidx[e] < G, "cs_gene_moments: rowidx[%lld]=%d outside [0,%d)", (long long)e, rowidx[e], G);
    DevBuf<int32_t> d_i;
    DevBuf<double> d_x, d_s, d_q;
    CS_CUDA(d_i.alloc((size_t)nnz));
    CS_CUDA(d_x.alloc((size_t)nnz));
    CS_CUDA(d_s.alloc((size_t)G));
    CS_CUDA(d_q.alloc((size_t)G));
    CS_CUDA(cudaMemset(d_s.p, 0, sizeof(double) * (size_t)(G > 0 ? G : 1)));
    CS_CUDA(cudaMemset(d_q.p, 0, sizeof(double) * (size_t)(G > 0 ? G : 1)));
    if (nnz) {
        if (int rc = cs_h2d_staged(d_i.p, rowidx, sizeof(int32_t) * (size_t)nnz)) return rc;
        if (int rc = cs_h2d_staged(d_x.p, x, sizeof(double) * (size_t)nnz)) return rc;
        gene_moments_kernel<<<cs_num_sms() * 8, 256>>>(nnz, d_i.p, d_x.p, d_s.p, d_q.p);
        CS_CUDA(cudaGetLastError());
    }
    if (G) {
        CS_CUDA(cudaMemcpy(sum, d_s.p, sizeof(double) * (size_t)G, cudaMemcpyDeviceToHost));
        CS_CUDA(cudaMemcpy(sumsq, d_q.p, sizeof(double) * (size_t)G, cudaMemcpyDeviceToHost));
    }
    return CS_OK;
}

extern "C" int cs_cluster_stats_dev(int N, int R, int K, const double *d_X, const int *d_Z, double *d_sum_x,
                                    double *d_cnt_x, double *d_nk, void *stream)
{
    CS_REQUIRE(N >= 0 && R >= 1 && K >= 1 && d_sum_x && d_cnt_x && d_nk, "cs_cluster_stats_dev: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaMemsetAsync(d_sum_x, 0, sizeof(double) * (size_t)K * R, st));
    CS_CUDA(cudaMemsetAsync(d_cnt_x, 0, sizeof(double) * (size_t)K * R, st));
    CS_CUDA(cudaMemsetAsync(d_nk, 0, sizeof(double) * (size_t)K, st));
    if (N == 0) return CS_OK;
    const size_t smem = sizeof(double) * 2 * (size_t)K * R;
    const int use_smem = smem <= 160 * 1024;
    if (use_smem && smem > 48 * 1024)
        CS_CUDA(cudaFuncSetAttribute(cluster_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cluster_stats_kernel<<<(N + ST_CHUNK - 1) / ST_CHUNK, st_threads(R), use_smem ? smem : 0, st>>>(
        N, R, K, d_X, d_Z, d_sum_x, d_cnt_x, d_nk, use_smem);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int cs_sigma_stats_dev(int N, int R, int K, const double *d_X, const int *d_Z, const double *d_U,
                                  double *d_sq, double *d_cnt, void *stream)
{
    CS_REQUIRE(N >= 0 && R >= 1 && K >= 1 && d_sq && d_cnt, "cs_sigma_stats_dev: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaMemsetAsync(d_sq, 0, sizeof(double) * (size_t)R, st));
    CS_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(double) * (size_t)R, st));
    if (N == 0) return CS_OK;
    sigma_stats_kernel<<<(N + ST_CHUNK - 1) / ST_CHUNK, st_threads(R), 0, st>>>(N, R, K, d_X, d_Z, d_U, d_sq, d_cnt);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// ---- host (R layout) forms: what a `.Call` stub can bind (AssignCluster :46-53, :67-77) ---------
template <typename T>
int cs_transpose(int rows, int cols, const T *d_in, T *d_out, cudaStream_t st);

// X: N x R column-major; Z: N labels (labels outside 1..K are skipped, like cells AssignCluster has
// set aside); sum_x / cnt_x: K x R column-major; n_k: K
extern "C" int cs_cluster_stats(int N, int R, int K, const double *X, const int *Z, double *sum_x, double *cnt_x,
                                double *n_k)
{
    CS_REQUIRE(N >= 0 && R >= 1 && K >= 1 && sum_x && cnt_x && n_k && (N == 0 || (X && Z)), "cs_cluster_stats: bad arguments");
    DevBuf<double> dXc, dX, dS, dC, dSc, dCc, dN;
    DevBuf<int> dZ;
    const size_t nx = (size_t)N * R, nk = (size_t)K * R;
    CS_CUDA(dXc.alloc(nx));
    CS_CUDA(dX.alloc(nx));
    CS_CUDA(dZ.alloc(N));
    CS_CUDA(dS.alloc(nk));
    CS_CUDA(dC.alloc(nk));
    CS_CUDA(dSc.alloc(nk));
    CS_CUDA(dCc.alloc(nk));
    CS_CUDA(dN.alloc(K));
    int rc;
    if (N) {
        if (int rc0 = cs_h2d_staged(dXc.p, X, sizeof(double) * nx)) return rc0;
        CS_CUDA(cudaMemcpy(dZ.p, Z, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
        if ((rc = cs_transpose<double>(R, N, dXc.p, dX.p, 0))) return rc; // column-major N x R -> cell-major
    }
    if ((rc = cs_cluster_stats_dev(N, R, K, dX.p, dZ.p, dS.p, dC.p, dN.p, nullptr))) return rc;
    // row-major K x R -> column-major K x R (== row-major R x K)
    if ((rc = cs_transpose<double>(K, R, dS.p, dSc.p, 0))) return rc;
    if ((rc = cs_transpose<double>(K, R, dC.p, dCc.p, 0))) return rc;
    CS_CUDA(cudaMemcpy(sum_x, dSc.p, sizeof(double) * nk, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(cnt_x, dCc.p, sizeof(double) * nk, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(n_k, dN.p, sizeof(double) * (size_t)K, cudaMemcpyDeviceToHost));
    return CS_OK;
}

// U: K x R column-major; sq[r] = sum over cells with a label in 1..K of (x - U[z, r])^2 (NA skipped),
// cnt[r] = number of non-NA x in segment r over ALL cells (R/BayesNonparCluster.R:234)
extern "C" int cs_sigma_stats(int N, int R, int K, const double *X, const int *Z, const double *U, double *sq,
                              double *cnt)
{
    CS_REQUIRE(N >= 0 && R >= 1 && K >= 1 && U && sq && cnt && (N == 0 || (X && Z)), "cs_sigma_stats: bad arguments");
    DevBuf<double> dXc, dX, dUc, dU, dQ, dC;
    DevBuf<int> dZ;
    const size_t nx = (size_t)N * R, nk = (size_t)K * R;
    CS_CUDA(dXc.alloc(nx));
    CS_CUDA(dX.alloc(nx));
    CS_CUDA(dZ.alloc(N));
    CS_CUDA(dUc.alloc(nk));
    CS_CUDA(dU.alloc(nk));
    CS_CUDA(dQ.alloc(R));
    CS_CUDA(dC.alloc(R));
    int rc;
    if (N) {
        if (int rc0 = cs_h2d_staged(dXc.p, X, sizeof(double) * nx)) return rc0;
        CS_CUDA(cudaMemcpy(dZ.p, Z, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
        if ((rc = cs_transpose<double>(R, N, dXc.p, dX.p, 0))) return rc;
    }
    CS_CUDA(cudaMemcpy(dUc.p, U, sizeof(double) * nk, cudaMemcpyHostToDevice));
    if ((rc = cs_transpose<double>(R, K, dUc.p, dU.p, 0))) return rc;
    if ((rc = cs_sigma_stats_dev(N, R, K, dX.p, dZ.p, dU.p, dQ.p, dC.p, nullptr))) return rc;
    CS_CUDA(cudaMemcpy(sq, dQ.p, sizeof(double) * (size_t)R, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(cnt, dC.p, sizeof(double) * (size_t)R, cudaMemcpyDeviceToHost));
    return CS_OK;
}

// allele variant lives in cs_allele.cu
