// cs_regioncov.cu — EstRegionCov on a device-resident gene x cell matrix.
//
// Reference: R/EstRegionCov.R:36-48 (align the cells, per-gene variance filter) and :80-159 (per
// segment: column sums of the genes in the segment, cells with a positive sum, library sizes,
// the two medians and the ratios).  The reference densifies the matrix and loops over segments;
// here the matrix is uploaded ONCE (cs_csc_upload, optionally gathering a subset / permutation of
// the columns on the device), and the moments (cs_csc_gene_moments) and the whole per-segment
// computation (cs_csc_region_cov: aggregation through the binning plan, library sizes, exact
// medians by radix select, ratios) run on that copy.
//
// Exactness: counts are integers, so every sum is exact and order-free; the medians are exact
// order statistics (the mean of the two middle ones for an even count, like R's median); the
// ratios are the same IEEE double divisions in the same order as the reference's statements.
#include "cs_common.cuh"

#include <math_constants.h>
#include <vector>

struct cs_csc {
    int G = 0, N = 0;
    int64_t nnz = 0;
    DevBuf<int64_t> colptr;
    DevBuf<int32_t> rowidx;
    DevBuf<double> x;
};

namespace {

__global__ void rc_validate_rows_kernel(int64_t nnz, const int32_t *__restrict__ rowidx, int G, long long *__restrict__ bad)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const int g = rowidx[e];
        if (g < 0 || g >= G) atomicMax((unsigned long long *)bad, (unsigned long long)(e + 1));
    }
}

// column j of the output = column sel[j] of the input (one warp per column)
__global__ void __launch_bounds__(256)
rc_gather_columns_kernel(int Nout, const int32_t *__restrict__ sel, const int64_t *__restrict__ in_colptr,
                         const int32_t *__restrict__ in_rowidx, const double *__restrict__ in_x,
                         const int64_t *__restrict__ out_colptr, int32_t *__restrict__ out_rowidx, double *__restrict__ out_x)
{
    const int lane = threadIdx.x & 31, nw = gridDim.x * (blockDim.x >> 5);
    for (int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < Nout; j += nw) {
        const int c = sel[j];
        const int64_t s = in_colptr[c], n = in_colptr[c + 1] - s, d = out_colptr[j];
        for (int64_t e = lane; e < n; e += 32) {
            out_rowidx[d + e] = in_rowidx[s + e];
            out_x[d + e] = in_x[s + e];
        }
    }
}

__global__ void rc_moments_kernel(int64_t nnz, const int32_t *__restrict__ rowidx, const double *__restrict__ x,
                                  double *__restrict__ sum, double *__restrict__ sumsq)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const double v = x[e];
        const int g = rowidx[e];
        atomicAdd(sum + g, v);
        atomicAdd(sumsq + g, v * v);
    }
}

// sparse S x N (per-cell columns) -> dense sums[s * N + c]
__global__ void __launch_bounds__(256)
rc_scatter_dense_kernel(int N, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                        const double *__restrict__ x, double *__restrict__ dense)
{
    const int lane = threadIdx.x & 31, nw = gridDim.x * (blockDim.x >> 5);
    for (int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < N; c += nw) {
        const int64_t e0 = colptr[c], e1 = colptr[c + 1];
        for (int64_t e = e0 + lane; e < e1; e += 32) dense[(size_t)rowidx[e] * N + c] = x[e];
    }
}

// library size of every cell: the sum of its counts over the flagged genes (one warp per column)
__global__ void __launch_bounds__(256)
rc_library_kernel(int N, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                  const double *__restrict__ x, const unsigned char *__restrict__ flag, double *__restrict__ alpha)
{
    const int lane = threadIdx.x & 31, nw = gridDim.x * (blockDim.x >> 5);
    for (int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < N; c += nw) {
        const int64_t e0 = colptr[c], e1 = colptr[c + 1];
        double s = 0.0;
        for (int64_t e = e0 + lane; e < e1; e += 32)
            if (flag[rowidx[e]]) s += x[e];
#pragma unroll
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(CS_FULL, s, off);
        if (lane == 0) alpha[c] = s;
    }
}

__device__ __forceinline__ unsigned long long rc_key(double v)
{
    // order-preserving map of the doubles onto unsigned integers (-0 == +0 is not needed here)
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double rc_unkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    return __longlong_as_double((long long)b);
}

constexpr int RC_THREADS = 512;

// The median of {val(c) : pred(c)} over c in [0, N) by an 8-pass radix select on the keys, done
// by one CTA.  `val` may be evaluated several times per cell (it reads two L2-resident arrays).
// Returns NaN when a selected value is NaN (R's median without na.rm) or when nothing is selected.
template <typename Pred, typename Val>
__device__ double rc_cta_median(int N, Pred pred, Val val, unsigned *hist /* 256 */, unsigned long long *sh /* 4 */)
{
    const int tid = threadIdx.x;
    // count, and look for NaN
    if (tid < 4) sh[tid] = 0ull;
    __syncthreads();
    unsigned cnt = 0, nan = 0;
    for (int c = tid; c < N; c += RC_THREADS)
        if (pred(c)) {
            cnt++;
            const double v = val(c);
            nan |= (v != v);
        }
    atomicAdd(&sh[0], (unsigned long long)cnt);
    if (nan) atomicOr(&sh[1], 1ull);
    __syncthreads();
    const unsigned long long n = sh[0];
    const bool any_nan = sh[1] != 0ull;
    __syncthreads();
    if (n == 0ull || any_nan) return CUDART_NAN;
    unsigned long long k = (n - 1) >> 1; // lower middle, 0-based
    unsigned long long prefix = 0ull;
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int i = tid; i < 256; i += RC_THREADS) hist[i] = 0u;
        __syncthreads();
        for (int c = tid; c < N; c += RC_THREADS)
            if (pred(c)) {
                const unsigned long long key = rc_key(val(c));
                if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
        __syncthreads();
        if (tid == 0) {
            unsigned long long acc = 0ull;
            int d = 0;
            for (; d < 256; d++) {
                if (acc + hist[d] > k) break;
                acc += hist[d];
            }
            sh[2] = (unsigned long long)d;
            sh[3] = acc;
        }
        __syncthreads();
        prefix |= sh[2] << shift;
        k -= sh[3];
        __syncthreads();
    }
    const double lo = rc_unkey(prefix);
    if (n & 1ull) return lo;
    // even count: the next order statistic is `lo` again when enough copies of it exist, else the
    // smallest selected value above it
    if (tid == 0) {
        sh[0] = 0ull;
        sh[1] = ~0ull;
    }
    __syncthreads();
    unsigned le = 0;
    unsigned long long mn = ~0ull;
    for (int c = tid; c < N; c += RC_THREADS)
        if (pred(c)) {
            const unsigned long long key = rc_key(val(c));
            if (key <= prefix) le++;
            else mn = min(mn, key);
        }
    atomicAdd(&sh[0], (unsigned long long)le);
    atomicMin(&sh[1], mn);
    __syncthreads();
    const unsigned long long nle = sh[0], up = sh[1];
    __syncthreads();
    const double hi = (nle > ((n - 1) >> 1) + 1) ? lo : rc_unkey(up);
    return (lo + hi) / 2.0;
}

// One CTA per segment (R/EstRegionCov.R:88-159):
//   sel    = cells with sums > 0                                   (:93)
//   med    = median(alpha[sel])                                    (:141,:146)
//   csum   = sums[normal & sel] / (alpha / med)                    (:153)
//   deltas = (sums[out & sel] / (alpha / med)) / median(csum)      (:154-159), NaN elsewhere
// celltype: 0 normal, 1 tumor, anything else neither; include: 0 tumor, 1 control, 2 all.
__global__ void __launch_bounds__(RC_THREADS)
rc_segment_kernel(int S, int N, const double *__restrict__ sums, const double *__restrict__ alpha,
                  const int *__restrict__ celltype, int include, double *__restrict__ deltas, int *__restrict__ nsel)
{
    __shared__ unsigned hist[256];
    __shared__ unsigned long long sh[4];
    for (int s = blockIdx.x; s < S; s += gridDim.x) {
        const double *__restrict__ row = sums + (size_t)s * N;
        auto sel = [&](int c) { return row[c] > 0.0; };
        const double med = rc_cta_median(N, sel, [&](int c) { return alpha[c]; }, hist, sh);
        auto ratio = [&](int c) { return row[c] / (alpha[c] / med); };
        const double medc = rc_cta_median(N, [&](int c) { return row[c] > 0.0 && celltype[c] == 0; }, ratio, hist, sh);
        int cnt = 0;
        for (int c = threadIdx.x; c < N; c += RC_THREADS) {
            const bool in = row[c] > 0.0;
            cnt += in;
            const int t = celltype[c];
            const bool out = in && (include == 2 || (include == 0 && t == 1) || (include == 1 && t == 0));
            deltas[(size_t)s * N + c] = out ? ratio(c) / medc : CUDART_NAN;
        }
        if (threadIdx.x == 0) sh[0] = 0ull;
        __syncthreads();
        atomicAdd(&sh[0], (unsigned long long)cnt);
        __syncthreads();
        if (threadIdx.x == 0) nsel[s] = (int)sh[0];
        __syncthreads();
    }
}

} // namespace

extern "C" int cs_csc_upload(int G, int N, const int32_t *colptr, const int32_t *rowidx, const double *x, int Nsel,
                             const int32_t *col_sel, cs_csc **out)
{
    CS_REQUIRE(out, "cs_csc_upload: NULL output argument");
    *out = nullptr;
    CS_REQUIRE(G >= 0 && N >= 0 && colptr && colptr[0] == 0, "cs_csc_upload: bad dimensions / colptr[0] must be 0");
    for (int c = 0; c < N; c++) CS_REQUIRE(colptr[c] <= colptr[c + 1], "colptr not monotone at column %d", c);
    const int64_t nnz = colptr[N];
    CS_REQUIRE(nnz == 0 || (rowidx && x), "cs_csc_upload: NULL slots");
    CS_REQUIRE(col_sel || Nsel == 0 || Nsel == N, "cs_csc_upload: Nsel without col_sel");
    cs_csc *m = new cs_csc();
    struct Guard {
        cs_csc *p;
        ~Guard() { delete p; }
    } guard{m};
    std::vector<int64_t> cp64((size_t)N + 1);
    for (int c = 0; c <= N; c++) cp64[c] = colptr[c];
    DevBuf<int64_t> d_cp;
    DevBuf<int32_t> d_ri;
    DevBuf<double> d_x;
    CS_CUDA(d_cp.alloc((size_t)N + 1));
    CS_CUDA(d_ri.alloc((size_t)nnz));
    CS_CUDA(d_x.alloc((size_t)nnz));
    CS_CUDA(cudaMemcpy(d_cp.p, cp64.data(), sizeof(int64_t) * cp64.size(), cudaMemcpyHostToDevice));
    int rc;
    if (nnz) {
        if ((rc = cs_h2d_staged(d_ri.p, rowidx, sizeof(int32_t) * (size_t)nnz))) return rc;
        if ((rc = cs_h2d_staged(d_x.p, x, sizeof(double) * (size_t)nnz))) return rc;
        DevBuf<long long> d_bad;
        CS_CUDA(d_bad.alloc(1));
        CS_CUDA(cudaMemset(d_bad.p, 0, sizeof(long long)));
        rc_validate_rows_kernel<<<cs_num_sms() * 8, 256>>>(nnz, d_ri.p, G, d_bad.p);
        long long bad = 0;
        CS_CUDA(cudaMemcpy(&bad, d_bad.p, sizeof(long long), cudaMemcpyDeviceToHost));
        CS_REQUIRE(bad == 0, "rowidx[%lld]=%d outside [0,%d)", bad - 1, rowidx[bad - 1], G);
    }
    m->G = G;
    if (!col_sel) {
        m->N = N;
        m->nnz = nnz;
        std::swap(m->colptr.p, d_cp.p), std::swap(m->colptr.n, d_cp.n);
        std::swap(m->rowidx.p, d_ri.p), std::swap(m->rowidx.n, d_ri.n);
        std::swap(m->x.p, d_x.p), std::swap(m->x.n, d_x.n);
    } else {
        // the selected columns, in the caller's order (mtx[, match(celltype0[, 1], barcodes[, 1])])
        std::vector<int64_t> ocp((size_t)Nsel + 1, 0);
        for (int j = 0; j < Nsel; j++) {
            CS_REQUIRE(col_sel[j] >= 0 && col_sel[j] < N, "cs_csc_upload: col_sel[%d]=%d outside [0,%d)", j, col_sel[j], N);
            ocp[j + 1] = ocp[j] + (cp64[col_sel[j] + 1] - cp64[col_sel[j]]);
        }
        DevBuf<int32_t> d_sel;
        CS_CUDA(d_sel.alloc((size_t)Nsel));
        CS_CUDA(m->colptr.alloc((size_t)Nsel + 1));
        CS_CUDA(m->rowidx.alloc((size_t)ocp[Nsel]));
        CS_CUDA(m->x.alloc((size_t)ocp[Nsel]));
        CS_CUDA(cudaMemcpy(m->colptr.p, ocp.data(), sizeof(int64_t) * ocp.size(), cudaMemcpyHostToDevice));
        if (Nsel) {
            CS_CUDA(cudaMemcpy(d_sel.p, col_sel, sizeof(int32_t) * (size_t)Nsel, cudaMemcpyHostToDevice));
            rc_gather_columns_kernel<<<cs_num_sms() * 8, 256>>>(Nsel, d_sel.p, d_cp.p, d_ri.p, d_x.p, m->colptr.p,
                                                                m->rowidx.p, m->x.p);
            CS_CUDA(cudaGetLastError());
            CS_CUDA(cudaDeviceSynchronize());
        }
        m->N = Nsel;
        m->nnz = ocp[Nsel];
    }
    *out = m;
    guard.p = nullptr;
    return CS_OK;
}

extern "C" void cs_csc_free(cs_csc *m) { delete m; }

extern "C" int cs_csc_gene_moments(cs_csc *m, double *sum, double *sumsq)
{
    CS_REQUIRE(m && sum && sumsq, "cs_csc_gene_moments: NULL argument");
    const size_t G = (size_t)m->G;
    DevBuf<double> d_s, d_q;
    CS_CUDA(d_s.alloc(G));
    CS_CUDA(d_q.alloc(G));
    CS_CUDA(cudaMemset(d_s.p, 0, sizeof(double) * (G ? G : 1)));
    CS_CUDA(cudaMemset(d_q.p, 0, sizeof(double) * (G ? G : 1)));
    if (m->nnz) {
        rc_moments_kernel<<<cs_num_sms() * 8, 256>>>(m->nnz, m->rowidx.p, m->x.p, d_s.p, d_q.p);
        CS_CUDA(cudaGetLastError());
    }
    if (G) {
        CS_CUDA(cudaMemcpy(sum, d_s.p, sizeof(double) * G, cudaMemcpyDeviceToHost));
        CS_CUDA(cudaMemcpy(sumsq, d_q.p, sizeof(double) * G, cudaMemcpyDeviceToHost));
    }
    return CS_OK;
}

// ---- device-resident halves of cs_csc_region_cov, for cell-sharded runs (the medians need every
// cell: the shards' sums and library sizes are all-gathered in between, clonalscope_b200/distributed.py)

// per-segment column sums (d_sums: S x N, row s contiguous) and library sizes (d_alpha: N) of the
// handle's cells
extern "C" int cs_csc_region_sums_dev(cs_csc *m, int S, const int32_t *map_ptr, const int32_t *map_seg,
                                      const unsigned char *lib_flag, double *d_sums, double *d_alpha, void *stream)
{
    CS_REQUIRE(m && S >= 1 && map_ptr && lib_flag && d_sums && d_alpha, "cs_csc_region_sums_dev: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = m->N, G = m->G;
    if (N == 0) return CS_OK;
    cs_bin_plan *plan = nullptr;
    int rc = cs_bin_plan_create(G, S, map_ptr, map_seg, &plan);
    if (rc) return rc;
    struct PlanGuard {
        cs_bin_plan *p;
        ~PlanGuard() { cs_bin_plan_destroy(p); }
    } guard{plan};
    DevBuf<int64_t> d_ocp;
    DevBuf<int32_t> d_ori;
    DevBuf<double> d_ox;
    DevBuf<unsigned char> d_flag;
    CS_CUDA(d_ocp.alloc((size_t)N + 1));
    int64_t need = 0;
    if ((rc = cs_bin_counts_dev(plan, N, m->colptr.p, m->rowidx.p, m->x.p, d_ocp.p, nullptr, nullptr, 0, &need, stream))) return rc;
    for (int attempt = 0;; attempt++) {
        CS_CUDA(d_ori.alloc((size_t)need));
        CS_CUDA(d_ox.alloc((size_t)need));
        int64_t got = 0;
        rc = cs_bin_counts_dev(plan, N, m->colptr.p, m->rowidx.p, m->x.p, d_ocp.p, d_ori.p, d_ox.p, need, &got, stream);
        if (rc == CS_ERR_UCAP && got > need && attempt == 0) { // (columns with unsorted rows: the first count was void)
            need = got;
            continue;
        }
        if (rc) return rc;
        break;
    }
    CS_CUDA(d_flag.alloc((size_t)(G ? G : 1)));
    CS_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(double) * (size_t)S * N, st));
    if (G) CS_CUDA(cudaMemcpyAsync(d_flag.p, lib_flag, (size_t)G, cudaMemcpyHostToDevice, st));
    const int grid = cs_num_sms() * 8;
    rc_scatter_dense_kernel<<<grid, 256, 0, st>>>(N, d_ocp.p, d_ori.p, d_ox.p, d_sums);
    rc_library_kernel<<<grid, 256, 0, st>>>(N, m->colptr.p, m->rowidx.p, m->x.p, d_flag.p, d_alpha);
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaStreamSynchronize(st)); // (the scratch above is freed on return)
    return CS_OK;
}

// R/EstRegionCov.R:141-159 for S segments over N cells whose sums and library sizes lie on the device
extern "C" int cs_region_ratios_dev(int S, int N, const double *d_sums, const double *d_alpha, const int *d_celltype,
                                    int include, double *d_deltas, int *d_nsel, void *stream)
{
    CS_REQUIRE(S >= 1 && N >= 0 && d_sums && d_alpha && d_celltype && d_deltas && d_nsel, "cs_region_ratios_dev: NULL argument");
    CS_REQUIRE(include >= 0 && include <= 2, "cs_region_ratios_dev: include must be 0 (tumor), 1 (control) or 2 (all)");
    if (N == 0) return CS_OK;
    rc_segment_kernel<<<S < 4 * cs_num_sms() ? S : 4 * cs_num_sms(), RC_THREADS, 0, (cudaStream_t)stream>>>(
        S, N, d_sums, d_alpha, d_celltype, include, d_deltas, d_nsel);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int cs_csc_region_cov(cs_csc *m, int S, const int32_t *map_ptr, const int32_t *map_seg,
                                 const unsigned char *lib_flag, const int *celltype, int include, double *deltas,
                                 double *sums, double *alpha, int *nsel)
{
    CS_REQUIRE(m && S >= 1 && map_ptr && lib_flag && celltype && deltas && alpha && nsel, "cs_csc_region_cov: NULL argument");
    CS_REQUIRE(include >= 0 && include <= 2, "cs_csc_region_cov: include must be 0 (tumor), 1 (control) or 2 (all)");
    const int N = m->N;
    if (N == 0) return CS_OK;
    const size_t SN = (size_t)S * N;
    DevBuf<double> d_sums, d_deltas, d_alpha;
    DevBuf<int> d_type, d_nsel;
    CS_CUDA(d_sums.alloc(SN));
    CS_CUDA(d_deltas.alloc(SN));
    CS_CUDA(d_alpha.alloc((size_t)N));
    CS_CUDA(d_type.alloc((size_t)N));
    CS_CUDA(d_nsel.alloc((size_t)S));
    CS_CUDA(cudaMemcpy(d_type.p, celltype, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
    int rc;
    if ((rc = cs_csc_region_sums_dev(m, S, map_ptr, map_seg, lib_flag, d_sums.p, d_alpha.p, nullptr))) return rc;
    if ((rc = cs_region_ratios_dev(S, N, d_sums.p, d_alpha.p, d_type.p, include, d_deltas.p, d_nsel.p, nullptr))) return rc;
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaDeviceSynchronize());
    if ((rc = cs_d2h_staged(deltas, d_deltas.p, sizeof(double) * SN))) return rc;
    if (sums && (rc = cs_d2h_staged(sums, d_sums.p, sizeof(double) * SN))) return rc;
    CS_CUDA(cudaMemcpy(alpha, d_alpha.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost));
    CS_CUDA(cudaMemcpy(nsel, d_nsel.p, sizeof(int) * (size_t)S, cudaMemcpyDeviceToHost));
    return CS_OK;
}
