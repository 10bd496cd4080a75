// cs_loglik.cu — fixed-state cell x cluster Gaussian log-likelihood matrix, layout
// transposes, and the posterior majority-vote tally.
//
//   LL[i,k] = sum_r dnorm(X[i,r]; U[k,r], sigma[r], log=TRUE)   (NA segments skipped)
// reference: R/BayesNonparCluster.R:91-96 (Z0 init), :193 (scoring), R/AssignCluster.R:58.
// One warp per cell: lanes stride the segments (coalesced 256 B row reads), the row kept in
// registers and scored against four clusters per pass, FP64 accumulation, folded warp sums.  Not a tensor-core workload: the NaN mask and the
// 1e-6 criterion rule out the expanded-square GEMM form.
#include "cs_common.cuh"

#include <math_constants.h>
#include <stdlib.h>
#include <vector>

namespace {

constexpr int LL_THREADS = 256;
constexpr int LL_KT = 4; // clusters scored per pass over a cell's registers

// any R: a pass over the row per cluster (the row comes from L1 after the first)
__global__ void __launch_bounds__(LL_THREADS)
loglik_matrix_generic_kernel(int N, int R, int K, const double *__restrict__ X, const double *__restrict__ U,
                             const double *__restrict__ sigma, double *__restrict__ LL)
{
    extern __shared__ double s_ll[]; // isig[R], cterm[R] = log(sigma)+ln sqrt(2 pi)
    double *s_isig = s_ll, *s_c = s_ll + R;
    for (int r = threadIdx.x; r < R; r += LL_THREADS) {
        const double s = sigma[r];
        s_isig[r] = 1.0 / s;
        s_c[r] = CS_LN_SQRT_2PI + log(s);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = LL_THREADS / 32;
    for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < N; i += gridDim.x * warps_per_block) {
        const double *x = X + (size_t)i * R;
        // constant part: -(sum over non-NA segments of log sigma + ln sqrt(2 pi))
        double cpart = 0.0;
        for (int r = lane; r < R; r += 32)
            if (!cs_isna(x[r])) cpart += s_c[r];
        cpart = cs_warp_sum_bfly(cpart);
        for (int k = 0; k < K; k++) {
            const double *u = U + (size_t)k * R;
            double q = 0.0;
            for (int r = lane; r < R; r += 32) {
                const double xv = x[r];
                if (!cs_isna(xv)) {
                    const double d = (xv - __ldg(u + r)) * s_isig[r];
                    q += d * d;
                }
            }
            q = cs_warp_sum_bfly(q);
            if (lane == 0) LL[(size_t)i * K + k] = -(cpart + 0.5 * q);
        }
    }
}

// R <= 32 * NCH: the cell's row and 1/sigma stay in registers, LL_KT clusters per pass for independent FMA
// chains, and the LL_KT warp sums are folded together (6 shuffles instead of 5 per cluster).
template <int NCH>
__global__ void __launch_bounds__(LL_THREADS)
loglik_matrix_kernel(int N, int R, int K, const double *__restrict__ X, const double *__restrict__ U,
                     const double *__restrict__ sigma, double *__restrict__ LL)
{
    extern __shared__ double s_ll[];
    double *s_isig = s_ll, *s_c = s_ll + R;
    for (int r = threadIdx.x; r < R; r += LL_THREADS) {
        const double s = sigma[r];
        s_isig[r] = 1.0 / s;
        s_c[r] = CS_LN_SQRT_2PI + log(s);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = LL_THREADS / 32;
    const int kslot = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1); // which of the LL_KT sums this lane ends up with
    for (int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5); i < N; i += gridDim.x * warps_per_block) {
        const double *xrow = X + (size_t)i * R;
        double x[NCH], is[NCH];
        double cpart = 0.0;
        unsigned okmask = 0u; // chunks whose segment exists and is not NA for this cell
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int r = lane + 32 * c;
            const double xv = (r < R) ? __ldcs(xrow + r) : CUDART_NAN;
            const bool ok = !cs_isna(xv);
            x[c] = xv;
            is[c] = ok ? s_isig[r] : 0.0;
            if (ok) {
                cpart += s_c[r];
                okmask |= 1u << c;
            }
        }
        cpart = cs_warp_sum_bfly(cpart);
        for (int k0 = 0; k0 < K; k0 += LL_KT) {
            double q[LL_KT];
#pragma unroll
            for (int j = 0; j < LL_KT; j++) q[j] = 0.0;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const int r = min(lane + 32 * c, R - 1);
#pragma unroll
                for (int j = 0; j < LL_KT; j++) {
                    const int k = min(k0 + j, K - 1);
                    const double d = (x[c] - __ldg(U + (size_t)k * R + r)) * is[c];
                    if (okmask >> c & 1u) q[j] = fma(d, d, q[j]);
                }
            }
            // fold: halves of the warp keep two sums each, then quarters one each
            const bool up16 = lane & 16, up8 = lane & 8;
            const double s0 = up16 ? q[0] : q[2], s1 = up16 ? q[1] : q[3]; // what goes to the other half
            double a0 = (up16 ? q[2] : q[0]) + __shfl_xor_sync(CS_FULL, s0, 16);
            double a1 = (up16 ? q[3] : q[1]) + __shfl_xor_sync(CS_FULL, s1, 16);
            double b = (up8 ? a1 : a0) + __shfl_xor_sync(CS_FULL, up8 ? a0 : a1, 8);
            b += __shfl_xor_sync(CS_FULL, b, 4);
            b += __shfl_xor_sync(CS_FULL, b, 2);
            b += __shfl_xor_sync(CS_FULL, b, 1);
            const int k = k0 + kslot;
            if ((lane & 7) == 0 && k < K) LL[(size_t)i * K + k] = -(cpart + 0.5 * b);
        }
    }
}

// Two cells per warp pass: every cluster value loaded from L1 serves both cells (with many clusters
// the one-cell kernel is bound by its load instructions as much as by the FP64 pipe).  Per cell the
// additions happen in the same order as in loglik_matrix_kernel: the results are bit-identical.
template <int NCH>
__global__ void __launch_bounds__(LL_THREADS)
loglik_matrix2_kernel(int N, int R, int K, const double *__restrict__ X, const double *__restrict__ U,
                      const double *__restrict__ sigma, double *__restrict__ LL)
{
    extern __shared__ double s_ll[];
    double *s_isig = s_ll, *s_c = s_ll + R;
    for (int r = threadIdx.x; r < R; r += LL_THREADS) {
        const double s = sigma[r];
        s_isig[r] = 1.0 / s;
        s_c[r] = CS_LN_SQRT_2PI + log(s);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = LL_THREADS / 32;
    const int kslot = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    const int npairs = (N + 1) / 2;
    auto fold = [&](const double (&q)[LL_KT]) {
        const bool up16 = lane & 16, up8 = lane & 8;
        const double s0 = up16 ? q[0] : q[2], s1 = up16 ? q[1] : q[3];
        double a0 = (up16 ? q[2] : q[0]) + __shfl_xor_sync(CS_FULL, s0, 16);
        double a1 = (up16 ? q[3] : q[1]) + __shfl_xor_sync(CS_FULL, s1, 16);
        double b = (up8 ? a1 : a0) + __shfl_xor_sync(CS_FULL, up8 ? a0 : a1, 8);
        b += __shfl_xor_sync(CS_FULL, b, 4);
        b += __shfl_xor_sync(CS_FULL, b, 2);
        b += __shfl_xor_sync(CS_FULL, b, 1);
        return b;
    };
    for (int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < npairs; p += gridDim.x * warps_per_block) {
        const int i0 = 2 * p, i1 = min(2 * p + 1, N - 1); // (an odd N: the last pair scores its cell twice, writes it once)
        const double *xr0 = X + (size_t)i0 * R, *xr1 = X + (size_t)i1 * R;
        double x0[NCH], x1[NCH], is[NCH];
        double c0 = 0.0, c1 = 0.0;
        unsigned ok0 = 0u, ok1 = 0u;
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            const int r = lane + 32 * c;
            x0[c] = (r < R) ? __ldcs(xr0 + r) : CUDART_NAN;
            x1[c] = (r < R) ? __ldcs(xr1 + r) : CUDART_NAN;
            is[c] = (r < R) ? s_isig[r] : 0.0;
            if (!cs_isna(x0[c])) {
                c0 += s_c[r];
                ok0 |= 1u << c;
            }
            if (!cs_isna(x1[c])) {
                c1 += s_c[r];
                ok1 |= 1u << c;
            }
        }
        c0 = cs_warp_sum_bfly(c0);
        c1 = cs_warp_sum_bfly(c1);
        for (int k0 = 0; k0 < K; k0 += LL_KT) {
            double q0[LL_KT], q1[LL_KT];
#pragma unroll
            for (int j = 0; j < LL_KT; j++) q0[j] = q1[j] = 0.0;
#pragma unroll
            for (int c = 0; c < NCH; c++) {
                const int r = min(lane + 32 * c, R - 1);
#pragma unroll
                for (int j = 0; j < LL_KT; j++) {
                    const int k = min(k0 + j, K - 1);
                    const double u = __ldg(U + (size_t)k * R + r);
                    const double d0 = (x0[c] - u) * is[c], d1 = (x1[c] - u) * is[c];
                    if (ok0 >> c & 1u) q0[j] = fma(d0, d0, q0[j]);
                    if (ok1 >> c & 1u) q1[j] = fma(d1, d1, q1[j]);
                }
            }
            const double b0 = fold(q0), b1 = fold(q1);
            const int k = k0 + kslot;
            if ((lane & 7) == 0 && k < K) {
                LL[(size_t)i0 * K + k] = -(c0 + 0.5 * b0);
                if (i1 != i0) LL[(size_t)i1 * K + k] = -(c1 + 0.5 * b1);
            }
        }
    }
}

// out[c*rows + r] = in[r*cols + c] : generic tiled transpose (rows x cols row-major in)
template <typename T>
__global__ void transpose_kernel(int rows, int cols, const T *__restrict__ in, T *__restrict__ out)
{
    __shared__ T tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = in[(size_t)r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][j];
    }
}

// ---- majority vote --------------------------------------------------------
// one thread per cell; element (t,i) at Z[t*st + i*si].  One pass over the T labels with the
// distinct labels seen so far and their counts in MV_SLOTS registers (a converged chain shows
// two or three per cell); a cell with more distinct labels than that falls back to counting each
// first occurrence against the whole column.  Most frequent label, ties -> the smallest label
// (names(table(x))[which.max(table(x))], R/AssignCluster.R:37).
constexpr int MV_SLOTS = 6, MV_SLOTS2 = 24;

// counts of up to SLOTS distinct labels of one cell in registers; false when there are more
template <int SLOTS>
__device__ __forceinline__ bool mv_tally(int T, const int *__restrict__ z, int64_t st, int &best_out)
{
    int lab[SLOTS], cnt[SLOTS];
#pragma unroll
    for (int s = 0; s < SLOTS; s++) {
        lab[s] = 0;
        cnt[s] = 0;
    }
    int used = 0;
    bool overflow = false;
#pragma unroll 4
    for (int t = 0; t < T; t++) {
        const int l = __ldcs(z + (size_t)t * st);
        bool hit = false;
#pragma unroll
        for (int s = 0; s < SLOTS; s++) {
            const bool m = s < used && lab[s] == l;
            cnt[s] += m ? 1 : 0;
            hit |= m;
        }
        if (!hit) {
#pragma unroll
            for (int s = 0; s < SLOTS; s++)
                if (s == used) {
                    lab[s] = l;
                    cnt[s] = 1;
                }
            overflow |= used >= SLOTS;
            used = min(used + 1, SLOTS);
        }
    }
    if (overflow) return false;
    int best = lab[0], bc = cnt[0];
#pragma unroll
    for (int s = 1; s < SLOTS; s++)
        if (s < used && (cnt[s] > bc || (cnt[s] == bc && lab[s] < best))) {
            bc = cnt[s];
            best = lab[s];
        }
    best_out = best;
    return true;
}

// thread per cell.  A cell of a converged chain visits a handful of labels: six register slots
// decide it in one pass over its T labels; a cell with more takes a second pass with 24 slots,
// and only beyond that the quadratic count of every first occurrence against the whole column.
__global__ void __launch_bounds__(128)
majority_vote_kernel(int T, int N, const int *__restrict__ Z, int64_t st, int64_t si, int *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int *z = Z + (size_t)i * si;
    int best = 0;
    if (mv_tally<MV_SLOTS>(T, z, st, best)) {
        out[i] = best;
        return;
    }
    if (mv_tally<MV_SLOTS2>(T, z, st, best)) {
        out[i] = best;
        return;
    }
    int bc = 0;
    best = 0;
    for (int a = 0; a < T; a++) {
        const int la = z[(size_t)a * st];
        int c = 0;
        bool seen = false;
        for (int t = 0; t < T; t++) {
            const int lt = z[(size_t)t * st];
            if (lt == la) {
                if (t < a) { seen = true; break; }
                c++;
            }
        }
        if (seen) continue;
        if (c > bc || (c == bc && la < best)) {
            bc = c;
            best = la;
        }
    }
    out[i] = best;
}

} // namespace

template <typename T>
int cs_transpose(int rows, int cols, const T *d_in, T *d_out, cudaStream_t st)
{
    if (rows == 0 || cols == 0) return CS_OK;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
    transpose_kernel<T><<<grid, block, 0, st>>>(rows, cols, d_in, d_out);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}
template int cs_transpose<double>(int, int, const double *, double *, cudaStream_t);
template int cs_transpose<int>(int, int, const int *, int *, cudaStream_t);

extern "C" int cs_loglik_matrix_dev(int N, int R, int K, const double *d_X, const double *d_U,
                                    const double *d_sigma, double *d_LL, void *stream)
{
    CS_REQUIRE(N >= 0 && R >= 1 && K >= 0, "cs_loglik_matrix: bad dimensions N=%d R=%d K=%d", N, R, K);
    if (N == 0 || K == 0) return CS_OK;
    const size_t smem = sizeof(double) * 2 * (size_t)R;
    CS_REQUIRE(smem <= 200 * 1024, "cs_loglik_matrix: R=%d too large", R);
    int grid = (N + LL_THREADS / 32 - 1) / (LL_THREADS / 32);
    const int maxgrid = cs_num_sms() * 8;
    if (grid > maxgrid) grid = maxgrid;
    void (*kern)(int, int, int, const double *, const double *, const double *, double *) = loglik_matrix_generic_kernel;
    if (R <= 32) kern = loglik_matrix_kernel<1>;
    else if (R <= 64) kern = loglik_matrix_kernel<2>;
    else if (R <= 128) kern = loglik_matrix_kernel<4>;
    else if (R <= 192) kern = loglik_matrix_kernel<6>;
    else if (R <= 256) kern = loglik_matrix_kernel<8>;
    else if (R <= 320) kern = loglik_matrix_kernel<10>;
    else if (R <= 384) kern = loglik_matrix_kernel<12>;
    else if (R <= 512) kern = loglik_matrix_kernel<16>;
    // many clusters: two cells per pass (CS_LL_PAIR=0 keeps the one-cell kernel)
    static const bool pair_ok = !(getenv("CS_LL_PAIR") && atoi(getenv("CS_LL_PAIR")) == 0);
    if (pair_ok && K >= 8 && N >= 2) {
        if (R <= 32) kern = loglik_matrix2_kernel<1>;
        else if (R <= 64) kern = loglik_matrix2_kernel<2>;
        else if (R <= 128) kern = loglik_matrix2_kernel<4>;
        else if (R <= 192) kern = loglik_matrix2_kernel<6>;
        else if (R <= 256) kern = loglik_matrix2_kernel<8>;
        else if (R <= 320) kern = loglik_matrix2_kernel<10>;
    }
    CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, LL_THREADS, smem, (cudaStream_t)stream>>>(N, R, K, d_X, d_U, d_sigma, d_LL);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int cs_loglik_matrix(int N, int R, int K, const double *X, const double *U, const double *sigma,
                                double *LL)
{
    CS_REQUIRE(X && U && sigma && LL, "cs_loglik_matrix: NULL argument");
    CS_REQUIRE(N >= 0 && R >= 1 && K >= 0, "cs_loglik_matrix: bad dimensions N=%d R=%d K=%d", N, R, K);
    if (N == 0 || K == 0) return CS_OK;
    DevBuf<double> dXc, dX, dUc, dU, dS, dL, dLc;
    const size_t nx = (size_t)N * R, nu = (size_t)K * R, nl = (size_t)N * K;
    CS_CUDA(dXc.alloc(nx));
    CS_CUDA(dX.alloc(nx));
    CS_CUDA(dUc.alloc(nu));
    CS_CUDA(dU.alloc(nu));
    CS_CUDA(dS.alloc(R));
    CS_CUDA(dL.alloc(nl));
    CS_CUDA(dLc.alloc(nl));
    if (int rc0 = cs_h2d_staged(dXc.p, X, sizeof(double) * nx)) return rc0;
    CS_CUDA(cudaMemcpy(dUc.p, U, sizeof(double) * nu, cudaMemcpyHostToDevice));
    CS_CUDA(cudaMemcpy(dS.p, sigma, sizeof(double) * R, cudaMemcpyHostToDevice));
    int rc;
    // column-major N x R == row-major R x N -> transpose to row-major N x R
    if ((rc = cs_transpose<double>(R, N, dXc.p, dX.p, 0))) return rc;
    if ((rc = cs_transpose<double>(R, K, dUc.p, dU.p, 0))) return rc;
    if ((rc = cs_loglik_matrix_dev(N, R, K, dX.p, dU.p, dS.p, dL.p, nullptr))) return rc;
    // row-major N x K -> column-major N x K (== row-major K x N)
    if ((rc = cs_transpose<double>(N, K, dL.p, dLc.p, 0))) return rc;
    CS_CUDA(cudaDeviceSynchronize());
    if ((rc = cs_d2h_staged(LL, dLc.p, sizeof(double) * nl))) return rc;
    return CS_OK;
}

extern "C" int cs_majority_vote_dev(int T, int N, const int *d_Zall, int64_t stride_t, int64_t stride_i,
                                    int *d_zmaj, void *stream)
{
    CS_REQUIRE(T >= 1 && N >= 0 && d_Zall && d_zmaj, "cs_majority_vote: bad arguments T=%d N=%d", T, N);
    if (N == 0) return CS_OK;
    majority_vote_kernel<<<(N + 127) / 128, 128, 0, (cudaStream_t)stream>>>(T, N, d_Zall, stride_t, stride_i, d_zmaj);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

extern "C" int cs_majority_vote(int T, int N, const int *Zall, int *zmaj)
{
    CS_REQUIRE(T >= 1 && N >= 0 && Zall && zmaj, "cs_majority_vote: bad arguments T=%d N=%d", T, N);
    if (N == 0) return CS_OK;
    DevBuf<int> dZ, dZt, dO;
    const size_t n = (size_t)T * N;
    CS_CUDA(dZ.alloc(n));
    CS_CUDA(dZt.alloc(n));
    CS_CUDA(dO.alloc(N));
    if (int rc0 = cs_h2d_staged(dZ.p, Zall, sizeof(int) * n)) return rc0;
    // T x N column-major == row-major N x T; transpose to sweep-major (T x N row-major)
    // so that consecutive threads (cells) read consecutive addresses
    int rc = cs_transpose<int>(N, T, dZ.p, dZt.p, 0);
    if (rc) return rc;
    rc = cs_majority_vote_dev(T, N, dZt.p, (int64_t)N, 1, dO.p, nullptr);
    if (rc) return rc;
    CS_CUDA(cudaMemcpy(zmaj, dO.p, sizeof(int) * N, cudaMemcpyDeviceToHost));
    return CS_OK;
}
