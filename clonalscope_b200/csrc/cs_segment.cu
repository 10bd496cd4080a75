// cs_segment.cu — the caller side of the gene -> bin aggregation: bin x cluster pooling of the
// bin x cell counts and the per-chromosome HMM segmentation of the pooled profiles.
//
// Reference: R/CreateSegtableNoWGS.R:62-69 (bin_by_cluster[, k] = rowSums(bin_mtx[, Zest == k]),
// :77 ref_counts = rowSums(bin_mtx[, normal cells])) and R/Segmentation_bulk.R:150-166
// (caTools::runmean of the normalised coverage, then HiddenMarkov::dthmm + Viterbi with four
// Gaussian states).  caTools and HiddenMarkov are third-party packages that are not part of the
// reference tree; their published algorithms are restated here (runmean: window of k values,
// right half k %/% 2, the window shrinking at both ends — endrule "mean"; Viterbi.dthmm: log-space
// forward maxima over nu[i-1, j] + log Pi[j, k], first maximum wins, backtracking on
// log Pi[, y[i+1]] + nu[i, ]).  Compiled without FMA contraction so the log-space sums round like
// R's.
#include "cs_common.cuh"

#include <math_constants.h>
#include <vector>

namespace {

// out[label[c] * B + bin] += x  for every stored entry of column c; columns whose label is outside
// [0, K) are skipped (NA, cells that are not in the pool).  Counts are integers carried in doubles:
// the FP64 atomic sums are exact and order-free below 2^53.
__global__ void __launch_bounds__(256)
rowsum_by_label_kernel(int B, int N, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                       const double *__restrict__ x, const int *__restrict__ label, int K, double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < N; c += nwarps) {
        const int k = label[c];
        if (k < 0 || k >= K) continue;
        double *__restrict__ o = out + (size_t)k * B;
        const int64_t e0 = colptr[c], e1 = colptr[c + 1];
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int b = __ldcs(rowidx + e);
            const double v = __ldcs(x + e);
            if (b >= 0 && b < B) atomicAdd(o + b, v);
        }
    }
}

struct HmmParams {
    double mean[4], sd[4], logdelta[4], logPi[16]; // logPi[j * 4 + k]: from state j to state k
};

// One CTA per sequence (one chromosome of one pooled profile).
//   smooth[i] = mean of the window around i: k values, k2 = k / 2 of them to the right, the
//               window clipped at both ends (k = min(nmean, n));
//   nu        the Viterbi forward table (n x 4, kept for the backtracking), by thread 0;
//   state[i]  in 0..3.
__global__ void __launch_bounds__(128)
segment_hmm_kernel(int nseq, const int64_t *__restrict__ off, const double *__restrict__ val, int nmean, HmmParams P,
                   double *__restrict__ smooth, double *__restrict__ nu, int *__restrict__ state, int *__restrict__ fail)
{
    const int s = blockIdx.x;
    if (s >= nseq) return;
    const int64_t o = off[s];
    const int n = (int)(off[s + 1] - o);
    if (n <= 0) return;
    const double *__restrict__ v = val + o;
    double *__restrict__ sm = smooth + o;
    const int k = nmean <= 1 ? 1 : (nmean > n ? n : nmean);
    const int k2 = k >> 1, k1 = k - k2 - 1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        if (k == 1) {
            sm[i] = v[i];
            continue;
        }
        const int lo = max(0, i - k1), hi = min(n - 1, i + k2);
        // compensated sum in index order (the reference's running sum carries an error term as well)
        double sum = 0.0, err = 0.0;
        int num = 0;
        for (int j = lo; j <= hi; j++) {
            const double y = v[j];
            if (isfinite(y)) {
                const double t = sum + y;
                err += (fabs(sum) >= fabs(y)) ? (sum - t) + y : (y - t) + sum;
                sum = t;
                num++;
            }
        }
        sm[i] = num ? (sum + err) / num : CUDART_NAN;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double *__restrict__ tab = nu + 4 * o;
    int *__restrict__ st = state + o;
    double prev[4], ls[4];
#pragma unroll
    for (int a = 0; a < 4; a++) ls[a] = log(P.sd[a]);
    auto dens = [&](double xv, int a) {
        const double z = (xv - P.mean[a]) / P.sd[a];
        return -(CS_LN_SQRT_2PI + 0.5 * z * z + ls[a]);
    };
#pragma unroll
    for (int a = 0; a < 4; a++) {
        prev[a] = P.logdelta[a] + dens(sm[0], a);
        tab[a] = prev[a];
    }
    for (int i = 1; i < n; i++) {
        const double xv = sm[i];
        double cur[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
            double best = prev[0] + P.logPi[b];
#pragma unroll
            for (int a = 1; a < 4; a++) best = fmax(best, prev[a] + P.logPi[a * 4 + b]);
            cur[b] = best + dens(xv, b);
        }
#pragma unroll
        for (int a = 0; a < 4; a++) {
            prev[a] = cur[a];
            tab[4 * (size_t)i + a] = cur[a];
        }
    }
    bool under = false;
    int y = 0;
    for (int a = 0; a < 4; a++) {
        under |= prev[a] == -CUDART_INF || prev[a] != prev[a];
        if (prev[a] > prev[y]) y = a;
    }
    if (under) atomicExch(fail, s + 1); // ("Problems With Underflow")
    st[n - 1] = y;
    for (int i = n - 2; i >= 0; i--) {
        int b = 0;
        double bv = P.logPi[0 * 4 + y] + tab[4 * (size_t)i];
        for (int a = 1; a < 4; a++) {
            const double t = P.logPi[a * 4 + y] + tab[4 * (size_t)i + a];
            if (t > bv) {
                bv = t;
                b = a;
            }
        }
        y = b;
        st[i] = y;
    }
}

} // namespace

extern "C" int cs_rowsum_by_label_dev(int B, int N, const int64_t *d_colptr, const int32_t *d_rowidx, const double *d_x,
                                      const int *d_label, int K, double *d_out, void *stream)
{
    CS_REQUIRE(B >= 1 && N >= 0 && K >= 1 && d_colptr && d_label && d_out, "cs_rowsum_by_label: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CS_CUDA(cudaMemsetAsync(d_out, 0, sizeof(double) * (size_t)B * K, st));
    if (N == 0) return CS_OK;
    int grid = (N + 7) / 8;
    const int maxgrid = cs_num_sms() * 8;
    if (grid > maxgrid) grid = maxgrid;
    rowsum_by_label_kernel<<<grid, 256, 0, st>>>(B, N, d_colptr, d_rowidx, d_x, d_label, K, d_out);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// Host form on dgCMatrix slots (bin x cell, 0-based): out is B x K column-major (R's matrix).
extern "C" int cs_rowsum_by_label(int B, int N, const int32_t *colptr, const int32_t *rowidx, const double *x,
                                  const int *label, int K, double *out)
{
    CS_REQUIRE(B >= 1 && N >= 0 && K >= 1 && colptr && label && out, "cs_rowsum_by_label: bad arguments");
    CS_REQUIRE(colptr[0] == 0, "cs_rowsum_by_label: colptr[0] must be 0");
    for (int c = 0; c < N; c++) CS_REQUIRE(colptr[c] <= colptr[c + 1], "colptr not monotone at column %d", c);
    const int64_t nnz = colptr[N];
    std::vector<int64_t> cp64((size_t)N + 1);
    for (int c = 0; c <= N; c++) cp64[c] = colptr[c];
    DevBuf<int64_t> d_cp;
    DevBuf<int32_t> d_i;
    DevBuf<double> d_x, d_o;
    DevBuf<int> d_l;
    CS_CUDA(d_cp.alloc((size_t)N + 1));
    CS_CUDA(d_i.alloc((size_t)nnz));
    CS_CUDA(d_x.alloc((size_t)nnz));
    CS_CUDA(d_l.alloc((size_t)N));
    CS_CUDA(d_o.alloc((size_t)B * K));
    CS_CUDA(cudaMemcpy(d_cp.p, cp64.data(), sizeof(int64_t) * cp64.size(), cudaMemcpyHostToDevice));
    if (N) CS_CUDA(cudaMemcpy(d_l.p, label, sizeof(int) * (size_t)N, cudaMemcpyHostToDevice));
    int rc;
    if (nnz) {
        if ((rc = cs_h2d_staged(d_i.p, rowidx, sizeof(int32_t) * (size_t)nnz))) return rc;
        if ((rc = cs_h2d_staged(d_x.p, x, sizeof(double) * (size_t)nnz))) return rc;
    }
    if ((rc = cs_rowsum_by_label_dev(B, N, d_cp.p, d_i.p, d_x.p, d_l.p, K, d_o.p, nullptr))) return rc;
    CS_CUDA(cudaMemcpy(out, d_o.p, sizeof(double) * (size_t)B * K, cudaMemcpyDeviceToHost)); // k-major == column-major B x K
    return CS_OK;
}

// nseq sequences stored back to back in val (sequence s = val[off[s] .. off[s+1])): running mean
// of width nmean and the Viterbi path of the 4-state Gaussian HMM (means / sds, initial
// distribution delta, transition matrix Pi row-major: from state j to state k).
// Outputs: smooth (as val), state (0-based index into means).  Host pointers.
extern "C" int cs_segment_hmm(int nseq, const int64_t *off, const double *val, int nmean, const double *means,
                              const double *sds, const double *delta, const double *Pi, double *smooth, int *state)
{
    CS_REQUIRE(nseq >= 0 && off && means && sds && delta && Pi && state, "cs_segment_hmm: NULL argument");
    if (nseq == 0) return CS_OK;
    const int64_t total = off[nseq];
    for (int s = 0; s < nseq; s++) CS_REQUIRE(off[s] <= off[s + 1], "cs_segment_hmm: offsets not monotone at %d", s);
    CS_REQUIRE(off[0] == 0 && val != nullptr, "cs_segment_hmm: off[0] must be 0");
    HmmParams P;
    for (int a = 0; a < 4; a++) {
        P.mean[a] = means[a];
        P.sd[a] = sds[a];
        CS_REQUIRE(sds[a] > 0.0, "cs_segment_hmm: sd[%d] must be positive", a);
        P.logdelta[a] = log(delta[a]);
        for (int b = 0; b < 4; b++) P.logPi[a * 4 + b] = log(Pi[a * 4 + b]);
    }
    DevBuf<int64_t> d_off;
    DevBuf<double> d_val, d_sm, d_nu;
    DevBuf<int> d_st, d_fail;
    CS_CUDA(d_off.alloc((size_t)nseq + 1));
    CS_CUDA(d_val.alloc((size_t)total));
    CS_CUDA(d_sm.alloc((size_t)total));
    CS_CUDA(d_nu.alloc((size_t)total * 4));
    CS_CUDA(d_st.alloc((size_t)total));
    CS_CUDA(d_fail.alloc(1));
    CS_CUDA(cudaMemset(d_fail.p, 0, sizeof(int)));
    CS_CUDA(cudaMemcpy(d_off.p, off, sizeof(int64_t) * ((size_t)nseq + 1), cudaMemcpyHostToDevice));
    if (total) CS_CUDA(cudaMemcpy(d_val.p, val, sizeof(double) * (size_t)total, cudaMemcpyHostToDevice));
    segment_hmm_kernel<<<nseq, 128>>>(nseq, d_off.p, d_val.p, nmean, P, d_sm.p, d_nu.p, d_st.p, d_fail.p);
    CS_CUDA(cudaGetLastError());
    int fail = 0;
    CS_CUDA(cudaMemcpy(&fail, d_fail.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (fail) {
        cs_set_error("Problems With Underflow (sequence %d)", fail - 1);
        return CS_ERR_STATE;
    }
    if (total) {
        if (smooth) CS_CUDA(cudaMemcpy(smooth, d_sm.p, sizeof(double) * (size_t)total, cudaMemcpyDeviceToHost));
        CS_CUDA(cudaMemcpy(state, d_st.p, sizeof(int) * (size_t)total, cudaMemcpyDeviceToHost));
    }
    return CS_OK;
}
