// cs_binning.cu — gene -> bin aggregation of a sparse gene x cell count matrix.
//
// Replaces the hot loop of Gen_bin_cell_rna_filtered (reference
// R/Gene_bin_cell_rna_filtered.R:63-81): the reference densifies genes x cells (:69)
// and then, for every bin, sums the rows of the genes overlapping it (:71-77) before
// converting back to a sparse Matrix (:79-81).  Here each CSC column (cell) is
// aggregated independently on one CTA:
//
//   pass 1  bin_count_kernel   rowidx only (4 B/nz): marks touched bins in a shared
//                              bitmap, popcounts -> upper bound of nnz per column.
//           bin_count_stream_kernel: when genes are in genomic order over sorted tiles
//                              (first bins non-decreasing, hits contiguous) a column is a
//                              stream of sorted intervals and one warp counts it with a
//                              prefix-max scan — no shared memory, 64 warps per SM
//   scan    exclusive scan of the per-column counts -> out_colptr
//   pass 2  bin_fill_stream_kernel (regular maps): sums accumulated in RANK space — the rank
//                              of every bin follows from two prefix scans over the column
//                              (see the kernel) — no bin-indexed accumulator, no compaction
//           bin_fill_kernel    (any map) 128-bit loads of rowidx/x (first group fetched one column
//                              ahead), gene->bin lookup through a packed per-gene word
//                              (first_bin | nhits<<20), shared memory integer accumulators
//                              for all bins — one ATOMS.ADD per hit and nothing else; the
//                              touched bins are recovered from the accumulators themselves
//                              (conflict-free rotated scan into one 32-bin mask per
//                              register), block scan, staged u16 ranks, coalesced stores
//
// Integer counts carried in doubles are summed exactly in int32 accumulators; every
// value is checked for integrality / range and every column for overflow, and any
// violation re-runs the column set through the FP64-accumulator instantiation.
// A touched bin whose sum is zero (explicit zeros, cancelling negatives) is dropped
// as Matrix(sparse=TRUE) does; that rare case triggers a compaction pass.
#include "cs_common.cuh"

#include <utility>

#include <limits.h>
#include <stdlib.h>
#include <vector>

namespace {

constexpr uint32_t GMAP_ESC = 0xFFFFFFFFu;
constexpr int FILL_THREADS = 512;
constexpr int LIST_CAP = 6144; // ranks staged per output round (u16 each)
constexpr int COUNT_THREADS = 128;

enum { FLAG_INEXACT = 0, FLAG_SHRUNK = 1, FLAG_CAP = 2, FLAG_IRREG = 3, NFLAGS = 4 };

// ---- pass 1 ---------------------------------------------------------------
__global__ void __launch_bounds__(COUNT_THREADS)
bin_count_kernel(int N, int B, const int64_t *__restrict__ colptr,
                 const int32_t *__restrict__ rowidx, const uint32_t *__restrict__ gmap,
                 const int32_t *__restrict__ map_ptr, const int32_t *__restrict__ map_bin,
                 int *__restrict__ counts)
{
    extern __shared__ uint32_t s_bits[];
    __shared__ int s_warp[COUNT_THREADS / 32];
    const int words = (B + 31) >> 5;
    for (int w = threadIdx.x; w < words; w += COUNT_THREADS) s_bits[w] = 0;
    __syncthreads();
    for (int c = blockIdx.x; c < N; c += gridDim.x) {
        const int64_t e0 = colptr[c], e1 = colptr[c + 1];
        for (int64_t e = e0 + threadIdx.x; e < e1; e += COUNT_THREADS) {
            const int g = rowidx[e];
            const uint32_t w = __ldg(gmap + g);
            if (w != GMAP_ESC) {
                const int b0 = w & 0xFFFFF, nh = w >> 20;
                for (int t = 0; t < nh; t++) {
                    const int b = b0 + t;
                    atomicOr(&s_bits[b >> 5], 1u << (b & 31));
                }
            } else {
                for (int m = map_ptr[g]; m < map_ptr[g + 1]; m++) {
                    const int b = map_bin[m];
                    atomicOr(&s_bits[b >> 5], 1u << (b & 31));
                }
            }
        }
        __syncthreads();
        int cnt = 0;
        for (int w = threadIdx.x; w < words; w += COUNT_THREADS) {
            cnt += __popc(s_bits[w]);
            s_bits[w] = 0;
        }
        for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(CS_FULL, cnt, off);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int i = 0; i < COUNT_THREADS / 32; i++) t += s_warp[i];
            counts[c] = t;
        }
        __syncthreads();
    }
}

// ---- exclusive scan int32[N] -> int64[N+1] ---------------------------------
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t *s_warp, int64_t &total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int64_t inc = v;
    for (int off = 1; off < 32; off <<= 1) {
        int64_t t = __shfl_up_sync(CS_FULL, inc, off);
        if (lane >= off) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int64_t w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
        int64_t winc = w;
        for (int off = 1; off < 32; off <<= 1) {
            int64_t t = __shfl_up_sync(CS_FULL, winc, off);
            if (lane >= off) winc += t;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    total = s_warp[32];
    int64_t r = s_warp[wid] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tiles_kernel(int N, const int *__restrict__ counts, int64_t *__restrict__ out,
                  int64_t *__restrict__ tile_sums)
{
    __shared__ int64_t s_warp[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < N) ? counts[base + i] : 0;
        s += v[i];
    }
    int64_t total;
    int64_t ex = block_exclusive_scan(s, s_warp, total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < N) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_sums_kernel(int ntiles, int64_t *__restrict__ tile_sums, int64_t *__restrict__ grand)
{
    __shared__ int64_t s_warp[33];
    int64_t carry = 0;
    for (int base = 0; base < ntiles; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        int64_t v = (i < ntiles) ? tile_sums[i] : 0, total;
        int64_t ex = block_exclusive_scan(v, s_warp, total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add_kernel(int N, int64_t *__restrict__ out, const int64_t *__restrict__ tile_sums,
                const int64_t *__restrict__ grand)
{
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    const int64_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < N) out[base + i] += add;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[N] = *grand;
}

// ---- pass 2 ---------------------------------------------------------------
template <typename Acc>
struct AccOps;
template <>
struct AccOps<int> {
    // returns false when v cannot be accumulated exactly as an int32
    static __device__ __forceinline__ bool conv(double v, int &iv)
    {
        iv = __double2int_rz(v); // saturating
        return __int2double_rn(iv) == v && (unsigned)(iv + (1 << 30)) < (1u << 31);
    }
    static __device__ __forceinline__ unsigned mag(int iv) { return (unsigned)abs(iv); }
};
template <>
struct AccOps<double> {
    static __device__ __forceinline__ bool conv(double v, double &iv)
    {
        iv = v;
        return true;
    }
    static __device__ __forceinline__ unsigned mag(double) { return 0u; }
};

// Shared-memory atomics retire about one lane every two cycles on this part, which makes them
// the budget of this kernel (~2,500 per column): one ATOMS.ADD per hit and nothing else — the
// touched bins are recovered afterwards from the accumulators themselves.
template <typename Acc>
__device__ __forceinline__ void hit(int b, Acc iv, Acc *s_acc)
{
    atomicAdd(&s_acc[b], iv);
}

// genes spanning three or more bins, or with non-contiguous bins (rare): rolled loops so that
// the common path stays a handful of instructions
template <typename Acc>
__device__ __forceinline__ void scatter_wide(int g, uint32_t w, Acc iv, Acc *s_acc,
                                          const int32_t *__restrict__ map_ptr,
                                          const int32_t *__restrict__ map_bin, unsigned long long &magsum)
{
    if (w != GMAP_ESC) {
        const int b0 = w & 0xFFFFF, nh = w >> 20;
#pragma unroll 1
        for (int t = 0; t < nh; t++) hit<Acc>(b0 + t, iv, s_acc);
        magsum += (unsigned long long)AccOps<Acc>::mag(iv) * (unsigned)nh;
    } else {
        const int m0 = map_ptr[g], m1 = map_ptr[g + 1];
#pragma unroll 1
        for (int m = m0; m < m1; m++) hit<Acc>(map_bin[m], iv, s_acc);
        magsum += (unsigned long long)AccOps<Acc>::mag(iv) * (unsigned)(m1 - m0);
    }
}

template <typename Acc>
__device__ __forceinline__ void scatter_entry(int g, double v, Acc *s_acc,
                                              const uint32_t *__restrict__ gmap,
                                              const int32_t *__restrict__ map_ptr,
                                              const int32_t *__restrict__ map_bin, bool &inexact,
                                              unsigned long long &magsum)
{
    Acc iv;
    if (!AccOps<Acc>::conv(v, iv)) inexact = true;
    const uint32_t w = __ldg(gmap + g);
    const unsigned nh = w >> 20;
    if (nh - 1u < 2u) { // one or two contiguous bins: 98 % of the genes over 200-kb tiles
        const int b0 = w & 0xFFFFF;
        hit<Acc>(b0, iv, s_acc);
        if (nh == 2u) hit<Acc>(b0 + 1, iv, s_acc);
        magsum += (unsigned long long)AccOps<Acc>::mag(iv) * nh;
    } else if (w != 0u) {
        scatter_wide<Acc>(g, w, iv, s_acc, map_ptr, map_bin, magsum);
    }
}

template <typename Acc>
__global__ void __launch_bounds__(FILL_THREADS, 3)
bin_fill_kernel(int N, int B, const int64_t *__restrict__ colptr,
                const int32_t *__restrict__ rowidx, const double *__restrict__ x,
                const uint32_t *__restrict__ gmap, const int32_t *__restrict__ map_ptr,
                const int32_t *__restrict__ map_bin, const int64_t *__restrict__ out_colptr,
                int32_t *__restrict__ out_rowidx, double *__restrict__ out_x, int64_t out_cap,
                int *__restrict__ counts_actual, int *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    // every thread owns a contiguous run of PER bins (a multiple of 32) of the accumulator array
    const int PER = (((B + FILL_THREADS - 1) / FILL_THREADS) + 31) & ~31;
    const int BP = (B + 31) & ~31; // accumulators, padded to a multiple of 32
    Acc *s_acc = reinterpret_cast<Acc *>(s_raw);
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_acc + BP);
    __shared__ int s_warp[FILL_THREADS / 32 + 1];
    __shared__ unsigned long long s_mag[FILL_THREADS / 32];

    if (out_colptr[N] > out_cap) { // caller's buffers too small: report, write nothing
        if (blockIdx.x == 0 && threadIdx.x == 0) flags[FLAG_CAP] = 1;
        return;
    }
    for (int b = threadIdx.x; b < BP; b += FILL_THREADS) s_acc[b] = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int MAXW = 4; // PER <= 128 bins per thread (B <= 65535)
    const int nw = PER >> 5;
    const int bin0 = threadIdx.x * PER;

    // the first 128-bit group of every thread is fetched one column ahead, so that its HBM
    // latency hides behind the compaction of the current column
    int4 pg4 = make_int4(0, 0, 0, 0);
    double2 pv01 = make_double2(0, 0), pv23 = make_double2(0, 0);
    auto prefetch = [&](int col) {
        if (col < N) {
            const int64_t a0 = colptr[col], a1 = colptr[col + 1];
            const int64_t ab = min((a0 + 3) & ~(int64_t)3, a1), ae = max(ab, a1 & ~(int64_t)3);
            const int64_t e = ab + 4 * (int64_t)threadIdx.x;
            if (e < ae) {
                pg4 = __ldcs(reinterpret_cast<const int4 *>(rowidx + e));
                pv01 = __ldcs(reinterpret_cast<const double2 *>(x + e));
                pv23 = __ldcs(reinterpret_cast<const double2 *>(x + e + 2));
            }
        }
    };
    prefetch(blockIdx.x);

    for (int c = blockIdx.x; c < N; c += gridDim.x) {
        const int64_t e0 = colptr[c], e1 = colptr[c + 1];
        bool inexact = false;
        unsigned long long magsum = 0;
        // 128-bit body: 4 row indices + 4 values per thread per step
        const int64_t eb = min((e0 + 3) & ~(int64_t)3, e1), ee = max(eb, e1 & ~(int64_t)3);
        for (int64_t e = eb + 4 * (int64_t)threadIdx.x; e < ee; e += 4 * FILL_THREADS) {
            int4 g4 = pg4;
            double2 v01 = pv01, v23 = pv23;
            if (e >= eb + 4 * (int64_t)FILL_THREADS) { // columns longer than one round of the CTA
                g4 = __ldcs(reinterpret_cast<const int4 *>(rowidx + e));
                v01 = __ldcs(reinterpret_cast<const double2 *>(x + e));
                v23 = __ldcs(reinterpret_cast<const double2 *>(x + e + 2));
            }
            scatter_entry<Acc>(g4.x, v01.x, s_acc, gmap, map_ptr, map_bin, inexact, magsum);
            scatter_entry<Acc>(g4.y, v01.y, s_acc, gmap, map_ptr, map_bin, inexact, magsum);
            scatter_entry<Acc>(g4.z, v23.x, s_acc, gmap, map_ptr, map_bin, inexact, magsum);
            scatter_entry<Acc>(g4.w, v23.y, s_acc, gmap, map_ptr, map_bin, inexact, magsum);
        }
        { // unaligned head and tail (at most 3 + 3 entries)
            const int nhead = (int)(eb - e0), ntail = (int)(e1 - ee);
            if ((int)threadIdx.x < nhead + ntail) {
                const int64_t e = ((int)threadIdx.x < nhead) ? e0 + threadIdx.x
                                                             : ee + (threadIdx.x - nhead);
                scatter_entry<Acc>(rowidx[e], x[e], s_acc, gmap, map_ptr, map_bin, inexact,
                                   magsum);
            }
        }
        // exactness guard: every value integral and the column cannot overflow int32
        for (int off = 16; off >= 1; off >>= 1) magsum += __shfl_xor_sync(CS_FULL, magsum, off);
        if (lane == 0) s_mag[wid] = magsum;
        if (inexact) flags[FLAG_INEXACT] = 1;
        prefetch(c + gridDim.x);
        __syncthreads();

        // compaction: the nonzero accumulators are the output (zero sums are dropped, as
        // Matrix(sparse=TRUE) does); one mask of 32 bins per register
        int cnt = 0;
        uint32_t mask[MAXW];
#pragma unroll
        for (int j = 0; j < MAXW; j++) {
            mask[j] = 0;
            if (j < nw && bin0 + 32 * j < BP) {
                const Acc *a = s_acc + bin0 + 32 * j;
#pragma unroll
                for (int t = 0; t < 32; t++) { // rotated per lane: the 32 lanes read 32 different banks
                    const int q = (t + lane) & 31;
                    mask[j] |= (a[q] != (Acc)0) ? (1u << q) : 0u;
                }
                cnt += __popc(mask[j]);
            }
        }
        int inc = cnt;
        for (int off = 1; off < 32; off <<= 1) {
            int t = __shfl_up_sync(CS_FULL, inc, off);
            if (lane >= off) inc += t;
        }
        if (lane == 31) s_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = (lane < FILL_THREADS / 32) ? s_warp[lane] : 0, winc = w;
            for (int off = 1; off < 32; off <<= 1) {
                int t = __shfl_up_sync(CS_FULL, winc, off);
                if (lane >= off) winc += t;
            }
            if (lane < FILL_THREADS / 32) s_warp[lane] = winc - w;
            if (lane == FILL_THREADS / 32 - 1) s_warp[FILL_THREADS / 32] = winc;
            unsigned long long m = (lane < FILL_THREADS / 32) ? s_mag[lane] : 0ull;
            for (int off = 16; off >= 1; off >>= 1) m += __shfl_xor_sync(CS_FULL, m, off);
            if (lane == 0 && m >= 2147483647ull) flags[FLAG_INEXACT] = 1;
        }
        __syncthreads();
        const int rank0 = s_warp[wid] + inc - cnt;
        const int total = s_warp[FILL_THREADS / 32];
        const int64_t o = out_colptr[c];
        // the staging list holds LIST_CAP ranks; columns with more outputs go out in rounds
        for (int base = 0; base < total; base += LIST_CAP) {
            int rank = rank0 - base;
#pragma unroll
            for (int j = 0; j < MAXW; j++) {
                uint32_t bits = mask[j];
                while (bits) {
                    const int t = __ffs(bits) - 1;
                    bits &= bits - 1;
                    if ((unsigned)rank < (unsigned)LIST_CAP) s_list[rank] = (uint16_t)(bin0 + 32 * j + t);
                    rank++;
                }
            }
            __syncthreads();
            const int n = min(LIST_CAP, total - base);
            for (int i = threadIdx.x; i < n; i += FILL_THREADS) {
                const int b = s_list[i];
                __stcs(out_rowidx + o + base + i, b);
                __stcs(out_x + o + base + i, (double)s_acc[b]);
                s_acc[b] = (Acc)0;
            }
            __syncthreads();
        }
        // touched bins whose sum is zero were dropped: their accumulators already hold zero
        if (threadIdx.x == 0) {
            counts_actual[c] = total;
            if (total != (int)(out_colptr[c + 1] - o)) flags[FLAG_SHRUNK] = 1;
        }
        __syncthreads();
    }
}

// ---- streaming count (regular maps) ------------------------------------------
// When every gene's bins are contiguous and their first bin is non-decreasing in gene
// order (genes in genomic / GTF order over sorted tiles — what Cell Ranger matrices and
// the reference's 200-kb tiles give), a CSC column is a stream of intervals [b0, b0+nh)
// sorted by b0, and the number of distinct bins it touches is the length of their union:
// an interval adds the bins no earlier interval covered (prefix max of the interval ends).
// One warp counts one column with shuffles only.
constexpr int STREAM_THREADS = 512;
constexpr int STREAM_WARPS = STREAM_THREADS / 32;

__device__ __forceinline__ int warp_incl_max(int v, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(CS_FULL, v, off);
        if (lane >= off) v = max(v, t);
    }
    return v;
}
constexpr int COUNT_UNROLL = 4;
__global__ void __launch_bounds__(STREAM_THREADS)
bin_count_stream_kernel(int N, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                        const uint32_t *__restrict__ gmap, int *__restrict__ counts)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * STREAM_WARPS;
    for (int c = blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5); c < N; c += nwarps) {
        const int64_t e0 = colptr[c], e1 = colptr[c + 1];
        int cov = 0, cnt = 0;
        for (int64_t e = e0; e < e1; e += 32 * COUNT_UNROLL) {
            int g[COUNT_UNROLL];
            uint32_t w[COUNT_UNROLL];
#pragma unroll
            for (int u = 0; u < COUNT_UNROLL; u++) {
                const int64_t idx = e + 32 * u + lane;
                g[u] = (idx < e1) ? __ldcs(rowidx + idx) : -1;
            }
#pragma unroll
            for (int u = 0; u < COUNT_UNROLL; u++) w[u] = (g[u] >= 0) ? __ldg(gmap + g[u]) : 0u;
#pragma unroll
            for (int u = 0; u < COUNT_UNROLL; u++) {
                const int nh = w[u] >> 20, b0 = w[u] & 0xFFFFF;
                const int hi = nh ? b0 + nh : 0;
                const int pm = warp_incl_max(hi, lane);
                int prev = __shfl_up_sync(CS_FULL, pm, 1);
                if (lane == 0) prev = 0;
                prev = max(prev, cov);
                if (nh) cnt += max(0, hi - max(b0, prev));
                cov = max(cov, __shfl_sync(CS_FULL, pm, 31));
            }
        }
        for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(CS_FULL, cnt, off);
        if (lane == 0) counts[c] = cnt;
    }
}

// ---- streaming fill (regular maps, sorted columns) -----------------------------
// With contiguous hits whose first bin is non-decreasing along the column (see
// bin_count_stream_kernel) the RANK of every bin a column touches follows from two prefix
// scans over its entries, M = max end of the intervals in front of an entry and C = number
// of distinct bins they cover:
//     bin b of interval [f, f+nh) is new when b >= M: rank C + (b - max(f, M));
//     otherwise it lies in [f, M), which is covered without gaps: rank C - (M - b).
// So the sums are accumulated directly in RANK space — a dense array of the column's output
// length — and neither a bin-indexed accumulator nor its compaction exists.  One CTA per
// column, one warp per tile of 256 entries (eight consecutive entries per lane, 128-bit
// loads); tiles are chained through shared memory in super-steps of 16 tiles: tile maxima
// -> barrier -> new-bin counts -> barrier -> scatter (one ATOMS.ADD per hit; the bin id of
// a new bin is stored straight to the output).  Values are int32-exact as in
// bin_fill_kernel (same guard, same FP64 re-run by the caller).  A column whose row indices
// are not ascending raises FLAG_IRREG and the caller falls back to the generic kernels.
#ifndef SF_THREADS_N
#define SF_THREADS_N 256
#endif
#ifndef SF_CAP_N
#define SF_CAP_N 8192
#endif
constexpr int SF_THREADS = SF_THREADS_N;
constexpr int SF_WARPS = SF_THREADS / 32;
constexpr int SF_CAP = SF_CAP_N;      // ranks accumulated per pass over a column (longer columns take more passes)
constexpr int SF_EPL = 8;             // consecutive entries per lane (two 128-bit groups)
constexpr int SF_TILE = 32 * SF_EPL;  // entries per warp tile

__global__ void __launch_bounds__(SF_THREADS, 1024 / SF_THREADS)
bin_fill_stream_kernel(int N, int B, const int64_t *__restrict__ colptr,
                       const int32_t *__restrict__ rowidx, const double *__restrict__ x,
                       const uint32_t *__restrict__ gmap, const int64_t *__restrict__ out_colptr,
                       int32_t *__restrict__ out_rowidx, double *__restrict__ out_x, int64_t out_cap,
                       int *__restrict__ counts_actual, int *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    int *s_acc = reinterpret_cast<int *>(s_raw); // rank-indexed sums: a window of SF_CAP ranks
    __shared__ int s_tmax[32], s_tnew[32], s_glast[32];
    __shared__ unsigned long long s_mag[SF_WARPS];
    __shared__ int s_nz[SF_WARPS];

    if (out_colptr[N] > out_cap) { // caller's buffers too small: report, write nothing
        if (blockIdx.x == 0 && threadIdx.x == 0) flags[FLAG_CAP] = 1;
        return;
    }
    for (int b = threadIdx.x; b < SF_CAP; b += SF_THREADS) s_acc[b] = 0;
    if (threadIdx.x < 32) { // (entries past the warps of the CTA stay neutral)
        s_tmax[threadIdx.x] = 0;
        s_tnew[threadIdx.x] = 0;
        s_glast[threadIdx.x] = -1;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool before_me = lane < wid; // lanes that stand for the tiles in front of this warp's

    // (the offsets of the next column are fetched while the current one is processed)
    int64_t ne0 = 0, ne1 = 0, no0 = 0, no1 = 0;
    if ((int)blockIdx.x < N) {
        ne0 = colptr[blockIdx.x], ne1 = colptr[blockIdx.x + 1];
        no0 = out_colptr[blockIdx.x], no1 = out_colptr[blockIdx.x + 1];
    }
    for (int c = blockIdx.x; c < N; c += gridDim.x) {
        const int64_t e0 = ne0, e1 = ne1;
        const int64_t o = no0;
        const int expect = (int)(no1 - o);
        if (c + (int)gridDim.x < N) {
            ne0 = colptr[c + gridDim.x], ne1 = colptr[c + gridDim.x + 1];
            no0 = out_colptr[c + gridDim.x], no1 = out_colptr[c + gridDim.x + 1];
        }
        int32_t *__restrict__ ocol_i = out_rowidx + o;
        double *__restrict__ ocol_x = out_x + o;
        const int64_t tb = e0 & ~(int64_t)3;
        const int ntiles = (int)((e1 - tb + SF_TILE - 1) / SF_TILE);
        bool inexact = false, irregular = false;
        unsigned long long magsum = 0;
        int nz = 0;
        // (a column with more than SF_CAP distinct bins is streamed once per window of ranks)
        for (int rb = 0; rb == 0 || rb < expect; rb += SF_CAP) {
        int carryM = 0, carryC = 0, prev_glast = -1;
        if (rb) magsum = 0;
        for (int t0 = 0; t0 < ntiles; t0 += SF_WARPS) {
            // ---- this warp's tile: SF_EPL consecutive entries per lane (warps without one only
            // keep the barriers company)
            const int64_t tile0 = tb + (int64_t)(t0 + wid) * SF_TILE;
            const bool has = tile0 < e1;
            const int64_t base = tile0 + SF_EPL * lane;
            uint32_t w[SF_EPL];
            int iv[SF_EPL];
            int incl = 0, tfirst = INT_MAX, mbase = 0, cinc = 0, lnew = 0;
            if (has) {
                int g[SF_EPL];
                double v[SF_EPL];
#pragma unroll
                for (int q = 0; q < SF_EPL; q += 4) {
                    const int64_t b4 = base + q;
                    g[q] = g[q + 1] = g[q + 2] = g[q + 3] = -1;
                    v[q] = v[q + 1] = v[q + 2] = v[q + 3] = 0.0;
                    if (b4 >= e0 && b4 + 4 <= e1) {
                        const int4 g4 = __ldcs(reinterpret_cast<const int4 *>(rowidx + b4));
                        const double2 v01 = __ldcs(reinterpret_cast<const double2 *>(x + b4));
                        const double2 v23 = __ldcs(reinterpret_cast<const double2 *>(x + b4 + 2));
                        g[q] = g4.x, g[q + 1] = g4.y, g[q + 2] = g4.z, g[q + 3] = g4.w;
                        v[q] = v01.x, v[q + 1] = v01.y, v[q + 2] = v23.x, v[q + 3] = v23.y;
                    } else if (b4 + 4 > e0 && b4 < e1) { // the column's ragged ends
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            if (b4 + i >= e0 && b4 + i < e1) {
                                g[q + i] = rowidx[b4 + i];
                                v[q + i] = x[b4 + i];
                            }
                    }
                }
                int lmax = 0, gfirst = INT_MAX, glast = -1;
                unsigned magl = 0; // sum of |v| x min(hits, 2) over this lane's entries (16 x 2^24 fits)
#pragma unroll
                for (int i = 0; i < SF_EPL; i++) {
                    w[i] = (g[i] >= 0) ? __ldg(gmap + g[i]) : 0u;
                    const int f = (int)(w[i] & 0xFFFFF), nh = (int)(w[i] >> 20);
                    lmax = max(lmax, nh ? f + nh : 0);
                    iv[i] = 0;
                    if (g[i] >= 0) {
                        // (integral and below 2^24 in magnitude, or the FP64 path takes the matrix)
                        iv[i] = __double2int_rz(v[i]);
                        if (!(__int2double_rn(iv[i]) == v[i] && (unsigned)(iv[i] + (1 << 24)) < (1u << 25))) inexact = true;
                        magl += (unsigned)abs(iv[i]) << (nh >= 2 ? 1 : 0);
                        if (g[i] < glast) irregular = true; // (row indices must ascend)
                        gfirst = min(gfirst, g[i]);
                        glast = g[i];
                    }
                }
                magsum += magl;
                // ascending across lanes (interior lanes are all valid: the neighbour's last entry is enough)
                const int before = __shfl_up_sync(CS_FULL, glast, 1);
                if (lane > 0 && gfirst < before) irregular = true;
                incl = warp_incl_max(lmax, lane);
                tfirst = __reduce_min_sync(CS_FULL, gfirst);
                const int tlast = __reduce_max_sync(CS_FULL, glast);
                if (lane == 31) {
                    s_tmax[wid] = incl;
                    s_glast[wid] = tlast;
                }
            } else if (lane == 31) {
                s_tmax[wid] = 0;
                s_glast[wid] = -1;
                s_tnew[wid] = 0;
            }
            __syncthreads();
            // ---- prefix max of the interval ends in front of this lane's entries (lane l reads tile l's)
            const int tm = s_tmax[lane], tl = s_glast[lane];
            const int step_max = max(carryM, __reduce_max_sync(CS_FULL, tm));
            const int gall = max(prev_glast, __reduce_max_sync(CS_FULL, tl));
            if (has) {
                const int m_in = max(carryM, __reduce_max_sync(CS_FULL, before_me ? tm : 0));
                const int gl = max(prev_glast, __reduce_max_sync(CS_FULL, before_me ? tl : -1));
                if (tfirst < gl) irregular = true;
                int excl = __shfl_up_sync(CS_FULL, incl, 1);
                if (lane == 0) excl = 0;
                mbase = max(m_in, excl);
                int m = mbase;
#pragma unroll
                for (int i = 0; i < SF_EPL; i++) {
                    const int f = (int)(w[i] & 0xFFFFF), nh = (int)(w[i] >> 20);
                    if (nh) {
                        lnew += max(0, f + nh - max(f, m));
                        m = max(m, f + nh);
                    }
                }
                cinc = lnew;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const int tprev = __shfl_up_sync(CS_FULL, cinc, off);
                    if (lane >= off) cinc += tprev;
                }
                if (lane == 31) s_tnew[wid] = cinc;
            }
            __syncthreads();
            const int tn = s_tnew[lane];
            const int step_new = (int)__reduce_add_sync(CS_FULL, (unsigned)tn);
            if (has) {
                const int c_in = carryC + (int)__reduce_add_sync(CS_FULL, before_me ? (unsigned)tn : 0u);
                // ---- scatter in rank space: bin b goes to b + (b < m ? C - m : C - max(f, m))
                int cr = c_in + cinc - lnew, m = mbase;
#pragma unroll
                for (int i = 0; i < SF_EPL; i++) {
                    const int f = (int)(w[i] & 0xFFFFF), nh = (int)(w[i] >> 20);
                    const int off_old = cr - m, off_new = cr - max(f, m), m0 = m;
                    if (nh) {
                        {
                            const bool isnew = f >= m0;
                            const int rank = f + (isnew ? off_new : off_old);
                            if ((unsigned)(rank - rb) < (unsigned)SF_CAP && rank < expect) {
                                atomicAdd(&s_acc[rank - rb], iv[i]);
                                if (isnew) ocol_i[rank] = f;
                            } else if ((unsigned)rank >= (unsigned)expect) {
                                irregular = true;
                            }
                        }
                        if (nh >= 2) {
                            const int b = f + 1;
                            const bool isnew = b >= m0;
                            const int rank = b + (isnew ? off_new : off_old);
                            if ((unsigned)(rank - rb) < (unsigned)SF_CAP && rank < expect) {
                                atomicAdd(&s_acc[rank - rb], iv[i]);
                                if (isnew) ocol_i[rank] = b;
                            } else if ((unsigned)rank >= (unsigned)expect) {
                                irregular = true;
                            }
                        }
                        cr += max(0, f + nh - max(f, m0));
                        m = max(m0, f + nh);
                    }
                    if (nh >= 3) { // genes over three bins or more (2 % of them)
                        magsum += (unsigned long long)(unsigned)abs(iv[i]) * (unsigned)(nh - 2);
#pragma unroll 1
                        for (int k = 2; k < nh; k++) {
                            const int b = f + k;
                            const bool isnew = b >= m0;
                            const int rank = b + (isnew ? off_new : off_old);
                            if ((unsigned)(rank - rb) < (unsigned)SF_CAP && rank < expect) {
                                atomicAdd(&s_acc[rank - rb], iv[i]);
                                if (isnew) ocol_i[rank] = b;
                            } else if ((unsigned)rank >= (unsigned)expect) {
                                irregular = true;
                            }
                        }
                    }
                }
            }
            carryM = step_max;
            carryC += step_new;
            prev_glast = gall;
            if (t0 + SF_WARPS < ntiles) __syncthreads(); // (the tile slots are rewritten by the next super-step)
        }
        // exactness guard: every value integral and the column cannot overflow int32
        for (int off = 16; off >= 1; off >>= 1) magsum += __shfl_xor_sync(CS_FULL, magsum, off);
        if (lane == 0) s_mag[wid] = magsum;
        if (inexact) flags[FLAG_INEXACT] = 1;
        if (irregular || carryC != expect) flags[FLAG_IRREG] = 1;
        __syncthreads();
        if (wid == 0) {
            unsigned long long mm = (lane < SF_WARPS) ? s_mag[lane] : 0ull;
            for (int off = 16; off >= 1; off >>= 1) mm += __shfl_xor_sync(CS_FULL, mm, off);
            if (lane == 0 && mm >= 2147483647ull) flags[FLAG_INEXACT] = 1;
        }
        // ---- the sums, in rank order; zero sums are kept here and closed up by the caller
        const int total = min(min(carryC, expect) - rb, SF_CAP);
        for (int r = threadIdx.x; r < total; r += SF_THREADS) {
            const int sv = s_acc[r];
            s_acc[r] = 0;
            __stcs(ocol_x + rb + r, (double)sv);
            nz += (sv != 0) ? 1 : 0;
        }
        if (rb + SF_CAP < expect) __syncthreads(); // (the window is refilled by the next pass)
        } // rank windows
        nz = (int)__reduce_add_sync(CS_FULL, (unsigned)nz);
        if (lane == 0) s_nz[wid] = nz;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tnz = 0;
            for (int w2 = 0; w2 < SF_WARPS; w2++) tnz += s_nz[w2];
            counts_actual[c] = tnz;
            if (tnz != expect) flags[FLAG_SHRUNK] = 1;
        }
    }
}

// ---- one-pass streaming aggregation (regular maps, sorted columns) --------------------------
// The same rank-space idea as bin_fill_stream_kernel, organised so that (i) the matrix is read
// from DRAM once and the result written once — no separate count kernel, no column scan: a
// column's output offset comes from a decoupled look-back over the per-column counts — and
// (ii) there is no shared-memory atomic per hit (ATOMS retires one lane every two cycles:
// at ~2,500 hits per column it alone costs more than the column's share of a 60 %-of-HBM
// budget).  Almost every hit opens a bin of its own (2 genes per 200-kb tile, 7 % of them
// expressed in a cell), so the NEW part of every interval — [max(f, M), e) — is written with
// plain stores, and only the part that falls on bins an earlier entry opened — [f, min(e, M)),
// a tenth of the hits — is added atomically, after the stores of the tile are visible.
//
// One WARP per column (columns are handed out by a ticket, so a column's predecessors are
// always held by running warps), no CTA barrier anywhere.  A column is traversed twice in tiles
// of 256 entries (eight consecutive entries per lane, 256-bit loads):
//   count  row indices only: interval ends, prefix max M across the tile (lane-serial over the
//          eight entries + one warp scan), number of new bins -> the column's count; published;
//   fill   row indices again (L2 hits: they were read microseconds ago) + values.  The ranks a
//          tile opens are appended to a LINEAR staging buffer of the warp ({sum, bin} pairs in
//          rank order): a lane walks its eight entries with a running shared-memory address —
//          new bins are stores at immediate offsets, the old part of an entry is one address
//          kept for after the stores — no ring arithmetic.  Ranks below C - (M - f_last) can no
//          longer be touched (first bins do not decrease); when the next tile would not fit they
//          are written to the output in rank order (coalesced) and the at most OP_MAXNH ranks
//          behind them move to the front of the buffer.  The exclusive prefix over the columns
//          in front is collected (look-back) only at the first such write, by which time they
//          have long published their counts.
// Exactness as in the other kernels: values must be integers (checked), and a column whose
// n_entries x 2^bits(max |v|) could overflow an int32 sum raises FLAG_INEXACT (FP64 re-run).
constexpr int OP_THREADS = 256;
constexpr int OP_WARPS = OP_THREADS / 32;
constexpr int OP_SMAP_THREADS = 512; // one CTA per SM: 16 warps' buffers (80 KB) + the map table (120 KB for 30k genes), 128 registers
constexpr int OP_CAP = 640;  // ranks a warp stages ({sum, bin}: 8 bytes each)
constexpr int OP_EPL = 8;
constexpr int OP_TILE = 32 * OP_EPL;
constexpr int OP_MAXNH = 32; // widest gene, in bins, this kernel takes
constexpr unsigned long long OP_FLAG_A = 1ull << 62, OP_FLAG_P = 2ull << 62, OP_VALMASK = (1ull << 62) - 1;

__device__ __forceinline__ void op_load_g8(const int32_t *__restrict__ p, int (&g)[8], bool last_use)
{
    if (last_use)
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(g[0]), "=r"(g[1]), "=r"(g[2]), "=r"(g[3]), "=r"(g[4]), "=r"(g[5]), "=r"(g[6]), "=r"(g[7])
                     : "l"(p));
    else
        asm volatile("ld.global.nc.L1::no_allocate.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(g[0]), "=r"(g[1]), "=r"(g[2]), "=r"(g[3]), "=r"(g[4]), "=r"(g[5]), "=r"(g[6]), "=r"(g[7])
                     : "l"(p));
}
__device__ __forceinline__ void op_load_x4(const double *__restrict__ p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
                 : "l"(p));
}
__device__ __forceinline__ void op_prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long op_ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void op_st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ int warp_incl_sum(int v, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(CS_FULL, v, off);
        if (lane >= off) v += t;
    }
    return v;
}
// (shfl.up hands a lane below the offset its own value back: max needs no lane test)
__device__ __forceinline__ int warp_incl_max_nt(int v)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) v = max(v, __shfl_up_sync(CS_FULL, v, off));
    return v;
}

// row indices of this lane's eight entries and their packed map words.  Outside the column:
// -1 in front of it, INT_MAX behind it (so that "ascending" can be tested on all eight), word 0.
// row indices of this lane's eight entries: -1 in front of the column, INT_MAX behind it.  A group
// that straddles an end of the column is still one 256-bit load when it lies inside the array
// (`total` entries): the neighbours' entries it brings along are masked out.
template <bool EDGE>
__device__ __forceinline__ void op_load_rows(const int32_t *__restrict__ rowidx, int64_t gb, int64_t e0, int64_t e1,
                                             int64_t total, bool last_use, int (&g)[8])
{
    if (gb >= e0 && gb + OP_EPL <= e1) {
        op_load_g8(rowidx + gb, g, last_use);
    } else if (EDGE && gb < e1 && gb + OP_EPL > e0 && gb + OP_EPL <= total) {
        op_load_g8(rowidx + gb, g, false);
#pragma unroll
        for (int i = 0; i < OP_EPL; i++) g[i] = gb + i < e0 ? -1 : (gb + i >= e1 ? INT_MAX : g[i]);
    } else {
        // (predicated loads straight into the registers: nothing waits for them here)
#pragma unroll
        for (int i = 0; i < OP_EPL; i++) {
            g[i] = gb + i < e0 ? -1 : INT_MAX;
            if (gb + i >= e0 && gb + i < e1)
                asm volatile("ld.global.nc.s32 %0, [%1];" : "+r"(g[i]) : "l"(rowidx + gb + i));
        }
    }
}
__device__ __forceinline__ void op_load_x4_keep(const double *__restrict__ p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
// SMAP: the map words come from the CTA's shared-memory copy of the table (no L1 / L2 round trip)
template <bool SMAP>
__device__ __forceinline__ uint32_t op_map_word(const uint32_t *__restrict__ gmap, const uint32_t *s_gmap, int g)
{
    if (SMAP) return s_gmap[g];
    return __ldg(gmap + g);
}
// the staging buffer by 32-bit shared address ({sum, bin} pairs of 8 bytes)
__device__ __forceinline__ void op_st_pair_if(int cond_gt, int than, unsigned addr, int sum, int bin)
{
    asm volatile("{ .reg .pred p; setp.gt.s32 p, %0, %1; @p st.shared.v2.s32 [%2], {%3, %4}; }" ::"r"(cond_gt), "r"(than),
                 "r"(addr), "r"(sum), "r"(bin)
                 : "memory");
}
__device__ __forceinline__ void op_st_pair(unsigned addr, int sum, int bin)
{
    asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(addr), "r"(sum), "r"(bin) : "memory");
}
__device__ __forceinline__ void op_red_add(unsigned addr, int v)
{
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void op_ld_pair(unsigned addr, int &sum, int &bin)
{
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(sum), "=r"(bin) : "r"(addr) : "memory");
}

// the packed map words of eight rows (0 for the sentinels outside the column and for rows outside the table)
template <bool SMAP>
__device__ __forceinline__ void op_words_of(const uint32_t *__restrict__ gmap, const uint32_t *s_gmap, int G,
                                            const int (&g)[8], uint32_t (&w)[8])
{
#pragma unroll
    for (int i = 0; i < OP_EPL; i++) w[i] = ((unsigned)g[i] < (unsigned)G) ? op_map_word<SMAP>(gmap, s_gmap, g[i]) : 0u;
}

// THREADS x CTAS_PER_SM: 256 x {2,3,4} with the map read through L1, or — SMAP — one CTA of
// OP_SMAP_THREADS per SM that holds the whole map table in shared memory next to its warps' buffers
// OPT: 1 = count the next column before filling this one, 2 = rows of the next tile in flight during
// the fill, 4 = 256-bit loads on tiles that straddle a column end, 8 = four predicated stores per
// entry, 16 = the next tile's values prefetched into L1 (instead of two tiles ahead into L2)
template <int THREADS, int CTAS_PER_SM, bool SMAP, int OPT>
__global__ void __launch_bounds__(THREADS, CTAS_PER_SM)
bin_onepass_kernel(int N, int G, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                   const double *__restrict__ x, const uint32_t *__restrict__ gmap,
                   int64_t *__restrict__ out_colptr, int32_t *__restrict__ out_rowidx, double *__restrict__ out_x,
                   int64_t out_cap, int do_fill, unsigned long long *__restrict__ status, int *__restrict__ ticket,
                   int *__restrict__ counts_actual, int *__restrict__ flags, long long *__restrict__ total_out)
{
    extern __shared__ __align__(16) unsigned char s_op_raw[];
    int2 *s_buf_all = reinterpret_cast<int2 *>(s_op_raw);
    const uint32_t *s_gmap = reinterpret_cast<const uint32_t *>(s_op_raw + (size_t)(THREADS / 32) * OP_CAP * sizeof(int2));
    if (SMAP) {
        uint32_t *dst = const_cast<uint32_t *>(s_gmap);
        for (int i = threadIdx.x; i < G; i += THREADS) dst[i] = __ldg(gmap + i);
        __syncthreads();
    }
    constexpr bool PIPE = (OPT & 1) != 0, FILLPF = (OPT & 2) != 0, EDGE = (OPT & 4) != 0, ST4 = (OPT & 8) != 0,
                   XL1 = (OPT & 16) != 0, DEPTH2 = false;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned buf = (unsigned)__cvta_generic_to_shared(s_buf_all + (size_t)wid * OP_CAP);
    __shared__ long long s_total; // entries in the whole matrix (read back where a tile straddles a column end)
    if (threadIdx.x == 0) s_total = colptr[N];
    __syncthreads();
#define total ((int64_t)s_total)
    bool inexact = false, irregular = false;

    auto take_ticket = [&]() {
        int t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1);
        return __shfl_sync(CS_FULL, t, 0);
    };
    // ---------------- count: number of distinct bins column [e0, e1) touches; published at once
    // (rows two tiles ahead are in flight in registers, three tiles ahead on their way into L2)
    auto count_column = [&](int col, int64_t e0, int64_t e1) {
        const int64_t tb = e0 & ~(int64_t)(OP_EPL - 1);
        int cnt = 0, carryM = 0, carryG = -1;
        int gn[OP_EPL], gn2[OP_EPL];
        op_load_rows<EDGE>(rowidx, tb + OP_EPL * lane, e0, e1, total, !do_fill, gn);
        if (DEPTH2 && tb + OP_TILE < e1) op_load_rows<EDGE>(rowidx, tb + OP_TILE + OP_EPL * lane, e0, e1, total, !do_fill, gn2);
        for (int64_t t0 = tb; t0 < e1; t0 += OP_TILE) {
            int g[OP_EPL];
            uint32_t w[OP_EPL];
#pragma unroll
            for (int i = 0; i < OP_EPL; i++) {
                g[i] = gn[i];
                if (DEPTH2) gn[i] = gn2[i];
            }
            if (lane < 8 && t0 + 3 * OP_TILE + 32 * lane < e1) op_prefetch_l2(rowidx + t0 + 3 * OP_TILE + 32 * lane);
            if (DEPTH2) {
                if (t0 + 2 * OP_TILE < e1) op_load_rows<EDGE>(rowidx, t0 + 2 * OP_TILE + OP_EPL * lane, e0, e1, total, !do_fill, gn2);
            } else {
                if (t0 + OP_TILE < e1) op_load_rows<EDGE>(rowidx, t0 + OP_TILE + OP_EPL * lane, e0, e1, total, !do_fill, gn);
            }
            op_words_of<SMAP>(gmap, s_gmap, G, g, w);
            int before = __shfl_up_sync(CS_FULL, g[OP_EPL - 1], 1);
            if (lane == 0) before = carryG;
            bool bad = g[0] < before; // (row indices must ascend)
            int lmax = 0;
#pragma unroll
            for (int i = 0; i < OP_EPL; i++) {
                lmax = max(lmax, (int)(w[i] & 0xFFFFF) + (int)(w[i] >> 20));
                if (i) bad |= g[i] < g[i - 1];
            }
            irregular |= bad;
            carryG = __shfl_sync(CS_FULL, g[OP_EPL - 1], 31);
            const int incl = warp_incl_max_nt(lmax);
            int m = __shfl_up_sync(CS_FULL, incl, 1);
            if (lane == 0) m = 0;
            m = max(m, carryM);
#pragma unroll
            for (int i = 0; i < OP_EPL; i++) {
                const int f = (int)(w[i] & 0xFFFFF), e = f + (int)(w[i] >> 20);
                cnt += max(0, e - max(f, m));
                m = max(m, e);
            }
            carryM = max(carryM, __shfl_sync(CS_FULL, incl, 31));
        }
        cnt = (int)__reduce_add_sync(CS_FULL, (unsigned)cnt);
        if (lane == 0) op_st_status(status + col, OP_FLAG_A | (unsigned long long)cnt);
        return cnt;
    };

    // Columns are handed out by a ticket.  A warp counts (and publishes) its NEXT column before it
    // fills the current one: by the time the fill needs the prefix over the columns in front
    // (look-back, at its first write-out) every warp that holds one of them has long published it.
    int c = take_ticket();
    int64_t e0 = 0, e1 = 0;
    int cnt = 0;
    if (c < N) {
        e0 = colptr[c];
        e1 = colptr[c + 1];
        cnt = count_column(c, e0, e1);
    }
    while (c < N) {
        int cn = N, ncnt = 0; // (only these two stay live across the fill: the bounds are read again)
        if (PIPE && do_fill) {
            cn = take_ticket();
            if (cn < N) ncnt = count_column(cn, colptr[cn], colptr[cn + 1]);
        }
        const int64_t tb = e0 & ~(int64_t)(OP_EPL - 1);
        long long excl = 0;
        auto look_back = [&]() {
            int j = c - 1;
            while (j >= 0) {
                const int idx = j - lane;
                unsigned long long sv = OP_FLAG_P; // (in front of column 0: an inclusive prefix of zero)
                if (idx >= 0) {
                    sv = op_ld_status(status + idx);
                    while ((sv >> 62) == 0ull) {
                        __nanosleep(200);
                        sv = op_ld_status(status + idx);
                    }
                }
                const unsigned pm = __ballot_sync(CS_FULL, (sv >> 62) == 2ull);
                const int first = pm ? __ffs(pm) - 1 : 32; // nearest column whose inclusive prefix is known
                long long v = (lane <= first) ? (long long)(sv & OP_VALMASK) : 0ll;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(CS_FULL, v, off);
                excl += v;
                if (pm) break;
                j -= 32;
            }
            if (lane == 0) {
                op_st_status(status + c, OP_FLAG_P | (unsigned long long)(excl + cnt));
                out_colptr[c] = excl;
                if (c == N - 1) {
                    out_colptr[N] = excl + cnt;
                    *total_out = excl + cnt;
                }
            }
        };
        if (!do_fill) {
            look_back();
            c = take_ticket();
            if (c < N) {
                e0 = colptr[c];
                e1 = colptr[c + 1];
                cnt = count_column(c, e0, e1);
            }
            continue;
        }
        // ---------------- fill
        bool have_excl = false, skip = false;
        int32_t *__restrict__ ocol_i = out_rowidx; // (advance with every write-out)
        double *__restrict__ ocol_x = out_x;
        int carryM = 0, carryF = 0, held = 0, written = 0, nz = 0;
        unsigned orbits = 0;
        // ranks [0, nfin) of the buffer go to the output; the rest moves to the front
        auto write_out = [&](int nfin) {
            if (!have_excl) {
                look_back();
                have_excl = true;
                if (excl + cnt > out_cap) { // caller's buffers too small: report, write nothing
                    if (lane == 0) flags[FLAG_CAP] = 1;
                    skip = true;
                    return;
                }
                ocol_i += excl;
                ocol_x += excl;
            }
            for (int r = lane; r < nfin; r += 32) {
                int sv, bv;
                op_ld_pair(buf + 8u * r, sv, bv);
                __stcs(ocol_x + r, (double)sv);
                __stcs(ocol_i + r, bv);
                nz += (sv != 0) ? 1 : 0;
            }
            ocol_i += nfin;
            ocol_x += nfin;
            written += nfin;
            const int rem = held - nfin; // (<= OP_MAXNH)
            __syncwarp();
            for (int r0 = 0; r0 < rem; r0 += 32) {
                int sv = 0, bv = 0;
                if (r0 + lane < rem) op_ld_pair(buf + 8u * (nfin + r0 + lane), sv, bv);
                __syncwarp();
                if (r0 + lane < rem) op_st_pair(buf + 8u * (r0 + lane), sv, bv);
            }
            __syncwarp();
            held = rem;
        };
        int gn[OP_EPL]; // the next tile's rows, in flight while this one is worked on
        if (FILLPF) op_load_rows<EDGE>(rowidx, tb + OP_EPL * lane, e0, e1, total, true, gn);
        for (int64_t t0 = tb; t0 < e1 && !skip; t0 += OP_TILE) {
            const int64_t gb = t0 + OP_EPL * lane;
            int g[OP_EPL], iv[OP_EPL];
            uint32_t w[OP_EPL];
            double v[OP_EPL];
            if (FILLPF) {
#pragma unroll
                for (int i = 0; i < OP_EPL; i++) g[i] = gn[i];
            } else {
                op_load_rows<EDGE>(rowidx, gb, e0, e1, total, true, g);
            }
            // this tile's values are asked for now and converted only when the placement needs them
            const bool full = gb >= e0 && gb + OP_EPL <= e1;
            if (full) {
                op_load_x4(x + gb, v[0], v[1], v[2], v[3]);
                op_load_x4(x + gb + 4, v[4], v[5], v[6], v[7]);
            } else if (EDGE && gb < e1 && gb + OP_EPL > e0 && gb + OP_EPL <= total) { // (the neighbours' values are masked out)
                op_load_x4_keep(x + gb, v[0], v[1], v[2], v[3]);
                op_load_x4_keep(x + gb + 4, v[4], v[5], v[6], v[7]);
#pragma unroll
                for (int i = 0; i < OP_EPL; i++) v[i] = (gb + i >= e0 && gb + i < e1) ? v[i] : 0.0;
            } else {
#pragma unroll
                for (int i = 0; i < OP_EPL; i++) {
                    v[i] = 0.0;
                    if (gb + i >= e0 && gb + i < e1) asm volatile("ld.global.cs.f64 %0, [%1];" : "+d"(v[i]) : "l"(x + gb + i));
                }
            }
            if (XL1) {
                if (lane < 16 && t0 + OP_TILE + 16 * lane < e1) asm volatile("prefetch.global.L1 [%0];" ::"l"(x + t0 + OP_TILE + 16 * lane));
            } else {
                if (lane < 16 && t0 + 2 * OP_TILE + 16 * lane < e1) op_prefetch_l2(x + t0 + 2 * OP_TILE + 16 * lane);
            }
            if (FILLPF && t0 + OP_TILE < e1) op_load_rows<EDGE>(rowidx, gb + OP_TILE, e0, e1, total, true, gn);
            op_words_of<SMAP>(gmap, s_gmap, G, g, w);
            // interval ends, first bins, the prefix max M in front of this lane's entries
            int lmax = 0, lf = 0;
#pragma unroll
            for (int i = 0; i < OP_EPL; i++) {
                lmax = max(lmax, (int)(w[i] & 0xFFFFF) + (int)(w[i] >> 20));
                lf = max(lf, (int)(w[i] & 0xFFFFF)); // (starts do not decrease: the last one with a bin is the largest)
            }
            const int incl = warp_incl_max_nt(lmax);
            int m0 = __shfl_up_sync(CS_FULL, incl, 1);
            if (lane == 0) m0 = 0;
            m0 = max(m0, carryM);
            int ln = 0; // bins this lane's entries open
            {
                int m = m0;
#pragma unroll
                for (int i = 0; i < OP_EPL; i++) {
                    const int f = (int)(w[i] & 0xFFFFF), e = f + (int)(w[i] >> 20);
                    ln += max(0, e - max(f, m));
                    m = max(m, e);
                }
            }
            const int cincl = warp_incl_sum(ln, lane);
            const int need = __shfl_sync(CS_FULL, cincl, 31);
            if (held + need > OP_CAP) { // make room: everything no later entry can touch goes out
                write_out(held - max(0, carryM - carryF));
                if (skip) break;
                if (held + need > OP_CAP) { // (a tile of very wide genes: not this kernel's case)
                    irregular = true;
                    skip = true;
                    break;
                }
            }
#pragma unroll
            for (int i = 0; i < OP_EPL; i++) {
                iv[i] = __double2int_rz(v[i]);
                inexact |= __int2double_rn(iv[i]) != v[i];
                orbits |= (unsigned)abs(iv[i]);
            }
            // placement: new bins are stores at the lane's running address, the old part of an entry
            // (bins an earlier entry opened: ranks just below the running one) is remembered
            unsigned oa[OP_EPL];
            bool multi = false;
            {
                int m = m0;
                unsigned addr = buf + 8u * (unsigned)(held + cincl - ln);
#pragma unroll
                for (int i = 0; i < OP_EPL; i++) {
                    const int f = (int)(w[i] & 0xFFFFF), e = f + (int)(w[i] >> 20);
                    const int s = max(f, m), nn = e - s, old = min(e, m) - f;
                    op_st_pair_if(nn, 0, addr, iv[i], s);
                    op_st_pair_if(nn, 1, addr + 8u, iv[i], s + 1);
                    if (ST4) {
                        op_st_pair_if(nn, 2, addr + 16u, iv[i], s + 2);
                        op_st_pair_if(nn, 3, addr + 24u, iv[i], s + 3);
                    }
                    if (nn > (ST4 ? 4 : 2)) {
#pragma unroll 1
                        for (int k = (ST4 ? 4 : 2); k < nn; k++) op_st_pair(addr + 8u * k, iv[i], s + k);
                    }
                    oa[i] = old > 0 ? addr - 8u * (unsigned)(m - f) : 0u;
                    multi |= old > 1;
                    addr += 8u * (unsigned)max(nn, 0);
                    m = max(m, e);
                }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < OP_EPL; i++)
                if (oa[i]) op_red_add(oa[i], iv[i]);
            if (__any_sync(CS_FULL, multi)) { // entries that reach over two or more old bins (rare)
                int m = m0;
#pragma unroll
                for (int i = 0; i < OP_EPL; i++) {
                    const int f = (int)(w[i] & 0xFFFFF), e = f + (int)(w[i] >> 20);
                    const int old = min(e, m) - f;
#pragma unroll 1
                    for (int k = 1; k < old; k++) op_red_add(oa[i] + 8u * k, iv[i]);
                    m = max(m, e);
                }
            }
            __syncwarp();
            held += need;
            carryM = max(carryM, __shfl_sync(CS_FULL, incl, 31));
            carryF = max(carryF, (int)__reduce_max_sync(CS_FULL, (unsigned)lf));
            if (written + held > cnt) { // (cannot happen for ascending rows: the count pass saw the same entries)
                irregular = true;
                skip = true;
            }
        }
        if (!skip) write_out(held);
        if (!have_excl) look_back(); // (empty columns, or columns given up on)
        nz = (int)__reduce_add_sync(CS_FULL, (unsigned)nz);
        orbits = __reduce_or_sync(CS_FULL, orbits);
        if (lane == 0 && !skip) {
            counts_actual[c] = nz;
            if (nz != cnt) flags[FLAG_SHRUNK] = 1;
            // an int32 sum takes at most (e1 - e0) values below 2^bits(orbits)
            const int bits = 32 - __clz(orbits);
            if (bits >= 31 || ((unsigned long long)(e1 - e0) << bits) > 2147483647ull) flags[FLAG_INEXACT] = 1;
        }
        if (!PIPE) {
            cn = take_ticket();
            if (cn < N) ncnt = count_column(cn, colptr[cn], colptr[cn + 1]);
        }
        c = cn;
        cnt = ncnt;
        if (c < N) {
            e0 = colptr[c];
            e1 = colptr[c + 1];
        }
    }
    if (inexact) flags[FLAG_INEXACT] = 1;
    if (irregular) flags[FLAG_IRREG] = 1;
#undef total
}

constexpr int OP_DEFAULT_OPT = 11; // (measured: tools/sweep_binning.py)
using op_kernel_t = decltype(&bin_onepass_kernel<OP_SMAP_THREADS, 1, true, 0>);
template <int T, int... I>
static op_kernel_t op_pick(int opt, std::integer_sequence<int, I...>)
{
    static const op_kernel_t table[] = {&bin_onepass_kernel<T, 1, true, I>...};
    return table[opt];
}

// closes the gaps left by dropped zero sums: column c's first nwritten entries (all of its
// padded range when nwritten == NULL) are copied, zeros skipped, to the compact offsets
__global__ void compact_columns_kernel(int N, const int64_t *__restrict__ old_colptr,
                                       const int *__restrict__ nwritten,
                                       const int64_t *__restrict__ new_colptr,
                                       const int32_t *__restrict__ src_i,
                                       const double *__restrict__ src_x,
                                       int32_t *__restrict__ dst_i, double *__restrict__ dst_x)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    for (int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < N; c += nwarps) {
        const int64_t so = old_colptr[c];
        int64_t d = new_colptr[c];
        const int n = nwritten ? nwritten[c] : (int)(old_colptr[c + 1] - so);
        for (int i = 0; i < n; i += 32) {
            const bool ok = i + lane < n;
            const double v = ok ? src_x[so + i + lane] : 0.0;
            const bool keep = ok && v != 0.0;
            const unsigned m = __ballot_sync(CS_FULL, keep);
            if (keep) {
                const int64_t pos = d + __popc(m & ((1u << lane) - 1u));
                dst_i[pos] = src_i[so + i + lane];
                dst_x[pos] = v;
            }
            d += __popc(m);
        }
    }
}

} // namespace

// ----------------------------------------------------------------- plan / API
struct cs_bin_plan {
    int G = 0, B = 0;
    bool regular = false; // contiguous hits with non-decreasing first bin: streaming kernels apply
    bool stream_fill = true; // (cleared for good when a matrix with unsorted columns shows up)
    bool onepass = false; // regular and no gene wider than OP_MAXNH bins: the one-pass kernel applies
    DevBuf<unsigned long long> status; // one-pass kernel: per-column look-back words
    DevBuf<int> ticket;
    DevBuf<uint32_t> gmap;
    DevBuf<int32_t> map_ptr, map_bin;
    DevBuf<int> counts, counts2, flags;
    DevBuf<int64_t> tile_sums, grand, colptr2;
    DevBuf<int32_t> tmp_i;
    DevBuf<double> tmp_x;
    int *h_pinned = nullptr; // NFLAGS ints + 1 int64 (nnz)
    const void *op_kern = nullptr; // launch configuration of the one-pass kernel, found once
    size_t op_smem = 0;
    int op_per_sm = 1;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}; // count | scan | fill boundaries
    bool timed = false;
    ~cs_bin_plan()
    {
        if (h_pinned) cudaFreeHost(h_pinned);
        for (auto e : ev)
            if (e) cudaEventDestroy(e);
    }
};

static size_t fill_smem_bytes(int B, size_t acc_size)
{
    return (size_t)((B + 31) & ~31) * acc_size + (size_t)(B < LIST_CAP ? B : LIST_CAP) * 2 + 16;
}

extern "C" int cs_bin_plan_create(int G, int B, const int32_t *map_ptr, const int32_t *map_bin,
                                  cs_bin_plan **plan_out)
{
    CS_REQUIRE(plan_out != nullptr, "cs_bin_plan_create: plan is NULL");
    *plan_out = nullptr;
    CS_REQUIRE(G >= 0 && B >= 1 && map_ptr, "cs_bin_plan_create: bad G/B/map_ptr");
    CS_REQUIRE(B <= 65535, "cs_bin_plan_create: B=%d bins exceeds the 65535 this build supports", B);
    if (fill_smem_bytes(B, sizeof(int)) > 227 * 1024) {
        cs_set_error("cs_bin_plan_create: B=%d bins need more than 227 KB of shared memory", B);
        return CS_ERR_UNSUPPORTED;
    }
    const int nhits = map_ptr[G];
    for (int g = 0; g < G; g++) CS_REQUIRE(map_ptr[g] <= map_ptr[g + 1], "map_ptr not monotone at gene %d", g);
    for (int m = 0; m < nhits; m++)
        CS_REQUIRE(map_bin[m] >= 0 && map_bin[m] < B, "map_bin[%d]=%d outside [0,%d)", m, map_bin[m], B);
    std::vector<uint32_t> gm((size_t)G > 0 ? G : 1);
    for (int g = 0; g < G; g++) {
        const int m0 = map_ptr[g], nh = map_ptr[g + 1] - m0;
        bool contig = nh < 4096;
        for (int t = 1; t < nh && contig; t++) contig = map_bin[m0 + t] == map_bin[m0] + t;
        gm[g] = nh == 0 ? 0u : (contig ? ((uint32_t)nh << 20) | (uint32_t)map_bin[m0] : GMAP_ESC);
    }
    cs_bin_plan *p = new cs_bin_plan();
    p->G = G;
    p->B = B;
    {
        bool reg = true;
        int last_b0 = 0;
        for (int g = 0; g < G && reg; g++) {
            if (gm[g] == 0u) continue; // no hits
            if (gm[g] == GMAP_ESC) { reg = false; break; }
            const int b0 = gm[g] & 0xFFFFF;
            if (b0 < last_b0) reg = false;
            last_b0 = b0;
        }
        p->regular = reg;
        int maxnh = 0;
        for (int g = 0; g < G; g++)
            if (gm[g] != GMAP_ESC) maxnh = maxnh > (int)(gm[g] >> 20) ? maxnh : (int)(gm[g] >> 20);
        p->onepass = reg && maxnh <= OP_MAXNH;
    }
    int rc = CS_OK;
    do {
#define TRY(call) if ((call) != cudaSuccess) { cs_set_error("cs_bin_plan_create: %s", cudaGetErrorString(cudaGetLastError())); rc = CS_ERR_CUDA; break; }
        TRY(p->gmap.alloc(G));
        TRY(p->map_ptr.alloc(G + 1));
        TRY(p->map_bin.alloc(nhits));
        TRY(p->flags.alloc(NFLAGS));
        TRY(p->grand.alloc(1));
        TRY(p->ticket.alloc(1));
        TRY(cudaMallocHost((void **)&p->h_pinned, 64));
        for (int i = 0; i < 4; i++) TRY(cudaEventCreate(&p->ev[i]));
        TRY(cudaMemcpy(p->gmap.p, gm.data(), sizeof(uint32_t) * G, cudaMemcpyHostToDevice));
        TRY(cudaMemcpy(p->map_ptr.p, map_ptr, sizeof(int32_t) * (G + 1), cudaMemcpyHostToDevice));
        if (nhits) TRY(cudaMemcpy(p->map_bin.p, map_bin, sizeof(int32_t) * nhits, cudaMemcpyHostToDevice));
#undef TRY
    } while (0);
    if (rc != CS_OK) {
        delete p;
        return rc;
    }
    *plan_out = p;
    return CS_OK;
}

extern "C" void cs_bin_plan_destroy(cs_bin_plan *plan) { delete plan; }

extern "C" int cs_bin_plan_timing(cs_bin_plan *p, double *count_ms, double *scan_ms, double *fill_ms)
{
    CS_REQUIRE(p && p->timed, "cs_bin_plan_timing: no completed count+fill call to report");
    float ms;
    CS_CUDA(cudaEventElapsedTime(&ms, p->ev[0], p->ev[1]));
    if (count_ms) *count_ms = ms;
    CS_CUDA(cudaEventElapsedTime(&ms, p->ev[1], p->ev[2]));
    if (scan_ms) *scan_ms = ms;
    CS_CUDA(cudaEventElapsedTime(&ms, p->ev[2], p->ev[3]));
    if (fill_ms) *fill_ms = ms;
    return CS_OK;
}

template <typename Acc>
static int launch_fill(cs_bin_plan *p, int N, const int64_t *d_colptr, const int32_t *d_rowidx,
                       const double *d_x, const int64_t *d_out_colptr, int32_t *d_out_rowidx,
                       double *d_out_x, int64_t out_cap, cudaStream_t st)
{
    const size_t smem = fill_smem_bytes(p->B, sizeof(Acc));
    if (smem > 227 * 1024) {
        cs_set_error("cs_bin_counts: B=%d bins with FP64 accumulators need %zu B of shared memory", p->B, smem);
        return CS_ERR_UNSUPPORTED;
    }
    CS_CUDA(cudaFuncSetAttribute(bin_fill_kernel<Acc>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bin_fill_kernel<Acc>, FILL_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int grid = cs_num_sms() * per_sm;
    if (grid > N) grid = N;
    bin_fill_kernel<Acc><<<grid, FILL_THREADS, smem, st>>>(
        N, p->B, d_colptr, d_rowidx, d_x, p->gmap.p, p->map_ptr.p, p->map_bin.p, d_out_colptr,
        d_out_rowidx, d_out_x, out_cap, p->counts2.p, p->flags.p);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

static int run_scan(cs_bin_plan *p, int N, const int *counts, int64_t *out, cudaStream_t st)
{
    const int ntiles = (N + SCAN_TILE - 1) / SCAN_TILE;
    CS_CUDA(p->tile_sums.reserve(ntiles));
    scan_tiles_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(N, counts, out, p->tile_sums.p);
    scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(ntiles, p->tile_sums.p, p->grand.p);
    scan_add_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(N, out, p->tile_sums.p, p->grand.p);
    CS_CUDA(cudaGetLastError());
    return CS_OK;
}

// what follows a fill whose flags are in p->h_pinned: the FP64 re-run for values an int32 sum cannot
// carry, and the compaction of bins whose sum is zero (zeros_inline: the fill left them in place)
static int finish_fill(cs_bin_plan *p, int N, const int64_t *d_colptr, const int32_t *d_rowidx, const double *d_x,
                       int64_t *d_out_colptr, int32_t *d_out_rowidx, double *d_out_x, int64_t out_cap,
                       int64_t *nnz_out, cudaStream_t st, bool streamed)
{
    int rc;
    int64_t *h_nnz = reinterpret_cast<int64_t *>(p->h_pinned + NFLAGS);
    bool zeros_inline = streamed; // the streaming fill leaves zero sums in place, the generic one drops them
    if (p->h_pinned[FLAG_INEXACT]) { // non-integer / huge values: exact-order-free FP64 accumulators
        zeros_inline = false;
        CS_CUDA(cudaMemsetAsync(p->flags.p, 0, sizeof(int) * NFLAGS, st));
        rc = launch_fill<double>(p, N, d_colptr, d_rowidx, d_x, d_out_colptr, d_out_rowidx, d_out_x, out_cap, st);
        if (rc) return rc;
        CS_CUDA(cudaMemcpyAsync(p->h_pinned, p->flags.p, sizeof(int) * NFLAGS, cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
    }
    if (p->h_pinned[FLAG_SHRUNK]) { // some touched bins summed to zero: close the gaps
        const int64_t padded = *h_nnz;
        CS_CUDA(p->colptr2.reserve((size_t)N + 1));
        CS_CUDA(p->tmp_i.reserve((size_t)padded));
        CS_CUDA(p->tmp_x.reserve((size_t)padded));
        rc = run_scan(p, N, p->counts2.p, p->colptr2.p, st);
        if (rc) return rc;
        compact_columns_kernel<<<cs_num_sms() * 8, 128, 0, st>>>(N, d_out_colptr, zeros_inline ? nullptr : p->counts2.p,
                                                                 p->colptr2.p, d_out_rowidx, d_out_x, p->tmp_i.p,
                                                                 p->tmp_x.p);
        CS_CUDA(cudaGetLastError());
        CS_CUDA(cudaMemcpyAsync(h_nnz, p->grand.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
        *nnz_out = *h_nnz;
        CS_CUDA(cudaMemcpyAsync(d_out_colptr, p->colptr2.p, sizeof(int64_t) * ((size_t)N + 1), cudaMemcpyDeviceToDevice, st));
        CS_CUDA(cudaMemcpyAsync(d_out_rowidx, p->tmp_i.p, sizeof(int32_t) * (size_t)*h_nnz, cudaMemcpyDeviceToDevice, st));
        CS_CUDA(cudaMemcpyAsync(d_out_x, p->tmp_x.p, sizeof(double) * (size_t)*h_nnz, cudaMemcpyDeviceToDevice, st));
        CS_CUDA(cudaStreamSynchronize(st));
    }
    return CS_OK;
}

extern "C" int cs_bin_counts_dev(cs_bin_plan *p, int N, const int64_t *d_colptr,
                                 const int32_t *d_rowidx, const double *d_x,
                                 int64_t *d_out_colptr, int32_t *d_out_rowidx, double *d_out_x,
                                 int64_t out_cap, int64_t *nnz_out, void *stream)
{
    CS_REQUIRE(p && N >= 0 && d_colptr && d_out_colptr && nnz_out, "cs_bin_counts_dev: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        CS_CUDA(cudaMemsetAsync(d_out_colptr, 0, sizeof(int64_t), st));
        CS_CUDA(cudaStreamSynchronize(st));
        *nnz_out = 0;
        return CS_OK;
    }
    CS_CUDA(p->counts.reserve(N));
    CS_CUDA(p->counts2.reserve(N));
    if (d_out_rowidx) CS_REQUIRE(d_out_x != nullptr, "cs_bin_counts_dev: d_out_x is NULL");
    // one pass (count, look-back and fill in one kernel) when the map is regular and the slots are
    // aligned for 256-bit loads; CS_BIN_ONEPASS=0 keeps the count / scan / fill kernels
    const bool onepass_env = !(getenv("CS_BIN_ONEPASS") && atoi(getenv("CS_BIN_ONEPASS")) == 0);
    if (p->onepass && p->stream_fill && onepass_env && ((uintptr_t)d_rowidx % 32 == 0) && ((uintptr_t)d_x % 32 == 0)) {
        // one memset and one copy back per call: the flags, the total and the ticket live behind the
        // per-column status words, laid out like p->h_pinned (NFLAGS ints, then the int64 total)
        CS_CUDA(p->status.reserve((size_t)N + 4));
        CS_CUDA(cudaEventRecord(p->ev[0], st)); // (the timing of the call's device work starts with its memset)
        CS_CUDA(cudaMemsetAsync(p->status.p, 0, sizeof(unsigned long long) * ((size_t)N + 4), st));
        int *d_flags = reinterpret_cast<int *>(p->status.p + N);
        long long *d_total = reinterpret_cast<long long *>(p->status.p + N) + NFLAGS / 2;
        int *d_ticket = reinterpret_cast<int *>(d_total + 1);
        // (CS_BIN_OP_CTAS = 2 | 3 | 4 CTAs per SM: registers per thread against warps in flight)
        // (CS_BIN_OP_SMAP=0: the 256-thread CTAs that read the map through L1)
        int ctas = getenv("CS_BIN_OP_CTAS") ? atoi(getenv("CS_BIN_OP_CTAS")) : 3;
        const int sthreads = OP_SMAP_THREADS;
        const size_t smap_bytes = (size_t)(sthreads / 32) * OP_CAP * sizeof(int2) + sizeof(uint32_t) * (size_t)p->G;
        const bool smap = smap_bytes <= 227u * 1024u && !(getenv("CS_BIN_OP_SMAP") && atoi(getenv("CS_BIN_OP_SMAP")) == 0);
        int opt = getenv("CS_BIN_OP_OPT") ? (atoi(getenv("CS_BIN_OP_OPT")) & 31) : OP_DEFAULT_OPT;
        auto kern = smap ? op_pick<OP_SMAP_THREADS>(opt, std::make_integer_sequence<int, 32>{})
                         : (ctas == 2 ? bin_onepass_kernel<OP_THREADS, 2, false, 2> : bin_onepass_kernel<OP_THREADS, 3, false, 2>);
        const int threads = smap ? sthreads : OP_THREADS;
        const size_t smem = smap ? smap_bytes : (size_t)OP_WARPS * OP_CAP * sizeof(int2);
        if (p->op_kern != (const void *)kern || p->op_smem != smem) { // (the attribute and the occupancy query once per plan)
            CS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int q = 0;
            CS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, threads, smem));
            p->op_per_sm = q < 1 ? 1 : q;
            p->op_kern = (const void *)kern;
            p->op_smem = smem;
        }
        int grid = cs_num_sms() * p->op_per_sm;
        if (grid > (N + threads / 32 - 1) / (threads / 32)) grid = (N + threads / 32 - 1) / (threads / 32);
        p->timed = false;
        CS_CUDA(cudaEventRecord(p->ev[1], st));
        CS_CUDA(cudaEventRecord(p->ev[2], st));
        kern<<<grid, threads, smem, st>>>(N, p->G, d_colptr, d_rowidx, d_x, p->gmap.p, d_out_colptr, d_out_rowidx, d_out_x, out_cap,
                                          d_out_rowidx ? 1 : 0, p->status.p, d_ticket, p->counts2.p, d_flags, d_total);
        CS_CUDA(cudaGetLastError());
        CS_CUDA(cudaEventRecord(p->ev[3], st));
        int64_t *h_nnz1 = reinterpret_cast<int64_t *>(p->h_pinned + NFLAGS);
        CS_CUDA(cudaMemcpyAsync(p->h_pinned, d_flags, sizeof(int) * NFLAGS + sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
        *nnz_out = *h_nnz1;
        if (p->h_pinned[FLAG_IRREG]) { // a column with unsorted rows (or a gene too wide): the generic kernels from now on
            p->stream_fill = false;
            p->regular = false;
            return cs_bin_counts_dev(p, N, d_colptr, d_rowidx, d_x, d_out_colptr, d_out_rowidx, d_out_x, out_cap, nnz_out,
                                     stream);
        }
        if (!d_out_rowidx) return CS_OK;
        p->timed = true;
        if (p->h_pinned[FLAG_CAP]) {
            cs_set_error("cs_bin_counts_dev: output capacity %lld < %lld entries needed", (long long)out_cap,
                         (long long)*h_nnz1);
            return CS_ERR_UCAP;
        }
        return finish_fill(p, N, d_colptr, d_rowidx, d_x, d_out_colptr, d_out_rowidx, d_out_x, out_cap, nnz_out, st, true);
    }
    CS_CUDA(cudaMemsetAsync(p->flags.p, 0, sizeof(int) * NFLAGS, st));
    const int words = (p->B + 31) / 32;
    int grid1 = cs_num_sms() * 16;
    if (grid1 > N) grid1 = N;
    p->timed = false;
    CS_CUDA(cudaEventRecord(p->ev[0], st));
    if (p->regular) {
        int grid = cs_num_sms() * 4;
        if (grid > (N + STREAM_WARPS - 1) / STREAM_WARPS) grid = (N + STREAM_WARPS - 1) / STREAM_WARPS;
        bin_count_stream_kernel<<<grid, STREAM_THREADS, 0, st>>>(N, d_colptr, d_rowidx, p->gmap.p, p->counts.p);
    } else {
        bin_count_kernel<<<grid1, COUNT_THREADS, words * 4, st>>>(
            N, p->B, d_colptr, d_rowidx, p->gmap.p, p->map_ptr.p, p->map_bin.p, p->counts.p);
    }
    CS_CUDA(cudaGetLastError());
    CS_CUDA(cudaEventRecord(p->ev[1], st));
    int rc = run_scan(p, N, p->counts.p, d_out_colptr, st);
    if (rc) return rc;
    CS_CUDA(cudaEventRecord(p->ev[2], st));
    int64_t *h_nnz = reinterpret_cast<int64_t *>(p->h_pinned + NFLAGS);
    if (!d_out_rowidx) {
        CS_CUDA(cudaMemcpyAsync(h_nnz, p->grand.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        CS_CUDA(cudaStreamSynchronize(st));
        *nnz_out = *h_nnz;
        return CS_OK;
    }
    CS_REQUIRE(d_out_x != nullptr, "cs_bin_counts_dev: d_out_x is NULL");
    const bool streamed = p->regular && p->stream_fill;
    if (streamed) {
        const size_t smem = (size_t)SF_CAP * sizeof(int);
        CS_CUDA(cudaFuncSetAttribute(bin_fill_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bin_fill_stream_kernel, SF_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
        int grid = cs_num_sms() * per_sm;
        if (grid > N) grid = N;
        bin_fill_stream_kernel<<<grid, SF_THREADS, smem, st>>>(N, p->B, d_colptr, d_rowidx, d_x, p->gmap.p, d_out_colptr,
                                                              d_out_rowidx, d_out_x, out_cap, p->counts2.p, p->flags.p);
        CS_CUDA(cudaGetLastError());
    } else {
        rc = launch_fill<int>(p, N, d_colptr, d_rowidx, d_x, d_out_colptr, d_out_rowidx, d_out_x, out_cap, st);
        if (rc) return rc;
    }
    CS_CUDA(cudaEventRecord(p->ev[3], st));
    p->timed = true;
    CS_CUDA(cudaMemcpyAsync(h_nnz, p->grand.p, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaMemcpyAsync(p->h_pinned, p->flags.p, sizeof(int) * NFLAGS, cudaMemcpyDeviceToHost, st));
    CS_CUDA(cudaStreamSynchronize(st));
    *nnz_out = *h_nnz;
    if (p->h_pinned[FLAG_CAP]) {
        cs_set_error("cs_bin_counts_dev: output capacity %lld < %lld entries needed", (long long)out_cap,
                     (long long)*h_nnz);
        return CS_ERR_UCAP;
    }
    if (p->h_pinned[FLAG_IRREG]) {
        // row indices of some column do not ascend (the streaming count presumed they do, so the
        // offsets are void as well): the generic kernels from now on.  CS_ERR_UCAP from here means
        // the caller sized the output from a streaming count; *nnz_out holds what is needed.
        p->stream_fill = false;
        p->regular = false;
        return cs_bin_counts_dev(p, N, d_colptr, d_rowidx, d_x, d_out_colptr, d_out_rowidx, d_out_x, out_cap, nnz_out,
                                 stream);
    }
    return finish_fill(p, N, d_colptr, d_rowidx, d_x, d_out_colptr, d_out_rowidx, d_out_x, out_cap, nnz_out, st, streamed);
}

// ------------------------------------------------------- host (R-layout) form
// the position of some row index outside [0, G) (-1 as an unsigned maximum when there is none)
__global__ void validate_rows_kernel(int64_t nnz, const int32_t *__restrict__ rowidx, int G, long long *__restrict__ bad)
{
    long long mine = -1;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
        const int g = rowidx[e];
        if ((g < 0 || g >= G) && mine < 0) mine = e;
    }
    if (mine >= 0) atomicMin(reinterpret_cast<unsigned long long *>(bad), (unsigned long long)mine);
}

struct cs_bin_job {
    int N = 0, B = 0;
    int64_t nnz = 0;
    DevBuf<int64_t> out_colptr;
    DevBuf<int32_t> out_rowidx;
    DevBuf<double> out_x;
};

extern "C" int cs_bin_counts_begin(int G, int N, int B, const int32_t *colptr, const int32_t *rowidx,
                                   const double *x, const int32_t *map_ptr, const int32_t *map_bin,
                                   cs_bin_job **job_out, int64_t *nnz_out)
{
    CS_REQUIRE(job_out && nnz_out, "cs_bin_counts_begin: NULL output argument");
    *job_out = nullptr;
    CS_REQUIRE(G >= 0 && N >= 0 && B >= 1 && colptr && map_ptr, "cs_bin_counts_begin: bad dimensions");
    CS_REQUIRE(colptr[0] == 0, "cs_bin_counts_begin: colptr[0] must be 0");
    const int64_t nnz = colptr[N];
    for (int c = 0; c < N; c++) CS_REQUIRE(colptr[c] <= colptr[c + 1], "colptr not monotone at column %d", c);
    cs_bin_plan *plan = nullptr;
    int rc = cs_bin_plan_create(G, B, map_ptr, map_bin, &plan);
    if (rc) return rc;
    struct PlanGuard {
        cs_bin_plan *p;
        ~PlanGuard() { cs_bin_plan_destroy(p); }
    } guard{plan};
    std::vector<int64_t> cp64((size_t)N + 1);
    for (int c = 0; c <= N; c++) cp64[c] = colptr[c];
    DevBuf<int64_t> d_colptr;
    DevBuf<int32_t> d_rowidx;
    DevBuf<double> d_x;
    cs_bin_job *job = new cs_bin_job();
    struct JobGuard {
        cs_bin_job *j;
        ~JobGuard() { delete j; }
    } jguard{job};
    job->N = N;
    job->B = B;
    CS_CUDA(d_colptr.alloc((size_t)N + 1));
    CS_CUDA(d_rowidx.alloc((size_t)nnz));
    CS_CUDA(d_x.alloc((size_t)nnz));
    CS_CUDA(job->out_colptr.alloc((size_t)N + 1));
    CS_CUDA(cudaMemcpy(d_colptr.p, cp64.data(), sizeof(int64_t) * ((size_t)N + 1), cudaMemcpyHostToDevice));
    if (nnz) {
        if ((rc = cs_h2d_staged(d_rowidx.p, rowidx, sizeof(int32_t) * (size_t)nnz))) return rc;
        if ((rc = cs_h2d_staged(d_x.p, x, sizeof(double) * (size_t)nnz))) return rc;
        // row indices are validated on the device (a host loop over them costs as much as the copy)
        long long *d_bad = reinterpret_cast<long long *>(job->out_colptr.p); // (free until the count runs)
        CS_CUDA(cudaMemset(d_bad, 0xFF, sizeof(long long)));
        validate_rows_kernel<<<cs_num_sms() * 8, 256>>>(nnz, d_rowidx.p, G, d_bad);
        long long bad = -1;
        CS_CUDA(cudaMemcpy(&bad, d_bad, sizeof(long long), cudaMemcpyDeviceToHost));
        CS_REQUIRE(bad < 0, "rowidx[%lld]=%d outside [0,%d)", bad, rowidx[bad], G);
    }
    int64_t need = 0;
    rc = cs_bin_counts_dev(plan, N, d_colptr.p, d_rowidx.p, d_x.p, job->out_colptr.p, nullptr, nullptr, 0, &need, nullptr);
    if (rc) return rc;
    CS_REQUIRE(need <= 2147483647LL, "cs_bin_counts_begin: result has %lld nonzeros, more than a dgCMatrix holds", (long long)need);
    CS_CUDA(job->out_rowidx.alloc((size_t)need));
    CS_CUDA(job->out_x.alloc((size_t)need));
    rc = cs_bin_counts_dev(plan, N, d_colptr.p, d_rowidx.p, d_x.p, job->out_colptr.p, job->out_rowidx.p,
                           job->out_x.p, need, &job->nnz, nullptr);
    if (rc == CS_ERR_UCAP && job->nnz > need) { // (columns with unsorted rows: the first count was void)
        need = job->nnz;
        CS_REQUIRE(need <= 2147483647LL, "cs_bin_counts_begin: result has %lld nonzeros, more than a dgCMatrix holds", (long long)need);
        CS_CUDA(job->out_rowidx.alloc((size_t)need));
        CS_CUDA(job->out_x.alloc((size_t)need));
        rc = cs_bin_counts_dev(plan, N, d_colptr.p, d_rowidx.p, d_x.p, job->out_colptr.p, job->out_rowidx.p,
                               job->out_x.p, need, &job->nnz, nullptr);
    }
    if (rc) return rc;
    *nnz_out = job->nnz;
    *job_out = job;
    jguard.j = nullptr;
    return CS_OK;
}

extern "C" int cs_bin_counts_fetch(cs_bin_job *job, int32_t *out_colptr, int32_t *out_rowidx, double *out_x)
{
    CS_REQUIRE(job && out_colptr, "cs_bin_counts_fetch: NULL argument");
    struct JobGuard {
        cs_bin_job *j;
        ~JobGuard() { delete j; }
    } jguard{job};
    std::vector<int64_t> cp64((size_t)job->N + 1);
    CS_CUDA(cudaMemcpy(cp64.data(), job->out_colptr.p, sizeof(int64_t) * cp64.size(), cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < cp64.size(); c++) out_colptr[c] = (int32_t)cp64[c];
    if (job->nnz) {
        CS_REQUIRE(out_rowidx && out_x, "cs_bin_counts_fetch: NULL output slots");
        int rc;
        if ((rc = cs_d2h_staged(out_rowidx, job->out_rowidx.p, sizeof(int32_t) * (size_t)job->nnz))) return rc;
        if ((rc = cs_d2h_staged(out_x, job->out_x.p, sizeof(double) * (size_t)job->nnz))) return rc;
    }
    return CS_OK;
}

extern "C" void cs_bin_counts_free(cs_bin_job *job) { delete job; }

// bin x cluster pooling of a job's bin x cell counts where they lie on the device (the job stays
// valid): out is B x K column-major, see cs_rowsum_by_label
extern "C" int cs_bin_job_rowsum_by_label(cs_bin_job *job, const int *label, int K, double *out)
{
    CS_REQUIRE(job && label && out && K >= 1, "cs_bin_job_rowsum_by_label: bad arguments");
    DevBuf<int> d_l;
    DevBuf<double> d_o;
    CS_CUDA(d_l.alloc((size_t)job->N));
    CS_CUDA(d_o.alloc((size_t)job->B * K));
    if (job->N) CS_CUDA(cudaMemcpy(d_l.p, label, sizeof(int) * (size_t)job->N, cudaMemcpyHostToDevice));
    int rc = cs_rowsum_by_label_dev(job->B, job->N, job->out_colptr.p, job->out_rowidx.p, job->out_x.p, d_l.p, K,
                                    d_o.p, nullptr);
    if (rc) return rc;
    CS_CUDA(cudaMemcpy(out, d_o.p, sizeof(double) * (size_t)job->B * K, cudaMemcpyDeviceToHost));
    return CS_OK;
}
