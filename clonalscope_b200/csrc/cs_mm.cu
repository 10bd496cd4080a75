// cs_mm.cu — MatrixMarket coordinate text -> dgCMatrix slots.
//
// Replaces Matrix::readMM + as(., "dgCMatrix") at the head of the reference's tutorials
// (samples/P5931/scRNA/README.md:72, R/Gene_bin_cell_rna_filtered.R:23): the 10x count matrix
// `matrix.mtx(.gz)`.  The text is parsed on the host — the body is cut at line ends into one piece
// per thread, each piece counted and then parsed in place with an integer fast path — and the
// triplets become compressed columns on the device: range check, order check, column pointer by
// binary search when the file is already column-major with ascending rows (what cellranger
// writes); otherwise a radix sort of (column, row) keys (cub) and a segmented sum of duplicates.
#include "cs_common.cuh"

#include <algorithm>
#include <atomic>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

struct cs_mm_job {
    int64_t rows = 0, cols = 0, nnz = 0;
    DevBuf<int64_t> colptr;
    DevBuf<int32_t> rowidx;
    DevBuf<double> x;
};

namespace {

// flags[0]: an index out of range (1-based entry number), flags[1]: not column-major / rows not ascending
__global__ void mm_check_kernel(int64_t n, const int32_t *__restrict__ r, const int32_t *__restrict__ c, int64_t rows,
                                int64_t cols, unsigned long long *__restrict__ flags)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int ri = r[e], ci = c[e];
        if (ri < 1 || ri > rows || ci < 1 || ci > cols) atomicMax(&flags[0], (unsigned long long)(e + 1));
        if (e > 0) {
            const int rp = r[e - 1], cp = c[e - 1];
            if (cp > ci || (cp == ci && rp >= ri)) flags[1] = 1ull;
        }
    }
}

// colptr[j] = number of entries with (1-based) column <= j, by binary search in the sorted column list
__global__ void mm_colptr_kernel(int64_t n, const int32_t *__restrict__ c, int64_t cols, int64_t *__restrict__ colptr)
{
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= cols; j += (int64_t)gridDim.x * blockDim.x) {
        int64_t lo = 0, hi = n;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (c[mid] <= j) lo = mid + 1;
            else hi = mid;
        }
        colptr[j] = lo;
    }
}

__global__ void mm_shift_rows_kernel(int64_t n, const int32_t *__restrict__ r, int32_t *__restrict__ out)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) out[e] = r[e] - 1;
}

__global__ void mm_keys_kernel(int64_t n, const int32_t *__restrict__ r, const int32_t *__restrict__ c,
                               unsigned long long *__restrict__ key, unsigned *__restrict__ idx)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        key[e] = ((unsigned long long)(unsigned)c[e] << 32) | (unsigned)r[e];
        idx[e] = (unsigned)e;
    }
}

// head[e] = 1 where a new (column, row) pair starts in the sorted keys
__global__ void mm_heads_kernel(int64_t n, const unsigned long long *__restrict__ key, int64_t *__restrict__ head)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x)
        head[e] = (e == 0 || key[e] != key[e - 1]) ? 1 : 0;
}

// pos = inclusive scan of the heads - 1: the slot of every sorted entry; duplicates add up (counts: exact)
__global__ void mm_merge_kernel(int64_t n, const unsigned long long *__restrict__ key, const unsigned *__restrict__ idx,
                                const int64_t *__restrict__ pos, const double *__restrict__ v, int32_t *__restrict__ rowidx,
                                int32_t *__restrict__ col, double *__restrict__ x)
{
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = pos[e] - 1;
        atomicAdd(x + p, v[idx[e]]);
        if (e == 0 || key[e] != key[e - 1]) {
            rowidx[p] = (int32_t)(key[e] & 0xFFFFFFFFull) - 1;
            col[p] = (int32_t)(key[e] >> 32);
        }
    }
}

bool read_file(const char *path, std::string &buf, std::string &err)
{
    const size_t L = strlen(path);
    const bool gz = L > 3 && strcmp(path + L - 3, ".gz") == 0;
    if (gz) {
        gzFile f = gzopen(path, "rb");
        if (!f) {
            err = std::string("cannot open ") + path;
            return false;
        }
        gzbuffer(f, 1 << 20);
        std::vector<char> chunk(8u << 20);
        for (;;) {
            const int got = gzread(f, chunk.data(), (unsigned)chunk.size());
            if (got < 0) {
                err = std::string("gzip error in ") + path;
                gzclose(f);
                return false;
            }
            if (got == 0) break;
            buf.append(chunk.data(), (size_t)got);
        }
        gzclose(f);
        return true;
    }
    FILE *f = fopen(path, "rb");
    if (!f) {
        err = std::string("cannot open ") + path;
        return false;
    }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)sz);
    const size_t got = sz ? fread(&buf[0], 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) {
        err = std::string("short read on ") + path;
        return false;
    }
    return true;
}

inline const char *skip_ws(const char *p, const char *e)
{
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) p++;
    return p;
}

// one entry line "i j [v]"; returns false on a malformed line
inline bool parse_line(const char *p, const char *e, bool pattern, int32_t &i, int32_t &j, double &v)
{
    auto integer = [&](int64_t &out) {
        p = skip_ws(p, e);
        if (p >= e || *p < '0' || *p > '9') return false;
        int64_t a = 0;
        while (p < e && *p >= '0' && *p <= '9') a = a * 10 + (*p++ - '0');
        out = a;
        return true;
    };
    int64_t a, b;
    if (!integer(a) || !integer(b) || a > 2147483647LL || b > 2147483647LL) return false;
    i = (int32_t)a;
    j = (int32_t)b;
    if (pattern) {
        v = 1.0;
        return true;
    }
    p = skip_ws(p, e);
    if (p >= e) return false;
    // integer fast path (counts); anything else goes through strtod
    const char *q = p;
    bool neg = false;
    if (*q == '-' || *q == '+') neg = *q++ == '-';
    int64_t m = 0;
    int nd = 0;
    while (q < e && *q >= '0' && *q <= '9' && nd < 18) m = m * 10 + (*q++ - '0'), nd++;
    const char *r = skip_ws(q, e);
    if (nd > 0 && r == e) {
        v = neg ? -(double)m : (double)m;
        return true;
    }
    char tmp[64];
    const size_t len = std::min<size_t>((size_t)(e - p), sizeof(tmp) - 1);
    memcpy(tmp, p, len);
    tmp[len] = 0;
    char *end = nullptr;
    v = strtod(tmp, &end);
    return end != tmp;
}

} // namespace

extern "C" void cs_mm_read_free(cs_mm_job *job) { delete job; }

extern "C" int cs_mm_read_begin(const char *path, cs_mm_job **job_out, int64_t *dims)
{
    CS_REQUIRE(path && job_out && dims, "cs_mm_read_begin: NULL argument");
    *job_out = nullptr;
    std::string buf, err;
    if (!read_file(path, buf, err)) {
        cs_set_error("%s", err.c_str());
        return CS_ERR_ARG;
    }
    const char *p = buf.data(), *end = p + buf.size();
    auto line_end = [&](const char *s) {
        const char *nl = (const char *)memchr(s, '\n', (size_t)(end - s));
        return nl ? nl : end;
    };
    // banner
    const char *le = line_end(p);
    std::string banner(p, le);
    for (auto &ch : banner) ch = (char)tolower(ch);
    CS_REQUIRE(banner.rfind("%%matrixmarket", 0) == 0, "%s: not a MatrixMarket file", path);
    CS_REQUIRE(banner.find("coordinate") != std::string::npos, "%s: only the coordinate format is read", path);
    CS_REQUIRE(banner.find("complex") == std::string::npos, "%s: complex matrices are not read", path);
    const bool pattern = banner.find("pattern") != std::string::npos;
    const bool symmetric = banner.find("symmetric") != std::string::npos;
    CS_REQUIRE(!symmetric && banner.find("skew") == std::string::npos && banner.find("hermitian") == std::string::npos,
               "%s: only general matrices are read", path);
    p = le < end ? le + 1 : end;
    while (p < end && (*p == '%' || *p == '\n' || *p == '\r')) p = std::min(end, line_end(p) + 1);
    long long rows = 0, cols = 0, nnz_decl = 0;
    {
        le = line_end(p);
        std::string sz(p, le);
        CS_REQUIRE(sscanf(sz.c_str(), "%lld %lld %lld", &rows, &cols, &nnz_decl) == 3, "%s: bad size line", path);
        p = le < end ? le + 1 : end;
    }
    CS_REQUIRE(rows >= 0 && cols >= 0 && nnz_decl >= 0 && rows <= 2147483647LL && cols <= 2147483646LL, "%s: bad dimensions", path);
    // cut the body into pieces at line ends, count, then parse
    unsigned hw = std::thread::hardware_concurrency();
    int T = (int)std::max(1u, std::min(16u, hw ? hw : 4u));
    if ((size_t)(end - p) < (1u << 20)) T = 1;
    std::vector<const char *> cut(T + 1);
    cut[0] = p;
    cut[T] = end;
    for (int t = 1; t < T; t++) {
        const char *s = p + (size_t)(end - p) * t / T;
        s = line_end(s);
        cut[t] = s < end ? s + 1 : end;
    }
    for (int t = 1; t <= T; t++) cut[t] = std::max(cut[t], cut[t - 1]);
    std::vector<int64_t> cnt(T, 0);
    auto for_lines = [&](int t, auto &&fn) {
        const char *s = cut[t], *e = cut[t + 1];
        while (s < e) {
            const char *nl = (const char *)memchr(s, '\n', (size_t)(e - s));
            const char *l1 = nl ? nl : e;
            const char *q = skip_ws(s, l1);
            if (q < l1 && *q != '%') fn(q, l1);
            s = nl ? nl + 1 : e;
        }
    };
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] { for_lines(t, [&](const char *, const char *) { cnt[t]++; }); });
        for (auto &x : th) x.join();
    }
    std::vector<int64_t> off(T + 1, 0);
    for (int t = 0; t < T; t++) off[t + 1] = off[t] + cnt[t];
    const int64_t n = off[T];
    CS_REQUIRE(n == nnz_decl, "%s: the size line announces %lld entries, the file holds %lld", path, nnz_decl, (long long)n);
    std::vector<int32_t> hi((size_t)n), hj((size_t)n);
    std::vector<double> hv((size_t)n);
    std::atomic<long long> bad{0};
    {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++)
            th.emplace_back([&, t] {
                int64_t k = off[t];
                for_lines(t, [&](const char *a, const char *b) {
                    if (!parse_line(a, b, pattern, hi[(size_t)k], hj[(size_t)k], hv[(size_t)k])) bad.store(k + 1);
                    k++;
                });
            });
        for (auto &x : th) x.join();
    }
    CS_REQUIRE(bad.load() == 0, "%s: malformed entry line %lld", path, bad.load());
    buf.clear();
    buf.shrink_to_fit();
    // ---- device: triplets -> compressed columns
    cs_mm_job *job = new cs_mm_job();
    struct Guard {
        cs_mm_job *p;
        ~Guard() { delete p; }
    } guard{job};
    job->rows = rows;
    job->cols = cols;
    DevBuf<int32_t> d_i, d_j;
    DevBuf<double> d_v;
    DevBuf<unsigned long long> d_flags;
    CS_CUDA(d_i.alloc((size_t)n));
    CS_CUDA(d_j.alloc((size_t)n));
    CS_CUDA(d_v.alloc((size_t)n));
    CS_CUDA(d_flags.alloc(2));
    CS_CUDA(cudaMemset(d_flags.p, 0, 2 * sizeof(unsigned long long)));
    CS_CUDA(job->colptr.alloc((size_t)cols + 1));
    int rc;
    const int grid = cs_num_sms() * 8;
    unsigned long long flags[2] = {0, 0};
    if (n) {
        if ((rc = cs_h2d_staged(d_i.p, hi.data(), sizeof(int32_t) * (size_t)n))) return rc;
        if ((rc = cs_h2d_staged(d_j.p, hj.data(), sizeof(int32_t) * (size_t)n))) return rc;
        if ((rc = cs_h2d_staged(d_v.p, hv.data(), sizeof(double) * (size_t)n))) return rc;
        mm_check_kernel<<<grid, 256>>>(n, d_i.p, d_j.p, rows, cols, d_flags.p);
        CS_CUDA(cudaMemcpy(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost));
        CS_REQUIRE(flags[0] == 0, "%s: entry %llu (%d, %d) is outside the %lld x %lld matrix", path, flags[0],
                   hi[(size_t)flags[0] - 1], hj[(size_t)flags[0] - 1], rows, cols);
    }
    if (flags[1] == 0) { // column-major, rows ascending, no duplicates: the triplets are the slots
        CS_CUDA(job->rowidx.alloc((size_t)n));
        mm_shift_rows_kernel<<<grid, 256>>>(n, d_i.p, job->rowidx.p);
        mm_colptr_kernel<<<grid, 256>>>(n, d_j.p, cols, job->colptr.p);
        CS_CUDA(cudaGetLastError());
        CS_CUDA(cudaDeviceSynchronize());
        std::swap(job->x.p, d_v.p), std::swap(job->x.n, d_v.n);
        job->nnz = n;
    } else {
        DevBuf<unsigned long long> d_k, d_k2;
        DevBuf<unsigned> d_id, d_id2;
        DevBuf<int64_t> d_pos;
        DevBuf<int32_t> d_c2;
        CS_CUDA(d_k.alloc((size_t)n));
        CS_CUDA(d_k2.alloc((size_t)n));
        CS_CUDA(d_id.alloc((size_t)n));
        CS_CUDA(d_id2.alloc((size_t)n));
        CS_CUDA(d_pos.alloc((size_t)n));
        mm_keys_kernel<<<grid, 256>>>(n, d_i.p, d_j.p, d_k.p, d_id.p);
        size_t tmp_bytes = 0;
        CS_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_k.p, d_k2.p, d_id.p, d_id2.p, n));
        DevBuf<unsigned char> d_tmp;
        CS_CUDA(d_tmp.alloc(tmp_bytes));
        CS_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_k.p, d_k2.p, d_id.p, d_id2.p, n));
        mm_heads_kernel<<<grid, 256>>>(n, d_k2.p, d_pos.p);
        size_t scan_bytes = 0;
        CS_CUDA(cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, d_pos.p, d_pos.p, n));
        if (scan_bytes > tmp_bytes) CS_CUDA(d_tmp.alloc(scan_bytes));
        CS_CUDA(cub::DeviceScan::InclusiveSum(d_tmp.p, scan_bytes, d_pos.p, d_pos.p, n));
        int64_t m = 0;
        CS_CUDA(cudaMemcpy(&m, d_pos.p + (n - 1), sizeof(int64_t), cudaMemcpyDeviceToHost));
        CS_CUDA(job->rowidx.alloc((size_t)m));
        CS_CUDA(job->x.alloc((size_t)m));
        CS_CUDA(d_c2.alloc((size_t)m));
        CS_CUDA(cudaMemset(job->x.p, 0, sizeof(double) * (size_t)m));
        mm_merge_kernel<<<grid, 256>>>(n, d_k2.p, d_id2.p, d_pos.p, d_v.p, job->rowidx.p, d_c2.p, job->x.p);
        mm_colptr_kernel<<<grid, 256>>>(m, d_c2.p, cols, job->colptr.p);
        CS_CUDA(cudaGetLastError());
        CS_CUDA(cudaDeviceSynchronize());
        job->nnz = m;
    }
    CS_REQUIRE(job->nnz <= 2147483647LL, "%s: %lld nonzeros, more than a dgCMatrix holds", path, (long long)job->nnz);
    dims[0] = rows;
    dims[1] = cols;
    dims[2] = job->nnz;
    *job_out = job;
    guard.p = nullptr;
    return CS_OK;
}

extern "C" int cs_mm_read_fetch(cs_mm_job *job, int32_t *colptr, int32_t *rowidx, double *x)
{
    CS_REQUIRE(job && colptr, "cs_mm_read_fetch: NULL argument");
    struct Guard {
        cs_mm_job *p;
        ~Guard() { delete p; }
    } guard{job};
    std::vector<int64_t> cp((size_t)job->cols + 1);
    CS_CUDA(cudaMemcpy(cp.data(), job->colptr.p, sizeof(int64_t) * cp.size(), cudaMemcpyDeviceToHost));
    for (size_t c = 0; c < cp.size(); c++) colptr[c] = (int32_t)cp[c];
    if (job->nnz) {
        CS_REQUIRE(rowidx && x, "cs_mm_read_fetch: NULL output slots");
        int rc;
        if ((rc = cs_d2h_staged(rowidx, job->rowidx.p, sizeof(int32_t) * (size_t)job->nnz))) return rc;
        if ((rc = cs_d2h_staged(x, job->x.p, sizeof(double) * (size_t)job->nnz))) return rc;
    }
    return CS_OK;
}
