// cs_common.cuh — shared device/host helpers for libclonalscope_b200 (sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/clonalscope_b200.h"

// ---------------------------------------------------------------- host side
void cs_set_error(const char *fmt, ...);

#define CS_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess) {                                                       \
            cs_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,          \
                         cudaGetErrorString(e_));                                      \
            return CS_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)

#define CS_REQUIRE(cond, ...)                                                          \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            cs_set_error(__VA_ARGS__);                                                 \
            return CS_ERR_ARG;                                                         \
        }                                                                              \
    } while (0)

// RAII device buffer: freed when the owning C++ scope unwinds (never under Rf_error).
template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    // grow-only
    cudaError_t reserve(size_t count) {
        if (count <= n) return cudaSuccess;
        return alloc(count);
    }
};

int cs_num_sms();

// pageable host memory <-> device through pinned slots and a few copy threads (cs_stage.cu);
// both return after the copy has completed
int cs_h2d_staged(void *d_dst, const void *h_src, size_t bytes);
int cs_d2h_staged(void *h_dst, const void *d_src, size_t bytes);

// -------------------------------------------------------------- device side
#define CS_LN_SQRT_2PI 0.918938533204672741780329736406
#define CS_FULL 0xffffffffu

__device__ __forceinline__ bool cs_isna(double x) { return x != x; }

// Philox4x32-10 (Salmon et al., SC'11).  Counter layout used by the sampler:
//   ctr = (cell | cluster id, segment, sweep, tag<<24 | attempt), key = (seed, 'CLON')
struct Philox4 {
    uint32_t v[4];
};
__device__ __forceinline__ Philox4 cs_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 r;
    r.v[0] = c0; r.v[1] = c1; r.v[2] = c2; r.v[3] = c3;
    return r;
}

enum { CS_PH_PROPOSAL = 1, CS_PH_ASSIGN = 2, CS_PH_RNORM = 3, CS_PH_ALLELE = 4 };

// 53-bit uniform strictly inside (0,1)
__device__ __forceinline__ double cs_u53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6) + 0.5) / 9007199254740992.0;
}

__device__ __forceinline__ void cs_philox_draw(uint32_t seed, uint32_t c0, uint32_t c1, uint32_t c2,
                                               uint32_t tag, uint32_t attempt, double &ua, double &ub)
{
    Philox4 o = cs_philox4x32(c0, c1, c2, (tag << 24) | (attempt & 0xFFFFFFu), seed, 0x434C4F4Eu);
    ua = cs_u53(o.v[0], o.v[1]);
    ub = cs_u53(o.v[2], o.v[3]);
}

// Wichura (1988) AS 241, PPND16: rational approximations on three ranges.
static __device__ double cs_qnorm(double p)
{
    if (p != p) return p;
    if (p <= 0.0) return -INFINITY;
    if (p >= 1.0) return INFINITY;
    const double q = p - 0.5;
    double r, num, den;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        num = (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                   45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                133.14166789178437745) * r + 3.387132872796366608);
        den = (((((((r * 5226.495278852854561 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                   21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
                42.313330701600911252) * r + 1.0);
        return q * num / den;
    }
    r = (q < 0.0) ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        num = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r +
                   1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
                4.6303378461565452959) * r + 1.42343711074968357734);
        den = (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r +
                   0.14810397642748007459) * r + 0.68976733498510000455) * r + 1.6763848301838038494) * r +
                2.05319162663775882187) * r + 1.0);
    } else {
        r -= 5.0;
        num = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r +
                   0.026532189526576123093) * r + 0.29656057182850489123) * r + 1.7848265399172913358) * r +
                5.4637849111641143699) * r + 6.6579046435011037772);
        den = (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
                   7.868691311456132591e-4) * r + 0.0148753612908506148525) * r + 0.13692988092273580531) * r +
                0.59983220655588793769) * r + 1.0);
    }
    const double val = num / den;
    return (q < 0.0) ? -val : val;
}

// butterfly all-reduce of a double over a warp; every lane ends with the same bits
__device__ __forceinline__ double cs_warp_sum_bfly(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = __dadd_rn(v, __shfl_xor_sync(CS_FULL, v, off));
    return v;
}
__device__ __forceinline__ double cs_warp_min(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = fmin(v, __shfl_xor_sync(CS_FULL, v, off));
    return v;
}
__device__ __forceinline__ int cs_warp_min_i(int v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = min(v, __shfl_xor_sync(CS_FULL, v, off));
    return v;
}
