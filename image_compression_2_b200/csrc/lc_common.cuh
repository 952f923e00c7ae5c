// Shared definitions for the latent-codec kernels (sm_100a).  Included by device code and by the
// host side of the C ABI (include/latentcodec.h carries the public copies of the enums).
#pragma once
#include <stdint.h>

#ifdef LC_HOSTSIM
#include "cuda_emul.h"
#define LC_HD
#else
#include <cuda_runtime.h>
#define LC_HD __host__ __device__
#endif

// ---- status words written per stream (mirror of the reference's exceptions)
enum {
    LC_OK = 0,
    LC_ENC_BIT_OVERFLOW = 1, // reference: ValueError from bytearray.append (defect D3, verbatim mode)
    LC_DEC_SYMBOL_OOB = 2,   // reference: IndexError at cum[symbol+1]   (cabac_compression.py:291)
    LC_DEC_ZERO_RANGE = 3,   // reference: ZeroDivisionError             (cabac_compression.py:285)
    LC_DEC_NEG_SYMBOL = 4,   // (not produced since ABI 2: symbol -1 is followed through NumPy's negative indexing like the reference)
    LC_OUT_OVERFLOW = 5,     // per-stream output slot too small
    LC_BAD_SYMBOL = 6,       // input index outside [0, n_symbols)
    LC_POOL_OVERFLOW = 7     // internal scratch exhausted (sizing bug; never expected)
};

enum { LC_MODE_VERBATIM = 0, LC_MODE_REPAIRED = 1 };

// Every fp64 operation on the parity path is an explicitly rounded intrinsic: no FMA contraction,
// IEEE division (the reference is NumPy/Python float64; evaluation order is part of the spec).
#define LC_DADD(a, b) __dadd_rn((a), (b))
#define LC_DSUB(a, b) __dsub_rn((a), (b))
#define LC_DMUL(a, b) __dmul_rn((a), (b))
#define LC_DDIV(a, b) __ddiv_rn((a), (b))
#define LC_D2LL(a) __double2ll_rz((a)) // Python int(): truncate toward zero
#define LC_LL2D(a) __ll2double_rn((a))

#define LC_FULL_MASK 0xffffffffu

// Exact int64 -> float64 for |x| < 2^51 without I2F.F64.S64 (which measured ~150 cycles of latency on
// B200): add x to the bit pattern of 2^52+2^51 and subtract that constant -- one integer add, one DADD.
#ifdef LC_HOSTSIM
static inline double lc_ll2d_small(long long x)
{
    long long b = 0x4338000000000000LL + x;
    double d;
    memcpy(&d, &b, 8);
    return d - 6755399441055744.0;
}
// ~20-bit reciprocal seed refined by two Newton steps: relative error below 2^-49; NOT correctly
// rounded, so only used where a guard band absorbs the error
static inline double lc_rcp_fast(double x)
{
    double r = (double)(float)(1.0 / x);
    r = std::fma(r, std::fma(-x, r, 1.0), r);
    r = std::fma(r, std::fma(-x, r, 1.0), r);
    return r;
}
#else
static __device__ __forceinline__ double lc_ll2d_small(long long x)
{
    return __dadd_rn(__longlong_as_double(0x4338000000000000LL + x), -6755399441055744.0);
}
static __device__ __forceinline__ double lc_rcp_fast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
    r = __fma_rn(r, __fma_rn(-x, r, 1.0), r);
    return r;
}
#endif

// pull the 128-byte line holding p into L1 ahead of a later load (no-op in the CPU emulator)
#ifdef LC_HOSTSIM
static inline void lc_prefetch_l1(const void *) {}
#else
static __device__ __forceinline__ void lc_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#endif

// Index arrays cross the C ABI as int32 (the reference's `.astype(np.int32)` codes, cabac_compression.py:471) or, to
// cut HBM traffic, as uint16 / uint8 (`idx_bytes` = 4, 2, 1 in include/latentcodec.h).  The kernels read and write them
// through these two views; the element size is warp-uniform, so the selects cost a predicated load or store.
struct LcCodes { // read-only
    const void *p;
    int eb;
    LC_HD LcCodes() : p(0), eb(4) {}
    LC_HD LcCodes(const int *q) : p(q), eb(4) {}
    LC_HD LcCodes(const void *q, int bytes) : p(q), eb(bytes) {}
    LC_HD int operator[](size_t i) const
    {
        if (eb == 4) return ((const int *)p)[i];
        if (eb == 1) return (int)((const unsigned char *)p)[i];
        return (int)((const unsigned short *)p)[i];
    }
    LC_HD LcCodes operator+(size_t off) const { return LcCodes((const char *)p + off * (size_t)eb, eb); }
};
struct LcIdxOut { // write-only (narrow types store the value truncated: decoded symbols are below n <= 2^(8*eb))
    void *p;
    int eb;
    LC_HD LcIdxOut() : p(0), eb(4) {}
    LC_HD LcIdxOut(int *q) : p(q), eb(4) {}
    LC_HD LcIdxOut(void *q, int bytes) : p(q), eb(bytes) {}
    LC_HD void store(size_t i, int v) const
    {
        if (eb == 4) ((int *)p)[i] = v;
        else if (eb == 1) ((unsigned char *)p)[i] = (unsigned char)v;
        else if (eb == 2) ((unsigned short *)p)[i] = (unsigned short)v; // eb == 0: the caller wants no index output
    }
    LC_HD LcIdxOut operator+(size_t off) const { return LcIdxOut((char *)p + off * (size_t)eb, eb); }
};

// Launch-time description of one coder launch.  All streams in a launch share it.
struct LcCoderCfg {
    int n;       // alphabet size, power of two in [2,1024]
    int R, C;    // rows per image, symbols per row
    int imgs;    // images per stream (they share coder + model, as the reference's batched call does)
    int has_ctx; // 1: (left,up) contexts of a 3-D shape; 0: single global context (non-3-D fallback)
    int mode;    // LC_MODE_*
    int total;   // imgs*R*C symbols per stream
    double rate; // ContextModel.adaptation_rate
    double delta; // guard band for the approximate cumulative sums, see lc_coder.cuh
    uint32_t slot_cap;   // hash slots per stream (power of two, >= 64)
    uint32_t slot_shift; // 32 - log2(slot_cap)
    uint32_t pool_bytes; // record pool per stream
    uint64_t scratch_stride; // bytes of scratch per resident warp (slots + pool)
    // shared-memory carve-up (bytes from the warp's base)
    uint32_t sm_dense, sm_lval, sm_lsym, sm_rows, sm_bytes;
    // NumPy pairwise-sum structure for this n
    int pw_len;    // block length min(n,128)
    int pw_steps;  // pw_len/8 terms per accumulator chain
    int pw_chains; // 8*(n/pw_len) chains; 0 when n<8 (plain sequential sum)
};

static inline LC_HD uint32_t lc_round_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

// Fills the derived fields of cfg from (n,R,C,imgs,has_ctx,mode,rate).  Returns 0 or a negative
// errno-style code for unsupported arguments.
static inline int lc_cfg_finalize(LcCoderCfg *c)
{
    if (c->n < 2 || c->n > 1024 || (c->n & (c->n - 1)) != 0) return -22;
    if (c->R < 1 || c->C < 1 || c->imgs < 1) return -22;
    int64_t total = (int64_t)c->imgs * c->R * c->C;
    if (total > (1 << 22)) return -22;
    if (c->has_ctx && c->C > 8192) return -22;
    c->total = (int)total;
    // guard band: |sequential float64 sum - approximate sum| <= (n+8)*2^-53*1.01 (all terms >= 0,
    // partial sums <= ~1); doubled.
    c->delta = (double)(c->n + 64) * 2.220446049250313e-16;
    int64_t max_ctx = c->has_ctx ? total : 1;
    int64_t all_ctx = (int64_t)(c->n + 1) * (c->n + 1);
    if (max_ctx > all_ctx) max_ctx = all_ctx;
    uint32_t cap = 64, lg = 6;
    while ((int64_t)cap < 2 * max_ctx) { cap <<= 1; lg++; }
    c->slot_cap = cap;
    c->slot_shift = 32 - lg;
    // record pool: <= 41 bytes per coded symbol under the doubling policy, rounded up
    c->pool_bytes = lc_round_up((uint32_t)(44 * total + 16384), 256);
    c->scratch_stride = (uint64_t)cap * 8 + c->pool_bytes;
    uint32_t off = 0;
    c->sm_dense = off; off += (uint32_t)c->n * 8;
    c->sm_lval = off;  off += (uint32_t)c->n * 8;
    c->sm_lsym = off;  off += lc_round_up((uint32_t)c->n * 2, 8);
    c->sm_rows = off;  off += c->has_ctx ? lc_round_up((uint32_t)c->C * 2 * 2, 8) : 0;
    c->sm_bytes = lc_round_up(off, 16);
    if (c->n < 8) { c->pw_len = c->n; c->pw_steps = 0; c->pw_chains = 0; }
    else {
        c->pw_len = c->n < 128 ? c->n : 128;
        c->pw_steps = c->pw_len / 8;
        c->pw_chains = 8 * (c->n / c->pw_len);
    }
    return 0;
}
