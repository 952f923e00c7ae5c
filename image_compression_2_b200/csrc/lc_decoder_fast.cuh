// Fast decoder: cabac_decode (cabac_compression.py:363-406) for the repaired coder mode and
// (left,up) contexts, one warp per stream.  Same arithmetic as lc_coder.cuh, restructured around
// what the profile of the first version showed (profiles/r01_*):
//
//  * Lazy first/second visits.  48 % of symbols open a fresh context and 23 % a context seen once
//    (config 1).  A context seen once or twice keeps only its symbols in the 64-bit table slot
//    (states A and B); its model is rebuilt on demand: after ONE update the model is
//    (u = u0*f[step(s1)], {s1: P1}) with f taken from a per-launch table, so contexts that are
//    never revisited cost no model arithmetic at all (45 % of the reference's updates are dead).
//  * From the third visit on (state C) the sparse record (u, sorted [(sym,val)]) lives in the
//    pool; it is held in REGISTERS while in use (lane j = entry j, at most 32 entries; a stream
//    that needs more is handed to the generic kernel), searched with warp scans, updated eagerly
//    after the symbol is known so that the update overlaps the next table probe.
//  * Software-pipelined probe: the table window of the NEXT context is requested as soon as the
//    symbol is known, and patched in registers if the current context's slot lands inside it.
//  * Closed-form renormalisation: in repaired mode low/high/code stay below 2^32, so the
//    bit-at-a-time loops (:295-309) are `clz` counts and shifts; bits are read many at a time.
//  * The exact sequential sums needed when a guard band is hit (range*cum is an exact integer in
//    real arithmetic for ~1/3 of second visits) are fixed-count chains over register values.
#pragma once
#include "lc_coder.cuh"

#define LC_NEEDS_GENERIC 100 // internal status: more than 32 distinct symbols in one context

// Optional per-state cycle accounting (build with -DLC_DEC_PROFILE; tools/dec_profile.py).  Row = context
// state 0..3 (+4: symbols that needed the exact search, +5: symbols that needed exact_at), columns =
// count, then cycles spent in probe / model load / search / interval / renorm+output / write-back.
#ifdef LC_DEC_PROFILE
__device__ unsigned long long lc_prof_global[8 * 8];
#define LCP_DECL __shared__ unsigned long long lcp_sm[8 * 8]; long long lcp_t = 0; int lcp_row = 0;
#define LCP_INIT() do { for (int i_ = F.lane; i_ < 64; i_ += 32) lcp_sm[i_] = 0ull; __syncwarp(); } while (0)
#define LCP_START() do { lcp_t = clock64(); } while (0)
#define LCP_ROW(r_) do { lcp_row = (r_); } while (0)
#define LCP_MARK(col_) do { const long long n_ = clock64(); if (F.lane == 0) lcp_sm[lcp_row * 8 + (col_)] += (unsigned long long)(n_ - lcp_t); lcp_t = n_; } while (0)
#define LCP_COUNT(r_, col_) do { if (F.lane == 0) lcp_sm[(r_) * 8 + (col_)] += 1ull; } while (0)
#define LCP_FLUSH() do { __syncwarp(); for (int i_ = F.lane; i_ < 64; i_ += 32) atomicAdd(&lc_prof_global[i_], lcp_sm[i_]); } while (0)
#else
#define LCP_DECL
#define LCP_INIT()
#define LCP_START()
#define LCP_ROW(r_)
#define LCP_MARK(col_)
#define LCP_COUNT(r_, col_)
#define LCP_FLUSH()
#endif

#define LCF_STATE(w) ((int)(((w) >> 22) & 3ull))
#define LCF_S1(w) ((int)(((w) >> 24) & 0x3FFull))
#define LCF_S2(w) ((int)(((w) >> 34) & 0x3FFull))
#define LCF_K(w) ((int)(((w) >> 24) & 0x3Full))
#define LCF_OFF(w) ((uint32_t)((w) >> 30))
#define LCF_PACK_A(key1, s1) ((unsigned long long)(key1) | (1ull << 22) | ((unsigned long long)(s1) << 24))
#define LCF_PACK_B(key1, s1, s2) \
    ((unsigned long long)(key1) | (2ull << 22) | ((unsigned long long)(s1) << 24) | ((unsigned long long)(s2) << 34))
#define LCF_PACK_C(key1, k, off) \
    ((unsigned long long)(key1) | (3ull << 22) | ((unsigned long long)(k) << 24) | ((unsigned long long)(off) << 30))

struct LcFast {
    int n, C, R, total, lane;
    double rate, delta, delta_v, tmargin, u0, P1;
    uint32_t slot_cap, slot_shift, pool_bytes, pool_top;
    int pw_len, pw_steps, pw_chains;
    unsigned long long *slots;
    char *pool;
    double *dense;        // smem: n doubles (pairwise image)
    double *u1tab;        // smem: u after the first update, by chain step (or by symbol when n < 8)
    unsigned short *rows; // smem: previous/current row of decoded symbols
    // model of the open context, in registers: lane j < k holds entry j (ascending symbol)
    int k, my_sym;
    double u, my_val;
};

// Shared-memory accesses of the decoder warp's hot loop go through 32-bit shared-window addresses held in
// registers (inline PTX), not through generic pointers: on sm_100a every access through a generic pointer to
// dynamic shared memory re-derives the window base (S2UR CgaCtaId / ULEA / LDCU, three to four instructions per
// access, measured 15 % of the decoder warp's instructions) and each pointer costs two registers.
#ifdef LC_HOSTSIM
typedef uintptr_t lcv_sa; // on the emulator a "shared address" is the host pointer
static inline lcv_sa lcv_sa_of(const void *p) { return (lcv_sa)p; }
static inline uint32_t lcv_sa_ld32(lcv_sa a) { return *(const volatile uint32_t *)a; }
static inline void lcv_sa_st32(lcv_sa a, uint32_t v) { *(volatile uint32_t *)a = v; }
static inline uint32_t lcv_sa_ld32_acq(lcv_sa a) { return *(const volatile uint32_t *)a; }
static inline int lcv_sa_ld8(lcv_sa a) { return (int)*(const volatile unsigned char *)a; }
static inline void lcv_sa_st8(lcv_sa a, int v) { *(volatile unsigned char *)a = (unsigned char)v; }
static inline int lcv_sa_ld16(lcv_sa a) { return (int)*(const volatile unsigned short *)a; }
static inline void lcv_sa_st16(lcv_sa a, int v) { *(volatile unsigned short *)a = (unsigned short)v; }
static inline double lcv_sa_ldf64(lcv_sa a) { return *(const volatile double *)a; }
static inline void lcv_sa_or32(lcv_sa a, uint32_t v) { atomicOr((uint32_t *)a, v); }
static inline void lcv_sa_bar_arrive(lcv_sa b) { *(volatile unsigned long long *)b += 1ull; }
// predicated forms (p != 0: do it): straight-line code on the GPU, no branch region around a one-lane store
static inline void lcv_sa_st32_if(uint32_t p, lcv_sa a, uint32_t v) { if (p) lcv_sa_st32(a, v); }
static inline void lcv_sa_st8_if(uint32_t p, lcv_sa a, int v) { if (p) lcv_sa_st8(a, v); }
static inline void lcv_sa_or32_if(uint32_t p, lcv_sa a, uint32_t v) { if (p) lcv_sa_or32(a, v); }
static inline void lcv_sa_bar_arrive_if(uint32_t p, lcv_sa b) { if (p) lcv_sa_bar_arrive(b); }
static inline void lcv_stcg32_if(uint32_t p, uint32_t *g, uint32_t v) { if (p) *g = v; }
#else
typedef uint32_t lcv_sa;
static __device__ __forceinline__ lcv_sa lcv_sa_of(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ uint32_t lcv_sa_ld32(lcv_sa a)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
static __device__ __forceinline__ void lcv_sa_st32(lcv_sa a, uint32_t v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v));
}
static __device__ __forceinline__ uint32_t lcv_sa_ld32_acq(lcv_sa a)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
static __device__ __forceinline__ int lcv_sa_ld8(lcv_sa a)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return (int)v;
}
static __device__ __forceinline__ void lcv_sa_st8(lcv_sa a, int v)
{
    asm volatile("st.volatile.shared.u8 [%0], %1;" ::"r"(a), "r"(v));
}
static __device__ __forceinline__ int lcv_sa_ld16(lcv_sa a)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return (int)v;
}
static __device__ __forceinline__ void lcv_sa_st16(lcv_sa a, int v)
{
    asm volatile("st.volatile.shared.u16 [%0], %1;" ::"r"(a), "r"(v));
}
static __device__ __forceinline__ double lcv_sa_ldf64(lcv_sa a) // per-launch constants: plain load
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
static __device__ __forceinline__ void lcv_sa_or32(lcv_sa a, uint32_t v)
{
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
static __device__ __forceinline__ void lcv_sa_bar_arrive(lcv_sa b) // release.cta
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory");
}
// predicated forms (p != 0: do it): straight-line code, no branch region (BSSY/BSYNC) around a one-lane store -- every
// such region cost the decoder warp 25-40 cycles per symbol (profiles/r02_decoder_regions.md)
static __device__ __forceinline__ void lcv_sa_st32_if(uint32_t p, lcv_sa a, uint32_t v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.volatile.shared.u32 [%1], %2; }" ::"r"(p), "r"(a), "r"(v));
}
static __device__ __forceinline__ void lcv_sa_st8_if(uint32_t p, lcv_sa a, int v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.volatile.shared.u8 [%1], %2; }" ::"r"(p), "r"(a), "r"(v));
}
static __device__ __forceinline__ void lcv_sa_or32_if(uint32_t p, lcv_sa a, uint32_t v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q red.shared.or.b32 [%1], %2; }" ::"r"(p), "r"(a), "r"(v) : "memory");
}
static __device__ __forceinline__ void lcv_sa_bar_arrive_if(uint32_t p, lcv_sa b) // release.cta
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q mbarrier.arrive.shared::cta.b64 _, [%1]; }" ::"r"(p), "r"(b) : "memory");
}
static __device__ __forceinline__ void lcv_stcg32_if(uint32_t p, uint32_t *g, uint32_t v)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q st.global.cg.u32 [%1], %2; }" ::"r"(p), "l"(g), "r"(v) : "memory");
}
#endif

__device__ __forceinline__ int lcf_tab_index(const LcFast &F, int s)
{
    return F.pw_chains == 0 ? s : ((s & (F.pw_len - 1)) >> 3);
}

// NumPy pairwise sum of F.dense (same routine as lc_pairwise_total, on the LcFast view)
__device__ __forceinline__ double lcf_pairwise_total(const LcFast &F)
{
    if (F.pw_chains == 0) {
        double r = 0.0;
        for (int i = 0; i < F.n; i++) r = LC_DADD(r, F.dense[i]);
        return r;
    }
    const int chains = F.pw_chains;
    const int c = F.lane & (chains - 1) & 31;
    const double *p = F.dense + (c >> 3) * F.pw_len + (c & 7);
    double r = lc_chain(p, F.pw_steps), r2 = 0.0;
    if (chains == 64) r2 = lc_chain(p + 4 * 128, F.pw_steps);
    for (int off = 1; off < chains && off < 32; off <<= 1) {
        r = LC_DADD(r, __shfl_xor_sync(LC_FULL_MASK, r, off));
        if (chains == 64) r2 = LC_DADD(r2, __shfl_xor_sync(LC_FULL_MASK, r2, off));
    }
    if (chains == 64) r = LC_DADD(r, r2);
    return r;
}

// P1 and the table of u after the first update (ContextModel.update_model on the uniform vector)
__device__ __forceinline__ void lcf_tables_init(LcFast &F)
{
    F.P1 = LC_DADD(F.u0, LC_DMUL(F.rate, LC_DSUB(1.0, F.u0)));
    const int entries = F.pw_chains == 0 ? F.n : F.pw_steps;
    for (int t = 0; t < entries; t++) {
        const int s = F.pw_chains == 0 ? t : 8 * t;
        for (int i = F.lane; i < F.n; i += 32) F.dense[i] = (i == s) ? F.P1 : F.u0;
        __syncwarp();
        const double total = lcf_pairwise_total(F);
        const double others = LC_DSUB(total, F.P1);
        const double f = (others > 0.0) ? LC_DDIV(LC_DSUB(1.0, F.P1), others) : 0.0;
        __syncwarp();
        if (F.lane == 0) F.u1tab[t] = LC_DMUL(F.u0, f);
    }
    __syncwarp();
}

// model after exactly one update with symbol s1
__device__ __forceinline__ void lcf_state_first(LcFast &F, int s1)
{
    F.k = 1;
    F.u = F.u1tab[lcf_tab_index(F, s1)];
    F.my_sym = (F.lane == 0) ? s1 : 0x7fffffff;
    F.my_val = (F.lane == 0) ? F.P1 : 0.0;
}

// ContextModel.update_model (:119-144) on the register-resident sparse model.
// Returns false when the context would need more than 32 entries.
__device__ __forceinline__ bool lcf_update(LcFast &F, int s)
{
    const bool valid = F.lane < F.k;
    const unsigned meq = __ballot_sync(LC_FULL_MASK, valid && F.my_sym == s);
    const unsigned mlt = __ballot_sync(LC_FULL_MASK, valid && F.my_sym < s);
    const int js = meq ? __ffs((int)meq) - 1 : -1;
    const int ins = __popc(mlt);
    if (js < 0 && F.k >= 32) return false;
    const double pv = __shfl_sync(LC_FULL_MASK, F.my_val, js < 0 ? 0 : js);
    const double p_old = js < 0 ? F.u : pv;
    const double p_new = LC_DADD(p_old, LC_DMUL(F.rate, LC_DSUB(1.0, p_old)));
    // dense image with p_new in place
    for (int i = F.lane; i < F.n; i += 32) F.dense[i] = F.u;
    __syncwarp();
    if (valid) F.dense[F.my_sym] = F.my_val;
    __syncwarp();
    if (F.lane == 0) F.dense[s] = p_new;
    __syncwarp();
    const double total = lcf_pairwise_total(F);
    __syncwarp();
    const double others = LC_DSUB(total, p_new);
    const double f = (others > 0.0) ? LC_DDIV(LC_DSUB(1.0, p_new), others) : 0.0;
    F.u = LC_DMUL(F.u, f);
    if (valid) F.my_val = (F.lane == js) ? p_new : LC_DMUL(F.my_val, f);
    if (js < 0) { // insert (s, p_new) at sorted position ins
        const int up_sym = __shfl_up_sync(LC_FULL_MASK, F.my_sym, 1);
        const double up_val = __shfl_up_sync(LC_FULL_MASK, F.my_val, 1);
        if (F.lane > ins && F.lane <= F.k) { F.my_sym = up_sym; F.my_val = up_val; }
        if (F.lane == ins) { F.my_sym = s; F.my_val = p_new; }
        F.k++;
    }
    return true;
}

// T after `count` sequential additions of u (np.cumsum over a run of never-observed symbols)
__device__ __forceinline__ double lcf_add_run(double T, double u, int count)
{
    for (; count >= 8; count -= 8) {
        T = LC_DADD(T, u); T = LC_DADD(T, u); T = LC_DADD(T, u); T = LC_DADD(T, u);
        T = LC_DADD(T, u); T = LC_DADD(T, u); T = LC_DADD(T, u); T = LC_DADD(T, u);
    }
    for (; count > 0; count--) T = LC_DADD(T, u);
    return T;
}

// exact np.cumsum values cum[s], cum[s+1] from the register model: fixed-count chains
__device__ __forceinline__ void lcf_exact_at(const LcFast &F, int s, LcInterval &out)
{
    double T = 0.0;
    int i = 0; // next symbol index to add
    double ps = F.u;
    for (int j = 0; j < F.k; j++) {
        const int sj = __shfl_sync(LC_FULL_MASK, F.my_sym, j);
        const double vj = __shfl_sync(LC_FULL_MASK, F.my_val, j);
        if (sj >= s) { if (sj == s) ps = vj; break; }
        T = lcf_add_run(T, F.u, sj - i);
        T = LC_DADD(T, vj);
        i = sj + 1;
    }
    T = lcf_add_run(T, F.u, s - i);
    out.sym = s; out.clo = T; out.chi = LC_DADD(T, ps); out.exact = 1;
}

// np.searchsorted(cum, v, 'left') - 1 with exact sums, from the register model
__device__ __forceinline__ void lcf_exact_search(const LcFast &F, double v, LcInterval &out)
{
    out.exact = 1;
    if (!(0.0 < v)) { out.sym = -1; out.clo = 0.0; out.chi = 0.0; return; }
    double T = 0.0;
    int i = 0;
    bool done = false;
    for (int j = 0; j <= F.k && !done; j++) {
        const int jj = j < F.k ? j : 0;
        int sj = __shfl_sync(LC_FULL_MASK, F.my_sym, jj);
        const double vj = __shfl_sync(LC_FULL_MASK, F.my_val, jj);
        if (j == F.k) sj = F.n; // tail run
        while (i < sj) {
            const double Tn = LC_DADD(T, F.u);
            if (Tn >= v) { out.sym = i; out.clo = T; out.chi = Tn; done = true; break; }
            T = Tn; i++;
        }
        if (done || j == F.k) break;
        const double Tn = LC_DADD(T, vj);
        if (Tn >= v) { out.sym = sj; out.clo = T; out.chi = Tn; done = true; break; }
        T = Tn; i = sj + 1;
    }
    if (!done) { out.sym = F.n; out.clo = T; out.chi = T; }
}

__device__ __forceinline__ bool lcf_gap_search(const LcFast &F, double v, double gbase, int gfirst, int glen, LcInterval &out)
{
    if (glen <= 0) return false;
    const double d = v - gbase;
    if (!(d > F.delta_v)) return false;
    const double t = d * lc_rcp_fast(F.u); // any error only makes the margin tests below fail
    if (!(t < (double)glen)) return false;
    const int m = (int)t;
    const double lo = gbase + (double)m * F.u;
    const double hi = gbase + (double)(m + 1) * F.u;
    if (!(v - lo > F.delta_v) || !(hi - v >= F.delta_v)) return false;
    out.sym = gfirst + m; out.clo = lo; out.chi = hi; out.exact = 0;
    return true;
}

// second visit: the model is (u, {s1: P1}) and every lane knows all of it -- no shuffles needed
__device__ __forceinline__ bool lcf_search_first(const LcFast &F, int s1, double v, LcInterval &out)
{
    const double A0 = (double)s1 * F.u; // ~cum[s1]
    const double B0 = A0 + F.P1;        // ~cum[s1+1]
    if (A0 - v >= F.delta_v) return lcf_gap_search(F, v, 0.0, 0, s1, out);
    if (v - B0 > F.delta_v) return lcf_gap_search(F, v, B0, s1 + 1, F.n - s1 - 1, out);
    if (v - A0 > F.delta_v && B0 - v >= F.delta_v) { out.sym = s1; out.clo = A0; out.chi = B0; out.exact = 0; return true; }
    return false;
}

// guarded approximate symbol search (see lc_search_fast); the whole list is in registers
__device__ __forceinline__ bool lcf_search(const LcFast &F, double v, LcInterval &out)
{
    const bool valid = F.lane < F.k;
    const int sv = valid ? F.my_sym : 0;
    const double vv = valid ? F.my_val : 0.0;
    double incl = vv;
    for (int off = 1; off < 32; off <<= 1) {
        const double t = __shfl_up_sync(LC_FULL_MASK, incl, off);
        if (F.lane >= off) incl += t;
    }
    const double A = (double)(sv - F.lane) * F.u + (incl - vv);
    const double Bv = A + vv;
    const unsigned hit = __ballot_sync(LC_FULL_MASK, valid && Bv >= v);
    if (hit) {
        const int l = __ffs((int)hit) - 1;
        const int lp = l > 0 ? l - 1 : 0;
        const double Al = __shfl_sync(LC_FULL_MASK, A, l);
        const double Bl = __shfl_sync(LC_FULL_MASK, Bv, l);
        const int sl = __shfl_sync(LC_FULL_MASK, sv, l);
        const double Bp = __shfl_sync(LC_FULL_MASK, Bv, lp);
        const int sp = __shfl_sync(LC_FULL_MASK, sv, lp);
        if (v - Al > F.delta_v) {
            if (!(Bl - v >= F.delta_v)) return false;
            out.sym = sl; out.clo = Al; out.chi = Bl; out.exact = 0;
            return true;
        }
        if (!(Al - v >= F.delta_v)) return false;
        const double gbase = l > 0 ? Bp : 0.0;
        const int gfirst = l > 0 ? sp + 1 : 0;
        return lcf_gap_search(F, v, gbase, gfirst, sl - gfirst, out);
    }
    const int last = F.k - 1;
    const double base = __shfl_sync(LC_FULL_MASK, Bv, last);
    const int g0 = __shfl_sync(LC_FULL_MASK, sv, last) + 1;
    return lcf_gap_search(F, v, base, g0, F.n - g0, out);
}

// ---- bit reader with a 64-bit window (>= 32 valid bits at the top whenever it is read)
struct LcBitReader64 {
    const unsigned char *src;
    long long nbytes;
    uint32_t cur, nxt;
    unsigned long long win;
    int widx, nwin;
};
__device__ __forceinline__ uint32_t lcf_br_load(const LcBitReader64 &b, long long chunk, int lane)
{
    const long long wi = chunk * 32 + lane;
    const long long byte0 = wi * 4;
    if (byte0 >= b.nbytes) return 0u;
    uint32_t w = __byte_perm(__ldg((const uint32_t *)b.src + wi), 0, 0x0123);
    const long long rem = b.nbytes - byte0;
    if (rem < 4) w &= 0xffffffffu << (8 * (4 - (int)rem));
    return w;
}
__device__ __forceinline__ uint32_t lcf_br_next_word(LcBitReader64 &b, int lane)
{
    const uint32_t w = __shfl_sync(LC_FULL_MASK, b.cur, b.widx & 31);
    b.widx++;
    if ((b.widx & 31) == 0) { b.cur = b.nxt; b.nxt = lcf_br_load(b, (long long)(b.widx >> 5) + 1, lane); }
    return w;
}
__device__ __forceinline__ void lcf_br_init(LcBitReader64 &b, const unsigned char *src, long long nbytes, int lane)
{
    b.src = src; b.nbytes = nbytes; b.widx = 0;
    b.cur = lcf_br_load(b, 0, lane);
    b.nxt = lcf_br_load(b, 1, lane);
    const unsigned long long w0 = lcf_br_next_word(b, lane);
    const unsigned long long w1 = lcf_br_next_word(b, lane);
    b.win = (w0 << 32) | w1;
    b.nwin = 64;
}
// next nb bits (0..32), MSB first
__device__ __forceinline__ uint32_t lcf_br_bits(LcBitReader64 &b, int nb, int lane)
{
    if (nb == 0) return 0u;
    const uint32_t v = (uint32_t)(b.win >> (64 - nb));
    b.win = nb == 64 ? 0ull : (b.win << nb);
    b.nwin -= nb;
    if (b.nwin <= 32) {
        const unsigned long long w = lcf_br_next_word(b, lane);
        b.win |= w << (32 - b.nwin);
        b.nwin += 32;
    }
    return v;
}


// decode_symbol (:272-292), search part: the symbol of the open context's model held in F (state 0: uniform
// vector, 1: one update with s1, >= 2: register record) for the coder state (lo, hi, code).
// Returns LC_OK (| 0x100 when the exact path was needed) or the fault status; num/rdv are (code-low+1) and range.
__device__ __forceinline__ int lcf_find_symbol(const LcFast &F, int state, int s1, uint32_t lo, uint32_t hi, uint32_t code,
                                               LcInterval &iv, double &num, double &rdv)
{
    const long long range = (long long)hi - (long long)lo + 1;
    if (range == 0) return LC_DEC_ZERO_RANGE;
    // scaled_value = (code-low+1)*1.0/range - 1e-10 (:285).  The IEEE quotient is only needed when
    // a decision falls inside the guard band; the searches run on a fast quotient whose error
    // (< 2^-46 absolute, v <= ~1) is part of F.delta_v.
    num = lc_ll2d_small((long long)code - (long long)lo + 1);
    rdv = lc_ll2d_small(range);
    const double va = num * lc_rcp_fast(rdv) - 1e-10;
    bool decided;
    int flags = 0;
    if (state == 0) { // uniform context: first i with i/n >= v, i.e. ceil(v*n) - 1
        iv.exact = 1;
        const double t = va * (double)F.n;
        const double tr = nearbyint(t);
        decided = (t > F.tmargin) && (fabs(t - tr) > F.tmargin) && (t < (double)F.n);
        if (decided) iv.sym = (int)ceil(t) - 1;
    } else {
        decided = state == 1 ? lcf_search_first(F, s1, va, iv) : lcf_search(F, va, iv);
    }
    if (!decided) { // exact quotient, exact sums
        flags = 0x100;
        double v = LC_DDIV(LC_DMUL(num, 1.0), rdv);
        v = LC_DSUB(v, 1e-10);
        if (state == 0) {
            iv.exact = 1;
            if (!(0.0 < v)) iv.sym = -1;
            else {
                const double t = LC_DMUL(v, (double)F.n);
                iv.sym = (t > (double)F.n) ? F.n : (int)(LC_D2LL(t) + ((double)LC_D2LL(t) < t ? 1 : 0)) - 1;
            }
        } else lcf_exact_search(F, v, iv);
    }
    if (state == 0) { iv.clo = LC_DMUL((double)iv.sym, F.u0); iv.chi = LC_DMUL((double)(iv.sym + 1), F.u0); }
    if (iv.sym >= F.n) return LC_DEC_SYMBOL_OOB;
    if (iv.sym < 0) return LC_NEEDS_GENERIC; // symbol -1: the generic kernel follows the reference through it
    return LC_OK | flags;
}

// decode_symbol (:291-292), interval part: high/low from the symbol's cumulative bounds
__device__ __forceinline__ void lcf_apply_symbol(const LcFast &F, LcInterval &iv, double num, double rdv, uint32_t &lo,
                                                 uint32_t &hi)
{
    long long low64 = lo, high64 = hi;
    if (!lc_interval_apply(iv, F.delta, low64, high64)) {
        // the symbol itself was decided with margin; only the exact bounds are missing
        lcf_exact_at(F, iv.sym, iv);
#ifdef LC_HOSTSIM
        { // emulator-only check of the margin argument: the exact sums must bracket the exact v
            const double vx = LC_DSUB(LC_DDIV(LC_DMUL(num, 1.0), rdv), 1e-10);
            if (!(iv.clo < vx) || !(vx <= iv.chi)) { fprintf(stderr, "lc_decoder_fast: guard argument violated\n"); abort(); }
        }
#endif
        lc_interval_apply(iv, F.delta, low64, high64);
    }
    (void)num; (void)rdv;
    lo = (uint32_t)low64; hi = (uint32_t)high64;
}

// decode_symbol (:272-292) in a context never seen before, in integer arithmetic (same derivation as decoder v2's
// state-0 path): the uniform model's bounds i/n are exact for a power-of-two n, so with a = (code-low+1)*n and
// E = 1e-10*n*range the symbol is the s with s*range < a - E <= (s+1)*range.  Candidate from a float quotient, decided
// only with integer margins (the reference's float64 evaluation is within 1e-3 of these units); returns false when
// they do not decide, and the caller takes the exact path.  lg_n = log2 n, eps_k = floor(1e-10 * n * 2^40).
#ifdef LC_HOSTSIM
static inline float lcf_rcp_f32(float x) { return 1.0f / x; }
#else
static __device__ __forceinline__ float lcf_rcp_f32(float x) // x >= 65535 here: no range handling needed
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif
__device__ __forceinline__ bool lcf_uniform_symbol(int n, uint32_t lg_n, uint32_t eps_k, uint32_t lo, uint32_t hi,
                                                   uint32_t code, int &s, uint32_t &nlo, uint32_t &nhi)
{
    const uint32_t rng1 = hi - lo, off = code - lo; // range-1, (code-low+1)-1
    if (!(hi >= lo && off <= rng1 && rng1 >= 0xffffu)) return false;
    int cand = (int)((float)off * lcf_rcp_f32((float)rng1) * (float)n);
    cand = cand > n - 1 ? n - 1 : cand;
    const unsigned long long below = (unsigned long long)(uint32_t)cand * rng1 + (uint32_t)cand; // cand*range
    const unsigned long long above = below + rng1 + 1ull;
    const unsigned long long a = ((unsigned long long)off + 1ull) << lg_n;
    const uint32_t e_lo = __umulhi(rng1, eps_k) >> 8; // <= E < e_lo + 3
    const unsigned long long dlt = a - below;         // e_lo + 4 <= a - below <= range + e_lo - 1, in 32 bits
    if (!((uint32_t)(dlt >> 32) == 0u && (uint32_t)dlt >= e_lo + 4u && (uint32_t)dlt - (e_lo + 4u) <= rng1 - 4u)) return false;
    s = cand;
    nlo = lo + (uint32_t)(below >> lg_n);
    nhi = lo + (uint32_t)(above >> lg_n) - 1u;
    return true;
}

// =================================================================================================
__device__ __forceinline__ void lc_fast_decode_stream(LcFast &F, const unsigned char *src, long long nbytes, LcIdxOut out,
                                                      const float *deq_table, float *deq_out, int *status_out,
                                                      int *fault_index)
{
    const uint32_t mask = F.slot_cap - 1;
    uint32_t lg_n = 0u;
    while ((1 << lg_n) < F.n) lg_n++;
    const bool pow2 = (1 << lg_n) == F.n;
    const uint32_t eps_k = (uint32_t)(1e-10 * (double)F.n * 1099511627776.0);
    LCP_DECL
    LCP_INIT();
    for (uint32_t i = F.lane; i < F.slot_cap; i += 32) __stcg(&F.slots[i], 0ull);
    F.pool_top = 0;
    __syncwarp();
    LcBitReader64 br; lcf_br_init(br, src, nbytes, F.lane);
    uint32_t lo = 0u, hi = 0xffffffffu;
    uint32_t code = lcf_br_bits(br, 32, F.lane); // start_decoding (:247-258)
    int status = LC_OK;
    int left = -1, my_out = 0;
    int pos = 0, r = 0, c = 0;
    lcv_sa row_cur = lcv_sa_of(F.rows), row_prev = row_cur + 2u * (uint32_t)F.C; // the row being decoded / the one above
    // software-pipelined probe: window of 32 slots for the context of position 0
    uint32_t key = 0u; // (left=-1, up=-1)
    uint32_t start = ((key * 2654435761u) >> F.slot_shift) & ~15u;
    unsigned long long w = __ldcg(&F.slots[(start + (uint32_t)F.lane) & mask]);
    for (; pos < F.total; pos++) {
        // ---- resolve the probe
        LCP_START();
        const uint32_t want = key + 1u;
        uint32_t slot_idx;
        unsigned long long word = 0ull;
        for (;;) {
            const uint32_t kk = (uint32_t)(w & 0x3FFFFFull);
            const unsigned mm = __ballot_sync(LC_FULL_MASK, kk == want);
            const unsigned me = __ballot_sync(LC_FULL_MASK, kk == 0u);
            if (mm) { const int l = __ffs((int)mm) - 1; slot_idx = (start + (uint32_t)l) & mask; word = __shfl_sync(LC_FULL_MASK, w, l); break; }
            if (me) { const int l = __ffs((int)me) - 1; slot_idx = (start + (uint32_t)l) & mask; break; }
            start += 32;
            w = __ldcg(&F.slots[(start + (uint32_t)F.lane) & mask]);
        }
        const int state = LCF_STATE(word);
        LCP_ROW(state); LCP_COUNT(state, 0); LCP_MARK(1);
        // ---- bring the model of this context into registers
        if (state == 1) lcf_state_first(F, LCF_S1(word));
        else if (state == 2) {
            lcf_state_first(F, LCF_S1(word));
            lcf_update(F, LCF_S2(word)); // k <= 2: cannot overflow
        } else if (state == 3) {
            F.k = LCF_K(word);
            const char *rec = F.pool + (size_t)LCF_OFF(word) * 16;
            int cl = 1; while ((1 << cl) < F.k) cl++;
            F.u = __ldcg((const double *)rec);
            const bool valid = F.lane < F.k;
            F.my_val = valid ? __ldcg((const double *)(rec + 8) + F.lane) : 0.0;
            F.my_sym = valid ? (int)__ldcg((const unsigned short *)(rec + 8 + (8 << cl)) + F.lane) : 0x7fffffff;
        }
        LCP_MARK(2);
        // ---- decode_symbol (:272-292)
        LcInterval iv;
        double num = 0.0, rdv = 0.0;
        int s = 0;
        uint32_t nlo = 0u, nhi = 0u;
        // a context never seen before (91 % of the symbols at 10 bits): integer path first
        const bool fresh = state == 0 && pow2 && lcf_uniform_symbol(F.n, lg_n, eps_k, lo, hi, code, s, nlo, nhi);
        if (!fresh) {
            const int fs = lcf_find_symbol(F, state, LCF_S1(word), lo, hi, code, iv, num, rdv);
#ifdef LC_DEC_PROFILE
            if (fs & 0x100) LCP_COUNT(4, state);
#endif
            if ((fs & 0xff) != LC_OK) { status = fs & 0xff; break; }
            s = iv.sym;
        }
        LCP_MARK(3);
        // ---- the symbol is known: request the table window of the NEXT position's context now, so
        // its latency overlaps the interval update, renormalisation and write-back below
        // (rows through shared-window addresses, base addresses swapped at row ends: see the note at lcv_sa_*)
        if (F.lane == 0) lcv_sa_st16(row_cur + 2u * (uint32_t)c, s);
        const bool last = c + 1 == F.C;
        const int c2 = last ? 0 : c + 1;
        const int r2 = last ? (r + 1 == F.R ? 0 : r + 1) : r;
        int up2 = -1;
        if (r2 > 0) up2 = (F.C == 1) ? s : lcv_sa_ld16(last ? row_cur : row_prev + 2u * (uint32_t)c2); // C==1: the element just written
        const uint32_t key2 = (uint32_t)((c2 > 0 ? s : -1) + 1) * (uint32_t)(F.n + 1) + (uint32_t)(up2 + 1);
        const uint32_t start2 = ((key2 * 2654435761u) >> F.slot_shift) & ~15u;
        unsigned long long w2 = __ldcg(&F.slots[(start2 + (uint32_t)F.lane) & mask]);

#ifdef LC_DEC_PROFILE
        if (!fresh) { long long l_ = lo, h_ = hi; if (!lc_interval_apply(iv, F.delta, l_, h_)) LCP_COUNT(5, state); }
#endif
        if (fresh) { lo = nlo; hi = nhi; }
        else lcf_apply_symbol(F, iv, num, rdv, lo, hi);
        LCP_MARK(4);
        // ---- renormalise (:295-303) and underflow (:306-309), closed form
        {
            const int d = __clz((int)(lo ^ hi));
            if (d) {
                const uint32_t bits = lcf_br_bits(br, d, F.lane);
                if (d == 32) { lo = 0u; hi = 0xffffffffu; code = bits; }
                else { lo <<= d; hi = (hi << d) | ((1u << d) - 1u); code = (code << d) | bits; }
            }
            const int e = __clz((int)~((lo & ~hi) << 1));
            if (e) {
                const uint32_t bits = lcf_br_bits(br, e, F.lane);
                lo = (lo << e) & 0x7fffffffu;
                hi = ((hi << e) & 0x7fffffffu) | 0x80000000u | ((1u << e) - 1u);
                code = ((code << e) ^ 0x80000000u) | bits;
            }
        }
        // ---- output
        if (F.lane == (pos & 31)) my_out = s;
        if ((pos & 31) == 31) {
            const int p = pos - 31 + F.lane;
            out.store(p, my_out);
            if (deq_out) deq_out[p] = __ldg(deq_table + my_out);
        }
        LCP_MARK(5);
        // ---- write back this context
        unsigned long long new_word;
        if (state == 0) new_word = LCF_PACK_A(want, s);
        else if (state == 1) new_word = LCF_PACK_B(want, LCF_S1(word), s);
        else {
            if (!lcf_update(F, s)) { status = LC_NEEDS_GENERIC; break; }
            int cl = 1; while ((1 << cl) < F.k) cl++;
            uint32_t off16 = LCF_OFF(word);
            bool alloc = (state == 2);
            if (state == 3) { int clo_ = 1; while ((1 << clo_) < LCF_K(word)) clo_++; alloc = cl != clo_; }
            if (alloc) {
                const uint32_t bytes = (8u + (10u << cl) + 15u) & ~15u;
                if (F.pool_top + bytes > F.pool_bytes) { status = LC_POOL_OVERFLOW; break; }
                off16 = F.pool_top >> 4;
                F.pool_top += bytes;
            }
            char *rec = F.pool + (size_t)off16 * 16;
            if (F.lane == 0) __stcg((double *)rec, F.u);
            if (F.lane < F.k) {
                __stcg((double *)(rec + 8) + F.lane, F.my_val);
                __stcg((unsigned short *)(rec + 8 + (8 << cl)) + F.lane, (unsigned short)F.my_sym);
            }
            new_word = LCF_PACK_C(want, F.k, off16);
        }
        if (F.lane == 0) __stcg(&F.slots[slot_idx], new_word);
        if (((start2 + (uint32_t)F.lane) & mask) == slot_idx) w2 = new_word; // the window was read before this store
        __syncwarp();
        key = key2; start = start2; w = w2;
        left = s; c = c2; r = r2;
        if (last) { const lcv_sa t_ = row_cur; row_cur = row_prev; row_prev = t_; }
        LCP_MARK(6);
    }
    LCP_FLUSH();
    *fault_index = pos;
    *status_out = status;
    {
        const int done = pos;
        const int p = (done & ~31) + F.lane;
        if (p < done) { out.store(p, my_out); if (deq_out) deq_out[p] = __ldg(deq_table + my_out); }
        for (int z = done + F.lane; z < F.total; z += 32) { out.store(z, 0); if (deq_out) deq_out[z] = 0.0f; }
    }
    (void)left;
}

// Block entry: one warp per block, persistent over streams.  smem: dense[n] | u1tab[max(16,n<8?n:..)] | rows
__device__ __forceinline__ void lc_fast_decode_block(const LcCoderCfg &cfg, const unsigned char *bytes,
                                                     const long long *offsets, const int *nbits, int B, LcIdxOut out,
                                                     const float *deq_table, float *deq_out, int *status, int *fault,
                                                     char *scratch, char *smem)
{
    LcFast F;
    F.n = cfg.n; F.C = cfg.C; F.R = cfg.R; F.total = cfg.total; F.lane = (int)(threadIdx.x & 31);
    F.rate = cfg.rate; F.delta = cfg.delta; F.u0 = LC_DDIV(1.0, (double)cfg.n);
    F.delta_v = cfg.delta + 1.5e-14;            // + error of the fast quotient (< 2^-46)
    F.tmargin = (double)cfg.n * 1.5e-14;        // the same error scaled by n, for the uniform closed form
    F.slot_cap = cfg.slot_cap; F.slot_shift = cfg.slot_shift; F.pool_bytes = cfg.pool_bytes; F.pool_top = 0;
    F.pw_len = cfg.pw_len; F.pw_steps = cfg.pw_steps; F.pw_chains = cfg.pw_chains;
    char *sc = scratch + (size_t)blockIdx.x * cfg.scratch_stride;
    F.slots = (unsigned long long *)sc;
    F.pool = sc + (size_t)cfg.slot_cap * 8;
    F.dense = (double *)(smem + cfg.sm_dense);
    F.u1tab = (double *)(smem + cfg.sm_lval); // the staged-list area of the generic coder is free here
    F.rows = (unsigned short *)(smem + cfg.sm_rows);
    F.k = 0; F.u = F.u0; F.my_sym = 0x7fffffff; F.my_val = 0.0;
    lcf_tables_init(F);
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        int fi = 0, st = 0;
        const long long nby = ((long long)nbits[sidx] + 7) >> 3;
        lc_fast_decode_stream(F, bytes + offsets[sidx], nby, out + (size_t)sidx * cfg.total, deq_table,
                              deq_out ? deq_out + (size_t)sidx * cfg.total : (float *)0, &st, &fi);
        if (F.lane == 0) { status[sidx] = st; fault[sidx] = fi; }
        __syncwarp();
    }
}
