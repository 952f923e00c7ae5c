// Decoder v2: cabac_decode (cabac_compression.py:363-406), repaired coder mode, (left,up) contexts,
// alphabets of at most 256 symbols.  Same results as lc_decoder_fast.cuh / lc_coder.cuh; restructured
// around the measured per-state cycle budget of the first two versions (profiles/r01_decoder_state_cycles_*):
// the serial chain per symbol was ~2800 cycles, spread evenly over table probe, search, interval and
// -- for contexts seen three times or more -- the model update (pairwise sum, IEEE division).
//
//  * Role-specialised warps.  A block is one stream: warp 0 (the DECODER warp) runs the strictly serial
//    part -- context lookup, symbol search, interval, renormalisation -- and never does model arithmetic;
//    warps 1..LCV_NU (UPDATER warps) run ContextModel.update_model (:119-144) behind it.  The decoder
//    posts (context, symbol) jobs into a shared-memory ring and goes on; a model is only needed again
//    when its context recurs, which the decoder detects from the keys of its last LCV_RING jobs (held
//    one per lane) and then waits for.  No cross-warp flag is read on the common path.
//  * Direct-mapped context table.  key = (left+1)*(n+1)+(up+1) < (n+1)^2 indexes a 2-bit state array in
//    SHARED memory (0 never seen, 1 seen once, 2 record inline, 3 record in the pool), a 4-byte word and
//    a 64-byte record in global memory.  48 % of the symbols of the benchmark open a fresh context and
//    are resolved from shared memory alone; the global arrays are never cleared (the state bits say what
//    is valid).
//  * Fresh contexts (state 0) are decoded in integer arithmetic: the uniform model's bounds i/n are
//    exact, so symbol and new low/high are 64-bit integer products, verified with integer margins.
//  * Contexts seen once (state 1) use a per-launch table of the EXACT np.cumsum values of the model
//    after one update (it depends only on the first symbol): no guard band, no sequential re-summation
//    (the truncation guard of the previous version tripped on a third of these symbols).
//  * Records of at most LCV_INLINE_K entries (94 % of the deeper visits) are searched as scalar code
//    from registers; larger ones use the lane-parallel search of lc_decoder_fast.cuh.
//  * The bit reader is warp-uniform (every lane loads the same word, one refill ahead, kept raw until consumed).
//  * What the next symbol needs is requested one symbol ahead: the context's 4-byte word into a register, its
//    64-byte inline record with cp.async into a shared-memory staging slot.  The hot loop addresses shared memory
//    through 32-bit shared-window addresses (lcv_sa_*, lc_decoder_fast.cuh).  Both came out of the per-instruction
//    profile of this loop (tools/dec_lines.py, tools/dec_stalls.py).
// Whenever a fast path cannot decide with its margins it falls back to lcf_find_symbol /
// lcf_apply_symbol, the same exact evaluation the other kernels use.
#pragma once
#include "lc_decoder_fast.cuh"

#define LCV_NU 1 // (the single completed-jobs counter of the ring assumes jobs finish in order: one updater warp)
#define LCV_WARPS (1 + LCV_NU)
#define LCV_RING 32
#define LCV_INLINE_K 6
#define LCV_SENTINEL 0xFFFFFFFFu
#define LCV_MAX_N 256
#define LCV_MAX_C 4096
#ifndef LCV_OPT_RINV
#define LCV_OPT_RINV 0 // carry 1/range across symbols (see LcvRinv)
#endif

// launch description (host fills it: lcv_cfg_make)
struct LcV2Cfg {
    uint32_t nkeys;       // (n+1)^2
    uint32_t lg_n;        // log2 n
    uint32_t eps_k;       // floor(1e-10 * n * 2^40): the -1e-10 of decode_symbol in units of range/2^40
    uint32_t pool_bytes;  // overflow-record pool per block
    // shared memory carve-up
    uint32_t sm_bits, sm_rows, sm_ring, sm_tab, sm_dense, sm_misc, sm_mail, sm_stage, sm_bytes;
    // per-block global scratch carve-up
    uint64_t g_word, g_rec, g_pool, g_stride;
};

static inline int lcv_eligible(const LcCoderCfg &c)
{
    return c.mode == LC_MODE_REPAIRED && c.has_ctx && c.n >= 8 && c.n <= LCV_MAX_N && c.C >= 4 && c.C <= LCV_MAX_C;
}

static inline void lcv_cfg_make(const LcCoderCfg &c, LcV2Cfg *v)
{
    v->nkeys = (uint32_t)(c.n + 1) * (uint32_t)(c.n + 1);
    v->lg_n = 0;
    while ((1 << v->lg_n) < c.n) v->lg_n++;
    v->eps_k = (uint32_t)(1e-10 * (double)c.n * 1099511627776.0);
    v->pool_bytes = c.pool_bytes;
    uint32_t off = 0;
    v->sm_bits = off;  off += lc_round_up((v->nkeys + 15) / 16 * 4, 16);
    v->sm_rows = off;  off += lc_round_up((uint32_t)c.C * 2, 16);
    v->sm_ring = off;  off += LCV_RING * 8 + LCV_RING * 12; // posted-barriers[32] (8 B) | key[32] | payload[32] | done[32]
    v->sm_tab = off;   off += 2 * 32 * 8;              // u1tab[32] | ru1tab[32]
    v->sm_dense = off; off += LCV_NU * (uint32_t)c.n * 8;
    v->sm_misc = off;  off += 16;                      // pool_top | abort
    v->sm_mail = off;  off += 96;                      // decoder v3: symbol word | packet header | pad | 64 B packet data
    v->sm_stage = off; off += 128;                     // decoder v2: two 64-byte slots the next context's record is copied into
    v->sm_bytes = lc_round_up(off, 16);
    uint64_t g = 0;
    v->g_word = g; g += ((uint64_t)v->nkeys * 4 + 255) & ~(uint64_t)255;
    v->g_rec = g;  g += (uint64_t)v->nkeys * 64;
    v->g_pool = g; g += ((uint64_t)v->pool_bytes + 255) & ~(uint64_t)255;
    v->g_stride = g;
}

// per-launch tables (global): u1tab[32] | ru1tab[32] | cum1[n][n+1]
static inline uint64_t lcv_tables_bytes(int n) { return (uint64_t)(64 + (uint64_t)n * (n + 1)) * 8; }

// "Job posted" signalling: one shared-memory mbarrier per ring slot, one arrival per use of the slot.  A waiting
// updater warp is suspended by the hardware (mbarrier.try_wait) instead of polling -- in the first version of
// this kernel the updaters' poll loops were half of all issued instructions and competed with the decoder warps.
// Use m of a slot (m = job / LCV_RING) completes phase m; waiting for it is a parity wait on (m & 1).  A slot is
// not reused before its job is finished, so the barrier is never more than one phase ahead of its waiter.
#ifdef LC_HOSTSIM
#define LCV_SPIN() emu::spin_yield()
#define LCV_FENCE() ((void)0)
static inline uint32_t lcv_ld_vol(const uint32_t *p) { return *(const volatile uint32_t *)p; }
static inline void lcv_st_vol(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
static inline uint32_t lcv_ld_acq(const uint32_t *p) { return *(const volatile uint32_t *)p; }
static inline void lcv_st_rel(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
static inline void lcv_bar_init(unsigned long long *b) { *(volatile unsigned long long *)b = 0ull; } // completed phases
static inline void lcv_bar_arrive(unsigned long long *b) { *(volatile unsigned long long *)b += 1ull; }
static inline void lcv_bar_wait(uintptr_t b, uint32_t parity)
{
    while (((uint32_t)*(volatile unsigned long long *)b & 1u) == parity) emu::spin_yield();
}
#else
#define LCV_SPIN() __nanosleep(32)
#define LCV_FENCE() __threadfence_block()
static __device__ __forceinline__ uint32_t lcv_ld_vol(const uint32_t *p) { return *(const volatile uint32_t *)p; }
static __device__ __forceinline__ void lcv_st_vol(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
// acquire load / release store at CTA scope (what the LCV_FENCE + volatile pairs express, without the full
// MEMBAR.SC the fence compiles to: that one also waits for the decoder's outstanding prefetch loads)
static __device__ __forceinline__ uint32_t lcv_ld_acq(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
static __device__ __forceinline__ void lcv_st_rel(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
static __device__ __forceinline__ void lcv_bar_init(unsigned long long *b)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
static __device__ __forceinline__ void lcv_bar_arrive(unsigned long long *b) // release.cta
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
// b: shared-window address.  The suspend-time hint lets the hardware keep the warp asleep until the phase completes
// instead of returning every ~80 cycles (the retry loop was 17 instructions per symbol on the updater warp, competing
// for issue slots with the decoder warps of the same sub-partition).
// LCV_OPT_SLEEP > 0: poll with a plain test_wait and sleep that many nanoseconds between polls instead of relying on
// try_wait's suspend hint (the hardware returns from it every ~80 cycles: the retry loop was 26 warp instructions per
// symbol, 8-20 % of everything the kernel issues).  A job is only needed when its context recurs, so a pick-up delay
// of a few hundred cycles is rarely on the decoder warp's path.
#ifndef LCV_OPT_SLEEP
#define LCV_OPT_SLEEP 0
#endif
static __device__ __forceinline__ void lcv_bar_wait(uint32_t b, uint32_t parity) // acquire.cta
{
#if LCV_OPT_SLEEP > 0
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(LCV_OPT_SLEEP);
    }
#else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LCV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra LCV_DONE;\n"
        "bra LCV_WAIT;\n"
        "LCV_DONE:\n"
        "}\n" ::"r"(b), "r"(parity), "r"(0x989680u) : "memory");
#endif
}
#endif

// Scheduling fence for one value: `v` cannot be touched before `after` exists.  ptxas otherwise places the first
// consumer of an early shared-memory load right behind it, where the in-order warp waits out the load's latency.
#ifdef LC_HOSTSIM
#define LCV_USE_AFTER(v, after) ((void)0)
#else
#define LCV_USE_AFTER(v, after) asm volatile("" : "+r"(v) : "r"(after))
#endif

// Asynchronous 64-byte copy global -> shared by lanes 0..3 (cp.async, L2 only), and its completion.  The decoder
// warp requests the next context's record with it one symbol ahead.  A register prefetch (four 128-bit loads into
// loop-carried registers) made the compiler copy the loaded registers right behind the loads, which stalled the
// warp for the full L2/HBM latency on every such symbol (10 % of its time in the ncu stall samples).
#ifdef LC_HOSTSIM
static inline void lcv_stage_copy(lcv_sa dst, const char *src, int lane)
{
    if (lane < 4) memcpy((char *)dst + 16 * lane, src + 16 * lane, 16);
}
static inline void lcv_stage_wait() {}
static inline double2 lcv_sa_ld128(lcv_sa a)
{
    double2 v;
    v.x = ((const volatile double *)a)[0]; v.y = ((const volatile double *)a)[1];
    return v;
}
#else
static __device__ __forceinline__ void lcv_stage_copy(lcv_sa dst, const char *src, int lane)
{
    if (lane < 4)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * (uint32_t)lane), "l"(src + 16 * lane) : "memory");
}
static __device__ __forceinline__ void lcv_stage_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }
static __device__ __forceinline__ double2 lcv_sa_ld128(lcv_sa a)
{
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
#endif

// ---- tables kernel body, one warp (block) per first symbol s1: u after the first update with s1 (its value
// depends only on s1's accumulator-chain step; the block of the step's first symbol publishes it with its
// reciprocal), and the exact np.cumsum (:346-347) of the model after that update.
__device__ __forceinline__ void lcv_tables_block(const LcCoderCfg &cfg, double *tables, char *smem)
{
    double *u1g = tables, *ru1g = tables + 32, *cum1 = tables + 64;
    const int n = cfg.n, lane = (int)(threadIdx.x & 31);
    LcFast F;
    F.n = n; F.lane = lane; F.rate = cfg.rate; F.u0 = LC_DDIV(1.0, (double)n);
    F.pw_len = cfg.pw_len; F.pw_steps = cfg.pw_steps; F.pw_chains = cfg.pw_chains;
    F.dense = (double *)smem;
    const double P1 = LC_DADD(F.u0, LC_DMUL(F.rate, LC_DSUB(1.0, F.u0)));
    for (int s1 = (int)blockIdx.x; s1 < n; s1 += (int)gridDim.x) {
        // ContextModel.update_model (:119-144) on the uniform vector
        for (int i = lane; i < n; i += 32) F.dense[i] = (i == s1) ? P1 : F.u0;
        __syncwarp();
        const double total = lcf_pairwise_total(F);
        const double others = LC_DSUB(total, P1);
        const double f = (others > 0.0) ? LC_DDIV(LC_DSUB(1.0, P1), others) : 0.0;
        const double u = LC_DMUL(F.u0, f);
        __syncwarp();
        const int t = cfg.pw_chains == 0 ? s1 : ((s1 & (cfg.pw_len - 1)) >> 3);
        const bool publishes = cfg.pw_chains == 0 ? true : (s1 == 8 * t);
        if (lane == 0 && publishes) { u1g[t] = u; ru1g[t] = lc_rcp_fast(u); }
        if (lane == 0) {
            double *row = cum1 + (size_t)s1 * (n + 1);
            double T = 0.0;
            row[0] = 0.0;
            for (int i = 0; i < n; i++) { T = LC_DADD(T, i == s1 ? P1 : u); row[i + 1] = T; }
        }
    }
}

// ---- warp-uniform bit reader: 64-bit window, the next word always already loaded
struct LcvBits {
    const uint32_t *w;
    uint32_t nbytes; // a stream is at most 2^22 symbols of at most ~80 bits
    unsigned long long win;
    int nwin;
    uint32_t widx;
    uint32_t nextw;
};
// word wi of the stream as loaded (little-endian memory order, garbage past the end) ...
__device__ __forceinline__ uint32_t lcv_br_load(const LcvBits &b, uint32_t wi)
{
    return wi * 4u < b.nbytes ? __ldg(b.w + wi) : 0u;
}
// ... and as the next 32 bits of the stream, MSB first, zeros past the end (the reference reads zeros there, :260-270).
// Split so that nothing touches the loaded register until the word is consumed one refill later.
__device__ __forceinline__ uint32_t lcv_br_fix(const LcvBits &b, uint32_t raw, uint32_t wi)
{
    uint32_t w = __byte_perm(raw, 0, 0x0123);
    const uint32_t byte0 = wi * 4u;
    if (byte0 + 4u > b.nbytes) w = byte0 >= b.nbytes ? 0u : (w & (0xffffffffu << (8u * (4u - (b.nbytes - byte0)))));
    return w;
}
__device__ __forceinline__ void lcv_br_init(LcvBits &b, const unsigned char *src, long long nbytes)
{
    b.w = (const uint32_t *)src; b.nbytes = (uint32_t)(nbytes > 0x7fffffffll ? 0x7fffffffll : nbytes);
    b.win = ((unsigned long long)lcv_br_fix(b, lcv_br_load(b, 0), 0) << 32) | lcv_br_fix(b, lcv_br_load(b, 1), 1);
    b.nwin = 64; b.widx = 2; b.nextw = lcv_br_load(b, 2);
}
// drop nb bits (0..32); the window keeps more than 32 valid bits at its top
__device__ __forceinline__ void lcv_br_skip(LcvBits &b, int nb)
{
    b.win <<= nb;
    b.nwin -= nb;
    if (b.nwin <= 32) {
        b.win |= (unsigned long long)lcv_br_fix(b, b.nextw, b.widx) << (32 - b.nwin);
        b.nwin += 32;
        b.widx++;
        b.nextw = lcv_br_load(b, b.widx);
    }
}
// next nb bits (1..32), MSB first
__device__ __forceinline__ uint32_t lcv_br_take(LcvBits &b, int nb)
{
    const uint32_t v = (uint32_t)(b.win >> (64 - nb));
    lcv_br_skip(b, nb);
    return v;
}

#ifdef LC_HOSTSIM
static inline float lcv_rcp_f32(float x) { return 1.0f / x; }
#else
static __device__ __forceinline__ float lcv_rcp_f32(float x) // x >= 65535 here: no range handling needed
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#endif

// ---- shared/global views of one block
struct LcV2 {
    uint32_t *sbits;          // 2-bit context states
    unsigned char *rows;      // previous/current row of decoded symbols
    unsigned long long *ring_bar; // "job posted" barrier per ring slot
    uint32_t *ring_key, *ring_pay, *ring_done;
    double *u1tab, *ru1tab;
    uint32_t *pool_top, *abort_code;
    uint32_t *gword;          // per context: first symbol (state 1) or K|offset of the pool record (state 3)
    char *grec;               // per context: 64-byte inline record  u | val[6] | sym[6] k pad
    char *pool;
    const double *cum1;
    const char *t2;           // per-launch records of the models after two visits (lcv_t2_block), or null
    uint32_t pool_bytes, eps_k, lg_n;
    // the same shared-memory areas as 32-bit shared addresses, for the decoder warp (lcv_view_sa)
    lcv_sa sa_bits, sa_rows, sa_ring_bar, sa_ring_key, sa_ring_pay, sa_ring_done, sa_tab, sa_abort, sa_stage;
};
__device__ __forceinline__ void lcv_view_sa(LcV2 &V)
{
    V.sa_bits = lcv_sa_of(V.sbits); V.sa_rows = lcv_sa_of(V.rows); V.sa_ring_bar = lcv_sa_of(V.ring_bar);
    V.sa_ring_key = lcv_sa_of(V.ring_key); V.sa_ring_pay = lcv_sa_of(V.ring_pay);
    V.sa_ring_done = lcv_sa_of(V.ring_done); V.sa_tab = lcv_sa_of(V.u1tab); V.sa_abort = lcv_sa_of(V.abort_code);
    V.sa_stage = 0;
}

#define LCV_PAY(s, st, s1) ((uint32_t)(s) | ((uint32_t)(st) << 10) | ((uint32_t)(s1) << 12))
#define LCV_PAY_S(w) ((int)((w) & 0x3FFu))
#define LCV_PAY_ST(w) ((int)(((w) >> 10) & 3u))
#define LCV_PAY_S1(w) ((int)(((w) >> 12) & 0x3FFu))
#define LCV_WORD_C(k, off16) ((uint32_t)(k) | ((uint32_t)(off16) << 6))
#define LCV_WORD_K(w) ((int)((w) & 0x3Fu))
#define LCV_WORD_OFF(w) ((uint32_t)(w) >> 6)

// lane-distributed register model (LcFast) from a pool record
__device__ __forceinline__ void lcv_load_pool(LcFast &F, const LcV2 &V, uint32_t word)
{
    F.k = LCV_WORD_K(word);
    if (F.k > 32) F.k = 32; // only reachable on a stream already flagged for the generic kernel
    const char *rec = V.pool + (size_t)LCV_WORD_OFF(word) * 16;
    int cl = 1; while ((1 << cl) < F.k) cl++;
    F.u = __ldcg((const double *)rec);
    const bool valid = F.lane < F.k;
    F.my_val = valid ? __ldcg((const double *)(rec + 8) + F.lane) : 0.0;
    F.my_sym = valid ? (int)__ldcg((const unsigned short *)(rec + 8 + (8 << cl)) + F.lane) : 0x7fffffff;
}

// store the register model (at most LCV_INLINE_K entries) as a 64-byte record  u | val[6] | sym[6] k
__device__ __forceinline__ void lcv_record_store(char *rec, const LcFast &F)
{
    if (F.lane == 0) { __stcg((double *)rec, F.u); __stcg((unsigned char *)rec + 62, (unsigned char)F.k); }
    if (F.lane < F.k) {
        __stcg((double *)(rec + 8) + F.lane, F.my_val);
        __stcg((unsigned char *)rec + 56 + F.lane, (unsigned char)F.my_sym);
    }
}

// Records of the model after the first TWO visits of a context, for every ordered pair (s1, s2): like cum1 they
// depend only on (n, rate), so n^2 updates once per launch replace one update per context (a quarter of the
// encoder's phase A, and the updater's job for every second visit -- which becomes a 64-byte copy).  n <= 256.
// Record (s1*n + s2) at t2 + 64*(s1*n + s2).  One warp per pair; smem: one n-double image per warp.
__device__ __forceinline__ void lcv_t2_block(const LcCoderCfg &cfg, const double *tables, char *t2, char *smem)
{
    const int warp = (int)(threadIdx.x >> 5), n_warps = (int)(blockDim.x >> 5);
    LcFast F;
    F.n = cfg.n; F.C = cfg.C; F.R = cfg.R; F.total = cfg.total; F.lane = (int)(threadIdx.x & 31);
    F.rate = cfg.rate; F.delta = cfg.delta; F.u0 = LC_DDIV(1.0, (double)cfg.n);
    F.delta_v = 0.0; F.tmargin = 0.0;
    F.P1 = LC_DADD(F.u0, LC_DMUL(F.rate, LC_DSUB(1.0, F.u0)));
    F.slot_cap = 0; F.slot_shift = 0; F.pool_bytes = 0; F.pool_top = 0;
    F.pw_len = cfg.pw_len; F.pw_steps = cfg.pw_steps; F.pw_chains = cfg.pw_chains;
    F.slots = (unsigned long long *)0; F.pool = (char *)0;
    F.dense = (double *)smem + (size_t)warp * cfg.n;
    F.u1tab = const_cast<double *>(tables); F.rows = (unsigned short *)0;
    F.k = 0; F.u = F.u0; F.my_sym = 0x7fffffff; F.my_val = 0.0;
    const int n = cfg.n, pairs = n * n;
    for (int e = (int)blockIdx.x * n_warps + warp; e < pairs; e += (int)gridDim.x * n_warps) {
        lcf_state_first(F, e / n);
        lcf_update(F, e % n); // k <= 2: cannot overflow
        lcv_record_store(t2 + (size_t)e * 64, F);
        __syncwarp();
    }
}

// lane-distributed register model from a 64-byte record
__device__ __forceinline__ void lcv_record_load(LcFast &F, const char *rec)
{
    int k = (int)__ldcg((const unsigned char *)rec + 62);
    if (k > LCV_INLINE_K) k = LCV_INLINE_K;
    F.k = k;
    F.u = __ldcg((const double *)rec);
    const bool valid = F.lane < k;
    F.my_val = valid ? __ldcg((const double *)(rec + 8) + F.lane) : 0.0;
    F.my_sym = valid ? (int)__ldcg((const unsigned char *)rec + 56 + F.lane) : 0x7fffffff;
}
__device__ __forceinline__ void lcv_load_inline(LcFast &F, const LcV2 &V, uint32_t key)
{
    const char *rec = V.grec + (size_t)key * 64;
    int k = (int)__ldcg((const unsigned char *)rec + 62);
    if (k > LCV_INLINE_K) k = LCV_INLINE_K;
    F.k = k;
    F.u = __ldcg((const double *)rec);
    const bool valid = F.lane < k;
    F.my_val = valid ? __ldcg((const double *)(rec + 8) + F.lane) : 0.0;
    F.my_sym = valid ? (int)__ldcg((const unsigned char *)rec + 56 + F.lane) : 0x7fffffff;
}

// =================================================================================================
// UPDATER warps
// =================================================================================================
// j: this warp's next job (jobs are numbered through all streams of the block; warp u takes j = u mod LCV_NU)
__device__ __forceinline__ void lcv_updater(LcFast &F, const LcV2 &V, uint32_t &j)
{
#ifdef LC_DEC_PROFILE // busy cycles and job counts per context state -> lc_prof_global[56..63]
    unsigned long long lcu_busy[4] = {0ull, 0ull, 0ull, 0ull}, lcu_jobs[4] = {0ull, 0ull, 0ull, 0ull};
    long long lcu_t = 0; int lcu_st = 0;
#define LCU_BEGIN(st_) do { lcu_st = (st_); lcu_t = clock64(); } while (0)
#define LCU_END() do { lcu_busy[lcu_st] += (unsigned long long)(clock64() - lcu_t); lcu_jobs[lcu_st] += 1ull; } while (0)
#else
#define LCU_BEGIN(st_)
#define LCU_END()
#endif
    for (;; j += LCV_NU) {
        const uint32_t slot = j & (LCV_RING - 1);
        lcv_bar_wait(V.sa_ring_bar + 8u * slot, (j / LCV_RING) & 1u);
        const uint32_t key = lcv_sa_ld32(V.sa_ring_key + 4u * slot);
        const uint32_t pay = lcv_sa_ld32(V.sa_ring_pay + 4u * slot);
        LCU_BEGIN(LCV_PAY_ST(pay));
        if (key == LCV_SENTINEL) {
#ifdef LC_DEC_PROFILE
            if (F.lane == 0) for (int i_ = 0; i_ < 4; i_++) { atomicAdd(&lc_prof_global[56 + i_], lcu_busy[i_]); atomicAdd(&lc_prof_global[60 + i_], lcu_jobs[i_]); }
#endif
            __syncwarp();
            if (F.lane == 0) lcv_sa_st32(V.sa_ring_done, j + 1u);
            j += LCV_NU;
            break;
        }
        if (lcv_sa_ld32(V.sa_abort) == 0u) {
            const int s = LCV_PAY_S(pay), st = LCV_PAY_ST(pay);
            const uint32_t shift = (key & 15u) * 2u;
            uint32_t word = 0u;
            if (st == 1 && V.t2) {
                // second visit: the model after (s1, s) is a per-launch record -- copy it
                if (F.lane < 4) {
                    const double2 v = __ldg((const double2 *)(V.t2 + ((size_t)LCV_PAY_S1(pay) * F.n + s) * 64) + F.lane);
                    __stcg((double2 *)(V.grec + (size_t)key * 64) + F.lane, v);
                }
                LCV_FENCE();
                if (F.lane == 0) atomicXor(V.sbits + (key >> 4), 3u << shift); // 01 -> 10
                __syncwarp();
                LCV_FENCE();
                if (F.lane == 0) lcv_sa_st32(V.sa_ring_done, j + 1u);
                LCU_END();
                continue;
            }
            if (st == 1) lcf_state_first(F, LCV_PAY_S1(pay));
            else if (st == 2) lcv_load_inline(F, V, key);
            else { word = __ldcg(V.gword + key); lcv_load_pool(F, V, word); }
            const int k_old = F.k;
            if (!lcf_update(F, s)) {
                if (F.lane == 0) atomicCAS(V.abort_code, 0u, (uint32_t)LC_NEEDS_GENERIC);
            } else if (F.k <= LCV_INLINE_K) {
                lcv_record_store(V.grec + (size_t)key * 64, F);
                LCV_FENCE();
                if (st == 1 && F.lane == 0) atomicXor(V.sbits + (key >> 4), 3u << shift); // 01 -> 10
            } else {
                int cl = 1; while ((1 << cl) < F.k) cl++;
                uint32_t off16 = LCV_WORD_OFF(word);
                bool alloc = st != 3;
                if (st == 3) { int clo_ = 1; while ((1 << clo_) < k_old) clo_++; alloc = cl != clo_; }
                bool ok = true;
                if (alloc) {
                    const uint32_t bytes = (8u + (10u << cl) + 15u) & ~15u;
                    uint32_t top = 0u;
                    if (F.lane == 0) top = atomicAdd(V.pool_top, bytes);
                    top = __shfl_sync(LC_FULL_MASK, top, 0);
                    if (top + bytes > V.pool_bytes) {
                        ok = false;
                        if (F.lane == 0) atomicCAS(V.abort_code, 0u, (uint32_t)LC_POOL_OVERFLOW);
                    }
                    off16 = top >> 4;
                }
                if (ok) {
                    char *rec = V.pool + (size_t)off16 * 16;
                    if (F.lane == 0) __stcg((double *)rec, F.u);
                    if (F.lane < F.k) {
                        __stcg((double *)(rec + 8) + F.lane, F.my_val);
                        __stcg((unsigned short *)(rec + 8 + (8 << cl)) + F.lane, (unsigned short)F.my_sym);
                    }
                    if (F.lane == 0) __stcg(V.gword + key, LCV_WORD_C(F.k, off16));
                    LCV_FENCE();
                    if (st != 3 && F.lane == 0) atomicXor(V.sbits + (key >> 4), (uint32_t)(st ^ 3) << shift);
                }
            }
        }
        __syncwarp();
        LCV_FENCE();
        if (F.lane == 0) lcv_sa_st32(V.sa_ring_done, j + 1u);
        LCU_END();
    }
}

// =================================================================================================
// DECODER warp
// =================================================================================================
struct LcvPost {
    uint32_t done_seen;  // a value of the updater's completed-jobs counter read earlier (it only grows)
    uint32_t njobs;      // jobs posted so far in this stream
    uint32_t my_key;     // lane l: key of the last job posted into ring slot l
    uint32_t my_job;     // ... and its index
};

__device__ __forceinline__ void lcv_post(const LcV2 &V, LcvPost &P, int lane, uint32_t key, uint32_t pay)
{
    const uint32_t j = P.njobs, slot = j & (LCV_RING - 1);
    // the job that used this slot (j - RING) must be finished before the slot is reused.  The updater warp finishes
    // jobs in order and publishes their count in one word (ring_done[0]); the decoder warp only re-reads it when the
    // value it saw last no longer proves that -- the updater is ahead nearly always, so the common case is a compare.
    if (j - P.done_seen >= (uint32_t)LCV_RING) {
        for (;;) {
            P.done_seen = lcv_sa_ld32_acq(V.sa_ring_done);
            if (j - P.done_seen < (uint32_t)LCV_RING) break;
            LCV_SPIN();
        }
    }
    if (lane == 0) {
        lcv_sa_st32(V.sa_ring_key + 4u * slot, key); lcv_sa_st32(V.sa_ring_pay + 4u * slot, pay);
        lcv_sa_bar_arrive(V.sa_ring_bar + 8u * slot);
    }
    if ((uint32_t)lane == slot) { P.my_key = key; P.my_job = j; }
    P.njobs = j + 1u;
}

// approximate symbol search in a gap of never-observed symbols (see lcf_gap_search), reciprocal supplied
__device__ __forceinline__ bool lcv_gap_search(double u, double ru, double dv, double v, double gbase, int gfirst,
                                               int glen, LcInterval &out)
{
    if (glen <= 0) return false;
    const double d = v - gbase;
    if (!(d > dv)) return false;
    const double t = d * ru; // any error only makes the margin tests below fail
    if (!(t < (double)glen)) return false;
    const int m = (int)t;
    const double lo = gbase + (double)m * u;
    const double hi = gbase + (double)(m + 1) * u;
    if (!(v - lo > dv) || !(hi - v >= dv)) return false;
    out.sym = gfirst + m; out.clo = lo; out.chi = hi; out.exact = 0;
    return true;
}

// lane-distributed register model (LcFast) from an inline record held in registers by every lane: only the exact
// paths need it
__device__ __forceinline__ void lcv_record_to_lanes(LcFast &F, const double2 &q0, const double2 &q1, const double2 &q2,
                                                    const double2 &q3)
{
    const unsigned long long sb = (unsigned long long)__double_as_longlong(q3.y);
    int k = (int)((sb >> 48) & 0xffu);
    if (k > LCV_INLINE_K) k = LCV_INLINE_K;
    const double val[LCV_INLINE_K] = {q0.y, q1.x, q1.y, q2.x, q2.y, q3.x};
    F.k = k; F.u = q0.x; F.my_sym = 0x7fffffff; F.my_val = 0.0;
#pragma unroll
    for (int j = 0; j < LCV_INLINE_K; j++)
        if (F.lane == j && j < k) { F.my_sym = (int)((sb >> (8 * j)) & 0xffu); F.my_val = val[j]; }
}

// The exact evaluation of one symbol (lcf_find_symbol / lcf_apply_symbol on the lane-distributed model) for whatever
// the fast paths could not decide, and for pool records (state 3).  About 3 % of the benchmark's symbols come here;
// inlined, this code made the decoder warp's loop 28 KB long and 5 % of its stall samples were instruction fetches, so
// it is a real function: the call costs ~100 cycles on those symbols only.
// Measured: as a call it helps the 10-streams-per-SM build (twenty warps at different program counters per SM:
// decode of 8192 streams 41.0 -> 40.3 ms) and costs the 8-per-SM build 2 % (the call and the copy of the model on 3 % of
// the symbols), so only the throughput build calls it; the latency build inlines the same body.
struct LcvCold { int status, s; uint32_t nlo, nhi; };
__device__ __forceinline__ LcvCold lcv_exact_symbol(LcFast &F, const char *pool, int st, int s1, uint32_t gw, const double2 &q0,
                                                    const double2 &q1, const double2 &q2, const double2 &q3, uint32_t lo,
                                                    uint32_t hi, uint32_t code)
{
    LcvCold r; r.status = LC_OK; r.s = 0; r.nlo = 0u; r.nhi = 0u;
    if (st == 1) lcf_state_first(F, s1);
    else if (st == 2) lcv_record_to_lanes(F, q0, q1, q2, q3);
    else if (st == 3) { // lane-distributed register model from the pool record (lcv_load_pool)
        F.k = LCV_WORD_K(gw);
        if (F.k > 32) F.k = 32; // only reachable on a stream already flagged for the generic kernel
        const char *rec = pool + (size_t)LCV_WORD_OFF(gw) * 16;
        int cl = 1; while ((1 << cl) < F.k) cl++;
        F.u = __ldcg((const double *)rec);
        const bool valid = F.lane < F.k;
        F.my_val = valid ? __ldcg((const double *)(rec + 8) + F.lane) : 0.0;
        F.my_sym = valid ? (int)__ldcg((const unsigned short *)(rec + 8 + (8 << cl)) + F.lane) : 0x7fffffff;
    }
    LcInterval iv;
    double num, rdv;
    const int fs = lcf_find_symbol(F, st == 3 ? 2 : st, s1, lo, hi, code, iv, num, rdv) & 0xff;
    if (fs != LC_OK) { r.status = fs; return r; }
    lcf_apply_symbol(F, iv, num, rdv, lo, hi);
    r.s = iv.sym; r.nlo = lo; r.nhi = hi;
    return r;
}
static __device__ __noinline__ LcvCold lcv_cold_symbol(LcFast *Fp, const char *pool, int st, int s1, uint32_t gw, double2 q0,
                                                       double2 q1, double2 q2, double2 q3, uint32_t lo, uint32_t hi,
                                                       uint32_t code)
{
    return lcv_exact_symbol(*Fp, pool, st, s1, gw, q0, q1, q2, q3, lo, hi, code);
}

// 1/range carried from symbol to symbol (LCV_OPT_RINV): the range after renormalisation is the range before it times
// 2^t exactly (low gets zeros shifted in, high ones; an underflow step flips the same bit of both), so its reciprocal
// is the reciprocal of the pre-renormalisation range with t subtracted from the exponent -- and that one can be
// computed while the leading-zero counts of the renormalisation are still in flight, instead of at the head of the
// next symbol's dependent chain (I2F + MUFU.RCP for fresh contexts, MUFU.RCP64H + four DFMA for the others).
struct LcvRinv { double d; float f; };
#ifdef LC_HOSTSIM
static inline double lcv_scale_down(double r, int t) { return std::ldexp(r, -t); }
static inline float lcv_scale_down_f(float r, int t) { return std::ldexp(r, -t); }
#else
static __device__ __forceinline__ double lcv_scale_down(double r, int t) // r * 2^-t, r in [2^-33, 2^-15], t <= 32
{
    return __hiloint2double(__double2hiint(r) - (t << 20), __double2loint(r));
}
static __device__ __forceinline__ float lcv_scale_down_f(float r, int t) { return __int_as_float(__float_as_int(r) - (t << 23)); }
#endif

// decode_symbol (:272-292) for one symbol: the symbol, and low/high after the interval update (before
// renormalisation), from the context's state st and the data that state needs (gw: the context word for states 1
// and 3; q0..q3: the inline record for state 2).  on_candidate(sym) is called as soon as a path has its candidate
// symbol (again with the final symbol if the exact evaluation was needed).  Returns LC_OK or the fault status.
// OUTLINE: the exact evaluation as a real function call (lcv_cold_symbol) instead of inlined code.
template <bool OUTLINE, class OnCandidate>
__device__ __forceinline__ int lcv_decode_symbol(LcFast &F, const LcV2 &V, int st, uint32_t gw, const double2 &q0,
                                                 const double2 &q1, const double2 &q2, const double2 &q3, uint32_t lo,
                                                 uint32_t hi, uint32_t code, int &s, int &s1, uint32_t &nlo, uint32_t &nhi,
                                                 int &fallback, OnCandidate on_candidate, const LcvRinv &rinv)
{
    const int lane = F.lane, n = F.n;
    const double cfix = 1e-10;
    fallback = 0;
    // ---- decode_symbol (:272-292)
    const uint32_t rng1 = hi - lo, off = code - lo; // range-1, (code-low+1)-1
    const bool pre_ok = hi >= lo && off <= rng1 && rng1 >= 0xffffu;
    bool done = false;
    s = 0; s1 = 0; nlo = 0u; nhi = 0u;
    if (st == 0) {
        if (pre_ok) {
            // uniform model: cum[i] = i/n exactly, so range*cum is exact and everything is integer work.
            // With a = (code-low+1)*n and E = 1e-10*n*range, the symbol is the s with
            // s*range < a - E (+- 3e-4) <= (s+1)*range; candidate from a float quotient, checked with margins.
            int cand = (int)((float)off * (LCV_OPT_RINV ? rinv.f : lcv_rcp_f32((float)rng1)) * (float)n);
            cand = cand > n - 1 ? n - 1 : cand; // (float)off rounds up to 2^32 at most: cand <= n
            on_candidate(cand);
            const unsigned long long below = (unsigned long long)(uint32_t)cand * rng1 + (uint32_t)cand; // cand*range
            const unsigned long long above = below + rng1 + 1ull;
            const unsigned long long a = ((unsigned long long)off + 1ull) << V.lg_n;
            const uint32_t e_lo = __umulhi(rng1, V.eps_k) >> 8; // <= E < e_lo + 3
            // below + e_lo + 4 <= a  and  a + 1 <= above + e_lo, i.e. e_lo + 4 <= a - below <= range + e_lo - 1, in 32
            // bits (a - below >= 2^32 would need the full 2^32 range and a top symbol: left to the exact path)
            const unsigned long long dlt = a - below;
            if ((uint32_t)(dlt >> 32) == 0u && (uint32_t)dlt >= e_lo + 4u && (uint32_t)dlt - (e_lo + 4u) <= rng1 - 4u) {
                s = cand; done = true;
                nlo = lo + (uint32_t)(below >> V.lg_n);
                nhi = lo + (uint32_t)(above >> V.lg_n) - 1u;
            }
        }
    } else if (st == 1) {
        s1 = (int)(gw & 0x3FFu);
        if (s1 >= n) s1 = n - 1; // only on a stream already flagged for the generic kernel
        if (pre_ok) { // model after one update: exact np.cumsum values from the per-launch table
            const double rd = lc_ll2d_small((long long)rng1 + 1), nd = lc_ll2d_small((long long)off + 1);
            const int t = lcf_tab_index(F, s1);
            const double u = lcv_sa_ldf64(V.sa_tab + 8u * (uint32_t)t), ru = lcv_sa_ldf64(V.sa_tab + 256u + 8u * (uint32_t)t);
            const double va = nd * (LCV_OPT_RINV ? rinv.d : lc_rcp_fast(rd)) - cfix;
            const double A0 = (double)s1 * u, B0 = A0 + F.P1;
            int sc;
            if (va < A0) sc = (int)(va * ru);
            else if (va <= B0) sc = s1;
            else sc = s1 + 1 + (int)((va - B0) * ru);
            sc = sc < 0 ? 0 : (sc > n - 1 ? n - 1 : sc);
            on_candidate(sc);
            const double *row = V.cum1 + (size_t)s1 * (n + 1);
            const double clo = __ldg(row + sc), chi = __ldg(row + sc + 1);
            const double xl = LC_DMUL(rd, clo), xh1 = LC_DMUL(rd, chi);
            const double tgt = nd - cfix * rd; // ~ v*range; |error| < 3e-6 for range <= 2^32
            if (tgt - xl > 1e-5 && xh1 - tgt >= 1e-5) { // cum[sc] < v <= cum[sc+1], decided with margin
                s = sc; done = true;
                nlo = lo + (uint32_t)LC_D2LL(xl);
                nhi = lo + (uint32_t)LC_D2LL(LC_DSUB(xh1, 1.0));
            }
        }
    } else if (st == 2) {
        const double u = q0.x;
        const unsigned long long sb = (unsigned long long)__double_as_longlong(q3.y);
        const uint32_t sb_lo = (uint32_t)sb, sb_hi = (uint32_t)(sb >> 32);
        int k = (int)((sb_hi >> 16) & 0xffu);
        if (k > LCV_INLINE_K) k = LCV_INLINE_K;
        bool decided = false;
        LcInterval iv; iv.sym = 0; iv.clo = 0.0; iv.chi = 0.0; iv.exact = 0;
        const double rd = lc_ll2d_small((long long)rng1 + 1), nd = lc_ll2d_small((long long)off + 1);
        if (pre_ok) {
            const double ru = lc_rcp_fast(u);
            const double va = nd * (LCV_OPT_RINV ? rinv.d : lc_rcp_fast(rd)) - cfix;
            // entries in ascending symbol order; approximate cum before (A) and after (B) each: stop at the first entry
            // whose upper bound reaches v.  Sequential with early exit -- a record holds 2-3 entries on average, and
            // the fully unrolled select form of this scan was a quarter of the warp's instructions on these symbols.
            double S = 0.0, Al = 0.0, Bl = 0.0, Bp = 0.0; // Bp/gf: end of the previous entry = start of the gap before l
            int sl = 0, gf = 0;
            bool found = false;
#define LCV_SCAN_STEP(j_, sym_expr_, val_)                                              \
            if (k <= (j_)) break;                                                        \
            {                                                                            \
                const int sy_ = (int)(sym_expr_);                                        \
                const double A_ = (double)(sy_ - (j_)) * u + S;                          \
                const double B_ = A_ + (val_);                                           \
                if (B_ >= va) { found = true; Al = A_; Bl = B_; sl = sy_; break; }       \
                Bp = B_; gf = sy_ + 1; S += (val_);                                      \
            }
            do {
                LCV_SCAN_STEP(0, sb_lo & 0xffu, q0.y)
                LCV_SCAN_STEP(1, (sb_lo >> 8) & 0xffu, q1.x)
                LCV_SCAN_STEP(2, (sb_lo >> 16) & 0xffu, q1.y)
                LCV_SCAN_STEP(3, sb_lo >> 24, q2.x)
                LCV_SCAN_STEP(4, sb_hi & 0xffu, q2.y)
                LCV_SCAN_STEP(5, (sb_hi >> 8) & 0xffu, q3.x)
            } while (0);
#undef LCV_SCAN_STEP
            if (found) {
                if (va - Al > F.delta_v) {
                    if (Bl - va >= F.delta_v) { iv.sym = sl; iv.clo = Al; iv.chi = Bl; decided = true; }
                } else if (Al - va >= F.delta_v) {
                    decided = lcv_gap_search(u, ru, F.delta_v, va, Bp, gf, sl - gf, iv);
                }
            } else {
                decided = lcv_gap_search(u, ru, F.delta_v, va, Bp, gf, n - gf, iv);
            }
        }
        if (decided) {
            on_candidate(iv.sym);
            long long low64 = lo, high64 = hi;
            if (OUTLINE) {
                // when the truncations are not stable under the bounds' error (4 % of these symbols) the exact
                // evaluation below redoes the symbol: same result, exact sums, no inlined copy of lcf_exact_at here
                if (lc_interval_apply(iv, F.delta, low64, high64)) {
                    nlo = (uint32_t)low64; nhi = (uint32_t)high64; s = iv.sym; done = true;
                }
            } else {
                if (!lc_interval_apply(iv, F.delta, low64, high64)) {
                    // the symbol itself was decided with margin; only the exact bounds are missing (lcf_apply_symbol)
                    lcv_record_to_lanes(F, q0, q1, q2, q3);
                    lcf_exact_at(F, iv.sym, iv);
                    lc_interval_apply(iv, F.delta, low64, high64);
                }
                nlo = (uint32_t)low64; nhi = (uint32_t)high64; s = iv.sym; done = true;
            }
        }
    } else {
        if (!OUTLINE) lcv_load_pool(F, V, gw);
    }
    if (!done) { // exact evaluation shared with the other kernels
        fallback = 1;
        if (OUTLINE) {
            LcFast Fc = F; // only the copy's address is taken: F itself stays in registers
            const LcvCold r = lcv_cold_symbol(&Fc, V.pool, st, s1, gw, q0, q1, q2, q3, lo, hi, code);
            if (r.status != LC_OK) return r.status;
            on_candidate(r.s);
            nlo = r.nlo; nhi = r.nhi; s = r.s;
        } else {
            if (st == 1) lcf_state_first(F, s1);
            if (st == 2) lcv_record_to_lanes(F, q0, q1, q2, q3);
            LcInterval iv;
            double num, rdv;
            const int fs = lcf_find_symbol(F, st == 3 ? 2 : st, s1, lo, hi, code, iv, num, rdv) & 0xff;
            if (fs != LC_OK) return fs;
            on_candidate(iv.sym);
            lcf_apply_symbol(F, iv, num, rdv, lo, hi);
            nlo = lo; nhi = hi; s = iv.sym;
        }
    }
    return LC_OK;
}

// issue the global loads the next visit of context `key` needs (state st): the 4-byte word (states 1, 3) or the
// inline record (state 2).  Only called when no job on that context can still be running.
#define LCV_PREFETCH(st_, key_, gw_, q0_, q1_, q2_, q3_)                                   \
    do {                                                                                    \
        if ((st_) == 2) {                                                                   \
            const double2 *q_ = (const double2 *)(V.grec + (size_t)(key_) * 64);            \
            q0_ = __ldcg(q_); q1_ = __ldcg(q_ + 1); q2_ = __ldcg(q_ + 2); q3_ = __ldcg(q_ + 3); \
        } else if ((st_) != 0) gw_ = __ldcg(V.gword + (key_));                              \
    } while (0)

// decoder v2: the same request, the inline record going to the staging slot in shared memory instead of registers
__device__ __forceinline__ void lcv_prefetch_staged(const LcV2 &V, int lane, int st, uint32_t key, uint32_t &gw, lcv_sa slot)
{
    if (st == 2) lcv_stage_copy(slot, V.grec + (size_t)key * 64, lane);
    else if (st != 0) gw = __ldcg(V.gword + key);
}

// write decoded symbols [first, first+count) of the row held in shared memory (and their dequantised values)
__device__ __forceinline__ void lcv_flush_row(const unsigned char *row, int first_col, int count, LcIdxOut out,
                                              const float *deq_table, float *deq_out, int lane)
{
    for (int i = lane; i < count; i += 32) {
        const int sy = (int)row[first_col + i];
        out.store(i, sy);
        if (deq_out) deq_out[i] = __ldg(deq_table + sy);
    }
}

// Compile-time switches of the decoder warp's loop skeleton (each measured on its own, tools/dec_variants.py;
// profiles/r02_decoder_variants.txt):
//   LCV_OPT_SPEC     request the next context's word AND record as soon as a candidate symbol exists, whatever the
//                    context's state turns out to be (a wasted 68-byte L2 read for fresh contexts; the record of a
//                    context seen twice or more arrives ~300 cycles earlier than when it is requested after the state
//                    lookup -- the decoder warp waited 188 cycles per such symbol for it)
//   LCV_OPT_PRED     the one-lane stores of the bookkeeping (row entry, state bits, context word, job post) as
//                    predicated instructions instead of branch regions
//   LCV_OPT_ROWLOOP  one loop per row inside a loop over rows: the row-end work leaves the per-symbol path
#ifndef LCV_OPT_SPEC
#define LCV_OPT_SPEC 0
#endif
#ifndef LCV_OPT_PRED
#define LCV_OPT_PRED 0
#endif
#ifndef LCV_OPT_ROWLOOP
#define LCV_OPT_ROWLOOP 1
#endif
//   LCV_OPT_UNROLL   symbols per trip of the row loop
#ifndef LCV_OPT_UNROLL
#define LCV_OPT_UNROLL 1
#endif
//   LCV_OPT_COLD     rarely executed parts of the loop (renormalisation by more than 32 bits, the wait for a pending
//                    job) as real function calls: the loop body is ~25 KB of SASS against a 6 KB L0 / 32 KB L1.5
//                    instruction cache, and every taken branch to a line that is not resident costs tens of cycles
#ifndef LCV_OPT_COLD
#define LCV_OPT_COLD 0
#endif
#define LCV_PRAGMA(x) _Pragma(#x)
#define LCV_UNROLL(n) LCV_PRAGMA(unroll n)

// lcv_post with predicated stores: posts the job when `doit` is set (warp-uniform), otherwise only passes through
__device__ __forceinline__ void lcv_post_if(const LcV2 &V, LcvPost &P, int lane, bool doit, uint32_t key, uint32_t pay)
{
    const uint32_t j = P.njobs, slot = j & (LCV_RING - 1);
    if (doit && j - P.done_seen >= (uint32_t)LCV_RING) { // (rare: the updater is ahead nearly always)
        for (;;) {
            P.done_seen = lcv_sa_ld32_acq(V.sa_ring_done);
            if (j - P.done_seen < (uint32_t)LCV_RING) break;
            LCV_SPIN();
        }
    }
    const uint32_t p0 = (doit && lane == 0) ? 1u : 0u;
    lcv_sa_st32_if(p0, V.sa_ring_key + 4u * slot, key);
    lcv_sa_st32_if(p0, V.sa_ring_pay + 4u * slot, pay);
    lcv_sa_bar_arrive_if(p0, V.sa_ring_bar + 8u * slot);
    const bool mine = doit && (uint32_t)lane == slot;
    P.my_key = mine ? key : P.my_key;
    P.my_job = mine ? j : P.my_job;
    P.njobs = j + (doit ? 1u : 0u);
}

// renormalisation by more than 32 bits (d + e > 32: a nearly empty range; only on streams about to fault)
static __device__ __noinline__ uint32_t lcv_cold_renorm(LcvBits *b, uint32_t code, int d, int e, uint32_t em)
{
    const uint32_t b1 = lcv_br_take(*b, d);
    code = __funnelshift_lc(0u, code, d) | b1;
    const uint32_t b2 = lcv_br_take(*b, e);
    return ((code << e) | b2) ^ em;
}
// the wait for a job on the next context that is still running, then that context's state
static __device__ __noinline__ int lcv_cold_pend(uint32_t sa_ring_done, uint32_t sa_bits_word, uint32_t shift2, bool mine,
                                                 uint32_t my_job)
{
    if (mine) while ((int)(lcv_sa_ld32(sa_ring_done) - (my_job + 1u)) < 0) LCV_SPIN();
    __syncwarp();
    LCV_FENCE();
    return (int)((lcv_sa_ld32(sa_bits_word) >> shift2) & 3u);
}

template <bool OUTLINE>
__device__ __forceinline__ void lcv_decode_stream(LcFast &F, const LcV2 &V, LcvPost &P, const unsigned char *src,
                                                  long long nbytes, LcIdxOut out, const float *deq_table, float *deq_out,
                                                  int *status_out, int *fault_index)
{
    const int lane = F.lane, n = F.n, C = F.C;
    LcvBits br; lcv_br_init(br, src, nbytes);
    uint32_t lo = 0u, hi = 0xffffffffu;
    uint32_t code = lcv_br_take(br, 32); // start_decoding (:247-258)
    int status = LC_OK;
    int pos = 0, r = 0, c = 0;
    lcv_sa row_cur = V.sa_rows, row_prev = V.sa_rows + (uint32_t)C; // the row being decoded and the one above it
    bool up_ok = false, next_ok = F.R > 1; // a row above this row / above the next row (same image)
    uint32_t key = 0u; // (left=-1, up=-1)
    int st = 0;        // its state; the data the state needs is requested one symbol ahead:
    uint32_t gw = 0u;  //   states 1, 3: the context's 4-byte word
    //   state 2: the inline record  u | val[6] | sym[6] k, copied to staging slot (pos & 1) in shared memory
    P.my_key = LCV_SENTINEL; // contexts of the previous stream are not this stream's
    LcvRinv rinv; rinv.d = 2.3283064365386963e-10; rinv.f = 2.3283064365386963e-10f; // 1/2^32: the initial range
    LCP_DECL
    LCP_INIT();
    // one symbol.  Returns false on a fault (status set).
    auto one_symbol = [&](const bool last) -> bool {
        LCP_START();
        LCP_ROW(st); LCP_COUNT(st, 0);
        const uint32_t shift = (key & 15u) * 2u;
        // next position and the symbol above it (written at least C-1 >= 3 symbols ago): requested early.  At the
        // end of a row the next position is column 0 of the next row, under column 0 of this one.
        const bool has_up2 = last ? next_ok : up_ok;
        const int up2 = has_up2 ? lcv_sa_ld8(last ? row_cur : row_prev + (uint32_t)(c + 1)) : -1;
        // ---- decode_symbol (:272-292).  The next position's context key needs only the symbol: as soon as a path
        // has its candidate, the state word of that context is requested from shared memory, so the load overlaps
        // the bounds arithmetic.
        int s = 0, s1 = 0, fell_back = 0;
        uint32_t nlo = 0u, nhi = 0u, key2 = 0u, w2 = 0u, spec_key = LCV_SENTINEL, gw2 = 0u;
        double2 q0 = {0.0, 0.0}, q1 = q0, q2 = q0, q3 = q0;
        if (st == 2) { // the record requested during the previous symbol
            const lcv_sa slot = V.sa_stage + 64u * (uint32_t)(pos & 1);
            lcv_stage_wait();
            __syncwarp();
            q0 = lcv_sa_ld128(slot); q1 = lcv_sa_ld128(slot + 16u); q2 = lcv_sa_ld128(slot + 32u); q3 = lcv_sa_ld128(slot + 48u);
        }
        {
            const int fs = lcv_decode_symbol<OUTLINE>(F, V, st, gw, q0, q1, q2, q3, lo, hi, code, s, s1, nlo, nhi, fell_back,
                                             [&](int sym_) {
                                                 int up_ = up2;
                                                 LCV_USE_AFTER(up_, sym_); // keeps the consumer of the early load here
                                                 key2 = (uint32_t)((last ? -1 : sym_) + 1) * (uint32_t)(n + 1) + (uint32_t)(up_ + 1);
                                                 w2 = lcv_sa_ld32(V.sa_bits + 4u * (key2 >> 4));
                                                 if (LCV_OPT_SPEC && spec_key == LCV_SENTINEL) {
                                                     // speculative: whatever the state is (valid unless a job on this
                                                     // context is still running -- then it is requested again below)
                                                     spec_key = key2;
                                                     lcv_stage_copy(V.sa_stage + 64u * (uint32_t)((pos + 1) & 1), V.grec + (size_t)key2 * 64, lane);
                                                     gw2 = __ldcg(V.gword + key2);
                                                 }
                                             }, rinv);
            if (fs != LC_OK) { status = fs; return false; }
            if (fell_back) LCP_COUNT(4, st);
        }
        lo = nlo; hi = nhi;
        if (LCV_OPT_RINV) { // reciprocal of the new range, before renormalisation scales it by 2^t
            const uint32_t w = nhi - nlo; // range - 1
            rinv.d = lc_rcp_fast(lc_ll2d_small((long long)w + 1));
            rinv.f = lcv_rcp_f32((float)w);
        }
        LCP_MARK(1);
        // ---- next position's context (get_context :78-117): the data its state needs is requested now (loaded
        // straight into the registers the next iteration reads) and arrives during renormalisation and write-back
        if (LCV_OPT_PRED) lcv_sa_st8_if(lane == 0 ? 1u : 0u, row_cur + (uint32_t)c, s);
        else if (lane == 0) lcv_sa_st8(row_cur + (uint32_t)c, s);
        const uint32_t shift2 = (key2 & 15u) * 2u;
        int st2 = (int)((w2 >> shift2) & 3u);
        bool pend2 = key2 == key; // this symbol's own update of the same context comes first
        const bool spec_ok = LCV_OPT_SPEC && spec_key == key2; // (a fallback may have changed the symbol)
        if (LCV_OPT_SPEC && !spec_ok) lcv_stage_wait(); // the slot must not receive two copies at once
        if (st2 != 0 && !pend2) {
            // a job on that context among the last LCV_RING posted ones may still be running
            pend2 = __ballot_sync(LC_FULL_MASK, P.my_key == key2) != 0u;
            if (!pend2) {
                if (spec_ok) gw = gw2;
                else lcv_prefetch_staged(V, lane, st2, key2, gw, V.sa_stage + 64u * (uint32_t)((pos + 1) & 1));
            }
        }
        // ---- renormalise (:295-303) and underflow (:306-309): closed form, the d+e new bits come straight from
        // the top of the bit window
        {
            const int d = __clz((int)(lo ^ hi)); // leading bits low and high share
            const uint32_t lo_d = __funnelshift_lc(0u, lo, d), hi_d = __funnelshift_lc(0xffffffffu, hi, d);
            const int e = __clz((int)~((lo_d & ~hi_d) << 1)); // underflow steps: low = 01.., high = 10..
            const int t = d + e;
            const uint32_t em = e ? 0x80000000u : 0u;
            if (t <= 32) {
                code = __funnelshift_lc((uint32_t)(br.win >> 32), code, t) ^ em;
                lcv_br_skip(br, t);
            } else if (LCV_OPT_COLD) {
                LcvBits tmp = br; // only the copy's address is taken: br itself stays in registers
                code = lcv_cold_renorm(&tmp, code, d, e, em);
                br = tmp;
            } else {
                const uint32_t b1 = lcv_br_take(br, d);
                code = __funnelshift_lc(0u, code, d) | b1;
                const uint32_t b2 = lcv_br_take(br, e);
                code = ((code << e) | b2) ^ em;
            }
            lo = __funnelshift_lc(0u, lo_d, e) & ~em;
            hi = __funnelshift_lc(0xffffffffu, hi_d, e) | em;
            if (LCV_OPT_RINV) {
                if (t <= 32 && hi - lo >= 0xffffu) { rinv.d = lcv_scale_down(rinv.d, t); rinv.f = lcv_scale_down_f(rinv.f, t); }
                else { // (only on streams that are about to fault: the fast paths reject such a range anyway)
                    rinv.d = lc_rcp_fast(lc_ll2d_small((long long)(hi - lo) + 1)); rinv.f = lcv_rcp_f32((float)(hi - lo));
                }
            }
        }
        LCP_MARK(2);
        // ---- this context's model moves on
        if (LCV_OPT_PRED) {
            const uint32_t p0 = (st == 0 && lane == 0) ? 1u : 0u;
            lcv_stcg32_if(p0, V.gword + key, (uint32_t)s);
            lcv_sa_or32_if(p0, V.sa_bits + 4u * (key >> 4), 1u << shift);
            lcv_post_if(V, P, lane, st != 0, key, LCV_PAY(s, st, s1));
        } else if (st == 0) {
            if (lane == 0) { __stcg(V.gword + key, (uint32_t)s); lcv_sa_or32(V.sa_bits + 4u * (key >> 4), 1u << shift); }
        } else {
            lcv_post(V, P, lane, key, LCV_PAY(s, st, s1));
        }
        __syncwarp(); // lane 0's writes (row, word, state bits) are ordered before the other lanes' next reads
        if (pend2) {
            LCP_COUNT(6, st2);
            const bool mine = P.my_key == key2;
            if (LCV_OPT_COLD) st2 = lcv_cold_pend(V.sa_ring_done, V.sa_bits + 4u * (key2 >> 4), shift2, mine, P.my_job);
            else {
                if (mine) while ((int)(lcv_sa_ld32(V.sa_ring_done) - (P.my_job + 1u)) < 0) LCV_SPIN();
                __syncwarp();
                LCV_FENCE();
                st2 = (int)((lcv_sa_ld32(V.sa_bits + 4u * (key2 >> 4)) >> shift2) & 3u);
            }
            if (LCV_OPT_SPEC) lcv_stage_wait(); // (the speculative copy may still be landing in the slot)
            lcv_prefetch_staged(V, lane, st2, key2, gw, V.sa_stage + 64u * (uint32_t)((pos + 1) & 1));
        }
        key = key2; st = st2;
        LCP_MARK(3);
        return true;
    };
    // the end of a row: write it out, see whether the updater gave up, swap the row buffers
    auto row_done = [&]() -> bool {
        lcv_flush_row(V.rows + (row_cur - V.sa_rows), 0, C, out + (pos + 1 - C), deq_table,
                      deq_out ? deq_out + (pos + 1 - C) : (float *)0, lane);
        const uint32_t ab = lcv_sa_ld32(V.sa_abort);
        if (ab) { status = (int)ab; return false; }
        const lcv_sa t_ = row_cur; row_cur = row_prev; row_prev = t_;
        r = r + 1 == F.R ? 0 : r + 1; // (the next image of the stream starts without a row above)
        up_ok = r > 0; next_ok = r + 1 != F.R;
        return true;
    };
    bool done_row = false; // the fault position's row is partly decoded unless the fault came at a row end
    if (LCV_OPT_ROWLOOP) {
        bool ok = true;
        while (ok && pos < F.total) {
            LCV_UNROLL(LCV_OPT_UNROLL)
            for (c = 0; c < C - 1; c++, pos++)
                if (!one_symbol(false)) { ok = false; break; }
            if (!ok) break;
            if (!one_symbol(true)) break;      // (c == C - 1)
            ok = row_done();
            pos++; c = 0;
            if (!ok) { done_row = true; break; }
        }
    } else {
        for (; pos < F.total; pos++) {
            const bool last = c + 1 == C;
            if (!one_symbol(last)) break;
            if (last) {
                const bool ok = row_done();
                c = 0;
                if (!ok) { pos++; done_row = true; break; }
            } else c++;
        }
    }
    LCP_FLUSH();
    // release the updaters
    for (int u = 0; u < LCV_NU; u++) lcv_post(V, P, lane, LCV_SENTINEL, 0u);
    *fault_index = pos;
    *status_out = status;
    __syncwarp();
    {
        // symbols of the unfinished row (c of them; none when the stream ended on a row boundary), zeros after a fault
        const int done = pos;
        const int part = done_row ? 0 : c;
        if (part > 0) lcv_flush_row(V.rows + (row_cur - V.sa_rows), 0, part, out + (done - part), deq_table,
                                    deq_out ? deq_out + (done - part) : (float *)0, lane);
        for (int z = done + lane; z < F.total; z += 32) { out.store(z, 0); if (deq_out) deq_out[z] = 0.0f; }
    }
}

// LCV_OPT_FLAT: the decoder warp runs the flat loop of lc_decoder_v2_flat.cuh (one dispatch, a branch-free tail, one
// rarely taken region per symbol) instead of lcv_decode_stream above
#ifndef LCV_OPT_FLAT
#define LCV_OPT_FLAT 1
#endif
#include "lc_decoder_v2_flat.cuh"

// Block entry: LCV_WARPS warps, persistent over streams.  FN/FC/FR > 0 fix the alphabet size and the image shape at
// compile time (one image per stream): keys, shifts, margins and loop bounds become immediates on the serial chain.
template <int FN, int FC, int FR, bool OUTLINE = false>
__device__ __forceinline__ void lcv_decode_block(const LcCoderCfg &cfg, const LcV2Cfg &vc, const unsigned char *bytes,
                                                 const long long *offsets, const int *nbits, int B, LcIdxOut out,
                                                 const float *deq_table, float *deq_out, int *status, int *fault,
                                                 char *scratch, const double *tables, const char *t2, char *smem)
{
    // LCV_OPT_ROLE_SWAP: odd blocks run the decoder on their second warp.  Warps go to the SM's four schedulers by
    // their slot number, a block's two warps take neighbouring slots, so with the decoder always on warp 0 every
    // decoder warp of an SM lands on two of the four schedulers and they compete for issue slots with each other.
#ifndef LCV_OPT_ROLE_SWAP
#define LCV_OPT_ROLE_SWAP 0
#endif
    const int warp = (LCV_OPT_ROLE_SWAP && LCV_WARPS == 2) ? (int)((threadIdx.x >> 5) ^ (blockIdx.x & 1u)) : (int)(threadIdx.x >> 5);
    LcV2 V;
    V.sbits = (uint32_t *)(smem + vc.sm_bits);
    V.rows = (unsigned char *)(smem + vc.sm_rows);
    V.ring_bar = (unsigned long long *)(smem + vc.sm_ring);
    V.ring_key = (uint32_t *)(V.ring_bar + LCV_RING);
    V.ring_pay = V.ring_key + LCV_RING; V.ring_done = V.ring_pay + LCV_RING;
    V.u1tab = (double *)(smem + vc.sm_tab); V.ru1tab = V.u1tab + 32;
    V.pool_top = (uint32_t *)(smem + vc.sm_misc); V.abort_code = V.pool_top + 1;
    char *sc = scratch + (size_t)blockIdx.x * vc.g_stride;
    V.gword = (uint32_t *)(sc + vc.g_word); V.grec = sc + vc.g_rec; V.pool = sc + vc.g_pool;
    V.cum1 = tables + 64;
    V.t2 = t2;
    V.pool_bytes = vc.pool_bytes; V.eps_k = vc.eps_k; V.lg_n = vc.lg_n;
    lcv_view_sa(V);
    V.sa_stage = lcv_sa_of(smem + vc.sm_stage);
    LcFast F;
    F.n = cfg.n; F.C = cfg.C; F.R = cfg.R; F.total = cfg.total; F.lane = (int)(threadIdx.x & 31);
    F.rate = cfg.rate; F.delta = cfg.delta; F.u0 = LC_DDIV(1.0, (double)cfg.n);
    F.delta_v = cfg.delta + 1.5e-14; F.tmargin = (double)cfg.n * 1.5e-14;
    F.P1 = LC_DADD(F.u0, LC_DMUL(F.rate, LC_DSUB(1.0, F.u0)));
    F.slot_cap = 0; F.slot_shift = 0; F.pool_bytes = vc.pool_bytes; F.pool_top = 0;
    F.pw_len = cfg.pw_len; F.pw_steps = cfg.pw_steps; F.pw_chains = cfg.pw_chains;
    F.slots = (unsigned long long *)0; F.pool = V.pool;
    F.dense = (double *)(smem + vc.sm_dense) + (size_t)(warp > 0 ? warp - 1 : 0) * cfg.n;
    F.u1tab = V.u1tab; F.rows = (unsigned short *)0;
    F.k = 0; F.u = F.u0; F.my_sym = 0x7fffffff; F.my_val = 0.0;
    if (FN > 0) {
        F.n = FN; F.C = FC; F.R = FR; F.total = FR * FC;
        F.pw_len = FN < 128 ? FN : 128; F.pw_steps = F.pw_len / 8; F.pw_chains = 8 * (FN / F.pw_len);
        int lg = 0; while ((1 << lg) < FN) lg++;
        V.lg_n = (uint32_t)lg;
        V.eps_k = (uint32_t)(1e-10 * (double)FN * 1099511627776.0);
    }
    if (threadIdx.x < 64) V.u1tab[threadIdx.x] = tables[threadIdx.x];
    if (threadIdx.x < LCV_RING) { lcv_bar_init(V.ring_bar + threadIdx.x); V.ring_done[threadIdx.x] = 0u; }
    LcvPost P; P.njobs = 0u; P.done_seen = 0u; P.my_key = LCV_SENTINEL; P.my_job = 0u;
    uint32_t ujob = (uint32_t)(warp > 0 ? warp - 1 : 0);
    const uint32_t nwords = (vc.nkeys + 15u) / 16u;
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        __syncthreads(); // the previous stream is finished by every warp
        for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) V.sbits[i] = 0u;
        if (threadIdx.x == 0) { *V.pool_top = 0u; *V.abort_code = 0u; }
        __syncthreads();
        if (warp == 0) {
            int fi = 0, st = 0;
            const long long nby = ((long long)nbits[sidx] + 7) >> 3;
            if (LCV_OPT_FLAT)
                lcv_decode_stream_flat<OUTLINE>(F, V, P, bytes + offsets[sidx], nby, out + (size_t)sidx * cfg.total, deq_table,
                                                deq_out ? deq_out + (size_t)sidx * cfg.total : (float *)0, &st, &fi);
            else
                lcv_decode_stream<OUTLINE>(F, V, P, bytes + offsets[sidx], nby, out + (size_t)sidx * cfg.total, deq_table,
                                           deq_out ? deq_out + (size_t)sidx * cfg.total : (float *)0, &st, &fi);
            if (F.lane == 0) { status[sidx] = st; fault[sidx] = fi; }
        } else {
            lcv_updater(F, V, ujob);
        }
    }
}
