// Decoder v2, "flat" decoder-warp loop (round 2).  Same data structures, same updater warp, same results as
// lcv_decode_stream (lc_decoder_v2.cuh); only the shape of the decoder warp's per-symbol code differs.
//
// Why: the per-instruction profile of the round-1 loop (profiles/r02_decoder_regions.md) showed the decoder warp
// executing ~188 instructions per symbol of which 35 were control instructions (BRA / BSSY / BSYNC): 16 branches on
// the path of a fresh-context symbol.  The warp is alone on its dependent chain, so every conditional branch costs the
// predicate-to-branch latency (13-14 cycles) plus a refetch after the join (10-14 cycles at the first instruction behind
// every BSYNC, 47 behind the loop's back edge): ~550 of the 1390 cycles per symbol were control flow, not arithmetic.
// Here a symbol is
//   1. ONE dispatch on the context's state into a region that holds everything that state needs: the search, the
//      verification, the interval, and the bookkeeping of that state (first visit: context word + state bit; later
//      visits: the job for the updater warp) -- or nothing at all when its margins cannot decide;
//   2. a straight-line tail without any branch: row entry, next context's state, pending-job ballot, predicated
//      prefetch of what that state needs, renormalisation, bit-window refill with a predicated load;
//   3. ONE rarely taken region for everything unusual, decided by a single flag: renormalisation by more than 32 bits,
//      a refill that touches the end of the stream, a job ring that is full, a next context whose job may still be
//      running (3 % of the symbols).
// What a state's region cannot decide goes to the exact evaluation shared with the other kernels (3 % of the symbols).
#pragma once

// Compile-time switches of the flat loop (each measured on its own with tools/dec_variants.py):
//   LCVF_BRANCHY_PF      the pending-job ballot and the prefetch of the next context's data inside a branch region that
//                        fresh contexts (48 % of the symbols) skip, instead of ~20 predicated-off instructions
//   LCVF_BRANCHY_REFILL  the bit window's refill (one symbol in four) inside a branch region
//   LCVF_CLZ_FLOAT       leading-zero counts of the renormalisation from the exponent of a round-toward-zero
//                        int->float conversion (ALU pipe) instead of FLO (quarter-rate pipe, ~20 cycles each, two in a row)
#ifndef LCVF_BRANCHY_PF
#define LCVF_BRANCHY_PF 1
#endif
#ifndef LCVF_BRANCHY_REFILL
#define LCVF_BRANCHY_REFILL 1
#endif
#ifndef LCVF_CLZ_FLOAT
#define LCVF_CLZ_FLOAT 1
#endif
//   LCVF_RENORM2         underflow count from the unshifted low/high words (two dependent instructions fewer)
//   LCVF_NO_SYNCWARP     no __syncwarp() between lane 0's shared-memory stores and the other lanes' later loads: the
//                        warp is converged and its shared-memory instructions execute in program order
#ifndef LCVF_RENORM2
#define LCVF_RENORM2 1
#endif
#ifndef LCVF_NO_SYNCWARP
#define LCVF_NO_SYNCWARP 0
#endif

// 16-byte asynchronous copy global -> shared when p is set (one predicated instruction, no branch region), and
// predicated 4-byte loads INTO the register that carries the value to the next symbol: written as `if (p) r = load`
// the compiler loads into a temporary and copies it under the predicate right away, and that copy waits out the
// whole L2 latency on the decoder warp's path.
#ifdef LC_HOSTSIM
static inline void lcvf_stage_copy_if(bool p, lcv_sa dst, const char *src) { if (p) memcpy((char *)dst, src, 16); }
static inline void lcvf_ldcg32_if(bool p, uint32_t &r, const uint32_t *src) { if (p) r = *(const volatile uint32_t *)src; }
static inline void lcvf_ldg32_if(bool p, uint32_t &r, const uint32_t *src) { if (p) r = *src; }
static inline lcv_sa lcvf_opaque(lcv_sa a) { return a; }
#else
static __device__ __forceinline__ void lcvf_ldcg32_if(bool p, uint32_t &r, const uint32_t *src)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q ld.global.cg.u32 %0, [%2]; }" : "+r"(r) : "r"((uint32_t)p), "l"(src) : "memory");
}
static __device__ __forceinline__ void lcvf_ldg32_if(bool p, uint32_t &r, const uint32_t *src) // read-only data
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %1, 0; @q ld.global.nc.u32 %0, [%2]; }" : "+r"(r) : "r"((uint32_t)p), "l"(src));
}
// a shared-window address the compiler cannot re-derive from the kernel's shared base (it otherwise recomputes it --
// S2R CgaCtaId, LEA, two adds -- at every use instead of keeping one register)
static __device__ __forceinline__ lcv_sa lcvf_opaque(lcv_sa a) { asm volatile("" : "+r"(a)); return a; }
static __device__ __forceinline__ void lcvf_stage_copy_if(bool p, lcv_sa dst, const char *src)
{
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %0, 0; @q cp.async.cg.shared.global [%1], [%2], 16; }" ::"r"((uint32_t)p), "r"(dst), "l"(src) : "memory");
}
#endif

// ---- bit reader of the flat loop: LcvBits with the end-of-stream handling off the per-symbol path.  The word waiting
// in `nextw` is already masked (bytes past the end of the stream are zero: the reference reads zeros there, :260-270),
// and a word that is not entirely inside the stream is loaded by the rare path.
struct LcvBitsF {
    const uint32_t *w;
    uint32_t nbytes, safe_words; // words that lie entirely inside the stream
    unsigned long long win;      // next bits of the stream, MSB first
    int nwin;                    // valid bits in win (> 32 between symbols)
    uint32_t widx;               // index of the word held in nextw
    uint32_t nextw;              // that word as loaded (memory order)
};
__device__ __forceinline__ uint32_t lcvf_masked_raw(const LcvBitsF &b, uint32_t wi)
{
    const uint32_t byte0 = wi * 4u;
    if (byte0 >= b.nbytes) return 0u;
    uint32_t raw = __ldg(b.w + wi);
    if (byte0 + 4u > b.nbytes) raw &= 0xffffffffu >> (8u * (4u - (b.nbytes - byte0)));
    return raw;
}
__device__ __forceinline__ void lcvf_br_init(LcvBitsF &b, const unsigned char *src, long long nbytes)
{
    b.w = (const uint32_t *)src; b.nbytes = (uint32_t)(nbytes > 0x7fffffffll ? 0x7fffffffll : nbytes);
    b.safe_words = b.nbytes >> 2;
    b.win = ((unsigned long long)__byte_perm(lcvf_masked_raw(b, 0), 0, 0x0123) << 32) | __byte_perm(lcvf_masked_raw(b, 1), 0, 0x0123);
    b.nwin = 64; b.widx = 2; b.nextw = lcvf_masked_raw(b, 2);
}
// drop nb bits (0..32), branch-free.  Returns true when the refill needs a word the rare path has to load
// (lcvf_br_fix): the one straddling the end of the stream, or one past it.
__device__ __forceinline__ bool lcvf_br_skip(LcvBitsF &b, int nb)
{
    b.win <<= nb;
    b.nwin -= nb;
#if LCVF_BRANCHY_REFILL
    bool rare_ = false;
    if (b.nwin <= 32) {
        b.win |= (unsigned long long)__byte_perm(b.nextw, 0, 0x0123) << (32 - b.nwin);
        b.nwin += 32;
        b.widx++;
        const bool inb_ = b.widx < b.safe_words;
        lcvf_ldg32_if(inb_, b.nextw, b.w + b.widx);
        rare_ = !inb_;
    }
    return rare_;
#endif
    const bool need = b.nwin <= 32;
    const unsigned long long add = (unsigned long long)__byte_perm(b.nextw, 0, 0x0123) << ((32 - b.nwin) & 63);
    b.win |= need ? add : 0ull;
    b.nwin += need ? 32 : 0;
    b.widx += need ? 1u : 0u;
    const bool inb = b.widx < b.safe_words;
    lcvf_ldg32_if(need && inb, b.nextw, b.w + b.widx);
    return need && !inb;
}
__device__ __forceinline__ void lcvf_br_fix(LcvBitsF &b) { b.nextw = lcvf_masked_raw(b, b.widx); }
// next nb bits (1..32), MSB first (rare path only)
__device__ __forceinline__ uint32_t lcvf_br_take(LcvBitsF &b, int nb)
{
    const uint32_t v = nb > 0 ? (uint32_t)(b.win >> (64 - nb)) : 0u;
    if (lcvf_br_skip(b, nb)) lcvf_br_fix(b);
    return v;
}

// post a job whose ring slot is known to be free (the flat loop checks that at the end of the previous symbol)
__device__ __forceinline__ void lcvf_post(const LcV2 &V, LcvPost &P, int lane, uint32_t key, uint32_t pay)
{
    const uint32_t j = P.njobs, slot = j & (LCV_RING - 1);
    if (lane == 0) {
        lcv_sa_st32(V.sa_ring_key + 4u * slot, key); lcv_sa_st32(V.sa_ring_pay + 4u * slot, pay);
        lcv_sa_bar_arrive(V.sa_ring_bar + 8u * slot);
    }
    if ((uint32_t)lane == slot) { P.my_key = key; P.my_job = j; }
    P.njobs = j + 1u;
}
// wait until the ring has a free slot for the next job
__device__ __forceinline__ void lcvf_ring_room(const LcV2 &V, LcvPost &P)
{
    while (P.njobs - P.done_seen >= (uint32_t)LCV_RING) {
        P.done_seen = lcv_sa_ld32_acq(V.sa_ring_done);
        if (P.njobs - P.done_seen < (uint32_t)LCV_RING) break;
        LCV_SPIN();
    }
}

#ifdef LC_HOSTSIM
static inline int lcvf_clz(uint32_t x) { return x ? __clz((int)x) : 32; }
static inline uint32_t lcvf_d2u(double x) { return (uint32_t)(long long)x; }
#else
static __device__ __forceinline__ uint32_t lcvf_d2u(double x) { return __double2uint_rz(x); } // x in [0, 2^32)
static __device__ __forceinline__ int lcvf_clz(uint32_t x) // x == 0 gives a value above 32 (the caller's "rare" case)
{
#if LCVF_CLZ_FLOAT
    float f;
    asm("cvt.rz.f32.u32 %0, %1;" : "=f"(f) : "r"(x));
    return 158 - (int)(__float_as_uint(f) >> 23);
#else
    return __clz((int)x);
#endif
}
#endif

template <bool OUTLINE>
__device__ __forceinline__ void lcv_decode_stream_flat(LcFast &F, const LcV2 &V, LcvPost &P, const unsigned char *src,
                                                       long long nbytes, LcIdxOut out, const float *deq_table,
                                                       float *deq_out, int *status_out, int *fault_index)
{
    const int lane = F.lane, n = F.n, C = F.C;
    LcvBitsF br; lcvf_br_init(br, src, nbytes);
    uint32_t lo = 0u, hi = 0xffffffffu;
    uint32_t code = lcvf_br_take(br, 32); // start_decoding (:247-258)
    int status = LC_OK;
    int pos = 0, r = 0, c = 0;
    lcv_sa row_cur = V.sa_rows, row_prev = V.sa_rows + (uint32_t)C; // the row being decoded and the one above it
    bool up_ok = false, next_ok = F.R > 1; // a row above this row / above the next row (same image)
    uint32_t key = 0u; // (left=-1, up=-1)
    int st = 0;        // its state; the data the state needs is requested one symbol ahead:
    uint32_t gw = 0u;  //   states 1, 3: the context's 4-byte word
    //   state 2: the inline record  u | val[6] | sym[6] k, copied to staging slot (pos & 1) in shared memory
    lcv_sa kw_addr = V.sa_bits; // address of the shared-memory word holding the context's state bits, and the
    uint32_t kbit = 1u;         // low bit of its 2-bit field (both computed when the context was looked up)
    const lcv_sa sa_stage = lcvf_opaque(V.sa_stage);
    P.my_key = LCV_SENTINEL; // contexts of the previous stream are not this stream's
    lcvf_ring_room(V, P);    // (the previous stream's release job may have filled the ring)
    const double cfix = 1e-10;
    LCP_DECL
    LCP_INIT();
    // one symbol.  Returns false on a fault (status set).
    auto one_symbol = [&](const bool last) -> bool {
        LCP_START();
        LCP_ROW(st); LCP_COUNT(st, 0);
        // next position and the symbol above it (written at least C-1 >= 3 symbols ago): requested early.  At the
        // end of a row the next position is column 0 of the next row, under column 0 of this one.
        const bool has_up2 = last ? next_ok : up_ok;
        const int up2 = has_up2 ? lcv_sa_ld8(last ? row_cur : row_prev + (uint32_t)(c + 1)) : -1;
        int s = 0, s1 = 0;
        uint32_t nlo = 0u, nhi = 0u, key2 = 0u, w2 = 0u;
        lcv_sa w2_addr = 0;
        bool done = false;
        // the next position's context key needs only the symbol: as soon as a path has its candidate, the state word
        // of that context is requested from shared memory, so the load overlaps the bounds arithmetic
        auto candidate = [&](int sym_) {
            int up_ = up2;
            LCV_USE_AFTER(up_, sym_); // keeps the consumer of the early load here
            key2 = (uint32_t)((last ? -1 : sym_) + 1) * (uint32_t)(n + 1) + (uint32_t)(up_ + 1);
            w2_addr = V.sa_bits + 4u * (key2 >> 4);
            w2 = lcv_sa_ld32(w2_addr);
        };
        double2 q0 = {0.0, 0.0}, q1 = q0, q2 = q0, q3 = q0;
        // ---- decode_symbol (:272-292): one region per context state
        const uint32_t rng1 = hi - lo, off = code - lo; // range-1, (code-low+1)-1
        const bool pre_ok = hi >= lo && off <= rng1 && rng1 >= 0xffffu;
        if (__builtin_expect(st == 0, 1)) {
            // uniform model: cum[i] = i/n exactly, so range*cum is exact and everything is integer work.
            // With a = (code-low+1)*n and E = 1e-10*n*range, the symbol is the s with
            // s*range < a - E (+- 3e-4) <= (s+1)*range; candidate from a float quotient, checked with margins.
            int cand = (int)((float)off * lcv_rcp_f32((float)rng1) * (float)n);
            cand = cand > n - 1 ? n - 1 : cand; // (float)off rounds up to 2^32 at most: cand <= n
            candidate(cand);
            const unsigned long long below = (unsigned long long)(uint32_t)cand * rng1 + (uint32_t)cand; // cand*range
            const unsigned long long above = below + rng1 + 1ull;
            const unsigned long long a = ((unsigned long long)off + 1ull) << V.lg_n;
            const uint32_t e_lo = __umulhi(rng1, V.eps_k) >> 8; // <= E < e_lo + 3
            // below + e_lo + 4 <= a  and  a + 1 <= above + e_lo, i.e. e_lo + 4 <= a - below <= range + e_lo - 1, in 32
            // bits (a - below >= 2^32 would need the full 2^32 range and a top symbol: left to the exact path)
            const unsigned long long dlt = a - below;
            const bool ok = pre_ok & ((uint32_t)(dlt >> 32) == 0u) & ((uint32_t)dlt >= e_lo + 4u) & ((uint32_t)dlt - (e_lo + 4u) <= rng1 - 4u);
            if (ok) {
                s = cand; done = true;
                nlo = lo + (uint32_t)(below >> V.lg_n);
                nhi = lo + (uint32_t)(above >> V.lg_n) - 1u;
                // first visit: the context word holds the symbol, the state goes 0 -> 1
                // (every lane stores the same word: one transaction, and no branch region around a one-lane store)
                __stcg(V.gword + key, (uint32_t)s);
                lcv_sa_or32_if(lane == 0 ? 1u : 0u, kw_addr, kbit);
            }
        } else { if (st == 1) {
            s1 = (int)(gw & 0x3FFu);
            if (s1 >= n) s1 = n - 1; // only on a stream already flagged for the generic kernel
            // model after one update: exact np.cumsum values from the per-launch table
            const double rd = lc_ll2d_small((long long)rng1 + 1), nd = lc_ll2d_small((long long)off + 1);
            const int t = lcf_tab_index(F, s1);
            const double u = lcv_sa_ldf64(V.sa_tab + 8u * (uint32_t)t), ru = lcv_sa_ldf64(V.sa_tab + 256u + 8u * (uint32_t)t);
            const double va = nd * lc_rcp_fast(rd) - cfix;
            const double A0 = (double)s1 * u, B0 = A0 + F.P1;
            int sc = va < A0 ? (int)(va * ru) : (va <= B0 ? s1 : s1 + 1 + (int)((va - B0) * ru));
            sc = sc < 0 ? 0 : (sc > n - 1 ? n - 1 : sc);
            candidate(sc);
            const double *row = V.cum1 + (size_t)s1 * (n + 1);
            const double clo = __ldg(row + sc), chi = __ldg(row + sc + 1);
            const double xl = LC_DMUL(rd, clo), xh1 = LC_DMUL(rd, chi);
            const double tgt = nd - cfix * rd; // ~ v*range; |error| < 3e-6 for range <= 2^32
            if (pre_ok && tgt - xl > 1e-5 && xh1 - tgt >= 1e-5) { // cum[sc] < v <= cum[sc+1], decided with margin
                s = sc; done = true;
                nlo = lo + lcvf_d2u(xl); // (both in [0, 2^32) here: 32-bit conversions, 19 cycles against 28)
                nhi = lo + lcvf_d2u(LC_DSUB(xh1, 1.0));
                lcvf_post(V, P, lane, key, LCV_PAY(s, 1, s1));
            }
        } else if (st == 2) {
            { // the record requested during the previous symbol
                const lcv_sa slot = sa_stage + 64u * (uint32_t)(pos & 1);
                lcv_stage_wait();
                __syncwarp();
                q0 = lcv_sa_ld128(slot); q1 = lcv_sa_ld128(slot + 16u); q2 = lcv_sa_ld128(slot + 32u); q3 = lcv_sa_ld128(slot + 48u);
            }
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
                const double va = nd * lc_rcp_fast(rd) - cfix;
                // entries in ascending symbol order; approximate cum before (A) and after (B) each: stop at the first entry
                // whose upper bound reaches v.  Sequential with early exit -- a record holds 2-3 entries on average.
                double S = 0.0, Al = 0.0, Bl = 0.0, Bp = 0.0; // Bp/gf: end of the previous entry = start of the gap before l
                int sl = 0, gf = 0;
                bool found = false;
#define LCVF_SCAN_STEP(j_, sym_expr_, val_)                                             \
                if (k <= (j_)) break;                                                    \
                {                                                                        \
                    const int sy_ = (int)(sym_expr_);                                    \
                    const double A_ = (double)(sy_ - (j_)) * u + S;                      \
                    const double B_ = A_ + (val_);                                       \
                    if (B_ >= va) { found = true; Al = A_; Bl = B_; sl = sy_; break; }   \
                    Bp = B_; gf = sy_ + 1; S += (val_);                                  \
                }
                do {
                    LCVF_SCAN_STEP(0, sb_lo & 0xffu, q0.y)
                    LCVF_SCAN_STEP(1, (sb_lo >> 8) & 0xffu, q1.x)
                    LCVF_SCAN_STEP(2, (sb_lo >> 16) & 0xffu, q1.y)
                    LCVF_SCAN_STEP(3, sb_lo >> 24, q2.x)
                    LCVF_SCAN_STEP(4, sb_hi & 0xffu, q2.y)
                    LCVF_SCAN_STEP(5, (sb_hi >> 8) & 0xffu, q3.x)
                } while (0);
#undef LCVF_SCAN_STEP
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
                candidate(iv.sym);
                long long low64 = lo, high64 = hi;
                // when the truncations are not stable under the bounds' error (4 % of these symbols) the exact
                // evaluation below redoes the symbol: same result, exact sums
                if (lc_interval_apply(iv, F.delta, low64, high64)) {
                    nlo = (uint32_t)low64; nhi = (uint32_t)high64; s = iv.sym; done = true;
                    lcvf_post(V, P, lane, key, LCV_PAY(s, 2, 0));
                }
            }
        } }
        if (!done) { // exact evaluation shared with the other kernels, and the bookkeeping of whatever state it was
            LCP_COUNT(4, st);
            if (OUTLINE) {
                LcFast Fc = F; // only the copy's address is taken: F itself stays in registers
                const LcvCold rc = lcv_cold_symbol(&Fc, V.pool, st, s1, gw, q0, q1, q2, q3, lo, hi, code);
                if (rc.status != LC_OK) { status = rc.status; return false; }
                nlo = rc.nlo; nhi = rc.nhi; s = rc.s;
            } else {
                if (st == 1) lcf_state_first(F, s1);
                if (st == 2) lcv_record_to_lanes(F, q0, q1, q2, q3);
                if (st == 3) lcv_load_pool(F, V, gw);
                LcInterval iv;
                double num, rdv;
                const int fs = lcf_find_symbol(F, st == 3 ? 2 : st, s1, lo, hi, code, iv, num, rdv) & 0xff;
                if (fs != LC_OK) { status = fs; return false; }
                uint32_t xlo = lo, xhi = hi;
                lcf_apply_symbol(F, iv, num, rdv, xlo, xhi);
                nlo = xlo; nhi = xhi; s = iv.sym;
            }
            candidate(s);
            if (st == 0) {
                if (lane == 0) { __stcg(V.gword + key, (uint32_t)s); lcv_sa_or32(kw_addr, kbit); }
            } else {
                lcvf_post(V, P, lane, key, LCV_PAY(s, st, s1));
            }
        }
        LCP_MARK(1);
        // ---- straight-line tail.  This position's row entry; the next position's context (get_context :78-117):
        // its state, whether a job on it may still be running (one of the last LCV_RING posted ones), and -- when not
        // -- the request for the data that state needs, which arrives during renormalisation
        if (lane == 0) lcv_sa_st8(row_cur + (uint32_t)c, s);
        const uint32_t shift2 = (key2 & 15u) * 2u;
        int st2 = (int)((w2 >> shift2) & 3u);
#if LCVF_BRANCHY_PF
        bool pend2 = key2 == key;
        if (st2 != 0 && !pend2) {
            pend2 = __ballot_sync(LC_FULL_MASK, P.my_key == key2) != 0u;
            const lcv_sa slot2 = sa_stage + 64u * (uint32_t)((pos + 1) & 1);
            lcvf_ldcg32_if(!pend2 & (st2 != 2), gw, V.gword + key2);
            lcvf_stage_copy_if(!pend2 & (st2 == 2) & (lane < 4), slot2 + 16u * (uint32_t)lane, V.grec + (size_t)key2 * 64 + 16 * lane);
        }
#else
        const bool pend2 = (__ballot_sync(LC_FULL_MASK, P.my_key == key2) != 0u) || key2 == key;
        {
            const bool pf = !pend2 & (st2 != 0);
            const lcv_sa slot2 = sa_stage + 64u * (uint32_t)((pos + 1) & 1);
            lcvf_ldcg32_if(pf & (st2 != 2), gw, V.gword + key2);
            lcvf_stage_copy_if(pf & (st2 == 2) & (lane < 4), slot2 + 16u * (uint32_t)lane, V.grec + (size_t)key2 * 64 + 16 * lane);
        }
#endif
        // ---- renormalise (:295-303) and underflow (:306-309): closed form, the d+e new bits come straight from
        // the top of the bit window
        const int d = lcvf_clz(nlo ^ nhi); // leading bits low and high share
#if LCVF_RENORM2
        // low = 01.., high = 10.. after those d bits and the one where they part: the underflow steps are the run of
        // (low = 1, high = 0) positions that follows, counted on the unshifted words (nothing waits for low << d)
        const uint32_t uf = nlo & ~nhi;
        const int e = lcvf_clz(~__funnelshift_lc(0u, uf, d + 1));
#else
        const uint32_t lo_d = __funnelshift_lc(0u, nlo, d), hi_d = __funnelshift_lc(0xffffffffu, nhi, d);
        const int e = lcvf_clz(~((lo_d & ~hi_d) << 1)); // underflow steps: low = 01.., high = 10..
#endif
        const int t = d + e;
        const uint32_t em = e ? 0x80000000u : 0u;
        const bool big = t > 32; // (a nearly empty range: only on streams about to fault)
        const int tt = big ? 0 : t;
        const uint32_t code_old = code;
        code = __funnelshift_lc((uint32_t)(br.win >> 32), code, tt) ^ em;
#if LCVF_RENORM2
        lo = __funnelshift_lc(0u, nlo, t) & ~em; // (shift counts above 32 clamp: zeros / ones, like the two steps)
        hi = __funnelshift_lc(0xffffffffu, nhi, t) | em;
#else
        lo = __funnelshift_lc(0u, lo_d, e) & ~em;
        hi = __funnelshift_lc(0xffffffffu, hi_d, e) | em;
#endif
        const bool refill_rare = lcvf_br_skip(br, tt);
        const bool rare = refill_rare | big | pend2 | (P.njobs - P.done_seen >= (uint32_t)LCV_RING);
        LCP_MARK(2);
#if !LCVF_NO_SYNCWARP || defined(LC_HOSTSIM)
        __syncwarp(); // lane 0's writes (row, word, state bits) are ordered before the other lanes' next reads
#endif
        if (rare) {
            if (refill_rare) lcvf_br_fix(br);
            if (big) { // renormalisation by more than 32 bits: the two steps one after the other
                const int dx = __clz((int)(nlo ^ nhi)); // (exact counts: 0..32)
                const int ex = __clz((int)~__funnelshift_lc(0u, nlo & ~nhi, dx + 1));
                const uint32_t b1 = lcvf_br_take(br, dx);
                uint32_t cv = __funnelshift_lc(0u, code_old, dx) | b1;
                const uint32_t b2 = lcvf_br_take(br, ex);
                code = (__funnelshift_lc(0u, cv, ex) | b2) ^ em;
            }
            lcvf_ring_room(V, P);
            if (pend2) {
                LCP_COUNT(6, st2);
                const bool mine = P.my_key == key2;
                if (mine) while ((int)(lcv_sa_ld32(V.sa_ring_done) - (P.my_job + 1u)) < 0) LCV_SPIN();
                __syncwarp();
                LCV_FENCE();
                st2 = (int)((lcv_sa_ld32(w2_addr) >> shift2) & 3u);
                lcv_prefetch_staged(V, lane, st2, key2, gw, sa_stage + 64u * (uint32_t)((pos + 1) & 1));
            }
        }
        key = key2; st = st2; kw_addr = w2_addr; kbit = 1u << shift2;
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
    {
        bool ok = true;
        while (ok && pos < F.total) {
            for (c = 0; c < C - 1; c++, pos++)
                if (!one_symbol(false)) { ok = false; break; }
            if (!ok) break;
            if (!one_symbol(true)) break;      // (c == C - 1)
            ok = row_done();
            pos++; c = 0;
            if (!ok) { done_row = true; break; }
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
