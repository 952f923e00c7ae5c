// Decoder v3: decoder v2 (lc_decoder_v2.cuh) with the serial warp split in two.
//
// v2's decoder warp executed ~250 in-order instructions per symbol, and the profile (stall sampling per SASS line,
// profiles/r01_ncu_all_kernels_v6.md) shows them waiting on one another rather than on memory: the chain
//   symbol -> next context key -> state bits -> recent-job check -> loads of that context's word/record
// and the chain
//   symbol -> interval -> renormalisation -> next symbol's search
// are independent after the symbol is known, but one in-order warp runs them back to back together with the
// bookkeeping (row buffer, job posting, state-bit updates, output).  v3 gives each chain its own warp:
//   warp 0  DECODER   search, interval arithmetic, renormalisation, bit window.  Touches no context state:
//                     it receives, per symbol, a PACKET (state + the word or 64-byte record that state needs) in
//                     shared memory and publishes the decoded symbol.
//   warp 1  CONTEXT   receives the symbol, derives the next context, waits for a pending update if there is one,
//                     fetches that context's data and hands the packet over -- while the decoder renormalises --
//                     then does the bookkeeping for the symbol just decoded (first-visit word and state bits, or
//                     the job for the updater), keeps the row buffer and writes finished rows out.
//   warp 2  UPDATER   unchanged: ContextModel.update_model behind both.
// The two hand-overs per symbol are single shared-memory words carrying the position as a sequence number.
// Everything else (direct-mapped contexts, state bits, tables, records, job ring, exactness fallbacks) is v2's.
//
// MEASURED (B200, 1024 streams): the decoder warp's own work drops (fresh contexts: 1440 -> 1050 cycles/symbol) but
// it then waits 250..1000 cycles for the packet (two hand-overs + the context warp's fetch), and with three warps per
// block the kernel is capped at 80 registers: 8.0 ms against v2's 7.6 ms.  v2 stays the default; this kernel is
// selected with LC_DECODER=v3 and is covered by the same emulator and GPU parity tests.
#pragma once
#include "lc_decoder_v2.cuh"

#define LC3_WARPS 3
#define LC3_STOP 0x3FFu // symbol value of the decoder's "stopped here" message (alphabets have at most 256 symbols)

#ifdef LC_HOSTSIM
#define LC3_SPIN() emu::spin_yield()
#else
#define LC3_SPIN() ((void)0)
#endif

// mailbox in shared memory (LcV2Cfg::sm_mail): symbol word | pad | 8-byte packet header | 64 bytes of packet data.
// Both hand-overs are single aligned words written by one lane (a 64-bit shared-memory store is not torn), so the
// common packets (states 0, 1, 3: the context word travels inside the header) need no ordering at all; the
// inline record of state 2 is stored by lanes 0..3 before the header by the same warp (shared-memory stores of a
// warp are performed in program order; __syncwarp() orders the lanes), and the decoder reads it after the header.
struct LcV3Mail {
    uint32_t *sym;            // decoder -> context: (pos << 10) | symbol
    unsigned long long *hdr;  // context -> decoder: pos << 32 | context word << 3 | abort << 2 | state
    uint32_t *data;           // 16 words: the inline record (state 2)
};
#ifdef LC_HOSTSIM
static inline unsigned long long lc3_ld_hdr(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }
static inline void lc3_st_hdr(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *)p = v; }
#else
static __device__ __forceinline__ unsigned long long lc3_ld_hdr(const unsigned long long *p) { return *(const volatile unsigned long long *)p; }
static __device__ __forceinline__ void lc3_st_hdr(unsigned long long *p, unsigned long long v) { *(volatile unsigned long long *)p = v; }
#endif

// =================================================================================================
// DECODER warp
// =================================================================================================
__device__ __forceinline__ void lc3_decoder(LcFast &F, const LcV2 &V, const LcV3Mail &M, const unsigned char *src,
                                            long long nbytes, int *status_out, int *fault_index)
{
    const int lane = F.lane;
    LcvBits br; lcv_br_init(br, src, nbytes);
    uint32_t lo = 0u, hi = 0xffffffffu;
    uint32_t code = lcv_br_take(br, 32); // start_decoding (:247-258)
    int status = LC_OK;
    int pos = 0;
    LCP_DECL
    LCP_INIT();
    for (; pos < F.total; pos++) {
        LCP_START();
        // ---- the packet of this position's context
        unsigned long long hdr;
        while ((uint32_t)((hdr = lc3_ld_hdr(M.hdr)) >> 32) != (uint32_t)pos) LC3_SPIN();
        if (hdr & 4ull) { status = (int)lcv_ld_vol(V.abort_code); break; }
        const int st = (int)(hdr & 3ull);
        LCP_ROW(st); LCP_COUNT(st, 0); LCP_MARK(1);
        const uint32_t gw = (uint32_t)hdr >> 3;
        double2 q0 = {0.0, 0.0}, q1 = q0, q2 = q0, q3 = q0;
        if (st == 2) {
            const volatile double2 *d = (const volatile double2 *)M.data;
            q0.x = d[0].x; q0.y = d[0].y; q1.x = d[1].x; q1.y = d[1].y;
            q2.x = d[2].x; q2.y = d[2].y; q3.x = d[3].x; q3.y = d[3].y;
        }
        // ---- decode_symbol (:272-292)
        int s = 0, s1 = 0, fell_back = 0;
        uint32_t nlo = 0u, nhi = 0u;
        LcvRinv rv3; // (this variant computes the reciprocal of the range at the head of every symbol)
        rv3.d = lc_rcp_fast(lc_ll2d_small((long long)(hi - lo) + 1)); rv3.f = lcv_rcp_f32((float)(hi - lo));
        const int fs = lcv_decode_symbol<false>(F, V, st, gw, q0, q1, q2, q3, lo, hi, code, s, s1, nlo, nhi, fell_back,
                                         [](int) {}, rv3);
        if (fs != LC_OK) { status = fs; break; }
        if (lane == 0) lcv_st_vol(M.sym, ((uint32_t)pos << 10) | (uint32_t)s); // the context warp takes it from here
        lo = nlo; hi = nhi;
        LCP_MARK(2);
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
            } else {
                const uint32_t b1 = lcv_br_take(br, d);
                code = __funnelshift_lc(0u, code, d) | b1;
                const uint32_t b2 = lcv_br_take(br, e);
                code = ((code << e) | b2) ^ em;
            }
            lo = __funnelshift_lc(0u, lo_d, e) & ~em;
            hi = __funnelshift_lc(0xffffffffu, hi_d, e) | em;
        }
        LCP_MARK(3);
    }
    LCP_FLUSH();
    if (pos < F.total && lane == 0) lcv_st_vol(M.sym, ((uint32_t)pos << 10) | LC3_STOP); // stopped before position pos
    *fault_index = pos;
    *status_out = status;
}

// =================================================================================================
// CONTEXT warp
// =================================================================================================
__device__ __forceinline__ void lc3_context(const LcFast &F, const LcV2 &V, const LcV3Mail &M, LcvPost &P, int *out,
                                            const float *deq_table, float *deq_out)
{
    const int lane = F.lane, n = F.n, C = F.C, total = F.total;
    int pos = 0, r = 0, c = 0;
    uint32_t key = 0u; // context of position pos: (left=-1, up=-1) ...
    int st = 0;        // ... its state ...
    uint32_t gw = 0u;  // ... and its word (states 1, 3)
    P.my_key = LCV_SENTINEL; // contexts of the previous stream are not this stream's
    if (lane == 0) lc3_st_hdr(M.hdr, 0ull); // packet of position 0: a fresh context
    int stop_pos = total;
    for (; pos < total; pos++) {
        const uint32_t shift = (key & 15u) * 2u;
        // next position and the symbol above it (written at least C-1 >= 3 symbols ago)
        int c2 = c + 1, r2 = r;
        if (c2 == C) { c2 = 0; if (++r2 == F.R) r2 = 0; }
        const int up2 = r2 > 0 ? (int)V.rows[((r2 - 1) & 1) * C + c2] : -1;
        // ---- the symbol of position pos
        uint32_t m;
        while (((m = lcv_ld_vol(M.sym)) >> 10) != (uint32_t)pos) LC3_SPIN();
        const int s = (int)(m & 0x3FFu);
        if ((uint32_t)s == LC3_STOP) { stop_pos = pos; break; }
        if (lane == 0) V.rows[(r & 1) * C + c] = (unsigned char)s;
        int s1 = (int)(gw & 0x3FFu);
        if (s1 >= n) s1 = n - 1;
        const bool more = pos + 1 < total;
        const uint32_t key2 = (uint32_t)((c2 > 0 ? s : -1) + 1) * (uint32_t)(n + 1) + (uint32_t)(up2 + 1);
        const uint32_t shift2 = (key2 & 15u) * 2u;
        // this symbol's own update of its context comes first when the next position has the same context
        const bool book_first = key2 == key || !more;
        int st2 = 0;
        uint32_t gw2 = 0u;
#define LC3_BOOKKEEP()                                                                                           \
        do {                                                                                                     \
            if (st == 0) {                                                                                       \
                if (lane == 0) { __stcg(V.gword + key, (uint32_t)s); atomicOr(V.sbits + (key >> 4), 1u << shift); } \
            } else lcv_post(V, P, lane, key, LCV_PAY(s, st, s1));                                                \
            __syncwarp();                                                                                        \
        } while (0)
        if (book_first) LC3_BOOKKEEP();
        if (more) {
            // ---- the next context: state, pending update, data, packet
            st2 = (int)((lcv_ld_vol(V.sbits + (key2 >> 4)) >> shift2) & 3u);
            if (st2 != 0) {
                const bool mine = P.my_key == key2; // a job on it among the last LCV_RING posted ones may still be running
                if (__ballot_sync(LC_FULL_MASK, mine)) {
                    if (mine) while ((int)(lcv_ld_acq(V.ring_done) - (P.my_job + 1u)) < 0) LCV_SPIN();
                    __syncwarp();
                    LCV_FENCE();
                    st2 = (int)((lcv_ld_vol(V.sbits + (key2 >> 4)) >> shift2) & 3u);
                }
                if (st2 == 2) {
                    if (lane < 4) {
                        const double2 v = __ldcg((const double2 *)(V.grec + (size_t)key2 * 64) + lane);
                        volatile double2 *d = (volatile double2 *)M.data + lane;
                        d->x = v.x; d->y = v.y;
                    }
                    __syncwarp();
                } else gw2 = __ldcg(V.gword + key2);
            }
            const uint32_t ab = lcv_ld_vol(V.abort_code) ? 4u : 0u;
            if (lane == 0)
                lc3_st_hdr(M.hdr, ((unsigned long long)(uint32_t)(pos + 1) << 32) | ((gw2 & 0x1fffffffu) << 3) | ab | (uint32_t)st2);
        }
        if (!book_first) LC3_BOOKKEEP();
#undef LC3_BOOKKEEP
        if (c2 == 0) // a row is complete: write it out
            lcv_flush_row(V.rows + (r & 1) * C, 0, C, out + (pos - (C - 1)), deq_table,
                          deq_out ? deq_out + (pos - (C - 1)) : (float *)0, lane);
        key = key2; c = c2; r = r2; st = st2; gw = gw2;
    }
    // release the updater
    for (int u = 0; u < LCV_NU; u++) lcv_post(V, P, lane, LCV_SENTINEL, 0u);
    __syncwarp();
    if (stop_pos < total) {
        // symbols of the unfinished row (c of them), zeros after the position the decoder stopped at
        if (c > 0) lcv_flush_row(V.rows + (r & 1) * C, 0, c, out + (stop_pos - c), deq_table,
                                 deq_out ? deq_out + (stop_pos - c) : (float *)0, lane);
        for (int z = stop_pos + lane; z < total; z += 32) { out[z] = 0; if (deq_out) deq_out[z] = 0.0f; }
    }
}

// Block entry: LC3_WARPS warps, persistent over streams.
__device__ __forceinline__ void lc3_decode_block(const LcCoderCfg &cfg, const LcV2Cfg &vc, const unsigned char *bytes,
                                                 const long long *offsets, const int *nbits, int B, int *out,
                                                 const float *deq_table, float *deq_out, int *status, int *fault,
                                                 char *scratch, const double *tables, char *smem)
{
    const int warp = (int)(threadIdx.x >> 5);
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
    V.t2 = (const char *)0;
    V.pool_bytes = vc.pool_bytes; V.eps_k = vc.eps_k; V.lg_n = vc.lg_n;
    lcv_view_sa(V);
    LcV3Mail M;
    M.sym = (uint32_t *)(smem + vc.sm_mail); M.hdr = (unsigned long long *)(M.sym + 2); M.data = M.sym + 4;
    LcFast F;
    F.n = cfg.n; F.C = cfg.C; F.R = cfg.R; F.total = cfg.total; F.lane = (int)(threadIdx.x & 31);
    F.rate = cfg.rate; F.delta = cfg.delta; F.u0 = LC_DDIV(1.0, (double)cfg.n);
    F.delta_v = cfg.delta + 1.5e-14; F.tmargin = (double)cfg.n * 1.5e-14;
    F.P1 = LC_DADD(F.u0, LC_DMUL(F.rate, LC_DSUB(1.0, F.u0)));
    F.slot_cap = 0; F.slot_shift = 0; F.pool_bytes = vc.pool_bytes; F.pool_top = 0;
    F.pw_len = cfg.pw_len; F.pw_steps = cfg.pw_steps; F.pw_chains = cfg.pw_chains;
    F.slots = (unsigned long long *)0; F.pool = V.pool;
    F.dense = (double *)(smem + vc.sm_dense);
    F.u1tab = V.u1tab; F.rows = (unsigned short *)0;
    F.k = 0; F.u = F.u0; F.my_sym = 0x7fffffff; F.my_val = 0.0;
    if (threadIdx.x < 64) V.u1tab[threadIdx.x] = tables[threadIdx.x];
    if (threadIdx.x < LCV_RING) { lcv_bar_init(V.ring_bar + threadIdx.x); V.ring_done[threadIdx.x] = 0u; }
    LcvPost P; P.njobs = 0u; P.done_seen = 0u; P.my_key = LCV_SENTINEL; P.my_job = 0u;
    uint32_t ujob = 0u;
    const uint32_t nwords = (vc.nkeys + 15u) / 16u;
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        __syncthreads(); // the previous stream is finished by every warp
        for (uint32_t i = threadIdx.x; i < nwords; i += blockDim.x) V.sbits[i] = 0u;
        if (threadIdx.x == 0) { *V.pool_top = 0u; *V.abort_code = 0u; *M.sym = 0xffffffffu; *M.hdr = ~0ull; }
        __syncthreads();
        if (warp == 0) {
            int fi = 0, st = 0;
            const long long nby = ((long long)nbits[sidx] + 7) >> 3;
            lc3_decoder(F, V, M, bytes + offsets[sidx], nby, &st, &fi);
            if (F.lane == 0) { status[sidx] = st; fault[sidx] = fi; }
        } else if (warp == 1) {
            lc3_context(F, V, M, P, out + (size_t)sidx * cfg.total, deq_table,
                        deq_out ? deq_out + (size_t)sidx * cfg.total : (float *)0);
        } else {
            lcv_updater(F, V, ujob);
        }
    }
}
