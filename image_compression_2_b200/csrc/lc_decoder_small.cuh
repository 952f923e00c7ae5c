// Decoder for SMALL alphabets (n <= 16): cabac_decode (cabac_compression.py:363-406), repaired coder mode,
// (left,up) contexts.
//
// Small alphabets revisit few contexts many times -- (n+1)^2 = 289 contexts for 8192 symbols at 4 bits -- so the
// sparse-record design of lc_decoder_v2.cuh (built for 8-bit latents, where half the symbols open a fresh context)
// is the wrong shape for them: at 4 bits 73 % of the symbols found a pool record of more than six entries, every
// symbol was an updater job, and decoder and updater warp took ~3000 cycles per symbol (profiles/r01_*, DESIGN.md
// section 5).  Here the model is what the reference keeps (ContextModel.context_models, :73): one DENSE float64
// vector per context (289 x 16 doubles = 37 KB per stream), and one warp per stream does everything with the lanes
// working on the vector elements.  The table lives in GLOBAL memory, one per resident warp: in shared memory it
// capped an SM at five streams and the kernel was latency bound at ~1560 cycles per symbol (35.8 ms for 4096
// streams of cfg3 at 4 bits); in global memory 32 warps fit an SM, the 128-byte vector reads and writes stay in
// the L2 (the tables of all resident warps are ~175 MB of address space, a stream touches its own 37 KB over
// and over), and their latency is covered by the other warps:
//   * open a context: every lane reads the whole vector (broadcast loads) and runs the strictly sequential
//     np.cumsum (:346-347) up to its own element, so lane i holds cum[i] and cum[i+1] EXACTLY as the reference
//     computes them -- no guard bands, no approximate sums;
//   * decode_symbol (:272-292): all lanes compare their cum[i+1] with the scaled code value at once (one ballot is
//     np.searchsorted), and every lane has already turned ITS interval into candidate low/high (the two float64
//     products and truncations of :291-292) while the scaled value was being computed -- the winner's pair is
//     fetched with two shuffles;
//   * update_model (:119-144) on the lanes: NumPy's pairwise sum is an 8-lane xor-butterfly over a[j]+a[j+8], the
//     scale factor one IEEE division, the new vector goes back to shared memory.
// The scaled value uses a fast reciprocal; whenever any cum[i] lies within its error bound of it the exact IEEE
// quotient decides instead.  Anything unusual (corrupt streams: symbol outside the alphabet, empty range) flags the
// stream for the generic kernel, which redoes it from the start and reports the reference's fault.
#pragma once
#include "lc_decoder_v2.cuh"

#define LCD_MAX_N 16

static inline int lcd_eligible(const LcCoderCfg &c)
{
    return c.mode == LC_MODE_REPAIRED && c.has_ctx && c.n >= 2 && c.n <= LCD_MAX_N && c.C <= 8192;
}
// per resident warp: the dense table in global scratch; shared memory: the previous/current row of decoded symbols
static inline LC_HD uint32_t lcd_tab_bytes(int n) { return lc_round_up((uint32_t)(n + 1) * (uint32_t)(n + 1) * (uint32_t)n * 8u, 256); }
static inline LC_HD uint32_t lcd_smem_bytes(int C) { return lc_round_up(2u * (uint32_t)C, 16); }
#define LCD_WARPS_PER_SM 32

// NumPy's pairwise float64 sum (call site :135) of the N values held one per lane (lane l < N holds a[l]); every lane
// gets the total.  N >= 8: accumulators r[j] = a[j] (+ a[j+8]), combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) -- an
// xor-butterfly inside groups of eight lanes (float64 addition is commutative, so every lane of a group ends with
// the same value).  N < 8: the plain loop res = 0.; res += a[i].
template <int N> __device__ __forceinline__ double lcd_pairwise_total(double a, int lane)
{
    if (N >= 8) {
        const int j = lane & 7;
        double r = __shfl_sync(LC_FULL_MASK, a, j);
        if (N == 16) r = LC_DADD(r, __shfl_sync(LC_FULL_MASK, a, j + 8));
        r = LC_DADD(r, __shfl_xor_sync(LC_FULL_MASK, r, 1));
        r = LC_DADD(r, __shfl_xor_sync(LC_FULL_MASK, r, 2));
        r = LC_DADD(r, __shfl_xor_sync(LC_FULL_MASK, r, 4));
        return r;
    }
    double t = __shfl_sync(LC_FULL_MASK, a, 0); // 0. + a[0] is a[0]
#pragma unroll
    for (int i = 1; i < N; i++) t = LC_DADD(t, __shfl_sync(LC_FULL_MASK, a, i));
    return t;
}

// The model of one context for the decoding warp: lane l < N holds p[l], cum[l] and cum[l+1] (sequential sums).
struct LcdCtx { double p, clo, chi; };

template <int N> __device__ __forceinline__ LcdCtx lcd_open(const double *vec, int lane)
{
    LcdCtx c;
    double q[N];
#pragma unroll
    for (int k = 0; k < N; k++) q[k] = vec[k]; // same address in every lane: broadcast
    double T = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) T = LC_DADD(T, k <= lane ? q[k] : 0.0); // x + 0. is x: lane l stops at cum[l+1]
    c.chi = T;
    c.clo = __shfl_up_sync(LC_FULL_MASK, T, 1);
    if (lane == 0) c.clo = 0.0;
    c.p = vec[lane < N ? lane : N - 1];
    return c;
}

// ContextModel.update_model (:119-144) on the lanes: lane l < N holds p[l] and gets its new value (lanes >= N: 0)
template <int N> __device__ __forceinline__ double lcd_update(double p, int s, int lane, double rate)
{
    const bool valid = lane < N;
    const double p_s = __shfl_sync(LC_FULL_MASK, p, s);
    const double p_new = LC_DADD(p_s, LC_DMUL(rate, LC_DSUB(1.0, p_s)));
    const double a = lane == s ? p_new : (valid ? p : 0.0);
    const double tot = lcd_pairwise_total<N>(a, lane);
    const double others = LC_DSUB(tot, p_new);
    const double f = (others > 0.0) ? LC_DDIV(LC_DSUB(1.0, p_new), others) : 0.0;
    return lane == s ? p_new : (valid ? LC_DMUL(p, f) : 0.0);
}

template <int N>
__device__ __forceinline__ void lcd_decode_stream(const LcCoderCfg &cfg, double *tab, unsigned char *rows, int lane,
                                                  const unsigned char *src, long long nbytes, LcIdxOut out,
                                                  const float *deq_table, float *deq_out, int *status_out, int *fault_index)
{
    const int C = cfg.C, R = cfg.R, total = cfg.total;
    const double rate = cfg.rate, u0 = LC_DDIV(1.0, (double)N);
    const double dv = 1.5e-14; // bound on |fast scaled value - exact scaled value| (values <= 1, reciprocal to 2^-49)
    const bool valid = lane < N;
    // fresh model: every context is ones(n)/n (:73)
    for (int i = lane; i < (N + 1) * (N + 1) * N; i += 32) tab[i] = u0;
    __syncwarp();
    LcvBits br; lcv_br_init(br, src, nbytes);
    uint32_t lo = 0u, hi = 0xffffffffu;
    uint32_t code = lcv_br_take(br, 32); // start_decoding (:247-258)
    int status = LC_OK;
    int pos = 0, r = 0, c = 0;
    unsigned char *row_cur = rows, *row_prev = rows + C;
    uint32_t key = 0u; // (left=-1, up=-1)
    LcdCtx M = lcd_open<N>(tab, lane);
    for (; pos < total; pos++) {
        // ---- decode_symbol (:272-292)
        const uint32_t rng1 = hi - lo, off = code - lo; // range-1, (code-low+1)-1
        if (!(hi >= lo && off <= rng1 && rng1 >= 0xffffu)) { status = LC_NEEDS_GENERIC; break; }
        const double rd = lc_ll2d_small((long long)rng1 + 1), nd = lc_ll2d_small((long long)off + 1);
        // every lane's own interval as new bounds (:291-292), independent of the symbol search
        const long long al = LC_D2LL(LC_DMUL(rd, M.clo)), ah = LC_D2LL(LC_DSUB(LC_DMUL(rd, M.chi), 1.0));
        const double va = nd * lc_rcp_fast(rd) - 1e-10;
        unsigned m = __ballot_sync(LC_FULL_MASK, valid && M.chi >= va);
        const unsigned close_call = __ballot_sync(LC_FULL_MASK, valid && fabs(M.chi - va) <= dv);
        if (close_call != 0u || m == 0u || !(va > dv)) {
            // a cum[i] within the reciprocal's error of the scaled value, or at an end of the table: the
            // reference's own expression decides (:285,288)
            const double v = LC_DSUB(LC_DDIV(LC_DMUL(nd, 1.0), rd), 1e-10);
            if (!(0.0 < v)) { status = LC_NEEDS_GENERIC; break; }         // symbol -1
            m = __ballot_sync(LC_FULL_MASK, valid && M.chi >= v);
            if (m == 0u) { status = LC_NEEDS_GENERIC; break; }             // symbol n: IndexError in the reference
        }
        const int s = __ffs((int)m) - 1; // first i with cum[i+1] >= v: np.searchsorted(cum, v) - 1
        hi = lo + __shfl_sync(LC_FULL_MASK, (uint32_t)ah, s); // high first: both use the old low (:291-292)
        lo = lo + __shfl_sync(LC_FULL_MASK, (uint32_t)al, s);
        // ---- next position's context (get_context :78-117)
        if (lane == 0) row_cur[c] = (unsigned char)s;
        const bool last = c + 1 == C;
        int rn = r, cn = c + 1;
        if (last) { cn = 0; rn = r + 1 == R ? 0 : r + 1; }
        // ---- renormalise (:295-303) and underflow (:306-309), closed form (as in lc_decoder_v2.cuh)
        {
            const int d = __clz((int)(lo ^ hi));
            const uint32_t lo_d = __funnelshift_lc(0u, lo, d), hi_d = __funnelshift_lc(0xffffffffu, hi, d);
            const int e = __clz((int)~((lo_d & ~hi_d) << 1));
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
        // ---- update_model (:119-144) on the lanes
        {
            const double pn = lcd_update<N>(M.p, s, lane, rate);
            if (valid) tab[(size_t)key * N + lane] = pn;
        }
        __syncwarp(); // the row entry and the new vector are visible to every lane
        if (last) { // a row is complete: write it out
            lcv_flush_row(row_cur, 0, C, out + (pos - (C - 1)), deq_table, deq_out ? deq_out + (pos - (C - 1)) : (float *)0, lane);
            unsigned char *t_ = row_cur; row_cur = row_prev; row_prev = t_;
        }
        if (pos + 1 < total) {
            const int left = cn > 0 ? s : -1;
            const int up = rn > 0 ? (int)row_prev[cn] : -1;
            key = (uint32_t)(left + 1) * (uint32_t)(N + 1) + (uint32_t)(up + 1);
            M = lcd_open<N>(tab + (size_t)key * N, lane);
        }
        r = rn; c = cn;
    }
    *fault_index = pos;
    *status_out = status;
    __syncwarp();
    if (status == LC_OK) { // symbols of an unfinished last row (none when the stream ends on a row boundary)
        if (c > 0) lcv_flush_row(row_cur, 0, c, out + (pos - c), deq_table, deq_out ? deq_out + (pos - c) : (float *)0, lane);
    }
}

// Block entry: one warp, persistent over streams.
template <int N>
__device__ __forceinline__ void lcd_decode_block(const LcCoderCfg &cfg, const unsigned char *bytes, const long long *offsets,
                                                 const int *nbits, int B, LcIdxOut out, const float *deq_table,
                                                 float *deq_out, int *status, int *fault, char *scratch, char *smem)
{
    const int lane = (int)(threadIdx.x & 31);
    double *tab = (double *)(scratch + (size_t)blockIdx.x * lcd_tab_bytes(N));
    unsigned char *rows = (unsigned char *)smem;
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        int fi = 0, st = 0;
        const long long nby = ((long long)nbits[sidx] + 7) >> 3;
        lcd_decode_stream<N>(cfg, tab, rows, lane, bytes + offsets[sidx], nby, out + (size_t)sidx * cfg.total, deq_table,
                             deq_out ? deq_out + (size_t)sidx * cfg.total : (float *)0, &st, &fi);
        if (lane == 0) { status[sidx] = st; fault[sidx] = fi; }
        __syncwarp();
    }
}

// ---- the encoder's phase A for the same alphabets (lc_encoder_sparse.cuh calls it): one warp evolves ONE context's
// dense vector through all its visits -- lane l holds p[l] in a register, the vector passes through a 128-byte
// shared-memory image once per visit so that every lane can run the sequential np.cumsum up to its own element -- and
// writes the exact (cum[s], cum[s+1]) of every visit from the third on (phase S already wrote the first two from
// closed forms and the per-launch table).  The sparse-record phase A spent 376 warp instructions per symbol at 4 bits
// (records of up to 16 entries, lane-parallel prefix sums with guard bands and exact re-evaluation); this is ~110.
template <int N>
__device__ __forceinline__ void lcd_enc_group(LcCodes codes, const uint32_t *skeys, const unsigned short *spos, int j0,
                                              int total, double rate, double *image, double *ivs, int lane)
{
    const uint32_t key = skeys[j0];
    double p = lane < N ? LC_DDIV(1.0, (double)N) : 0.0;
    p = lcd_update<N>(p, codes[spos[j0]], lane, rate);
    p = lcd_update<N>(p, codes[spos[j0 + 1]], lane, rate);
    for (int t = j0 + 2;; t++) {
        const int pos = spos[t];
        const int s = codes[pos];
        const bool last = (t + 1 >= total) || (skeys[t + 1] != key);
        __syncwarp();
        if (lane < N) image[lane] = p;
        __syncwarp();
        const LcdCtx M = lcd_open<N>(image, lane);
        if (lane == s) { ivs[2 * pos] = M.clo; ivs[2 * pos + 1] = M.chi; }
        if (last) break; // the update after the last visit is never read
        p = lcd_update<N>(p, s, lane, rate);
    }
}
