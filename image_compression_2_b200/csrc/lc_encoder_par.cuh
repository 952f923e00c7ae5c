// Parallel encoder (streams of at most LC_PAR_MAX_SYMBOLS symbols with (left,up) contexts).
//
// The encoder knows every symbol up front, and a context's probability vector depends only on the
// ordered symbols seen IN THAT CONTEXT (cabac_compression.py:128,143,157 are the only accesses to
// context_models[ctx]).  So cabac_encode (cabac_compression.py:315-359) splits into
//   phase S  per stream: context key of every position, stable sort of positions by key
//            (kernel lc_enc_sort_kernel in latentcodec.cu, CUB block radix sort);
//   phase A  per context group, in parallel over all groups of all streams: evolve the dense
//            float64 model exactly as ContextModel.update_model does (:119-144) and write, for
//            every position, the exact interval (cum[s], cum[s+1]) of np.cumsum (:346-347);
//            updates after a group's last visit are never read and are skipped;
//   phase B  per stream, serial: ArithmeticCoder.encode_symbol / renormalise / underflow /
//            finish (:189-245) consuming those intervals -- two multiplies, two truncations and
//            the bit emission per symbol.
// Phase A is the float64-heavy part and runs at full occupancy; phase B is a short dependent
// chain.  Results are bit-identical to the serial evaluation because every float64 operation is
// the same operation on the same operands in the same order.
#pragma once
#include "lc_coder.cuh"

#define LC_PAR_MAX_SYMBOLS 8192
#define LC_PAR_KEY_PAD 0xFFFFFFFFu

// exact np.cumsum prefix: sum of dense[0..s) in index order (all lanes compute the same value)
__device__ __forceinline__ double lc_dense_prefix(const double *dense, int s)
{
    double T = 0.0;
    int i = 0;
    for (; i + 4 <= s; i += 4) {
        const double a = dense[i], b = dense[i + 1], c = dense[i + 2], d = dense[i + 3];
        T = LC_DADD(T, a); T = LC_DADD(T, b); T = LC_DADD(T, c); T = LC_DADD(T, d);
    }
    for (; i < s; i++) T = LC_DADD(T, dense[i]);
    return T;
}

// ContextModel.update_model (:119-144) on the dense image held in shared memory.  A negative symbol is NumPy's
// negative index (the reference gets there after decoding symbol -1, :288-292,403): element n+s takes the increment
// and, because `i != symbol` is then true for every i, the scale factor is applied to ALL elements.
__device__ __forceinline__ void lc_dense_update(LcWarp &W, int s_signed)
{
    const int s = s_signed < 0 ? s_signed + W.n : s_signed;
    const double p_old = W.dense[s];
    const double p_new = LC_DADD(p_old, LC_DMUL(W.rate, LC_DSUB(1.0, p_old)));
    __syncwarp();
    if (W.lane == 0) W.dense[s] = p_new;
    __syncwarp();
    const double total = lc_pairwise_total(W);
    __syncwarp(); // every lane has read the image before it is scaled in place
    const double others = LC_DSUB(total, p_new);
    const double f = (others > 0.0) ? LC_DDIV(LC_DSUB(1.0, p_new), others) : 0.0;
    for (int i = W.lane; i < W.n; i += 32)
        if (i != s_signed) W.dense[i] = LC_DMUL(W.dense[i], f);
    __syncwarp();
}

// Phase A for one warp.  skeys/spos: this stream's positions sorted by (key, position); entries at
// index >= total are padding.  iv: float64 [2*total], iv[2p] = cum[s_p], iv[2p+1] = cum[s_p+1].
__device__ __forceinline__ void lc_enc_phase_a_warp(LcWarp &W, LcCodes codes,
                                                    const uint32_t *__restrict__ skeys,
                                                    const unsigned short *__restrict__ spos, double *iv,
                                                    int total, int warp_id, int n_warps)
{
    for (int chunk = warp_id; chunk * 32 < total; chunk += n_warps) {
        const int j = chunk * 32 + W.lane;
        const bool valid = j < total;
        const uint32_t kj = valid ? skeys[j] : 0u;
        const bool head = valid && (j == 0 || skeys[j - 1] != kj);
        if (head) { // first visit of a context: uniform model, cum[i] = i/n exactly
            const int p = spos[j];
            const int s = codes[p];
            iv[2 * p] = LC_DMUL((double)s, W.u0);
            iv[2 * p + 1] = LC_DMUL((double)(s + 1), W.u0);
        }
        const bool multi = head && (j + 1 < total) && (skeys[j + 1] == kj);
        unsigned m = __ballot_sync(LC_FULL_MASK, multi);
        while (m) {
            const int l = __ffs((int)m) - 1;
            m &= m - 1;
            int t = chunk * 32 + l;
            const uint32_t key = skeys[t];
            for (int i = W.lane; i < W.n; i += 32) W.dense[i] = W.u0;
            __syncwarp();
            for (;;) {
                const int p = spos[t];
                const int s = codes[p];
                const bool last = (t + 1 >= total) || (skeys[t + 1] != key);
                if (t != chunk * 32 + l) {
                    const double T = lc_dense_prefix(W.dense, s);
                    if (W.lane == 0) { iv[2 * p] = T; iv[2 * p + 1] = LC_DADD(T, W.dense[s]); }
                }
                if (last) break;
                lc_dense_update(W, s);
                t++;
            }
            __syncwarp();
        }
    }
}

// Phase B for one stream (one warp): the range coder over precomputed exact intervals.
__device__ __forceinline__ long long lc_enc_phase_b_stream(LcWarp &W, const double *__restrict__ ivs, int limit,
                                                           uint32_t *out, uint32_t cap_words, int *fault_index)
{
    LcBitWriter bw; lc_bw_init(bw, out, cap_words);
    long long low = 0, high = LC_FULL - 1, outstanding = 0;
    const long long fix = (W.mode == LC_MODE_VERBATIM) ? LC_FULL : LC_HALF; // defect D3
    double my_lo = 0.0, my_hi = 0.0;
    int pos = 0;
    W.status = LC_OK;
    for (; pos < limit; pos++) {
        const int l = pos & 31;
        if (l == 0) {
            const int p = pos + W.lane;
            if (p < limit) { my_lo = ivs[2 * p]; my_hi = ivs[2 * p + 1]; }
        }
        LcInterval iv;
        iv.clo = __shfl_sync(LC_FULL_MASK, my_lo, l);
        iv.chi = __shfl_sync(LC_FULL_MASK, my_hi, l);
        iv.exact = 1; iv.sym = 0;
        lc_interval_apply(iv, 0.0, low, high);
        if (W.mode == LC_MODE_REPAIRED) {
            // low and high stay below 2^32 in this mode, so the bit-at-a-time loops of
            // _renormalize_encoder / _handle_underflow (:189-210) have closed forms
            uint32_t lo = (uint32_t)low, hi = (uint32_t)high;
            const int d = __clz((int)(lo ^ hi)); // leading bits lo and hi share: that many bits are emitted
            if (d) {
                const int b1 = (int)(hi >> 31);
                lc_bw_put(bw, b1, 1, W.lane);
                if (outstanding > 0) { lc_bw_put(bw, 1 - b1, outstanding, W.lane); outstanding = 0; }
                if (d > 1) lc_bw_put_bits(bw, (hi << 1) >> (33 - d), d - 1, W.lane);
                if (d == 32) { lo = 0u; hi = 0xffffffffu; }
                else { lo <<= d; hi = (hi << d) | ((1u << d) - 1u); }
            }
            const int e = __clz((int)~((lo & ~hi) << 1)); // underflow steps: lo = 01.., hi = 10..
            if (e) {
                outstanding += e;
                lo = (lo << e) & 0x7fffffffu;
                hi = ((hi << e) & 0x7fffffffu) | 0x80000000u | ((1u << e) - 1u);
            }
            low = lo; high = hi;
        } else {
            while ((high & LC_HALF) == (low & LC_HALF)) {
                const long long bit = high >> 31;
                if (bit < 0 || bit > 1 || (outstanding > 0 && (1 - bit) < 0)) { W.status = LC_ENC_BIT_OVERFLOW; break; }
                lc_bw_put(bw, (int)bit, 1, W.lane);
                if (outstanding > 0) lc_bw_put(bw, (int)(1 - bit), outstanding, W.lane);
                outstanding = 0;
                low = (low << 1) & (LC_FULL - 1);
                high = ((high << 1) & (LC_FULL - 1)) | 1;
            }
            if (W.status != LC_OK) break;
            while ((low & LC_QUARTER) != 0 && (high & LC_QUARTER) == 0) {
                outstanding += 1;
                low = (low << 1) & (LC_HALF - 1);
                high = ((high << 1) & (LC_HALF - 1)) | fix | 1;
            }
        }
        if (bw.ovf) { W.status = LC_OUT_OVERFLOW; break; }
    }
    *fault_index = pos;
    if (W.status != LC_OK) return 0;
    outstanding += 1;
    const int first = (low & LC_QUARTER) != 0 ? 1 : 0;
    lc_bw_put(bw, first, 1, W.lane);
    lc_bw_put(bw, 1 - first, outstanding, W.lane);
    lc_bw_finish(bw, W.lane);
    if (bw.ovf) { W.status = LC_OUT_OVERFLOW; return 0; }
    return bw.nbits;
}

// Phase B, repaired mode: low/high stay below 2^32, so the state is two uint32, the
// renormalisation/underflow loops are clz counts and funnel shifts, and bits are appended many at a time.
// Everything is warp-uniform (every lane runs the same scalar chain; lane 0 stores): the interval of symbol
// pos+2 is requested with one 16-byte load by all lanes while symbol pos is coded.
struct LcBits32 {
    uint32_t *out;
    uint32_t cap_words, wpos;
    uint32_t acc; // the low `nacc` bits are pending output, nacc < 32 between calls
    int nacc, ovf;
};
// append the low nb bits of v (nb in 1..32, v < 2^nb)
__device__ __forceinline__ void lc_b32_put(LcBits32 &b, uint32_t v, int nb, int lane)
{
    const unsigned long long comb = ((unsigned long long)b.acc << nb) | v;
    const int tot = b.nacc + nb;
    if (tot >= 32) {
        const uint32_t word = (uint32_t)(comb >> (tot - 32));
        if (b.wpos < b.cap_words) { if (lane == 0) b.out[b.wpos] = __byte_perm(word, 0, 0x0123); }
        else b.ovf = 1;
        b.wpos++;
        b.nacc = tot - 32;
        b.acc = (uint32_t)comb & ((1u << b.nacc) - 1u);
    } else { b.acc = (uint32_t)comb; b.nacc = tot; }
}
__device__ __forceinline__ void lc_b32_put_run(LcBits32 &b, int bit, int count, int lane)
{
    while (count > 0) {
        const int take = count > 32 ? 32 : count;
        lc_b32_put(b, bit ? (take == 32 ? 0xffffffffu : ((1u << take) - 1u)) : 0u, take, lane);
        count -= take;
    }
}

__device__ __forceinline__ long long lc_enc_phase_b_repaired(int lane, const double *__restrict__ ivs, int limit,
                                                             uint32_t *out, uint32_t cap_words, int *status,
                                                             int *fault_index)
{
    LcBits32 bw;
    bw.out = out; bw.cap_words = cap_words; bw.wpos = 0; bw.acc = 0u; bw.nacc = 0; bw.ovf = 0;
    uint32_t lo = 0u, hi = 0xffffffffu;
    int outstanding = 0, nbits = 0;
    int pos = 0;
    const double2 *iv2 = (const double2 *)ivs; // (cum[s], cum[s+1]) per position, 16-byte aligned
    const double2 zero2 = {0.0, 0.0};
    double2 cur = limit > 0 ? __ldg(iv2) : zero2;
    double2 nx1 = limit > 1 ? __ldg(iv2 + 1) : zero2;
    for (; pos < limit; pos++) {
        const double2 nx2 = pos + 2 < limit ? __ldg(iv2 + pos + 2) : zero2;
        // encode_symbol (:220-224): high = low + int(range*c_hi - 1), low = low + int(range*c_lo)
        const double rd = lc_ll2d_small((long long)hi - (long long)lo + 1); // 0 when the interval has collapsed (hi = lo-1)
        const long long ah = LC_D2LL(LC_DSUB(LC_DMUL(rd, cur.y), 1.0));
        const long long al = LC_D2LL(LC_DMUL(rd, cur.x));
        hi = lo + (uint32_t)ah;
        lo = lo + (uint32_t)al;
        const int d = __clz((int)(lo ^ hi)); // leading bits low and high share: that many bits are emitted
        if (d) {
            nbits += d + outstanding;
            if (outstanding == 0) lc_b32_put(bw, hi >> (32 - d), d, lane);
            else {
                const int b1 = (int)(hi >> 31);
                lc_b32_put(bw, (uint32_t)b1, 1, lane);
                lc_b32_put_run(bw, 1 - b1, outstanding, lane);
                outstanding = 0;
                if (d > 1) lc_b32_put(bw, (hi << 1) >> (33 - d), d - 1, lane);
            }
        }
        const uint32_t lo_d = __funnelshift_lc(0u, lo, d), hi_d = __funnelshift_lc(0xffffffffu, hi, d);
        const int e = __clz((int)~((lo_d & ~hi_d) << 1)); // underflow steps: low = 01.., high = 10..
        const uint32_t em = e ? 0x80000000u : 0u;
        outstanding += e;
        lo = __funnelshift_lc(0u, lo_d, e) & ~em;
        hi = __funnelshift_lc(0xffffffffu, hi_d, e) | em;
        if (bw.ovf) break;
        cur = nx1; nx1 = nx2;
    }
    *fault_index = pos;
    if (bw.ovf) { *status = LC_OUT_OVERFLOW; return 0; }
    // finish_encoding (:230-245)
    outstanding += 1;
    const int first = (lo & 0x40000000u) != 0 ? 1 : 0;
    lc_b32_put(bw, (uint32_t)first, 1, lane);
    lc_b32_put_run(bw, 1 - first, outstanding, lane);
    nbits += 1 + outstanding;
    if (bw.nacc > 0) lc_b32_put(bw, 0u, 32 - bw.nacc, lane); // zero-pad the last word
    if (bw.ovf) { *status = LC_OUT_OVERFLOW; return 0; }
    *status = LC_OK;
    return (long long)nbits;
}

// ---- block entry points --------------------------------------------------------------------------

// Phase A: one block per stream, blockDim.x/32 warps share the stream's groups.  `smem` holds one
// dense image (n doubles) per warp.  first_bad[b] (= total when the stream is clean) is the position of
// the first out-of-range symbol found by phase S: only the positions before it are sorted (the
// rest carry padding keys) and coded, then phase B reports LC_BAD_SYMBOL there -- exactly where
// the serial encoder stops.
__device__ __forceinline__ void lc_enc_phase_a_block(const LcCoderCfg &cfg, LcCodes codes, int B,
                                                     const uint32_t *skeys, const unsigned short *spos,
                                                     const int *first_bad, double *ivs, char *smem)
{
    const int warp_id = (int)(threadIdx.x >> 5), n_warps = (int)(blockDim.x >> 5);
    LcWarp W;
    lc_warp_init(W, cfg, smem, (char *)0);
    W.dense = (double *)smem + (size_t)warp_id * cfg.n;
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        const size_t o = (size_t)sidx * LC_PAR_MAX_SYMBOLS;
        const int fb = first_bad[sidx];
        lc_enc_phase_a_warp(W, codes + (size_t)sidx * cfg.total, skeys + o, spos + o, ivs + 2 * o,
                            fb < cfg.total ? fb : cfg.total, warp_id, n_warps);
    }
}

// Phase B: one warp per stream (blockDim.x = 32).
__device__ __forceinline__ void lc_enc_phase_b_block(const LcCoderCfg &cfg, int B, const int *first_bad,
                                                     const double *ivs, unsigned char *out_slots,
                                                     uint32_t slot_bytes, int *nbits, int *status, int *fault)
{
    LcWarp W;
    lc_warp_init(W, cfg, (char *)0, (char *)0);
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        const size_t o = (size_t)sidx * LC_PAR_MAX_SYMBOLS;
        const int fb = first_bad[sidx];
        const int limit = fb < cfg.total ? fb : cfg.total;
        int fi = 0;
        long long nb;
        int st;
        if (cfg.mode == LC_MODE_REPAIRED) {
            nb = lc_enc_phase_b_repaired(W.lane, ivs + 2 * o, limit, (uint32_t *)(out_slots + (size_t)sidx * slot_bytes),
                                         slot_bytes / 4, &st, &fi);
        } else {
            nb = lc_enc_phase_b_stream(W, ivs + 2 * o, limit, (uint32_t *)(out_slots + (size_t)sidx * slot_bytes),
                                       slot_bytes / 4, &fi);
            st = W.status;
        }
        if (st == LC_OK && fb < cfg.total) { st = LC_BAD_SYMBOL; fi = fb; nb = 0; }
        if (W.lane == 0) { nbits[sidx] = (int)nb; status[sidx] = st; fault[sidx] = fi; }
        __syncwarp();
    }
}

#define LC_PAR_MAX_GROUPS (LC_PAR_MAX_SYMBOLS / 2)
