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

// ContextModel.update_model (:119-144) on the dense image held in shared memory
__device__ __forceinline__ void lc_dense_update(LcWarp &W, int s)
{
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
        if (i != s) W.dense[i] = LC_DMUL(W.dense[i], f);
    __syncwarp();
}

// Phase A for one warp.  skeys/spos: this stream's positions sorted by (key, position); entries at
// index >= total are padding.  iv: float64 [2*total], iv[2p] = cum[s_p], iv[2p+1] = cum[s_p+1].
__device__ __forceinline__ void lc_enc_phase_a_warp(LcWarp &W, const int *__restrict__ codes,
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
// renormalisation/underflow loops are clz counts, and bits are appended many at a time to a 64-bit
// accumulator.  Same arithmetic as lc_enc_phase_b_stream, ~4x fewer instructions on the serial chain.
struct LcBits64 {
    uint32_t *out;
    uint32_t cap_words, wpos;
    unsigned long long acc; // the low `nacc` bits are pending output, nacc < 32 between calls
    int nacc, ovf;
    long long nbits;
};
// append the low nb bits of v, nb in 1..32
__device__ __forceinline__ void lc_b64_put(LcBits64 &b, uint32_t v, int nb, int lane)
{
    b.acc = (b.acc << nb) | v;
    b.nacc += nb;
    b.nbits += nb;
    if (b.nacc >= 32) {
        b.nacc -= 32;
        const uint32_t word = (uint32_t)(b.acc >> b.nacc);
        if (b.wpos < b.cap_words) { if (lane == 0) b.out[b.wpos] = __byte_perm(word, 0, 0x0123); }
        else b.ovf = 1;
        b.wpos++;
    }
}
__device__ __forceinline__ void lc_b64_put_run(LcBits64 &b, int bit, long long count, int lane)
{
    while (count > 0) {
        const int take = count > 32 ? 32 : (int)count;
        lc_b64_put(b, bit ? (take == 32 ? 0xffffffffu : ((1u << take) - 1u)) : 0u, take, lane);
        count -= take;
    }
}

__device__ __forceinline__ long long lc_enc_phase_b_repaired(int lane, const double *__restrict__ ivs, int limit,
                                                             uint32_t *out, uint32_t cap_words, int *status,
                                                             int *fault_index)
{
    LcBits64 bw;
    bw.out = out; bw.cap_words = cap_words; bw.wpos = 0; bw.acc = 0ull; bw.nacc = 0; bw.ovf = 0; bw.nbits = 0;
    uint32_t lo = 0u, hi = 0xffffffffu;
    long long outstanding = 0;
    int pos = 0;
    // The intervals come from DRAM (phase A wrote them) and do not depend on the coder state: each lane
    // holds one symbol's pair of the current and of the next 32-symbol chunk (coalesced 16-byte loads
    // issued a whole chunk ahead), and the pair of symbol pos+1 is broadcast while symbol pos is coded.
    double cur_lo = 0.0, cur_hi = 0.0, nx_lo = 0.0, nx_hi = 0.0;
    if (lane < limit) { cur_lo = __ldg(ivs + 2 * lane); cur_hi = __ldg(ivs + 2 * lane + 1); }
    if (32 + lane < limit) { nx_lo = __ldg(ivs + 2 * (32 + lane)); nx_hi = __ldg(ivs + 2 * (32 + lane) + 1); }
    double c_lo = __shfl_sync(LC_FULL_MASK, cur_lo, 0), c_hi = __shfl_sync(LC_FULL_MASK, cur_hi, 0);
    for (; pos < limit; pos++) {
        const int pn = pos + 1;
        if ((pn & 31) == 0) { // entering the next chunk: rotate and request the one after it
            cur_lo = nx_lo; cur_hi = nx_hi;
            const int q = pn + 32 + lane;
            if (q < limit) { nx_lo = __ldg(ivs + 2 * q); nx_hi = __ldg(ivs + 2 * q + 1); }
        }
        const double n_lo = __shfl_sync(LC_FULL_MASK, cur_lo, pn & 31), n_hi = __shfl_sync(LC_FULL_MASK, cur_hi, pn & 31);
        // encode_symbol (:220-224): high = low + int(range*c_hi - 1), low = low + int(range*c_lo)
        const double rd = lc_ll2d_small((long long)hi - (long long)lo + 1); // 0 when the interval has collapsed (hi = lo-1)
        const long long ah = LC_D2LL(LC_DSUB(LC_DMUL(rd, c_hi), 1.0));
        const long long al = LC_D2LL(LC_DMUL(rd, c_lo));
        hi = lo + (uint32_t)ah;
        lo = lo + (uint32_t)al;
        const int d = __clz((int)(lo ^ hi));
        if (d) {
            if (outstanding == 0) lc_b64_put(bw, hi >> (32 - d), d, lane);
            else {
                const int b1 = (int)(hi >> 31);
                lc_b64_put(bw, (uint32_t)b1, 1, lane);
                lc_b64_put_run(bw, 1 - b1, outstanding, lane);
                outstanding = 0;
                if (d > 1) lc_b64_put(bw, (hi << 1) >> (33 - d), d - 1, lane);
            }
            if (d == 32) { lo = 0u; hi = 0xffffffffu; }
            else { lo <<= d; hi = (hi << d) | ((1u << d) - 1u); }
        }
        const int e = __clz((int)~((lo & ~hi) << 1));
        if (e) {
            outstanding += e;
            lo = (lo << e) & 0x7fffffffu;
            hi = ((hi << e) & 0x7fffffffu) | 0x80000000u | ((1u << e) - 1u);
        }
        if (bw.ovf) break;
        c_lo = n_lo; c_hi = n_hi;
    }
    *fault_index = pos;
    if (bw.ovf) { *status = LC_OUT_OVERFLOW; return 0; }
    // finish_encoding (:230-245)
    outstanding += 1;
    const int first = (lo & 0x40000000u) != 0 ? 1 : 0;
    lc_b64_put(bw, (uint32_t)first, 1, lane);
    lc_b64_put_run(bw, 1 - first, outstanding, lane);
    const long long nbits = bw.nbits;
    if (bw.nacc > 0) lc_b64_put(bw, 0u, 32 - bw.nacc, lane); // zero-pad the last word
    if (bw.ovf) { *status = LC_OUT_OVERFLOW; return 0; }
    *status = LC_OK;
    return nbits;
}

// ---- block entry points --------------------------------------------------------------------------

// Phase A: one block per stream, blockDim.x/32 warps share the stream's groups.  `smem` holds one
// dense image (n doubles) per warp.  first_bad[b] (= total when the stream is clean) is the position of
// the first out-of-range symbol found by phase S: only the positions before it are sorted (the
// rest carry padding keys) and coded, then phase B reports LC_BAD_SYMBOL there -- exactly where
// the serial encoder stops.
__device__ __forceinline__ void lc_enc_phase_a_block(const LcCoderCfg &cfg, const int *codes, int B,
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

// =================================================================================================
// Phase A, lane-per-group variant (n <= 256).
//
// In the warp-per-group kernel above all 32 lanes execute the same strictly sequential np.cumsum
// walk, so 31/32 of the issued work is redundant.  Here every LANE owns one context group: its
// dense model is a column of a [n][32] float64 tile in shared memory (element i of lane L at
// tile[i*32+L]: conflict-free for 64-bit accesses), and the lane runs the reference's arithmetic
// as plain scalar code -- NumPy's 8-accumulator pairwise sum, the sequential prefix, the scaling
// loop.  Lanes advance one visit per step in lock step and pick up the next group of the stream
// when theirs ends.  One warp (block) per stream; groups with a single visit never reach a lane.
// =================================================================================================

// NumPy pairwise sum of one lane's column (n >= 8: blocks of min(n,128), 8 accumulators, binary tree)
__device__ __forceinline__ double lc_col_pairwise(const double *col, int n, int pw_len, int pw_steps)
{
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; i++) r = LC_DADD(r, col[i * 32]);
        return r;
    }
    double bs[8];
    const int nblk = n / pw_len;
    for (int b = 0; b < nblk; b++) {
        const double *p = col + (size_t)b * pw_len * 32;
        double r0 = p[0], r1 = p[32], r2 = p[64], r3 = p[96], r4 = p[128], r5 = p[160], r6 = p[192], r7 = p[224];
        for (int t = 1; t < pw_steps; t++) {
            const double *q = p + t * 256;
            r0 = LC_DADD(r0, q[0]);   r1 = LC_DADD(r1, q[32]);  r2 = LC_DADD(r2, q[64]);  r3 = LC_DADD(r3, q[96]);
            r4 = LC_DADD(r4, q[128]); r5 = LC_DADD(r5, q[160]); r6 = LC_DADD(r6, q[192]); r7 = LC_DADD(r7, q[224]);
        }
        bs[b] = LC_DADD(LC_DADD(LC_DADD(r0, r1), LC_DADD(r2, r3)), LC_DADD(LC_DADD(r4, r5), LC_DADD(r6, r7)));
    }
    // recursive halving over the blocks (n2 = n/2 is a multiple of 128 for n >= 256)
    for (int w = 1; w < nblk; w <<= 1)
        for (int b = 0; b + w < nblk; b += 2 * w) bs[b] = LC_DADD(bs[b], bs[b + w]);
    return bs[0];
}

// ContextModel.update_model (:119-144) on one lane's column
__device__ __forceinline__ void lc_col_update(double *col, int n, int pw_len, int pw_steps, double rate, int s)
{
    const double p_old = col[s * 32];
    const double p_new = LC_DADD(p_old, LC_DMUL(rate, LC_DSUB(1.0, p_old)));
    col[s * 32] = p_new;
    const double total = lc_col_pairwise(col, n, pw_len, pw_steps);
    const double others = LC_DSUB(total, p_new);
    const double f = (others > 0.0) ? LC_DDIV(LC_DSUB(1.0, p_new), others) : 0.0;
    // in-place scaling, eight elements at a time: loads first, then the multiplies, then the stores
    // (the compiler cannot reorder shared-memory loads across the stores of a rolled loop)
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        double *q = col + i * 32;
        const double a0 = q[0], a1 = q[32], a2 = q[64], a3 = q[96], a4 = q[128], a5 = q[160], a6 = q[192], a7 = q[224];
        q[0] = LC_DMUL(a0, f);   q[32] = LC_DMUL(a1, f);  q[64] = LC_DMUL(a2, f);  q[96] = LC_DMUL(a3, f);
        q[128] = LC_DMUL(a4, f); q[160] = LC_DMUL(a5, f); q[192] = LC_DMUL(a6, f); q[224] = LC_DMUL(a7, f);
    }
    for (; i < n; i++) col[i * 32] = LC_DMUL(col[i * 32], f);
    col[s * 32] = p_new;
}

// exact np.cumsum prefix of one lane's column: loads in batches of eight ahead of the dependent adds
__device__ __forceinline__ double lc_col_prefix(const double *col, int s)
{
    double T = 0.0;
    int i = 0;
    for (; i + 8 <= s; i += 8) {
        const double *q = col + i * 32;
        const double a0 = q[0], a1 = q[32], a2 = q[64], a3 = q[96], a4 = q[128], a5 = q[160], a6 = q[192], a7 = q[224];
        T = LC_DADD(T, a0); T = LC_DADD(T, a1); T = LC_DADD(T, a2); T = LC_DADD(T, a3);
        T = LC_DADD(T, a4); T = LC_DADD(T, a5); T = LC_DADD(T, a6); T = LC_DADD(T, a7);
    }
    for (; i < s; i++) T = LC_DADD(T, col[i * 32]);
    return T;
}

// Group list of one stream, produced by phase S (lc_enc_sort_kernel) or, for the emulator, by
// lc_enc_group_list_warp below: glist[g] = sorted index of the first visit of the g-th context that is
// visited at least twice; first visits get their closed-form interval (uniform model: cum[i] = i/n).
#define LC_PAR_MAX_GROUPS (LC_PAR_MAX_SYMBOLS / 2)
#define LC_PAR_TASK_GROUPS 64

__device__ __forceinline__ int lc_enc_group_list_warp(int lane, const int *codes, const uint32_t *skeys,
                                                      const unsigned short *spos, int total, double u0, double *ivs,
                                                      unsigned short *glist)
{
    int ngroups = 0;
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        const bool valid = j < total;
        const uint32_t kj = valid ? skeys[j] : 0u;
        const bool head = valid && (j == 0 || skeys[j - 1] != kj);
        if (head) {
            const int p = spos[j];
            const int s = codes[p];
            ivs[2 * p] = LC_DMUL((double)s, u0);
            ivs[2 * p + 1] = LC_DMUL((double)(s + 1), u0);
        }
        const bool multi = head && (j + 1 < total) && (skeys[j + 1] == kj);
        const unsigned m = __ballot_sync(LC_FULL_MASK, multi);
        if (multi) glist[ngroups + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
        ngroups += __popc(m);
    }
    return ngroups;
}

// smem: tile[n*32] doubles | u1tab[32] doubles.
// Persistent warps pull tasks (stream, chunk of LC_PAR_TASK_GROUPS groups) from *task_counter.
__device__ __forceinline__ void lc_enc_phase_a_lanes_block(const LcCoderCfg &cfg, const int *codes_all, int B,
                                                           const uint32_t *skeys_all, const unsigned short *spos_all,
                                                           const int *first_bad, const unsigned short *glist_all,
                                                           const int *ngroups_all, double *ivs_all,
                                                           unsigned int *task_counter, char *smem)
{
    const int lane = (int)(threadIdx.x & 31);
    const int n = cfg.n, pw_len = cfg.pw_len, pw_steps = cfg.pw_steps;
    const double rate = cfg.rate;
    const double u0 = LC_DDIV(1.0, (double)n);
    const double P1 = LC_DADD(u0, LC_DMUL(rate, LC_DSUB(1.0, u0)));
    double *tile = (double *)smem;
    double *col = tile + lane;
    double *u1tab = tile + (size_t)n * 32;
    // u after the first update, by chain step of the symbol (by symbol when n < 8): each lane computes one entry
    {
        const int entries = cfg.pw_chains == 0 ? n : pw_steps;
        if (lane < entries) {
            const int s = cfg.pw_chains == 0 ? lane : 8 * lane;
            for (int i = 0; i < n; i++) col[i * 32] = u0;
            lc_col_update(col, n, pw_len, pw_steps, rate, s);
            u1tab[lane] = col[(s == 0 ? 1 : 0) * 32]; // any other symbol: u0 * f
        }
        __syncwarp();
    }
    const unsigned chunks_per_stream = LC_PAR_MAX_GROUPS / LC_PAR_TASK_GROUPS;
    const unsigned n_tasks = (unsigned)B * chunks_per_stream;
    for (;;) {
        unsigned task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(LC_FULL_MASK, task, 0);
        if (task >= n_tasks) break;
        const int sidx = (int)(task / chunks_per_stream);
        const int g_lo = (int)(task % chunks_per_stream) * LC_PAR_TASK_GROUPS;
        const int ngroups = ngroups_all[sidx];
        if (g_lo >= ngroups) continue;
        const int g_hi = (g_lo + LC_PAR_TASK_GROUPS < ngroups) ? g_lo + LC_PAR_TASK_GROUPS : ngroups;
        const size_t o = (size_t)sidx * LC_PAR_MAX_SYMBOLS;
        const int *codes = codes_all + (size_t)sidx * cfg.total;
        const uint32_t *skeys = skeys_all + o;
        const unsigned short *spos = spos_all + o;
        const unsigned short *glist = glist_all + (size_t)sidx * LC_PAR_MAX_GROUPS;
        double *ivs = ivs_all + 2 * o;
        const int fb = first_bad[sidx];
        const int total = fb < cfg.total ? fb : cfg.total;

        // every lane walks one group at a time, one visit per step
        int next = g_lo;   // next unassigned group of this task (warp-uniform)
        int t = -1;        // sorted index of the visit handled in this step, -1 = idle
        int visit = 0;     // 2 = second visit of the group (model still implicit), >= 3 = column is live
        int s1 = 0, p = 0, s = 0;
        bool last = false;
        double u1 = 0.0;
        uint32_t key = 0u;
        for (;;) {
            const unsigned need = __ballot_sync(LC_FULL_MASK, t < 0);
            if (need) {
                const int g = next + __popc(need & ((1u << lane) - 1u));
                if (t < 0 && g < g_hi) {
                    const int j0 = glist[g];
                    key = skeys[j0];
                    s1 = codes[spos[j0]];
                    u1 = u1tab[cfg.pw_chains == 0 ? s1 : ((s1 & (pw_len - 1)) >> 3)];
                    t = j0 + 1; // its second visit
                    visit = 2;
                    p = spos[t]; s = codes[p];
                    last = (t + 1 >= total) || (skeys[t + 1] != key);
                }
                next += __popc(need);
                if (next > g_hi) next = g_hi;
            }
            if (__ballot_sync(LC_FULL_MASK, t >= 0) == 0u) break;
            if (t >= 0) {
                // data of the following visit (if any): requested now, used in the next step
                const bool more = !last;
                int p_n = 0, s_n = 0;
                bool last_n = true;
                if (more) {
                    p_n = spos[t + 1];
                    last_n = (t + 2 >= total) || (skeys[t + 2] != key);
                }
                double T;
                double ps;
                if (visit == 2) {
                    // model after one update: u1 everywhere, P1 at s1 -- the exact np.cumsum prefix needs no column
                    T = 0.0;
                    const int run1 = s < s1 ? s : s1;
                    for (int i = 0; i < run1; i++) T = LC_DADD(T, u1);
                    if (s > s1) {
                        T = LC_DADD(T, P1);
                        for (int i = s1 + 1; i < s; i++) T = LC_DADD(T, u1);
                    }
                    ps = (s == s1) ? P1 : u1;
                } else {
                    T = lc_col_prefix(col, s);
                    ps = col[s * 32];
                }
                ivs[2 * p] = T;
                ivs[2 * p + 1] = LC_DADD(T, ps);
                if (more) s_n = codes[p_n];
                if (last) t = -1; // the update after the last visit is never read
                else {
                    if (visit == 2) { // the group goes on: materialise the column now
                        for (int i = 0; i < n; i++) col[i * 32] = u1;
                        col[s1 * 32] = P1;
                    }
                    lc_col_update(col, n, pw_len, pw_steps, rate, s);
                    t++; visit = 3;
                    p = p_n; s = s_n; last = last_n;
                }
            }
        }
        __syncwarp();
    }
}
