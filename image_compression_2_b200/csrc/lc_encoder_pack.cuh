// Phase B of the parallel encoder in two kernels (repaired coder mode).
//
// ArithmeticCoder.encode_symbol / _renormalize_encoder / _handle_underflow / finish_encoding
// (cabac_compression.py:189-245) have one truly serial part -- the (low, high) recurrence -- and one part that is
// serial only in the reference's formulation: appending the emitted bits.  Measured on the benchmark the fused
// loop ran at ~500 cycles/symbol for a recurrence whose dependent chain is ~120 cycles, because every in-order
// instruction of the bit writer sits between two steps of the chain.  So:
//   B1  one warp per stream, serial: the recurrence only.  Per symbol it leaves an 8-byte record
//       (high after the interval update | d << 32 | e << 40): the d leading bits low and high share are the bits
//       that symbol emits, e is its count of underflow steps.  Records overwrite the first half of the symbol's
//       (cum[s], cum[s+1]) pair, which has been consumed by then.
//   B2  one 256-thread block per stream, parallel: the emitted string of symbol i is
//       [b][pending_i x !b][d_i - 1 more bits] with pending_i = the e's accumulated since the previous emitting
//       symbol -- a segmented sum; bit offsets are a prefix sum of the lengths.  Two block scans, then every thread
//       ORs its 32 symbols' bits into a shared-memory image of the stream, which is written out coalesced
//       (MSB-first bytes, zero-padded).  This is the warp/block scan formulation of "bitstream-offset compaction".
// The bitstream is bit-identical to the serial writer's by construction (same bits, same order).
#pragma once
#include "lc_encoder_par.cuh"
#include "lc_decoder_v2.cuh" // shared-window address helpers (lcv_sa_*)

#define LC_B2_THREADS 256
#define LC_B2_ITEMS (LC_PAR_MAX_SYMBOLS / LC_B2_THREADS)

// ---- B1: the (low, high) recurrence.  `pairs`: in (cum[s], cum[s+1]) per position; out: record in the first 8 bytes.
// Returns the first bit of finish_encoding (low's second-highest bit).
//
// The pairs are streamed through a shared-memory ring with cp.async: every lane copies one pair, so one instruction
// moves 32 symbols (512 contiguous bytes), LC_B1_AHEAD chunks ahead of the chunk being coded.  The first version loaded
// each symbol's pair into registers one group of four symbols ahead (every lane the same 16 bytes); the register
// rotation at the end of a group then waited for loads that had had only ~0.5 us to arrive -- 18 % of the kernel's
// stall samples were that one MOV (long scoreboard).
#define LC_B1_CHUNK 32
#define LC_B1_AHEAD 2
#define LC_B1_SMEM ((LC_B1_AHEAD + 1) * LC_B1_CHUNK * 16)
#ifdef LC_HOSTSIM
static inline void lc_b1_async16(lcv_sa dst, const void *src) { memcpy((void *)dst, src, 16); }
static inline void lc_b1_commit() {}
static inline void lc_b1_wait_ahead() {}
#else
static __device__ __forceinline__ void lc_b1_async16(lcv_sa dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
static __device__ __forceinline__ void lc_b1_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
static __device__ __forceinline__ void lc_b1_wait_ahead() // all but the LC_B1_AHEAD most recent groups have landed
{
    asm volatile("cp.async.wait_group %0;" ::"n"(LC_B1_AHEAD) : "memory");
}
#endif
// (uint32_t)int(x) for x in [-1, 2^32): Python's int() truncates toward zero; -1 (a collapsed interval) wraps
#ifdef LC_HOSTSIM
static inline uint32_t lc_b1_d2u(double x) { return (uint32_t)(long long)x; }
#else
static __device__ __forceinline__ uint32_t lc_b1_d2u(double x) { return x <= -1.0 ? 0xffffffffu : __double2uint_rz(x); }
#endif
__device__ __forceinline__ int lc_enc_b1_stream(int lane, double *pairs, int limit, char *smem)
{
    uint32_t lo = 0u, hi = 0xffffffffu;
    unsigned long long *rec = (unsigned long long *)pairs;
    const lcv_sa ring = lcv_sa_of(smem);
    const int nchunks = (limit + LC_B1_CHUNK - 1) / LC_B1_CHUNK;
    // one symbol of the recurrence
    // The chain per symbol is what this kernel costs (one warp per stream, ~185 cycles per symbol in the first version):
    // leading-zero counts from the exponent of an int->float conversion instead of FLO (lcvf_clz: 12 cycles against 24,
    // twice per symbol), the underflow count from the unshifted words (two dependent instructions fewer), 32-bit
    // double->unsigned conversions instead of 64-bit ones (19 cycles against 28).  int(range*c_hi - 1) is -1 when the
    // interval has collapsed (range 0): that one value is kept by hand, everything else is in [0, 2^32).
#define LC_B1_STEP(iv_, pos_)                                                                                       \
    do {                                                                                                            \
        /* encode_symbol (:220-224): high = low + int(range*c_hi - 1), low = low + int(range*c_lo) */               \
        const double rd_ = lc_ll2d_small((long long)hi - (long long)lo + 1); /* 0 when the interval has collapsed */ \
        const double xh_ = LC_DSUB(LC_DMUL(rd_, (iv_).y), 1.0), xl_ = LC_DMUL(rd_, (iv_).x);                        \
        const uint32_t ah_ = lc_b1_d2u(xh_), al_ = lc_b1_d2u(xl_);                                                  \
        hi = lo + ah_;                                                                                              \
        lo = lo + al_;                                                                                              \
        int d_ = lcvf_clz(lo ^ hi); /* leading bits low and high share: that many bits are emitted */               \
        d_ = d_ > 32 ? 32 : d_;                                                                                     \
        int e_ = lcvf_clz(~__funnelshift_lc(0u, lo & ~hi, d_ + 1)); /* underflow steps: low = 01.., high = 10.. */   \
        e_ = e_ > 32 ? 32 : e_;                                                                                     \
        if (lane == 0)                                                                                              \
            rec[2 * (pos_)] = (unsigned long long)hi | ((unsigned long long)(uint32_t)(d_ | (e_ << 8)) << 32);      \
        const uint32_t em_ = e_ ? 0x80000000u : 0u;                                                                 \
        lo = __funnelshift_lc(0u, lo, d_ + e_) & ~em_; /* (counts above 32 clamp: zeros / ones, like two steps) */   \
        hi = __funnelshift_lc(0xffffffffu, hi, d_ + e_) | em_;                                                      \
    } while (0)
#define LC_B1_ISSUE(k_)                                                                                             \
    do {                                                                                                            \
        const int idx_ = (k_) * LC_B1_CHUNK + lane;                                                                 \
        if ((k_) < nchunks && idx_ < limit)                                                                         \
            lc_b1_async16(ring + (uint32_t)((((k_) % (LC_B1_AHEAD + 1)) * LC_B1_CHUNK + lane) * 16), pairs + 2 * (size_t)idx_); \
        lc_b1_commit(); /* (an empty group when there is nothing left: the group count stays uniform) */           \
    } while (0)
    for (int k = 0; k < LC_B1_AHEAD; k++) LC_B1_ISSUE(k);
    for (int k = 0; k < nchunks; k++) {
        LC_B1_ISSUE(k + LC_B1_AHEAD);
        lc_b1_wait_ahead();
        __syncwarp(); // every lane's copy of chunk k is visible to the warp
        const lcv_sa base = ring + (uint32_t)((k % (LC_B1_AHEAD + 1)) * LC_B1_CHUNK * 16);
        const int p0 = k * LC_B1_CHUNK;
        const int cnt = limit - p0 < LC_B1_CHUNK ? limit - p0 : LC_B1_CHUNK;
        double2 cur = lcv_sa_ld128(base);
        for (int i = 0; i < cnt; i++) {
            // the next pair is requested before this one is coded (shared memory: ~30 cycles against ~230)
            const double2 nxt = lcv_sa_ld128(base + (uint32_t)((i + 1 < cnt ? i + 1 : i) * 16));
            LC_B1_STEP(cur, p0 + i);
            cur = nxt;
        }
        __syncwarp(); // the slot is refilled LC_B1_AHEAD + 1 chunks later, after every lane has read it
    }
#undef LC_B1_ISSUE
#undef LC_B1_STEP
    return (lo & 0x40000000u) != 0u ? 1 : 0;
}

// ---- B2 helpers: block-wide scans over LC_B2_THREADS threads (warp shuffles + one shared array)
__device__ __forceinline__ int lc_b2_excl_sum(int v, int *sm /*[8]*/, int &total)
{
    const int lane = (int)(threadIdx.x & 31), wid = (int)(threadIdx.x >> 5);
    int incl = v;
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(LC_FULL_MASK, incl, off);
        if (lane >= off) incl += t;
    }
    __syncthreads();
    if (lane == 31) sm[wid] = incl;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < LC_B2_THREADS / 32; w++) { const int x = sm[w]; if (w < wid) base += x; tot += x; }
    total = tot;
    return base + incl - v;
}
// segmented sum: element = (flag: the segment restarts inside this element, sum: e's after the last restart, or all
// of them).  Returns the pending count carried INTO this thread's chunk; `all` = the count after the last chunk.
__device__ __forceinline__ int lc_b2_excl_seg(int flag, int sum, int *smf, int *sms /*[8] each*/, int &all)
{
    const int lane = (int)(threadIdx.x & 31), wid = (int)(threadIdx.x >> 5);
    int f = flag, s = sum; // inclusive scan within the warp
    for (int off = 1; off < 32; off <<= 1) {
        const int pf = __shfl_up_sync(LC_FULL_MASK, f, off), ps = __shfl_up_sync(LC_FULL_MASK, s, off);
        if (lane >= off) { if (!f) s += ps; f |= pf; }
    }
    __syncthreads();
    if (lane == 31) { smf[wid] = f; sms[wid] = s; }
    __syncthreads();
    int cf = 0, cs = 0, af = 0, as = 0; // carry into this warp; aggregate of all warps
    for (int w = 0; w < LC_B2_THREADS / 32; w++) {
        const int wf = smf[w], ws = sms[w];
        if (w < wid) { cs = wf ? ws : cs + ws; cf |= wf; }
        as = wf ? ws : as + ws; af |= wf;
    }
    all = as; (void)af; (void)cf;
    // exclusive value for this lane: inclusive value of the previous lane (or the warp carry for lane 0)
    int pf = __shfl_up_sync(LC_FULL_MASK, f, 1), ps = __shfl_up_sync(LC_FULL_MASK, s, 1);
    if (lane == 0) { pf = 0; ps = 0; }
    return pf ? ps : cs + ps;
}

// OR the low nb bits of v (nb in 1..32) into the MSB-first bit image at bit offset `off`
__device__ __forceinline__ void lc_b2_or_bits(uint32_t *img, int off, uint32_t v, int nb)
{
    const int w = off >> 5, sh = 32 - (off & 31) - nb;
    if (sh >= 0) atomicOr(img + w, v << sh);
    else { atomicOr(img + w, v >> (-sh)); atomicOr(img + w + 1, v << (32 + sh)); }
}
__device__ __forceinline__ void lc_b2_or_ones(uint32_t *img, int off, int count)
{
    while (count > 0) {
        const int take = count > 32 ? 32 : count;
        lc_b2_or_bits(img, off, take == 32 ? 0xffffffffu : ((1u << take) - 1u), take);
        off += take; count -= take;
    }
}

// ---- B2: one block per stream.  smem: uint32 image[cap_words + 1] | int scratch[24]
__device__ __forceinline__ void lc_enc_b2_block(const double *pairs, int limit, int first, uint32_t *slot,
                                                uint32_t cap_words, int *nbits_out, int *status_out, int *fault_out,
                                                char *smem)
{
    uint32_t *img = (uint32_t *)smem;
    int *sc = (int *)(img + cap_words + 1);
    const unsigned long long *rec = (const unsigned long long *)pairs;
    const int tid = (int)threadIdx.x;
    const int p0 = tid * LC_B2_ITEMS;
    for (uint32_t i = tid; i < cap_words + 1; i += LC_B2_THREADS) img[i] = 0u;
    // pass 1: this chunk's length (pending count carried in taken as 0), whether it emits, pending count it leaves
    // (the e's before the chunk's first emitting symbol -- `head` -- join the count carried in: that symbol's run
    // is carry + head, added once the carry is known)
    int len0 = 0, has = 0, tail = 0, head = 0;
    for (int i = 0; i < LC_B2_ITEMS; i++) {
        const int p = p0 + i;
        if (p >= limit) break;
        const uint32_t meta = (uint32_t)(rec[2 * p] >> 32);
        const int d = (int)(meta & 0xffu), e = (int)((meta >> 8) & 0xffu);
        if (d) {
            if (!has) { head = tail; len0 += d; } else len0 += d + tail;
            has = 1; tail = e;
        } else tail += e;
    }
    if (!has) head = tail;
    int pend_all = 0;
    const int carry = lc_b2_excl_seg(has, has ? tail : head, sc, sc + 8, pend_all);
    const int len = len0 + (has ? carry + head : 0);
    int body_bits = 0;
    int off = lc_b2_excl_sum(len, sc + 16, body_bits);
    // finish_encoding (:230-245): outstanding += 1; first bit, then `outstanding` copies of its inverse
    const int fin_run = pend_all + 1;
    const long long total_bits = (long long)body_bits + 1 + fin_run;
    const bool ovf = (total_bits + 31) / 32 > (long long)cap_words;
    __syncthreads(); // image cleared by everyone
    if (ovf) {
        if (tid == 0) { *status_out = LC_OUT_OVERFLOW; *nbits_out = 0; *fault_out = limit; }
        return;
    }
    // pass 2: place the bits
    int pend = carry;
    for (int i = 0; i < LC_B2_ITEMS; i++) {
        const int p = p0 + i;
        if (p >= limit) break;
        const unsigned long long r = rec[2 * p];
        const uint32_t hi = (uint32_t)r, meta = (uint32_t)(r >> 32);
        const int d = (int)(meta & 0xffu), e = (int)((meta >> 8) & 0xffu);
        if (d) {
            const uint32_t b1 = hi >> 31;
            if (pend == 0) lc_b2_or_bits(img, off, hi >> (32 - d), d);
            else {
                if (b1) lc_b2_or_bits(img, off, 1u, 1);
                else lc_b2_or_ones(img, off + 1, pend);
                if (d > 1) lc_b2_or_bits(img, off + 1 + pend, (hi << 1) >> (33 - d), d - 1);
            }
            off += d + pend;
            pend = e;
        } else pend += e;
    }
    if (tid == 0) {
        if (first) lc_b2_or_bits(img, body_bits, 1u, 1);
        else lc_b2_or_ones(img, body_bits + 1, fin_run);
    }
    __syncthreads();
    const uint32_t nwords = (uint32_t)((total_bits + 31) / 32);
    for (uint32_t i = tid; i < nwords; i += LC_B2_THREADS) slot[i] = __byte_perm(img[i], 0, 0x0123); // MSB-first bytes
    if (tid == 0) { *nbits_out = (int)total_bits; *status_out = LC_OK; *fault_out = limit; }
}

// ---- block entry points
// B1: one warp per stream (blockDim.x = 32); first_out[b] = first bit of finish_encoding
// smem: LC_B1_SMEM bytes, 16-byte aligned
__device__ __forceinline__ void lc_enc_phase_b1_block(const LcCoderCfg &cfg, int B, const int *first_bad, double *ivs,
                                                      int *first_out, char *smem)
{
    const int lane = (int)(threadIdx.x & 31);
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        const int fb = first_bad[sidx];
        const int first = lc_enc_b1_stream(lane, ivs + 2 * (size_t)sidx * LC_PAR_MAX_SYMBOLS, fb < cfg.total ? fb : cfg.total,
                                           smem);
        if (lane == 0) first_out[sidx] = first;
        __syncwarp();
    }
}
// B2: one block of LC_B2_THREADS per stream; nbits[b] holds B1's first bit on entry
__device__ __forceinline__ void lc_enc_phase_b2_block(const LcCoderCfg &cfg, int B, const int *first_bad, const double *ivs,
                                                      unsigned char *out_slots, uint32_t slot_bytes, int *nbits, int *status,
                                                      int *fault, char *smem)
{
    for (int sidx = (int)blockIdx.x; sidx < B; sidx += (int)gridDim.x) {
        const int fb = first_bad[sidx];
        const int limit = fb < cfg.total ? fb : cfg.total;
        const int first = nbits[sidx];
        __syncthreads(); // everyone has read the first bit (and is done with the previous stream's image)
        lc_enc_b2_block(ivs + 2 * (size_t)sidx * LC_PAR_MAX_SYMBOLS, limit, first,
                        (uint32_t *)(out_slots + (size_t)sidx * slot_bytes), slot_bytes / 4, nbits + sidx, status + sidx,
                        fault + sidx, smem);
        __syncthreads();
        if (threadIdx.x == 0 && status[sidx] == LC_OK && fb < cfg.total) {
            // the serial encoder stops at the first out-of-range symbol
            status[sidx] = LC_BAD_SYMBOL; fault[sidx] = fb; nbits[sidx] = 0;
        }
    }
}
