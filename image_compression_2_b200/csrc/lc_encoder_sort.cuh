// Phase S of the parallel encoder, second version: positions of one stream grouped by context with a two-pass stable
// radix sort built from warp primitives (ballot / popc), one 256-thread block per stream.
//
// The first version (lc_enc_sort_kernel: cub::BlockRadixSort over (key, position) pairs, 1024 threads, 8 items per
// thread, 63 registers) fills the register file of an SM with ONE block, so no other kernel's blocks -- in particular
// the decoder blocks of the previous batch when several batches are in flight (LatentPipeline.roundtrip_host_stream)
// -- can share the SM with it; it also sorts in five 4-bit passes.  Here:
//   * the context key of every position ((left+1)(n+1)+(up+1) < 2^21, get_context :78-117) is computed once into
//     shared memory; the sort moves 16-bit POSITIONS only and looks the key up;
//   * stable passes of 9 bits from the low end (two for alphabets up to 256 symbols): each warp owns a contiguous eighth of the
//     sequence and walks it 32 items at a time; nine ballots on the digit's bits give every item its peers in the tile,
//     the rank among them is a popcount, and the lowest peer advances the warp's own counter of that digit -- no atomics,
//     no block-wide scan per tile, stable by construction (tiles in order, lanes in order);
//   * between the counting and the scattering walk one block-wide exclusive scan over (digit, warp) turns the counts
//     into offsets;
//   * the group structure the rest of phase S needs (first visit / second visit / head of a context visited at least
//     three times) is one bit per sorted item -- "key differs from the item before" -- built with ballots.
// 256 threads x 40 registers and 58 KB of shared memory: ten kilo-registers, which fits beside seven resident decoder
// blocks.  Results (sorted keys, sorted positions, work list, first/second-visit intervals) are identical to the first
// version's: same stable order.
#pragma once

#define LCS2_THREADS 256
#define LCS2_WARPS (LCS2_THREADS / 32)
#define LCS2_N LC_PAR_MAX_SYMBOLS               // 8192 slots per stream (positions past the stream sort last)
#define LCS2_CHUNK (LCS2_N / LCS2_WARPS)        // items per warp
#define LCS2_SMEM (LCS2_N * 4 + LCS2_N * 2 + LCS2_WARPS * 512 * 2 + 256 * 4 + 256 * 4 + 64)

// lanes of the warp whose 9-bit digit equals this lane's: nine ballots, one per bit (independent, so they pipeline: ~50
// cycles).  `__match_any_sync` does the same in one instruction but takes time proportional to the number of distinct
// values -- ~800 cycles per call with ~30 distinct digits among 32 lanes, 35 % of the first version of this kernel.
__device__ __forceinline__ uint32_t lcs2_peers(uint32_t d)
{
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; b++) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(LC_FULL_MASK, bit);
        peers &= bit ? m : ~m;
    }
    return peers;
}

// one stable counting-sort pass over the block's items: item j of the input sequence is position `in[j]` (or j itself
// when in == nullptr), its digit is (keys[position] >> shift) & mask; writes the positions in sorted order to out
__device__ __forceinline__ void lcs2_pass(const uint32_t *keys, const unsigned short *in, unsigned short *out_smem,
                                          unsigned short *out_glob, uint32_t *out_keys, unsigned short *hist,
                                          uint32_t *scan_tmp, int shift, uint32_t mask, int nbins)
{
    const int tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned short *myh = hist + w * 512;
    for (int i = tid; i < LCS2_WARPS * 512; i += LCS2_THREADS) hist[i] = 0;
    __syncthreads();
    // ---- count: every warp its own contiguous chunk, tile by tile
    for (int t = 0; t < LCS2_CHUNK / 32; t++) {
        const int j = w * LCS2_CHUNK + t * 32 + lane;
        const int p = in ? (int)in[j] : j;
        const uint32_t d = (keys[p] >> shift) & mask;
        const uint32_t peers = lcs2_peers(d);
        if ((peers & lt) == 0u) myh[d] = (unsigned short)(myh[d] + __popc(peers)); // the lowest peer
        __syncwarp();
    }
    __syncthreads();
    // ---- offsets: exclusive scan in (digit, warp) order.  Thread t owns bins [t*per, (t+1)*per)
    {
        const int per = nbins / LCS2_THREADS > 0 ? nbins / LCS2_THREADS : 1; // 2 (512 bins) or 1 (256 bins)
        const int b0 = tid * per;
        uint32_t tot = 0u;
        if (b0 < nbins)
            for (int b = b0; b < b0 + per; b++)
                for (int ww = 0; ww < LCS2_WARPS; ww++) tot += hist[ww * 512 + b];
        // block-wide exclusive scan of tot
        uint32_t inc = tot;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(LC_FULL_MASK, inc, o); if (lane >= o) inc += x; }
        if (lane == 31) scan_tmp[w] = inc;
        __syncthreads();
        uint32_t base = 0u;
        for (int ww = 0; ww < w; ww++) base += scan_tmp[ww];
        uint32_t run = base + inc - tot;
        if (b0 < nbins)
            for (int b = b0; b < b0 + per; b++)
                for (int ww = 0; ww < LCS2_WARPS; ww++) {
                    const uint32_t c = hist[ww * 512 + b];
                    hist[ww * 512 + b] = (unsigned short)run;
                    run += c;
                }
    }
    __syncthreads();
    // ---- scatter
    for (int t = 0; t < LCS2_CHUNK / 32; t++) {
        const int j = w * LCS2_CHUNK + t * 32 + lane;
        const int p = in ? (int)in[j] : j;
        const uint32_t d = (keys[p] >> shift) & mask;
        const uint32_t peers = lcs2_peers(d);
        const uint32_t o = myh[d];
        __syncwarp();
        if ((peers & lt) == 0u) myh[d] = (unsigned short)(o + __popc(peers));
        __syncwarp();
        const uint32_t dst = o + __popc(peers & lt);
        if (out_smem) out_smem[dst] = (unsigned short)p;
        if (out_glob) out_glob[dst] = (unsigned short)p;
        if (out_keys) out_keys[dst] = keys[p];
    }
    __syncthreads();
}

// Phase S for the stream of this block.  Same outputs as lc_enc_sort_kernel.
__device__ __forceinline__ void lcs2_block(const LcCoderCfg &cfg, LcCodes codes, uint32_t *skeys_all,
                                           unsigned short *spos_all, int *__restrict__ first_bad,
                                           unsigned short *__restrict__ glist_all, int *__restrict__ ngroups,
                                           double *__restrict__ ivs_all, const double *__restrict__ tables, char *smem)
{
    uint32_t *keys = (uint32_t *)smem;                                   // [8192] key of position p
    unsigned short *posA = (unsigned short *)(keys + LCS2_N);            // [8192] positions after pass 1
    unsigned short *hist = posA + LCS2_N;                                // [8][512]
    uint32_t *flags = (uint32_t *)(hist + LCS2_WARPS * 512);             // [256] bit j&31 of word j>>5: item j opens a group
    uint32_t *cnt = flags + 256;                                         // [256] listed groups per tile, then their offsets
    uint32_t *scan_tmp = cnt + 256;                                      // [8]
    __shared__ int s_first_bad;
    const int tid = (int)threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int total = cfg.total, n = cfg.n, C = cfg.C, RC = cfg.R * cfg.C;
    const LcCodes c = codes + (size_t)blockIdx.x * total;
    if (tid == 0) s_first_bad = total;
    __syncthreads();
    for (int p = tid; p < total; p += LCS2_THREADS) {
        const int s = c[p];
        if (s < 0 || s >= n) atomicMin(&s_first_bad, p);
    }
    __syncthreads();
    const int fb = s_first_bad;
    // real keys are below (n+1)^2 < 2^key_bits; the key of a slot without a symbol (all ones in those bits) sorts last
    const int key_bits = 32 - __clz((n + 1) * (n + 1));
    const uint32_t pad = (1u << key_bits) - 1u;
    const bool one_image = RC >= total;                 // p % RC == p
    const int c_shift = (C & (C - 1)) == 0 ? 31 - __clz(C) : -1; // row length a power of two: shifts instead of divisions
    for (int p = tid; p < LCS2_N; p += LCS2_THREADS) {
        uint32_t k = pad;
        if (p < fb) {
            const int q = one_image ? p : p % RC;
            const int cc = c_shift >= 0 ? (q & (C - 1)) : q % C, rr = c_shift >= 0 ? (q >> c_shift) : q / C;
            const int left = cc > 0 ? c[p - 1] : -1;
            const int up = rr > 0 ? c[p - C] : -1;
            k = (uint32_t)(left + 1) * (uint32_t)(n + 1) + (uint32_t)(up + 1);
        }
        keys[p] = k;
    }
    __syncthreads();
    unsigned short *spos = spos_all + (size_t)blockIdx.x * LC_PAR_MAX_SYMBOLS;
    uint32_t *skeys = skeys_all + (size_t)blockIdx.x * LC_PAR_MAX_SYMBOLS;
    // stable passes of 9 bits from the low end: 2 for alphabets up to 256 symbols (keys < 2^18), 3 up to 1024; the
    // sequence alternates between the shared-memory buffer and the stream's slot of the global `spos` array
    {
        const unsigned short *cur = (const unsigned short *)0;
        bool in_smem = false;
        for (int shift = 0, i = 0; shift < key_bits; shift += 9, i++) {
            in_smem = (i & 1) == 0;
            uint32_t *ok = shift + 9 >= key_bits ? skeys : (uint32_t *)0; // the last pass also leaves the sorted keys
            if (in_smem) lcs2_pass(keys, cur, posA, (unsigned short *)0, ok, hist, scan_tmp, shift, 511u, 512);
            else lcs2_pass(keys, cur, (unsigned short *)0, spos, ok, hist, scan_tmp, shift, 511u, 512);
            cur = in_smem ? posA : spos;
        }
        if (in_smem) {
            for (int j = tid; j < LCS2_N; j += LCS2_THREADS) spos[j] = posA[j];
            __syncthreads();
        }
    }
    if (tid == 0) first_bad[blockIdx.x] = fb;
    // (the block reads back its own global writes: ordered by the __syncthreads that ends the pass)
    // ---- group boundaries: one bit per sorted item, and the sorted keys for phase A
    for (int T = w; T < LCS2_N / 32; T += LCS2_WARPS) {
        const int j = T * 32 + lane;
        const uint32_t k0 = skeys[j];
        const uint32_t km1 = j > 0 ? skeys[j - 1] : 0xFFFFFFFFu;
        const uint32_t m = __ballot_sync(LC_FULL_MASK, km1 != k0);
        if (lane == 0) flags[T] = m;
    }
    __syncthreads();
    for (int j = fb + tid; j < LCS2_N; j += LCS2_THREADS) skeys[j] = LC_PAR_KEY_PAD; // (slots without a symbol: the tail)
    const double u0 = LC_DDIV(1.0, (double)n);
    const double *cum1 = tables ? tables + 64 : (const double *)0;
    double *ivb = ivs_all + (size_t)blockIdx.x * 2 * LC_PAR_MAX_SYMBOLS;
    auto opens = [&](int j) -> bool { return (flags[j >> 5] >> (j & 31)) & 1u; };
    // ---- what needs no model (first visits: uniform model; second visits: table of exact cumsums after one update)
    // and the heads of the contexts phase A has to run (visited at least three times; twice without the table)
    for (int T = w; T < LCS2_N / 32; T += LCS2_WARPS) {
        const int j = T * 32 + lane;
        const bool valid = j < fb; // positions before the first bad symbol; the others sort behind them
        const bool head = valid && opens(j);
        bool listed;
        if (head) {
            const int p = spos[j];
            const int sy = c[p];
            reinterpret_cast<double2 *>(ivb)[p] = make_double2(LC_DMUL((double)sy, u0), LC_DMUL((double)(sy + 1), u0));
        }
        if (cum1) {
            const bool second = valid && !head && (j == 1 || opens(j - 1));
            if (second) { // model after one update with the first visit's symbol
                const int p = spos[j];
                const int sy = c[p], s1 = c[spos[j - 1]];
                const double *row = cum1 + (size_t)s1 * (n + 1);
                reinterpret_cast<double2 *>(ivb)[p] = make_double2(row[sy], row[sy + 1]);
            }
            listed = head && (j + 2 < fb) && !opens(j + 1) && !opens(j + 2);
        } else {
            listed = head && (j + 1 < fb) && !opens(j + 1);
        }
        const uint32_t m = __ballot_sync(LC_FULL_MASK, listed);
        if (lane == 0) cnt[T] = m;
    }
    __syncthreads();
    // ---- work list in ascending sorted index: exclusive scan of the tiles' counts (tile T = thread T)
    {
        const uint32_t m = cnt[tid];
        const uint32_t my = (uint32_t)__popc(m);
        uint32_t inc = my;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t x = __shfl_up_sync(LC_FULL_MASK, inc, o); if (lane >= o) inc += x; }
        if (lane == 31) scan_tmp[w] = inc;
        __syncthreads();
        uint32_t base = 0u, all = 0u;
        for (int ww = 0; ww < LCS2_WARPS; ww++) { if (ww < w) base += scan_tmp[ww]; all += scan_tmp[ww]; }
        uint32_t off = base + inc - my;
        unsigned short *gl = glist_all + (size_t)blockIdx.x * LC_PAR_MAX_GROUPS;
        uint32_t mm = m;
        while (mm) {
            const int b = __ffs((int)mm) - 1;
            mm &= mm - 1;
            gl[off++] = (unsigned short)(tid * 32 + b);
        }
        if (tid == 0) ngroups[blockIdx.x] = (int)all;
    }
}
