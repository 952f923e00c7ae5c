// Stateful coder: cabac_encode / cabac_decode (cabac_compression.py:315-406) starting from a GIVEN ContextModel
// and leaving behind the model the reference object would hold after the call.
//
// The reference keeps one ContextModel per CABACCompressor and mutates it in compress, decompress and across
// calls (cabac_compression.py:438,478,517; SURVEY.md defect D5).  The fast kernels start every stream from a
// fresh model; this kernel is the compatibility path for callers that rely on the shared object: one warp, one
// stream, DENSE float64 vectors exactly as the reference stores them, in a direct-mapped table in global memory
//   valid[nkeys] (uint8) | counts[nkeys] (int32) | vectors[nkeys][n] (float64),   key = (left+1)*(n+1) + (up+1)
// (a single entry for the global context of non-3-D data).  The host scatters the given model into the table
// before the call and gathers the valid entries afterwards.  Per symbol it does what the reference does -- load the
// vector, sequential np.cumsum walk, update_model with the pairwise sum, store the vector -- so it is O(n) per symbol
// like the reference (a few microseconds; this path is about fidelity, not throughput).  Both coder modes.
#pragma once
#include "lc_encoder_par.cuh"

struct LcStatefulTable {
    unsigned char *valid;
    int *counts;
    double *vecs;
};

static inline LC_HD uint64_t lc_stateful_nkeys(int n, int has_ctx) { return has_ctx ? (uint64_t)(n + 1) * (n + 1) : 1u; }
static inline LC_HD uint64_t lc_stateful_off_counts(int n, int has_ctx) { return (lc_stateful_nkeys(n, has_ctx) + 255) & ~(uint64_t)255; }
static inline LC_HD uint64_t lc_stateful_off_vecs(int n, int has_ctx)
{
    return lc_stateful_off_counts(n, has_ctx) + ((lc_stateful_nkeys(n, has_ctx) * 4 + 255) & ~(uint64_t)255);
}
static inline LC_HD uint64_t lc_stateful_bytes(int n, int has_ctx)
{
    return lc_stateful_off_vecs(n, has_ctx) + lc_stateful_nkeys(n, has_ctx) * (uint64_t)n * 8;
}

// defaultdict lookup (:73,157): the context's vector into the dense image (a missing context is ones(n)/n and exists from now on)
__device__ __forceinline__ void lcs_open(LcWarp &W, const LcStatefulTable &T, uint32_t key)
{
    const bool have = T.valid[key] != 0;
    const double *v = T.vecs + (size_t)key * W.n;
    for (int i = W.lane; i < W.n; i += 32) W.dense[i] = have ? v[i] : W.u0;
    __syncwarp();
}
// context_models[ctx] = new_probs (:143); `updated`: update_model ran (context_counts += 1, :144)
__device__ __forceinline__ void lcs_close(LcWarp &W, const LcStatefulTable &T, uint32_t key, bool updated)
{
    double *v = T.vecs + (size_t)key * W.n;
    for (int i = W.lane; i < W.n; i += 32) v[i] = W.dense[i];
    if (W.lane == 0) { T.valid[key] = 1; if (updated) T.counts[key] += 1; }
    __syncwarp();
}

__device__ __forceinline__ long long lcs_encode_stream(LcWarp &W, const LcStatefulTable &T, const int *codes, uint32_t *out,
                                                       uint32_t cap_words, int *fault_index)
{
    LcBitWriter bw; lc_bw_init(bw, out, cap_words);
    long long low = 0, high = LC_FULL - 1, outstanding = 0;
    const long long fix = (W.mode == LC_MODE_VERBATIM) ? LC_FULL : LC_HALF; // defect D3
    const int RC = W.R * W.C;
    W.status = LC_OK;
    int pos = 0;
    for (; pos < W.total; pos++) {
        const int s = codes[pos];
        const int q = pos % RC, c = q % W.C, r = q / W.C;
        const uint32_t key = lc_ctx_key(W, c > 0 ? codes[pos - 1] : -1, r > 0 ? codes[pos - W.C] : -1);
        lcs_open(W, T, key);
        // the lookup inserts the context before probs[symbol] can raise (:342-350)
        if (s < 0 || s >= W.n) { W.status = LC_BAD_SYMBOL; lcs_close(W, T, key, false); break; }
        LcInterval iv;
        iv.sym = s; iv.exact = 1;
        iv.clo = lc_dense_prefix(W.dense, s);
        iv.chi = LC_DADD(iv.clo, W.dense[s]);
        lc_interval_apply(iv, 0.0, low, high);
        // _renormalize_encoder (:189-202)
        while ((high & LC_HALF) == (low & LC_HALF)) {
            const long long bit = high >> 31;
            if (bit < 0 || bit > 1 || (outstanding > 0 && (1 - bit) < 0)) { W.status = LC_ENC_BIT_OVERFLOW; break; }
            lc_bw_put(bw, (int)bit, 1, W.lane);
            if (outstanding > 0) lc_bw_put(bw, (int)(1 - bit), outstanding, W.lane);
            outstanding = 0;
            low = (low << 1) & (LC_FULL - 1);
            high = ((high << 1) & (LC_FULL - 1)) | 1;
        }
        if (W.status == LC_OK) // _handle_underflow (:204-210)
            while ((low & LC_QUARTER) != 0 && (high & LC_QUARTER) == 0) {
                outstanding += 1;
                low = (low << 1) & (LC_HALF - 1);
                high = ((high << 1) & (LC_HALF - 1)) | fix | 1;
            }
        if (W.status == LC_OK && bw.ovf) W.status = LC_OUT_OVERFLOW;
        if (W.status != LC_OK) { lcs_close(W, T, key, false); break; } // the lookup already inserted the context
        lc_dense_update(W, s);
        lcs_close(W, T, key, true);
    }
    *fault_index = pos;
    if (W.status != LC_OK) return 0;
    // finish_encoding (:230-245)
    outstanding += 1;
    const int first = (low & LC_QUARTER) != 0 ? 1 : 0;
    lc_bw_put(bw, first, 1, W.lane);
    lc_bw_put(bw, 1 - first, outstanding, W.lane);
    lc_bw_finish(bw, W.lane);
    if (bw.ovf) { W.status = LC_OUT_OVERFLOW; return 0; }
    return bw.nbits;
}

__device__ __forceinline__ void lcs_decode_stream(LcWarp &W, const LcStatefulTable &T, const unsigned char *src,
                                                  long long nbytes, int *out, int *fault_index)
{
    LcBitReader br; lc_br_init(br, src, nbytes, W.lane);
    long long low = 0, high = LC_FULL - 1, code = 0;
    for (int i = 0; i < 32; i++) code = (code << 1) | lc_br_bit(br, W.lane); // start_decoding (:247-258)
    const long long fix = (W.mode == LC_MODE_VERBATIM) ? LC_FULL : LC_HALF;
    const int RC = W.R * W.C;
    W.status = LC_OK;
    int pos = 0;
    for (; pos < W.total; pos++) {
        const int q = pos % RC, c = q % W.C, r = q / W.C;
        // the decoded symbols so far are this warp's own writes (ordered by the __syncwarp of the previous close)
        const uint32_t key = lc_ctx_key(W, c > 0 ? out[pos - 1] : -1, r > 0 ? out[pos - W.C] : -1);
        lcs_open(W, T, key);
        // decode_symbol (:272-311)
        const long long range = high - low + 1;
        if (range == 0) { W.status = LC_DEC_ZERO_RANGE; lcs_close(W, T, key, false); break; }
        double v = LC_DDIV(LC_DMUL(LC_LL2D(code - low + 1), 1.0), LC_LL2D(range));
        v = LC_DSUB(v, 1e-10);
        LcInterval iv = lc_exact_search_dec(W.dense, W.n, v);
        if (iv.sym >= W.n) { W.status = LC_DEC_SYMBOL_OOB; lcs_close(W, T, key, false); break; }
        if (iv.sym < 0) { iv.clo = lc_exact_cum_total(W.dense, W.n); iv.chi = 0.0; } // cum[-1], cum[0] (:291-292)
        lc_interval_apply(iv, 0.0, low, high);
        while ((high & LC_HALF) == (low & LC_HALF)) {
            low = (low << 1) & (LC_FULL - 1);
            high = ((high << 1) & (LC_FULL - 1)) | 1;
            code = ((code << 1) & (LC_FULL - 1)) | lc_br_bit(br, W.lane);
        }
        while ((low & LC_QUARTER) != 0 && (high & LC_QUARTER) == 0) {
            low = (low << 1) & (LC_HALF - 1);
            high = ((high << 1) & (LC_HALF - 1)) | fix | 1;
            if (W.mode == LC_MODE_VERBATIM) code = ((code ^ LC_QUARTER) << 1) | lc_br_bit(br, W.lane);
            else code = (((code ^ LC_QUARTER) << 1) & (LC_FULL - 1)) | lc_br_bit(br, W.lane);
        }
        if (W.lane == 0) out[pos] = iv.sym;
        lc_dense_update(W, iv.sym);
        lcs_close(W, T, key, true); // (its __syncwarp also orders the symbol store before the next key reads)
    }
    *fault_index = pos;
    for (int z = pos + W.lane; z < W.total; z += 32) out[z] = 0;
}
