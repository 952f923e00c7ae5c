/*
 * latentcodec.h -- C ABI of liblatentcodec.so: the B200 (sm_100a) implementation of the latent
 * quantise -> entropy-code -> decode -> dequantise path of yubster4525/image_compression_2.
 *
 * The reference has no FFI; its boundary is the Python call surface.  Each
 * entry point below names the reference code it replaces (file:line under /root/reference).  The
 * Python drop-in classes in image_compression_2_b200/ bind these through ctypes; INTEGRATION.md
 * shows the binding a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name says host; sizes are in elements unless
 *    they say bytes; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *  - all calls are asynchronous on `stream`, launch on the CURRENT device (cudaSetDevice it to the device the
 *    pointers live on) and never synchronise.  The library keeps no mutable state between calls: the only statics are
 *    write-once caches of per-device facts (SM count, "opt-in shared-memory attributes set on this device") keyed by
 *    cudaGetDevice() and updated with atomics, and a thread-local launch counter (lc_debug_launch_count).  It reads no
 *    environment variable; behaviour switches are the `flags` argument of the _t entry points.  Any number of host
 *    threads may call into it concurrently, on the same or on different devices, each with its own scratch;
 *  - return value: 0 on success, a negative errno-style value (-EINVAL = -22) for arguments the
 *    implementation does not support, or -1000 - cudaError for a CUDA launch error;
 *  - per-stream faults of the coder are reported in `status[b]` / `fault_index[b]` (LC_STATUS_*),
 *    mirroring the exceptions the reference raises; kernels never trap.
 */
#ifndef LATENTCODEC_H
#define LATENTCODEC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LC_ABI_VERSION 2 /* 2: typed index arrays (_t entry points), flags, lc_debug_launch_count */

/* coder modes */
#define LC_CODER_VERBATIM 0 /* /root/reference/cabac_compression.py as shipped (defect D3 kept)      */
#define LC_CODER_REPAIRED 1 /* `| full_range` -> `| half_range` at :210/:308, 32-bit mask at :309    */

/* per-stream status words */
#define LC_STATUS_OK 0
#define LC_STATUS_ENC_BIT_OVERFLOW 1 /* reference: ValueError, bytearray.append out of range (:193-197) */
#define LC_STATUS_DEC_SYMBOL_OOB 2   /* reference: IndexError at cumulative_probs[symbol+1] (:291)       */
#define LC_STATUS_DEC_ZERO_RANGE 3   /* reference: ZeroDivisionError (:285)                              */
#define LC_STATUS_DEC_NEG_SYMBOL 4   /* not produced since ABI 2: symbol -1 is followed like the reference  */
#define LC_STATUS_OUT_OVERFLOW 5     /* per-stream output slot too small                                 */
#define LC_STATUS_BAD_SYMBOL 6       /* input index outside [0,n_symbols) (reference: IndexError :343)   */
#define LC_STATUS_POOL_OVERFLOW 7    /* internal scratch exhausted (never expected)                      */

/* `flags` of lc_encode_batch_t / lc_decode_batch_t: 0 = let the library choose.  The choices never change a result,
 * only which kernel produces it (tests use them to drive every kernel over the same inputs). */
#define LC_FLAG_DEC_LATENCY_BUILD 1     /* decoder v2: the 8-streams-per-SM build whatever the batch size            */
#define LC_FLAG_DEC_THROUGHPUT_BUILD 2  /* decoder v2: the 10-streams-per-SM build                                    */
#define LC_FLAG_DEC_GENERIC_SHAPE 4     /* decoder v2: not the kernel specialised for 8-bit [16,512] W+ latents       */
#define LC_FLAG_DEC_REGISTER_MODEL 8    /* the register-model decoder (the n > 256 kernel) also for n <= 256         */
#define LC_FLAG_DEC_SERIAL 16           /* the generic serial decoder only                                            */
#define LC_FLAG_ENC_SERIAL 32           /* the serial warp-per-stream encoder instead of the phase-split one          */
#define LC_FLAG_DEC_NO_SMALL 64         /* not the dense shared-memory decoder of alphabets up to 32 symbols          */
#define LC_FLAG_ENC_SORT_V1 128         /* phase S of the phase-split encoder with cub::BlockRadixSort (round 1)     */
#define LC_FLAG_DEBUG_DEC_V3 256        /* builds with -DLC_DEBUG_VARIANTS only: three-warp decoder                   */
#define LC_FLAG_DEBUG_ENC_DENSE_PHASE_A 512 /* builds with -DLC_DEBUG_VARIANTS only: dense warp-per-context phase A   */

int lc_version(void);

/* kernels launched by the calling host thread since the last reset (diagnostic; bench.py's gpu_launches) */
int64_t lc_debug_launch_count(int reset);

/* Index arrays: the reference hands the coder `.astype(np.int32)` codes (cabac_compression.py:471), so every entry
 * point exists in an int32 form.  The _t forms take `idx_bytes` = 4 (int32), 2 (uint16) or 1 (uint8; alphabets up to
 * 256 symbols): 1 or 2 bytes of HBM traffic per symbol instead of 4. */

/* ---- quantisers ------------------------------------------------------------------------------ */

/* Quantiser A + dequantiser A: StyleGAN3Compressor.compress, stylegan3_hvae_full.py:313-316.
 * idx = round_half_even(((w+1)*0.5)*(2^bits-1)) (no clamp), wq = idx/(2^bits-1)*2-1, five separately
 * rounded fp32 operations.  idx_out and wq_out may each be NULL. */
int lc_quantize_affine(const float *w, int64_t n_elem, int bits, int32_t *idx_out, float *wq_out, void *stream);
/* ... with uint16 / uint8 indices.  The narrow forms hold what the coder consumes -- the index CLAMPED to the alphabet
 * [0, 2^bits-1] (NaN -> 0, +-inf clamped) -- because the unclamped value of an out-of-range latent does not fit them; wq_out is
 * computed from the unclamped index exactly as above.  idx_bytes = 4 is lc_quantize_affine. */
int lc_quantize_affine_t(const float *w, int64_t n_elem, int bits, void *idx_out, int idx_bytes, float *wq_out,
                         void *stream);

/* Dequantiser A alone (same lines): w_out = idx/(2^bits-1)*2-1. */
int lc_dequantize_affine(const int32_t *idx, int64_t n_elem, int bits, float *w_out, void *stream);
int lc_dequantize_affine_t(const void *idx, int idx_bytes, int64_t n_elem, int bits, float *w_out, void *stream);

/* Quantiser B: GumbelSoftmaxDiscretization.forward -> encoding_indices,
 * gumbel_softmax_compression.py:97,118: argmin_k |z - codebook[k]| in fp32, first minimum.
 * `codebook` (n fp32, n <= 4096) comes from the host's torch.linspace (:49-52); `sorted_ascending`
 * != 0 enables the search path (otherwise all n entries are scanned).  deq_out (may be NULL)
 * receives codebook[idx] (cabac_compression.py:531). */
int lc_quantize_codebook(const float *z, int64_t n_elem, const float *codebook, int n, int sorted_ascending,
                         int32_t *idx_out, float *deq_out, void *stream);
int lc_quantize_codebook_t(const float *z, int64_t n_elem, const float *codebook, int n, int sorted_ascending,
                           void *idx_out, int idx_bytes, float *deq_out, void *stream);

/* Dequantiser B: codebook[idx], cabac_compression.py:531 / gumbel_softmax_compression.py:258.
 * Indices outside [0,n) produce NaN. */
int lc_dequantize_codebook(const int32_t *idx, int64_t n_elem, const float *codebook, int n, float *w_out,
                           void *stream);
int lc_dequantize_codebook_t(const void *idx, int idx_bytes, int64_t n_elem, const float *codebook, int n, float *w_out,
                             void *stream);

/* ---- entropy coder --------------------------------------------------------------------------- */

/* A launch codes B independent streams; each stream has `imgs` images of R x C symbols that share
 * one coder and one FRESH ContextModel (imgs > 1 reproduces the reference's batched
 * cabac_encode(data[B,R,C]) call, cabac_compression.py:330-337).  has_ctx = 0 selects the single
 * global context the reference uses for non-3-D shapes (:115-117).  n_symbols must be a power of
 * two in [2,1024]; imgs*R*C <= 2^22; C <= 8192 when has_ctx. */

/* bytes of device scratch lc_encode_batch / lc_decode_batch need for these arguments */
int64_t lc_coder_scratch_bytes(int B, int imgs, int R, int C, int n_symbols, int has_ctx);

/* recommended per-stream output slot size in bytes (multiple of 16) */
int64_t lc_encode_slot_bytes(int imgs, int R, int C, int n_symbols);

/* cabac_encode, cabac_compression.py:315-359 (+ ContextModel :60-162, ArithmeticCoder :166-245).
 *   idx        int32 [B][imgs*R*C]
 *   slots      workspace, B * slot_bytes bytes; stream b is written MSB-first packed at
 *              slots + b*slot_bytes (slot_bytes a multiple of 16)
 *   out_bytes  may be NULL.  Otherwise the streams are compacted into it: stream b starts at
 *              out_offsets[b] (16-byte aligned), out_offsets[B] = total bytes used; streams that
 *              faulted take no space.  If out_offsets[B] would exceed out_capacity nothing is
 *              copied for the streams that do not fit and their status becomes OUT_OVERFLOW.
 *   out_offsets int64 [B+1] (required iff out_bytes != NULL)
 *   out_nbits  int32 [B]: the number of bits the reference encoder appends; bytes = ceil(nbits/8)
 *   status, fault_index  int32 [B]; fault_index = symbols completely coded before the fault */
int lc_encode_batch(const int32_t *idx, int B, int imgs, int R, int C, int n_symbols, double adaptation_rate,
                    int mode, int has_ctx, void *scratch, int64_t scratch_bytes, uint8_t *slots, int64_t slot_bytes,
                    uint8_t *out_bytes, int64_t out_capacity, int64_t *out_offsets, int32_t *out_nbits,
                    int32_t *status, int32_t *fault_index, void *stream);
/* the same with idx of idx_bytes-byte elements and `flags` (LC_FLAG_ENC_*) */
int lc_encode_batch_t(const void *idx, int idx_bytes, int B, int imgs, int R, int C, int n_symbols,
                      double adaptation_rate, int mode, int has_ctx, void *scratch, int64_t scratch_bytes, uint8_t *slots,
                      int64_t slot_bytes, uint8_t *out_bytes, int64_t out_capacity, int64_t *out_offsets,
                      int32_t *out_nbits, int32_t *status, int32_t *fault_index, int flags, void *stream);

/* cabac_decode, cabac_compression.py:363-406 (+ ArithmeticCoder :247-311).
 *   bytes, offsets[B], nbits[B]: stream b = ceil(nbits[b]/8) bytes at bytes + offsets[b];
 *              offsets[b] must be a multiple of 4 and the buffer readable to the next multiple of 4
 *   idx_out    int32 [B][imgs*R*C], zeros from the fault position on.  A decoded symbol can be -1 (scaled value
 *              <= 0, from an already inconsistent coder state): the reference does not fault there, it goes on with
 *              NumPy's negative indexing (:288-292,403) and so do the kernels; -1 is stored (255 / 65535 in the
 *              narrow index types) and dequantises to the LAST table entry, as codebook[-1] does
 *   deq_table / deq_out: optional fused dequantiser B (deq_out = deq_table[idx], fp32), may be NULL */
int lc_decode_batch(const uint8_t *bytes, const int64_t *offsets, const int32_t *nbits, int B, int imgs, int R, int C,
                    int n_symbols, double adaptation_rate, int mode, int has_ctx, void *scratch,
                    int64_t scratch_bytes, int32_t *idx_out, const float *deq_table, float *deq_out, int32_t *status,
                    int32_t *fault_index, void *stream);
/* the same with idx_out of idx_bytes-byte elements -- or NULL when only the dequantised values are wanted (then
 * deq_out is required) -- and `flags` (LC_FLAG_DEC_*) */
int lc_decode_batch_t(const uint8_t *bytes, const int64_t *offsets, const int32_t *nbits, int B, int imgs, int R, int C,
                      int n_symbols, double adaptation_rate, int mode, int has_ctx, void *scratch,
                      int64_t scratch_bytes, void *idx_out, int idx_bytes, const float *deq_table, float *deq_out,
                      int32_t *status, int32_t *fault_index, int flags, void *stream);

/* ---- stateful coder: the reference's shared ContextModel (cabac_compression.py:438,478,517) --------------------- */

/* The reference mutates one ContextModel object in cabac_encode and cabac_decode and across calls.  These two entry
 * points code ONE stream (imgs images of R x C sharing coder and model, or has_ctx = 0) starting from the model held
 * in `table` and leave the model the reference object would hold afterwards in it.  Table layout (device memory,
 * lc_stateful_table_bytes bytes, key = (left+1)*(n+1) + (up+1) with -1 sentinels, one key when has_ctx = 0):
 *   uint8 valid[nkeys]                      at byte 0            1 = the context exists in context_models
 *   int32 counts[nkeys]                     at lc_stateful_table_offset(..., 1)   context_counts
 *   float64 vectors[nkeys][n]               at lc_stateful_table_offset(..., 2)   the probability vectors
 * The caller zeroes valid/counts, scatters the given model in, and gathers the valid entries after the call.
 * n_symbols: power of two up to 1024.  With (left,up) contexts the table is direct-mapped, (n+1)^2 * n * 8 bytes: 135 MB
 * at 256 symbols, 8.6 GB at 1024 -- sized for the 180 GB of a B200; only the vectors of contexts that exist are ever
 * touched (the valid flags say which), and nothing but valid[] and counts[] needs clearing. */
int64_t lc_stateful_table_bytes(int n_symbols, int has_ctx);
int64_t lc_stateful_table_offset(int n_symbols, int has_ctx, int which);

/* cabac_encode (:315-359) from/into the table's model.  slot: output, slot_bytes (multiple of 4) bytes. */
int lc_stateful_encode(const int32_t *idx, int imgs, int R, int C, int n_symbols, double adaptation_rate, int mode,
                       int has_ctx, void *table, int64_t table_bytes, uint8_t *slot, int64_t slot_bytes,
                       int32_t *out_nbits, int32_t *status, int32_t *fault_index, void *stream);

/* cabac_decode (:363-406) from/into the table's model.  bytes: 4-byte aligned, readable to a multiple of 4. */
int lc_stateful_decode(const uint8_t *bytes, int64_t nbytes, int imgs, int R, int C, int n_symbols,
                       double adaptation_rate, int mode, int has_ctx, void *table, int64_t table_bytes, int32_t *idx_out,
                       int32_t *status, int32_t *fault_index, void *stream);

/* ContextModel.update_model (cabac_compression.py:119-144) on ONE probability vector: vec[n_symbols] float64 in device
 * memory is replaced by the vector after observing `symbol` -- p[s] += rate*(1-p[s]), the others scaled by
 * (1-p[s]) / (pairwise-sum(p) - p[s]).  A negative symbol (>= -n_symbols) follows NumPy's negative indexing as the
 * reference does after decoding symbol -1 (:288-292,403): element n+symbol gets the increment and EVERY element is
 * scaled (`i != symbol` never excludes it).  n_symbols: power of two in [2,1024]. */
int lc_model_update(double *vec, int n_symbols, int symbol, double adaptation_rate, void *stream);

/* number of thread blocks (= resident streams) the coder kernels launch for B streams */
int lc_coder_grid(int B, int imgs, int R, int C, int n_symbols, int has_ctx);

#ifdef __cplusplus
}
#endif
#endif /* LATENTCODEC_H */
