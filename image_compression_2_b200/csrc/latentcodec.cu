// liblatentcodec.so -- CUDA kernels (sm_100a) and the C ABI declared in include/latentcodec.h.
// Build: see image_compression_2_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a
// --fmad=false -lineinfo).  No tensor cores: nothing on this path is a dense contraction.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <atomic>

#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

#include "../../include/latentcodec.h"
#include "lc_coder.cuh"
#include "lc_encoder_par.cuh"
#include "lc_decoder_fast.cuh"
#include "lc_decoder_v2.cuh"
#include "lc_decoder_small.cuh"
#ifdef LC_DEBUG_VARIANTS
#include "lc_decoder_v3.cuh"
#endif
#include "lc_encoder_sparse.cuh"
#include "lc_encoder_sort.cuh"
#include "lc_encoder_pack.cuh"
#include "lc_stateful.cuh"

// ---- per-device facts (immutable once read, so caching them is not mutable library state): SM count, and whether
// the kernels' opt-in attributes (more than 48 KB of dynamic shared memory) have been set on that device.  Keyed by
// cudaGetDevice(); atomics make concurrent first calls from several host threads benign (both write the same value).
#define LC_MAX_DEVICES 64
static std::atomic<int> g_sm_count[LC_MAX_DEVICES];
static std::atomic<int> g_attrs_set[LC_MAX_DEVICES];

static int lc_cur_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return -1; }
    return dev;
}

static int lc_num_sms()
{
    const int dev = lc_cur_device();
    if (dev >= 0 && dev < LC_MAX_DEVICES) {
        const int c = g_sm_count[dev].load(std::memory_order_relaxed);
        if (c > 0) return c;
    }
    int n = 0;
    if (dev < 0 || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148; // B200; keeps the sizing helpers usable on a machine without a GPU
    }
    if (dev < LC_MAX_DEVICES) g_sm_count[dev].store(n, std::memory_order_relaxed);
    return n;
}

// launches issued by the calling host thread (diagnostic: bench.py reports it as gpu_launches)
static thread_local long long t_launches = 0;
#define LC_LAUNCHED()                                                     \
    do {                                                                  \
        cudaError_t e__ = cudaGetLastError();                             \
        if (e__ != cudaSuccess) return -1000 - (int)e__;                  \
        t_launches++;                                                     \
    } while (0)

// =================================================================================================
// K1: quantiser A / dequantiser A  (stylegan3_hvae_full.py:313-316).  HBM-bound: 128-bit loads, four
// of them in flight per thread, streaming (evict-first) accesses; the index array leaves as int32,
// uint16 or uint8 (4 / 2 / 1 bytes per symbol of write traffic).
// =================================================================================================
#define LC_EW_UNROLL 4
__device__ __forceinline__ float lc_qa_round(float w, float scale)
{
    const float a = __fadd_rn(w, 1.0f);
    const float b = __fmul_rn(a, 0.5f);
    const float c = __fmul_rn(b, scale);
    return rintf(c); // torch.round: half to even
}
__device__ __forceinline__ float lc_qa_deq(float q, float scale)
{
    const float d = __fdiv_rn(q, scale);
    const float e = __fmul_rn(d, 2.0f);
    return __fsub_rn(e, 1.0f);
}
__device__ __forceinline__ int lc_qa_int(float q)
{
    return (q == q && fabsf(q) < 2.0e9f) ? (int)q : (int)0x80000000;
}

// four indices -> one vector store.  int32 keeps the unclamped value (the reference's quantiser A does not clamp);
// the narrow types hold what the coder consumes: the index clamped to the alphabet [0, hi] (NaN -> 0).
template <typename T> struct LcIdx4;
template <> struct LcIdx4<int> {
    typedef int4 V;
    static __device__ __forceinline__ int fix(float q, int) { return lc_qa_int(q); }
    static __device__ __forceinline__ V pack(int a, int b, int c, int d) { return make_int4(a, b, c, d); }
    static __device__ __forceinline__ int4 unpack(V v) { return v; }
};
template <> struct LcIdx4<unsigned short> {
    typedef ushort4 V;
    static __device__ __forceinline__ int fix(float q, int hi) { return (int)fminf(fmaxf(q, 0.0f), (float)hi); } // NaN -> 0
    static __device__ __forceinline__ V pack(int a, int b, int c, int d)
    {
        return make_ushort4((unsigned short)a, (unsigned short)b, (unsigned short)c, (unsigned short)d);
    }
    static __device__ __forceinline__ int4 unpack(V v) { return make_int4(v.x, v.y, v.z, v.w); }
};
template <> struct LcIdx4<unsigned char> {
    typedef uchar4 V;
    static __device__ __forceinline__ int fix(float q, int hi) { return (int)fminf(fmaxf(q, 0.0f), (float)hi); } // NaN -> 0
    static __device__ __forceinline__ V pack(int a, int b, int c, int d)
    {
        return make_uchar4((unsigned char)a, (unsigned char)b, (unsigned char)c, (unsigned char)d);
    }
    static __device__ __forceinline__ int4 unpack(V v) { return make_int4(v.x, v.y, v.z, v.w); }
};

template <typename T>
__global__ void __launch_bounds__(256) lc_quant_affine_kernel(const float *__restrict__ w, long long n_elem, float scale,
                                                              int hi, T *__restrict__ idx_out, float *__restrict__ wq_out)
{
    typedef typename LcIdx4<T>::V IV;
    const long long n4 = n_elem >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * LC_EW_UNROLL) {
        float4 v[LC_EW_UNROLL];
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < n4) v[u] = __ldcs(reinterpret_cast<const float4 *>(w) + i);
        }
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i >= n4) break;
            float4 q;
            q.x = lc_qa_round(v[u].x, scale); q.y = lc_qa_round(v[u].y, scale);
            q.z = lc_qa_round(v[u].z, scale); q.w = lc_qa_round(v[u].w, scale);
            if (idx_out)
                __stcs(reinterpret_cast<IV *>(idx_out) + i,
                       LcIdx4<T>::pack(LcIdx4<T>::fix(q.x, hi), LcIdx4<T>::fix(q.y, hi), LcIdx4<T>::fix(q.z, hi),
                                       LcIdx4<T>::fix(q.w, hi)));
            if (wq_out) {
                float4 o; o.x = lc_qa_deq(q.x, scale); o.y = lc_qa_deq(q.y, scale);
                o.z = lc_qa_deq(q.z, scale); o.w = lc_qa_deq(q.w, scale);
                __stcs(reinterpret_cast<float4 *>(wq_out) + i, o);
            }
        }
    }
    // tail (n_elem not a multiple of 4)
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const float q = lc_qa_round(w[i], scale);
        if (idx_out) idx_out[i] = (T)LcIdx4<T>::fix(q, hi);
        if (wq_out) wq_out[i] = lc_qa_deq(q, scale);
    }
}

static __device__ __noinline__ float lc_qa_deq_cold(float q, float scale) { return lc_qa_deq(q, scale); }

// Dequantiser A.  The value depends only on the index, and its IEEE division costs more instructions than the rest of
// the kernel together (71 % of the HBM roofline when every element divides): for alphabets of at most 4096 symbols every
// block tabulates the 2^bits values once in shared memory -- computed by the same three rounded operations -- and the
// elements become one shared-memory lookup each; indices outside the table (int32 input may hold anything) and wider
// alphabets are computed directly.
template <typename T>
__global__ void __launch_bounds__(256) lc_dequant_affine_kernel(const T *__restrict__ idx, long long n_elem, float scale,
                                                                   int tab_n, float *__restrict__ w_out)
{
    typedef typename LcIdx4<T>::V IV;
    extern __shared__ float lc_deq_tab[];
    for (int i = threadIdx.x; i < tab_n; i += blockDim.x) lc_deq_tab[i] = lc_qa_deq((float)i, scale);
    __syncthreads();
    auto deq = [&](int x) -> float { return (unsigned)x < (unsigned)tab_n ? lc_deq_tab[x] : lc_qa_deq_cold((float)x, scale); };
    const unsigned tn = (unsigned)tab_n;
    const long long n4 = n_elem >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * LC_EW_UNROLL) {
        IV v[LC_EW_UNROLL];
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < n4) v[u] = __ldcs(reinterpret_cast<const IV *>(idx) + i);
        }
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i >= n4) break;
            const int4 x = LcIdx4<T>::unpack(v[u]);
            float4 o;
            if (((unsigned)x.x < tn) & ((unsigned)x.y < tn) & ((unsigned)x.z < tn) & ((unsigned)x.w < tn)) {
                o.x = lc_deq_tab[x.x]; o.y = lc_deq_tab[x.y]; o.z = lc_deq_tab[x.z]; o.w = lc_deq_tab[x.w];
            } else {
                o.x = deq(x.x); o.y = deq(x.y); o.z = deq(x.z); o.w = deq(x.w);
            }
            __stcs(reinterpret_cast<float4 *>(w_out) + i, o);
        }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride)
        w_out[i] = deq((int)idx[i]);
}

// =================================================================================================
// K2: quantiser B (gumbel_softmax_compression.py:97,118) without the [N,n] distance matrix: the
// codebook sits in shared memory; for an ascending codebook the rounded fp32 distances are
// monotone either side of the nearest entry, so a lower-bound search plus a walk over equal
// distances gives torch.argmin's first minimum exactly.
// =================================================================================================
__device__ __forceinline__ float lc_dist(float z, float c) { return fabsf(__fsub_rn(z, c)); }

// cb: the codebook replicated `rep` times in shared memory, entry k of copy c at cb[k * rep + c] -- with rep = 32 every
// lane reads its own bank whatever k is (the data-dependent lookups of a single copy measured 3-4-way bank conflicts
// and made the kernel shared-memory bound at 20 % of the HBM roofline).
#define LC_CB(k) cb[(k) * rep + copy]

// the exact search: lower bound, nearer neighbour, then walks over equal rounded distances (first minimum)
__device__ __noinline__ int lc_argmin_sorted_slow(const float *cb, int rep, int copy, int n, float z, float guess_scale)
{
    int k;
    {
        float g = (z - LC_CB(0)) * guess_scale;
        g = g < 0.0f ? 0.0f : (g > (float)(n - 1) ? (float)(n - 1) : g);
        k = (int)g;
        while (k < n && LC_CB(k) < z) k++;
        while (k > 0 && LC_CB(k - 1) >= z) k--;
    }
    int best;
    if (k == 0) best = 0;
    else if (k == n) best = n - 1;
    else best = (lc_dist(z, LC_CB(k)) < lc_dist(z, LC_CB(k - 1))) ? k : k - 1;
    float d = lc_dist(z, LC_CB(best));
    while (best > 0 && lc_dist(z, LC_CB(best - 1)) <= d) { best--; d = lc_dist(z, LC_CB(best)); }
    while (best < n - 1 && lc_dist(z, LC_CB(best + 1)) < d) { best++; d = lc_dist(z, LC_CB(best)); }
    return best;
}

// Fast path: the rounded distances fl(|z - cb[k]|) of an ascending table are unimodal in k (rounding is monotone), so
// if the guessed entry k satisfies d(k-1) > d(k) < d(k+1) it is torch.argmin's first minimum -- three independent
// lookups, no loop.  A guess that is off (non-uniform table, |z| so large that distances tie over several entries,
// +-inf) takes the exact search.
__device__ __forceinline__ int lc_argmin_sorted(const float *cb, int rep, int copy, int n, float z, float guess_scale, float cb0)
{
    if (!(z == z)) return 0; // NaN: every distance is NaN, argmin returns 0
    float g = rintf((z - cb0) * guess_scale);
    g = g < 0.0f ? 0.0f : (g > (float)(n - 1) ? (float)(n - 1) : g);
    const int k = (int)g;
    const int kl = k > 0 ? k - 1 : 0, kr = k < n - 1 ? k + 1 : n - 1;
    const float dl = lc_dist(z, LC_CB(kl)), dm = lc_dist(z, LC_CB(k)), dr = lc_dist(z, LC_CB(kr));
    // STRICT on both sides: the rounded distances are non-increasing up to the nearest entry and non-decreasing after it,
    // so a strict local minimum is the global one -- but with duplicate entries, or entries closer together than an ulp
    // of their distance to z, the sequence has plateaus anywhere, and "d(k) <= d(k+1)" held on a plateau far from the
    // minimum (found by tests/test_gpu_quantizers.py::test_quantiser_b_sorted_non_uniform_tables).  An exact tie at the
    // minimum (z midway between two entries) now takes the exact search, which returns the first of the two.
    const bool left_ok = k == 0 || dl > dm, right_ok = k == n - 1 || dr > dm;
    if (left_ok && right_ok) return k;
    return lc_argmin_sorted_slow(cb, rep, copy, n, z, guess_scale);
}

// Two lookups for WELL-SEPARATED ascending tables (every |entry| < 8, every gap > 1e-5 -- checked per block; true of any
// linspace(-1, 1, n <= 4096) codebook): the rounded distances to neighbouring entries then differ by far more than an
// ulp, so ties only happen at exact midpoints, and once cb[k] <= z <= cb[k+1] is established the first minimum is k or
// k+1 (k on a tie; fl(|z - hi|) = fl(hi - z), rounding is symmetric).  Everything else -- below the first or above the
// last entry, NaN, a guess that is off by one -- leaves the hot loop.
__device__ __noinline__ int lc_argmin_separated_edge(const float *cb, int rep, int copy, int n, float z, float guess_scale,
                                                     float cb0, float cbl)
{
    if (fabsf(z) < 8.0f) {
        if (z < cb0) return 0;
        if (z > cbl) return n - 1;
    }
    return lc_argmin_sorted(cb, rep, copy, n, z, guess_scale, cb0);
}
__device__ __forceinline__ int lc_argmin_separated(const float *cb, int rep, int copy, int n, float z, float guess_scale,
                                                   float cb0, float cbl)
{
    const int k = __float2int_rd((z - cb0) * guess_scale); // floor; NaN gives 0 and fails the bracket test below
    if ((unsigned)k < (unsigned)(n - 1)) {
        const float lo = LC_CB(k), hi = LC_CB(k + 1);
        if (lo <= z && z <= hi) return __fsub_rn(hi, z) < __fsub_rn(z, lo) ? k + 1 : k;
    }
    return lc_argmin_separated_edge(cb, rep, copy, n, z, guess_scale, cb0, cbl);
}

__device__ __forceinline__ int lc_argmin_scan(const float *cb, int rep, int copy, int n, float z)
{
    float best = lc_dist(z, LC_CB(0));
    int bi = 0;
    for (int k = 1; k < n; k++) {
        const float d = lc_dist(z, LC_CB(k));
        if (d < best) { best = d; bi = k; }
    }
    return bi;
}

// No lookup at all for NEAR-UNIFORM well-separated tables (true of any torch.linspace table; every block measures the
// largest deviation `dev` of an entry from cb0 + k*step, in steps): t = (z - cb0) * (n-1)/span places z on the index
// axis with an error below 3e-7 * n (three fp32 roundings at |t| <= n), the decision boundary between entries k and k+1
// sits within dev of k + 0.5, and the rounded distances only tie within 1e-7 of it.  So whenever t is further than
// band = 2 * (dev + 3e-7 * n) from every half-integer, rint(t) IS torch.argmin's answer (band = 1.7e-4 for
// linspace(-1, 1, 256)); the values inside the band (0.03 %; a warp of 128 values calls the exact path every 25th
// vector -- at 0.4 % it was every other vector and the "rare" path cost as much as the loop), everything outside the
// table, NaN and +-inf take the two-lookup path, which decides exactly.  The rounding is done with the 1.5 * 2^23
// constant: integer and fraction from two FADDs, nothing on the conversion pipe: ~11 instructions per value instead
// of ~28 (profiles/r02_hbm_kernels.md).
__device__ __forceinline__ bool lc_argmin_uniform(float z, float cb0, float guess_scale, int n, float frac_max, int &k)
{
    const float t = __fmul_rn(__fsub_rn(z, cb0), guess_scale);
    const float tm = __fadd_rn(t, 12582912.0f);           // 1.5 * 2^23: the sum's low mantissa bits are rint(t)
    k = __float_as_int(tm) - 0x4B400000;
    const float frac = __fsub_rn(t, __fsub_rn(tm, 12582912.0f)); // t - rint(t), exact when k is inside the table
    // inside the table (k in [0, n) <=> tm in [1.5*2^23, 1.5*2^23 + n) <=> t in [-0.5, n - 0.5]: NaN, +-inf and huge values
    // fail here) and clear of the decision boundaries
    return (fabsf(frac) <= frac_max) & ((unsigned)k < (unsigned)n);
}
// the four values of a vector whose no-lookup test failed for at least one of them (one vector in ~60)
static __device__ __noinline__ int4 lc_k2_exact4(const float *cb, int rep, int copy, int n, float4 v, float guess_scale,
                                                 float cb0, float cbl)
{
    int4 o;
    o.x = lc_argmin_separated(cb, rep, copy, n, v.x, guess_scale, cb0, cbl);
    o.y = lc_argmin_separated(cb, rep, copy, n, v.y, guess_scale, cb0, cbl);
    o.z = lc_argmin_separated(cb, rep, copy, n, v.z, guess_scale, cb0, cbl);
    o.w = lc_argmin_separated(cb, rep, copy, n, v.w, guess_scale, cb0, cbl);
    return o;
}

#define LC_K2_THREADS 512 // one 32 KB table copy set per 512 threads
struct LcK2Args {
    const float *z; long long n_elem; const float *cb; int n, rep, copy;
    float cb0, cbl, guess_scale, frac_max;
    void *idx_out; float *deq_out;
};
// The vector loop of one search mode (0 full scan, 1 three-lookup search, 2 two-lookup search, 3 no lookup), one loop
// per mode: with the mode tested inside one unrolled loop the four unrolled copies of the mode in use lay 15 KB apart
// (the other modes' code between them) and the streaming loop did not stay in the instruction cache.  Inlined: as
// real functions the kernel's register count became the largest mode's (95) and halved the occupancy.
template <typename T, int MODE>
static __device__ __forceinline__ void lc_k2_loop(const LcK2Args &a)
{
    typedef typename LcIdx4<T>::V IV;
    const float *cb = a.cb;
    const int rep = a.rep, copy = a.copy, n = a.n;
    const float cb0 = a.cb0, cbl = a.cbl, guess_scale = a.guess_scale, frac_max = a.frac_max;
    const float *__restrict__ z = a.z;
    T *__restrict__ idx_out = (T *)a.idx_out;
    float *__restrict__ deq_out = a.deq_out;
    const long long n4 = a.n_elem >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * LC_EW_UNROLL) {
        float4 v[LC_EW_UNROLL];
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < n4) v[u] = __ldcs(reinterpret_cast<const float4 *>(z) + i);
        }
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i >= n4) break;
            int4 o;
            if (MODE == 3) {
                bool ok = lc_argmin_uniform(v[u].x, cb0, guess_scale, n, frac_max, o.x);
                ok &= lc_argmin_uniform(v[u].y, cb0, guess_scale, n, frac_max, o.y);
                ok &= lc_argmin_uniform(v[u].z, cb0, guess_scale, n, frac_max, o.z);
                ok &= lc_argmin_uniform(v[u].w, cb0, guess_scale, n, frac_max, o.w);
                if (!ok) o = lc_k2_exact4(cb, rep, copy, n, v[u], guess_scale, cb0, cbl);
            } else if (MODE == 2) {
                o.x = lc_argmin_separated(cb, rep, copy, n, v[u].x, guess_scale, cb0, cbl);
                o.y = lc_argmin_separated(cb, rep, copy, n, v[u].y, guess_scale, cb0, cbl);
                o.z = lc_argmin_separated(cb, rep, copy, n, v[u].z, guess_scale, cb0, cbl);
                o.w = lc_argmin_separated(cb, rep, copy, n, v[u].w, guess_scale, cb0, cbl);
            } else if (MODE == 1) {
                o.x = lc_argmin_sorted(cb, rep, copy, n, v[u].x, guess_scale, cb0);
                o.y = lc_argmin_sorted(cb, rep, copy, n, v[u].y, guess_scale, cb0);
                o.z = lc_argmin_sorted(cb, rep, copy, n, v[u].z, guess_scale, cb0);
                o.w = lc_argmin_sorted(cb, rep, copy, n, v[u].w, guess_scale, cb0);
            } else {
                o.x = lc_argmin_scan(cb, rep, copy, n, v[u].x); o.y = lc_argmin_scan(cb, rep, copy, n, v[u].y);
                o.z = lc_argmin_scan(cb, rep, copy, n, v[u].z); o.w = lc_argmin_scan(cb, rep, copy, n, v[u].w);
            }
            __stcs(reinterpret_cast<IV *>(idx_out) + i, LcIdx4<T>::pack(o.x, o.y, o.z, o.w));
            if (deq_out) {
                float4 d; d.x = LC_CB(o.x); d.y = LC_CB(o.y); d.z = LC_CB(o.z); d.w = LC_CB(o.w);
                __stcs(reinterpret_cast<float4 *>(deq_out) + i, d);
            }
        }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_elem; i += stride) {
        const int o = MODE ? lc_argmin_sorted(cb, rep, copy, n, z[i], guess_scale, cb0) : lc_argmin_scan(cb, rep, copy, n, z[i]);
        idx_out[i] = (T)o;
        if (deq_out) deq_out[i] = LC_CB(o);
    }
}

// Search mode for a table: 0 full scan (unsorted) / 1 three-lookup search / 2 two-lookup search / 3 no lookup; every
// block classifies the table itself (n <= 4096 values, L2-resident) while it fills its shared-memory copies.
// `only_if`: return at once with the mode when it differs (no table copies filled): -1 = fill for any mode but 3.
__device__ __forceinline__ int lc_k2_setup(const float *__restrict__ codebook, int n, int rep, int sorted, int only_if, float *cb,
                                           LcK2Args &a)
{
    __shared__ int s_separated;
    __shared__ unsigned s_dev; // largest deviation of an entry from cb0 + k*step, in steps (bit pattern of a float >= 0)
    if (threadIdx.x == 0) { s_separated = 1; s_dev = 0u; }
    __syncthreads();
    const float c0 = codebook[0], cl = codebook[n - 1], span = cl - c0;
    {
        const float step = span / (float)(n > 1 ? n - 1 : 1);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float c = codebook[i];
            if (!(fabsf(c) < 8.0f) || (i > 0 && !(c - codebook[i - 1] > 1e-5f))) s_separated = 0;
            float dev = fabsf(c - (c0 + (float)i * step)) / step;
            if (!(dev >= 0.0f && dev < 1.0f)) dev = 1.0f; // (NaN, step <= 0: not uniform)
            atomicMax(&s_dev, __float_as_uint(dev));
        }
    }
    __syncthreads();
    // half-width of the band around every half-integer that goes to the exact path: twice (the table's deviation + the
    // three fp32 roundings of t, 3e-7 * n), see lc_argmin_uniform
    const float band = 2.0f * (__uint_as_float(s_dev) + 3e-7f * (float)n) + 1e-6f;
    const int mode = !sorted ? 0 : (s_separated ? ((band < 0.05f && n >= 2 && n <= 1024) ? 3 : 2) : 1);
    if (only_if >= 0 ? mode != only_if : mode == 3) return mode;
    for (int i = threadIdx.x; i < n * rep; i += blockDim.x) cb[i] = codebook[i / rep];
    __syncthreads();
    a.cb = cb; a.n = n; a.rep = rep; a.copy = (int)(threadIdx.x & (unsigned)(rep - 1));
    a.cb0 = c0; a.cbl = cl;
    a.guess_scale = (sorted && span > 0.0f) ? (float)(n - 1) / span : 0.0f;
    a.frac_max = 0.5f - band;
    return mode;
}

// Two kernels, launched one after the other for a sorted table; each classifies the table and returns at once when the
// table is the other one's.  The no-lookup loop on its own needs 32 registers and no table per lane, so 2048 threads
// per SM keep 128 KB of loads in flight; inside the general kernel (64 registers for the search modes, 32 KB of table
// copies per block) it ran at half that occupancy and 65-76 % of the HBM roofline.
template <typename T>
__global__ void __launch_bounds__(256, 4) lc_quant_codebook_uniform_kernel(const float *__restrict__ z, long long n_elem,
                                                                           const float *__restrict__ codebook, int n, int rep,
                                                                           T *__restrict__ idx_out, float *__restrict__ deq_out)
{
    extern __shared__ float cb[];
    LcK2Args a;
    a.z = z; a.n_elem = n_elem; a.idx_out = idx_out; a.deq_out = deq_out;
    if (lc_k2_setup(codebook, n, rep, 1, 3, cb, a) != 3) return;
    lc_k2_loop<T, 3>(a);
}

template <typename T>
__global__ void __launch_bounds__(LC_K2_THREADS) lc_quant_codebook_kernel(const float *__restrict__ z, long long n_elem,
                                                                const float *__restrict__ codebook, int n, int rep,
                                                                int sorted, T *__restrict__ idx_out,
                                                                float *__restrict__ deq_out)
{
    extern __shared__ float cb[];
    LcK2Args a;
    a.z = z; a.n_elem = n_elem; a.idx_out = idx_out; a.deq_out = deq_out;
    const int mode = lc_k2_setup(codebook, n, rep, sorted, -1, cb, a);
    if (mode == 3) return; // (lc_quant_codebook_uniform_kernel, launched just before, did it)
    if (mode == 2) lc_k2_loop<T, 2>(a);
    else if (mode == 1) lc_k2_loop<T, 1>(a);
    else lc_k2_loop<T, 0>(a);
}
#undef LC_CB

template <typename T>
__global__ void __launch_bounds__(256) lc_dequant_codebook_kernel(const T *__restrict__ idx, long long n_elem,
                                                                  const float *__restrict__ codebook, int n,
                                                                  float *__restrict__ w_out)
{
    typedef typename LcIdx4<T>::V IV;
    extern __shared__ float cb[];
    for (int i = threadIdx.x; i < n; i += blockDim.x) cb[i] = codebook[i];
    __syncthreads();
    const float nanv = __int_as_float(0x7fc00000);
    const long long n4 = n_elem >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += stride * LC_EW_UNROLL) {
        IV v[LC_EW_UNROLL];
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i < n4) v[u] = __ldcs(reinterpret_cast<const IV *>(idx) + i);
        }
#pragma unroll
        for (int u = 0; u < LC_EW_UNROLL; u++) {
            const long long i = i0 + u * stride;
            if (i >= n4) break;
            const int4 x = LcIdx4<T>::unpack(v[u]);
            float4 o;
            o.x = (unsigned)x.x < (unsigned)n ? cb[x.x] : nanv; o.y = (unsigned)x.y < (unsigned)n ? cb[x.y] : nanv;
            o.z = (unsigned)x.z < (unsigned)n ? cb[x.z] : nanv; o.w = (unsigned)x.w < (unsigned)n ? cb[x.w] : nanv;
            __stcs(reinterpret_cast<float4 *>(w_out) + i, o);
        }
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
        const int x = (int)idx[i];
        w_out[i] = (unsigned)x < (unsigned)n ? cb[x] : nanv;
    }
}

// =================================================================================================
// K3 / K5: the coder kernels (one warp per block, persistent over streams) -- lc_coder.cuh
// =================================================================================================
__global__ void __launch_bounds__(32) lc_encode_kernel(LcCoderCfg cfg, LcCodes codes, int B,
                                                       unsigned char *slots, uint32_t slot_bytes, int *nbits, int *status,
                                                       int *fault, char *scratch)
{
    extern __shared__ __align__(16) char lc_smem[];
    lc_encode_block(cfg, codes, B, slots, slot_bytes, nbits, status, fault, scratch, lc_smem);
}

__global__ void __launch_bounds__(32) lc_decode_kernel(LcCoderCfg cfg, const unsigned char *__restrict__ bytes,
                                                       const long long *__restrict__ offsets, const int *__restrict__ nbits,
                                                       int B, LcIdxOut out, const float *__restrict__ deq_table, float *deq_out,
                                                       int *status, int *fault, char *scratch, int only_flagged)
{
    extern __shared__ __align__(16) char lc_smem[];
    lc_decode_block(cfg, bytes, offsets, nbits, B, out, deq_table, deq_out, status, fault, scratch, lc_smem, only_flagged);
}

__global__ void __launch_bounds__(32) lc_fast_decode_kernel(LcCoderCfg cfg, const unsigned char *__restrict__ bytes,
                                                            const long long *__restrict__ offsets,
                                                            const int *__restrict__ nbits, int B, LcIdxOut out,
                                                            const float *__restrict__ deq_table, float *deq_out,
                                                            int *status, int *fault, char *scratch)
{
    extern __shared__ __align__(16) char lc_smem[];
    lc_fast_decode_block(cfg, bytes, offsets, nbits, B, out, deq_table, deq_out, status, fault, scratch, lc_smem);
}

// Decoder v2 (lc_decoder_v2.cuh): one block per stream = one decoder warp + LCV_NU updater warps; a
// one-block kernel first fills the per-launch tables (u after the first update, exact cumsum rows).
__global__ void __launch_bounds__(32) lc_v2_tables_kernel(LcCoderCfg cfg, double *tables)
{
    extern __shared__ __align__(16) char lc_smem[];
    lcv_tables_block(cfg, tables, lc_smem);
}

// The decoder kernels exist in two register budgets.  Registers are allocated per warp in units that make 97..128
// registers per thread cost the same, so 8 blocks of 64 threads fill an SM's register file at <= 128 registers and 10
// blocks need <= 96.  LATENCY build (8 resident streams per SM, 106 registers): the dependent chain per symbol is
// shortest; used while every stream gets its own block in one wave (<= 8 x SMs streams: 6.1 ms for the benchmark).
// THROUGHPUT build (10 per SM, 90 registers, no spills): the chain is ~7 % slower but 25 % more streams are resident;
// used for larger batches (+3 % symbols/s at 8192 streams).  FN/FC/FR: generic shape, or the W+ latent shape of the
// reference (8-bit codes of a [16,512] latent, one image per stream) fixed at compile time.
#define LC_V2_LAT_PER_SM 8
#define LC_V2_THR_PER_SM 10
#define LC_V2_KERNEL(NAME, FN, FC, FR, PER_SM, OUTLINE)                                                                         \
    __global__ void __launch_bounds__(32 * LCV_WARPS, PER_SM)                                                          \
        NAME(LcCoderCfg cfg, LcV2Cfg vc, const unsigned char *__restrict__ bytes, const long long *__restrict__ offsets, \
             const int *__restrict__ nbits, int B, LcIdxOut out, const float *__restrict__ deq_table, float *deq_out, \
             int *status, int *fault, char *scratch, const double *tables, const char *t2)                             \
    {                                                                                                                  \
        extern __shared__ __align__(16) char lc_smem[];                                                                \
        lcv_decode_block<FN, FC, FR, OUTLINE>(cfg, vc, bytes, offsets, nbits, B, out, deq_table, deq_out, status,     \
                                              fault, scratch, tables, t2, lc_smem);                                    \
    }
#ifndef LCV_OPT_OUTLINE_LAT
#define LCV_OPT_OUTLINE_LAT false
#endif
LC_V2_KERNEL(lc_decode_v2_kernel, 0, 0, 0, LC_V2_LAT_PER_SM, LCV_OPT_OUTLINE_LAT)
LC_V2_KERNEL(lc_decode_v2_w8_kernel, 256, 512, 16, LC_V2_LAT_PER_SM, LCV_OPT_OUTLINE_LAT)
LC_V2_KERNEL(lc_decode_v2_thr_kernel, 0, 0, 0, LC_V2_THR_PER_SM, true)
LC_V2_KERNEL(lc_decode_v2_w8_thr_kernel, 256, 512, 16, LC_V2_THR_PER_SM, true)

// Small-alphabet decoder (lc_decoder_small.cuh): one warp per stream, dense model in shared memory, n <= 16
template <int N>
__global__ void __launch_bounds__(32) lc_decode_small_kernel(LcCoderCfg cfg, const unsigned char *__restrict__ bytes,
                                                             const long long *__restrict__ offsets,
                                                             const int *__restrict__ nbits, int B, LcIdxOut out,
                                                             const float *__restrict__ deq_table, float *deq_out,
                                                             int *status, int *fault, char *scratch)
{
    extern __shared__ __align__(16) char lc_smem[];
    lcd_decode_block<N>(cfg, bytes, offsets, nbits, B, out, deq_table, deq_out, status, fault, scratch, lc_smem);
}

#ifdef LC_DEBUG_VARIANTS
// Decoder v3 (lc_decoder_v3.cuh): decoder warp + context warp + updater warp per stream -- measured slower than v2
// (the two hand-overs per symbol cost what the decoder warp saves); debug builds only
__global__ void __launch_bounds__(32 * LC3_WARPS, 7) lc_decode_v3_kernel(LcCoderCfg cfg, LcV2Cfg vc,
                                                                          const unsigned char *__restrict__ bytes,
                                                                          const long long *__restrict__ offsets,
                                                                          const int *__restrict__ nbits, int B, int *out,
                                                                          const float *__restrict__ deq_table,
                                                                          float *deq_out, int *status, int *fault,
                                                                          char *scratch, const double *tables)
{
    extern __shared__ __align__(16) char lc_smem[];
    lc3_decode_block(cfg, vc, bytes, offsets, nbits, B, out, deq_table, deq_out, status, fault, scratch, tables, lc_smem);
}
#endif

// =================================================================================================
// K3-parallel: the encoder split by context group (lc_encoder_par.cuh)
//   phase S: keys + stable sort by key (one block per stream, CUB block radix sort)
//   phase A: exact intervals per position, one warp per context group, 8 warps per stream
//   phase B: the serial range coder over those intervals, one warp per stream
// =================================================================================================
#ifndef LC_SORT_THREADS
#define LC_SORT_THREADS 1024
#endif
#ifndef LC_SORT_RADIX_BITS
#define LC_SORT_RADIX_BITS 4
#endif
typedef cub::BlockRadixSort<uint32_t, LC_SORT_THREADS, LC_PAR_MAX_SYMBOLS / LC_SORT_THREADS, unsigned short, LC_SORT_RADIX_BITS>
    LcBlockSort;
typedef cub::BlockScan<int, LC_SORT_THREADS> LcBlockScan;

// Also emits what needs no model: the closed-form interval of every first visit (uniform model: cum[i] = i/n) and,
// when `tables` is given, the interval of every second visit (table of exact cumsums of the model after one update,
// lcv_tables_block), plus glist/ngroups = the contexts visited at least three times (the work items of phase A).
__global__ void __launch_bounds__(LC_SORT_THREADS) lc_enc_sort_kernel(LcCoderCfg cfg, LcCodes codes,
                                                          uint32_t *__restrict__ skeys, unsigned short *__restrict__ spos,
                                                          int *__restrict__ first_bad, unsigned short *__restrict__ glist,
                                                          int *__restrict__ ngroups, double *__restrict__ ivs,
                                                          const double *__restrict__ tables)
{
    extern __shared__ __align__(16) char lc_smem[];
    LcBlockSort::TempStorage &temp = *reinterpret_cast<LcBlockSort::TempStorage *>(lc_smem);
    __shared__ int s_first_bad;
    constexpr int ITEMS = LC_PAR_MAX_SYMBOLS / LC_SORT_THREADS;
    const int total = cfg.total, n = cfg.n, C = cfg.C, RC = cfg.R * cfg.C;
    const LcCodes c = codes + (size_t)blockIdx.x * total;
    if (threadIdx.x == 0) s_first_bad = total;
    __syncthreads();
    for (int p = threadIdx.x; p < total; p += LC_SORT_THREADS) {
        const int s = c[p];
        if (s < 0 || s >= n) atomicMin(&s_first_bad, p);
    }
    __syncthreads();
    const int fb = s_first_bad;
    uint32_t keys[ITEMS];
    unsigned short vals[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int p = (int)threadIdx.x * ITEMS + i; // blocked arrangement = position order (stable sort keeps it)
        vals[i] = (unsigned short)p;
        uint32_t k = LC_PAR_KEY_PAD;
        if (p < fb) {
            const int q = p % RC, cc = q % C, rr = q / C;
            const int left = cc > 0 ? c[p - 1] : -1;
            const int up = rr > 0 ? c[p - C] : -1;
            k = (uint32_t)(left + 1) * (uint32_t)(n + 1) + (uint32_t)(up + 1);
        }
        keys[i] = k;
    }
    // real keys are below (n+1)^2 < 2^key_bits; the padding key (all ones in those bits) sorts last
    const int key_bits = 32 - __clz((n + 1) * (n + 1));
    LcBlockSort(temp).Sort(keys, vals, 0, key_bits);
    uint32_t *ok = skeys + (size_t)blockIdx.x * LC_PAR_MAX_SYMBOLS + (size_t)threadIdx.x * ITEMS;
    unsigned short *ov = spos + (size_t)blockIdx.x * LC_PAR_MAX_SYMBOLS + (size_t)threadIdx.x * ITEMS;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) { ok[i] = keys[i]; ov[i] = vals[i]; }
    if (threadIdx.x == 0) first_bad[blockIdx.x] = fb;

    // neighbours across thread boundaries: sorted index j = tid*ITEMS + i
    __shared__ uint32_t s_last[LC_SORT_THREADS], s_last2[LC_SORT_THREADS], s_first[LC_SORT_THREADS], s_first2[LC_SORT_THREADS];
    __shared__ unsigned short s_last_val[LC_SORT_THREADS];
    __shared__ LcBlockScan::TempStorage scan_temp;
    s_last[threadIdx.x] = keys[ITEMS - 1];
    s_last2[threadIdx.x] = keys[ITEMS - 2];
    s_last_val[threadIdx.x] = vals[ITEMS - 1];
    s_first[threadIdx.x] = keys[0];
    s_first2[threadIdx.x] = keys[1];
    __syncthreads();
    const bool has_prev = threadIdx.x > 0, has_next = threadIdx.x < LC_SORT_THREADS - 1;
    const uint32_t prev1 = has_prev ? s_last[threadIdx.x - 1] : 0u, prev2 = has_prev ? s_last2[threadIdx.x - 1] : 0u;
    const unsigned short prev_val = has_prev ? s_last_val[threadIdx.x - 1] : (unsigned short)0;
    const uint32_t next1 = has_next ? s_first[threadIdx.x + 1] : LC_PAR_KEY_PAD;
    const uint32_t next2 = has_next ? s_first2[threadIdx.x + 1] : LC_PAR_KEY_PAD;
    const double u0 = LC_DDIV(1.0, (double)n);
    const double *cum1 = tables ? tables + 64 : (const double *)0;
    double *ivb = ivs + (size_t)blockIdx.x * 2 * LC_PAR_MAX_SYMBOLS;
    unsigned list_mask = 0u;
    int n_list = 0;
#pragma unroll
    for (int i = 0; i < ITEMS; i++) {
        const int j = (int)threadIdx.x * ITEMS + i;
        const uint32_t k0 = keys[i];
        const uint32_t km1 = i > 0 ? keys[i > 0 ? i - 1 : 0] : prev1;
        const uint32_t km2 = i > 1 ? keys[i > 1 ? i - 2 : 0] : (i == 1 ? prev1 : prev2);
        const uint32_t kp1 = i < ITEMS - 1 ? keys[i < ITEMS - 1 ? i + 1 : 0] : next1;
        const uint32_t kp2 = i < ITEMS - 2 ? keys[i < ITEMS - 2 ? i + 2 : 0] : (i == ITEMS - 2 ? next1 : next2);
        const bool valid = j < fb; // positions before the first bad symbol; padding keys sort behind them
        const bool head = valid && (j == 0 || km1 != k0);
        if (head) {
            const int p = vals[i];
            const int sy = c[p];
            ivb[2 * p] = LC_DMUL((double)sy, u0);
            ivb[2 * p + 1] = LC_DMUL((double)(sy + 1), u0);
        }
        if (cum1) {
            const bool second = valid && !head && (j == 1 || km2 != k0);
            if (second) { // model after one update with the first visit's symbol
                const int p = vals[i];
                const int sy = c[p], s1 = c[i > 0 ? vals[i > 0 ? i - 1 : 0] : prev_val];
                const double *row = cum1 + (size_t)s1 * (n + 1);
                ivb[2 * p] = row[sy];
                ivb[2 * p + 1] = row[sy + 1];
            }
            if (head && (j + 2 < fb) && kp1 == k0 && kp2 == k0) { list_mask |= 1u << i; n_list++; }
        } else if (head && (j + 1 < fb) && kp1 == k0) { list_mask |= 1u << i; n_list++; }
    }
    int g_off = 0, g_total = 0;
    LcBlockScan(scan_temp).ExclusiveSum(n_list, g_off, g_total);
    unsigned short *gl = glist + (size_t)blockIdx.x * LC_PAR_MAX_GROUPS;
    while (list_mask) {
        const int i = __ffs((int)list_mask) - 1;
        list_mask &= list_mask - 1;
        gl[g_off++] = (unsigned short)((int)threadIdx.x * ITEMS + i);
    }
    if (threadIdx.x == 0) ngroups[blockIdx.x] = g_total;
}

// Phase S, second version (lc_encoder_sort.cuh): two-pass radix sort from warp primitives, 256 threads per stream
__global__ void __launch_bounds__(LCS2_THREADS) lc_enc_sort2_kernel(LcCoderCfg cfg, LcCodes codes, uint32_t *skeys,
                                                                    unsigned short *spos, int *__restrict__ first_bad,
                                                                    unsigned short *__restrict__ glist,
                                                                    int *__restrict__ ngroups, double *__restrict__ ivs,
                                                                    const double *__restrict__ tables)
{
    extern __shared__ __align__(16) char lc_smem[];
    lcs2_block(cfg, codes, skeys, spos, first_bad, glist, ngroups, ivs, tables, lc_smem);
}

#ifdef LC_DEBUG_VARIANTS // the dense warp-per-group phase A (4.1 ms on the benchmark against 1.3 ms): debug builds only
__global__ void __launch_bounds__(256) lc_enc_phase_a_kernel(LcCoderCfg cfg, LcCodes codes, int B,
                                                             const uint32_t *__restrict__ skeys,
                                                             const unsigned short *__restrict__ spos,
                                                             const int *__restrict__ first_bad, double *ivs)
{
    extern __shared__ __align__(16) char lc_smem[];
    lc_enc_phase_a_block(cfg, codes, B, skeys, spos, first_bad, ivs, lc_smem);
}
#endif

#define LCS_BLOCK_WARPS 4
#ifndef LCS_BLOCKS_PER_SM
#define LCS_BLOCKS_PER_SM 12
#endif
__global__ void __launch_bounds__(32 * LCS_BLOCK_WARPS, LCS_BLOCKS_PER_SM) lc_enc_phase_a_sparse_kernel(
    LcCoderCfg cfg, LcCodes codes, int B, const uint32_t *__restrict__ skeys,
    const unsigned short *__restrict__ spos, const int *__restrict__ first_bad, const unsigned short *__restrict__ glist,
    const int *__restrict__ ngroups, double *ivs, unsigned int *task_counter, const double *__restrict__ tables,
    const char *__restrict__ t2)
{
    extern __shared__ __align__(16) char lc_smem[];
    lc_enc_phase_a_sparse_block(cfg, codes, B, skeys, spos, first_bad, glist, ngroups, ivs, task_counter, tables, t2,
                                lc_smem);
}

// per-launch records of the models after two visits (lcv_t2_block), used by the encoder's phase A and the
// decoder's updater warps
__global__ void __launch_bounds__(32 * LCS_BLOCK_WARPS) lc_t2_kernel(LcCoderCfg cfg, const double *__restrict__ tables,
                                                                     char *t2)
{
    extern __shared__ __align__(16) char lc_smem[];
    lcv_t2_block(cfg, tables, t2, lc_smem);
}

__global__ void __launch_bounds__(32) lc_enc_phase_b_kernel(LcCoderCfg cfg, int B, const int *__restrict__ first_bad,
                                                            const double *__restrict__ ivs, unsigned char *slots,
                                                            uint32_t slot_bytes, int *nbits, int *status, int *fault)
{
    lc_enc_phase_b_block(cfg, B, first_bad, ivs, slots, slot_bytes, nbits, status, fault);
}

// phase B split (lc_encoder_pack.cuh): B1 = the serial (low, high) recurrence, B2 = parallel bit placement
__global__ void __launch_bounds__(32) lc_enc_phase_b1_kernel(LcCoderCfg cfg, int B, const int *__restrict__ first_bad,
                                                             double *ivs, int *first_out)
{
    __shared__ __align__(16) char b1_ring[LC_B1_SMEM];
    lc_enc_phase_b1_block(cfg, B, first_bad, ivs, first_out, b1_ring);
}

__global__ void __launch_bounds__(LC_B2_THREADS) lc_enc_phase_b2_kernel(LcCoderCfg cfg, int B,
                                                                        const int *__restrict__ first_bad,
                                                                        const double *__restrict__ ivs, unsigned char *slots,
                                                                        uint32_t slot_bytes, int *nbits, int *status, int *fault)
{
    extern __shared__ __align__(16) char lc_smem[];
    lc_enc_phase_b2_block(cfg, B, first_bad, ivs, slots, slot_bytes, nbits, status, fault, lc_smem);
}

// stateful coder (lc_stateful.cuh): one warp, one stream, dense models in a direct-mapped table
__global__ void __launch_bounds__(32) lc_stateful_encode_kernel(LcCoderCfg cfg, const int *__restrict__ codes,
                                                                LcStatefulTable T, unsigned char *slot, uint32_t slot_bytes,
                                                                int *nbits, int *status, int *fault)
{
    extern __shared__ __align__(16) char lc_smem[];
    LcWarp W;
    lc_warp_init(W, cfg, lc_smem, (char *)0);
    int fi = 0;
    const long long nb = lcs_encode_stream(W, T, codes, (uint32_t *)slot, slot_bytes / 4, &fi);
    if (W.lane == 0) { *nbits = (int)nb; *status = W.status; *fault = fi; }
}

__global__ void __launch_bounds__(32) lc_stateful_decode_kernel(LcCoderCfg cfg, const unsigned char *__restrict__ bytes,
                                                                long long nbytes, LcStatefulTable T, int *out, int *status,
                                                                int *fault)
{
    extern __shared__ __align__(16) char lc_smem[];
    LcWarp W;
    lc_warp_init(W, cfg, lc_smem, (char *)0);
    int fi = 0;
    lcs_decode_stream(W, T, bytes, nbytes, out, &fi);
    if (W.lane == 0) { *status = W.status; *fault = fi; }
}

// ContextModel.update_model on one vector (the host-side ContextModel.update_model method)
__global__ void __launch_bounds__(32) lc_model_update_kernel(LcCoderCfg cfg, double *vec, int symbol)
{
    extern __shared__ __align__(16) char lc_smem[];
    LcWarp W;
    lc_warp_init(W, cfg, lc_smem, (char *)0);
    for (int i = W.lane; i < W.n; i += 32) W.dense[i] = vec[i];
    __syncwarp();
    lc_dense_update(W, symbol);
    for (int i = W.lane; i < W.n; i += 32) vec[i] = W.dense[i];
}

// =================================================================================================
// K4: stream compaction -- exclusive scan of the 16-byte-aligned stream sizes, then a word copy
// =================================================================================================
__global__ void __launch_bounds__(1024) lc_scan_sizes_kernel(const int *__restrict__ nbits, int *status, int B,
                                                             long long capacity, long long *offsets)
{
    __shared__ long long warp_sums[32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < B; base += 1024) {
        const int b = base + threadIdx.x;
        long long sz = 0;
        if (b < B && status[b] == LC_OK) sz = ((((long long)nbits[b] + 7) >> 3) + 15) & ~15ll;
        long long incl = sz;
        for (int off = 1; off < 32; off <<= 1) {
            const long long t = __shfl_up_sync(LC_FULL_MASK, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            long long ws = warp_sums[lane];
            for (int off = 1; off < 32; off <<= 1) {
                const long long t = __shfl_up_sync(LC_FULL_MASK, ws, off);
                if (lane >= off) ws += t;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const long long carry = carry_s;
        const long long excl = carry + (wid > 0 ? warp_sums[wid - 1] : 0) + incl - sz;
        if (b < B) {
            offsets[b] = excl;
            if (sz > 0 && excl + sz > capacity) status[b] = LC_OUT_OVERFLOW;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sums[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[B] = carry_s;
}

__global__ void __launch_bounds__(128) lc_compact_kernel(const unsigned char *__restrict__ slots, long long slot_bytes,
                                                         const int *__restrict__ nbits, const int *__restrict__ status,
                                                         const long long *__restrict__ offsets, unsigned char *out)
{
    const int b = blockIdx.x;
    if (status[b] != LC_OK) return;
    const long long nbytes = ((long long)nbits[b] + 7) >> 3;
    const int nwords = (int)((nbytes + 3) >> 2);       // the encoder zero-pads its last word
    const int npad = (int)(((nbytes + 15) & ~15ll) >> 2); // words of the aligned segment
    const uint32_t *src = reinterpret_cast<const uint32_t *>(slots + (long long)b * slot_bytes);
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + offsets[b]);
    for (int i = threadIdx.x; i < npad; i += blockDim.x) dst[i] = i < nwords ? src[i] : 0u;
}

// =================================================================================================
// C ABI
// =================================================================================================
static int lc_make_cfg(LcCoderCfg &cfg, int imgs, int R, int C, int n, double rate, int mode, int has_ctx)
{
    cfg = LcCoderCfg();
    cfg.n = n; cfg.R = R; cfg.C = C; cfg.imgs = imgs; cfg.has_ctx = has_ctx ? 1 : 0; cfg.mode = mode; cfg.rate = rate;
    if (mode != LC_MODE_VERBATIM && mode != LC_MODE_REPAIRED) return -22;
    return lc_cfg_finalize(&cfg);
}

static bool lc_idx_bytes_ok(int idx_bytes, int n) // can an element of idx_bytes bytes hold every symbol below n?
{
    return idx_bytes == 4 || (idx_bytes == 2 && n <= 65536) || (idx_bytes == 1 && n <= 256);
}

// Opt-in attributes (more than 48 KB of dynamic shared memory) are per device: set once per device, not per process.
static void lc_prepare_device()
{
    const int dev = lc_cur_device();
    if (dev >= 0 && dev < LC_MAX_DEVICES && g_attrs_set[dev].load(std::memory_order_acquire)) return;
    cudaFuncSetAttribute(lc_enc_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LcBlockSort::TempStorage));
    cudaFuncSetAttribute(lc_enc_sort2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LCS2_SMEM);
    cudaFuncSetAttribute(lc_enc_phase_b2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(lc_enc_phase_a_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LCS_BLOCK_WARPS * 1024 * 8);
    cudaFuncSetAttribute(lc_decode_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(lc_decode_v2_w8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(lc_decode_v2_thr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(lc_decode_v2_w8_thr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
#ifdef LC_DEBUG_VARIANTS
    cudaFuncSetAttribute(lc_enc_phase_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 8);
    cudaFuncSetAttribute(lc_decode_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
#endif
    cudaGetLastError();
    if (dev >= 0 && dev < LC_MAX_DEVICES) g_attrs_set[dev].store(1, std::memory_order_release);
}

static int lc_grid_for(const LcCoderCfg &cfg, int B)
{
    int per_sm = (int)((227u * 1024u) / (cfg.sm_bytes + 1024u));
    if (per_sm > 32) per_sm = 32;
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)lc_num_sms() * per_sm;
    if (g > B) g = B;
    return g < 1 ? 1 : (int)g;
}

// parallel-encoder workspace per stream: sorted keys (4 B) + positions (2 B) + two float64 bounds
#define LC_PAR_STREAM_BYTES ((int64_t)LC_PAR_MAX_SYMBOLS * (4 + 2 + 8 + 8) + LC_PAR_MAX_GROUPS * 2 + 32)
#define LC_PAR_TILE 8192

static bool lc_use_parallel_encoder(const LcCoderCfg &cfg, int flags)
{
    return cfg.has_ctx && cfg.total <= LC_PAR_MAX_SYMBOLS && !(flags & LC_FLAG_ENC_SERIAL);
}

// which kernel decodes: v2 (decoder + updater warp, n <= 256), the register-model kernel (lc_decoder_fast.cuh), or the
// generic serial kernel (verbatim mode, global-context streams); the caller's `flags` can force the slower ones
static bool lc_use_decoder_v2(const LcCoderCfg &cfg, int flags)
{
    return lcv_eligible(cfg) && !(flags & (LC_FLAG_DEC_REGISTER_MODEL | LC_FLAG_DEC_SERIAL));
}
// resident streams per SM for a batch of B: the latency build while one wave of it holds the batch
static int lc_v2_per_sm(const LcV2Cfg &vc, int B, int flags)
{
    int fit = (int)((227u * 1024u) / (vc.sm_bytes + 1024u));
    if (fit < 1) fit = 1;
    const int lat = fit < LC_V2_LAT_PER_SM ? fit : LC_V2_LAT_PER_SM;
    const int thr = fit < LC_V2_THR_PER_SM ? fit : LC_V2_THR_PER_SM;
    if (flags & LC_FLAG_DEC_LATENCY_BUILD) return lat;
    if (flags & LC_FLAG_DEC_THROUGHPUT_BUILD) return thr;
    return (long long)B <= (long long)lc_num_sms() * lat ? lat : thr;
}
static int lc_v2_grid(const LcV2Cfg &vc, int B, int flags)
{
    long long g = (long long)lc_num_sms() * lc_v2_per_sm(vc, B, flags);
    if (g > B) g = B;
    return g < 1 ? 1 : (int)g;
}
static int64_t lc_v2_scratch_need(const LcCoderCfg &cfg, int B)
{
    LcV2Cfg vc;
    lcv_cfg_make(cfg, &vc);
    // sized for the larger of the two grids, so the scratch size does not depend on the flags
    return (int64_t)lc_v2_grid(vc, B, LC_FLAG_DEC_THROUGHPUT_BUILD) * (int64_t)vc.g_stride + (int64_t)lcv_tables_bytes(cfg.n) +
           512 + (int64_t)cfg.n * cfg.n * 64;
}

static int lc_small_grid(int B)
{
    long long g = (long long)lc_num_sms() * LCD_WARPS_PER_SM;
    if (g > B) g = B;
    return g < 1 ? 1 : (int)g;
}

static int64_t lc_scratch_need(const LcCoderCfg &cfg, int B)
{
    int64_t need = (int64_t)lc_grid_for(cfg, B) * (int64_t)cfg.scratch_stride;
    if (lcd_eligible(cfg)) need += 256 + (int64_t)lc_small_grid(B) * (int64_t)lcd_tab_bytes(cfg.n);
    if (lcv_eligible(cfg)) {
        const int64_t v2 = lc_v2_scratch_need(cfg, B);
        if (v2 > need) need = v2;
    }
    if (lc_use_parallel_encoder(cfg, 0)) {
        const int64_t par = (int64_t)(B < LC_PAR_TILE ? B : LC_PAR_TILE) * LC_PAR_STREAM_BYTES + 512 +
                            (int64_t)lcv_tables_bytes(cfg.n) +
                            (cfg.n <= LCS_T2_MAX_N ? (int64_t)cfg.n * cfg.n * 64 : 0);
        if (par > need) need = par;
    }
    return need;
}

// grid of the elementwise kernels: one wave of resident blocks (a grid-stride loop gives every block the same share)
static int lc_ew_grid(long long n_elem, size_t smem_per_block = 0, int threads = 256)
{
    long long blocks = (n_elem / 4 + threads * LC_EW_UNROLL - 1) / (threads * LC_EW_UNROLL);
    int per_sm = 2048 / threads; // 2048 threads per SM
    if (smem_per_block) {
        const int fit = (int)((227u * 1024u) / (smem_per_block + 1024u));
        per_sm = fit < per_sm ? (fit < 1 ? 1 : fit) : per_sm;
    }
    const long long cap = (long long)lc_num_sms() * per_sm;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

extern "C" {

int lc_version(void) { return LC_ABI_VERSION; }

int64_t lc_debug_launch_count(int reset)
{
    const long long v = t_launches;
    if (reset) t_launches = 0;
    return (int64_t)v;
}

int lc_quantize_affine_t(const float *w, int64_t n_elem, int bits, void *idx_out, int idx_bytes, float *wq_out, void *stream)
{
    if (n_elem < 0 || bits < 1 || bits > 24 || !w) return -22;
    if (idx_out && !lc_idx_bytes_ok(idx_bytes, 1 << bits)) return -22;
    if (n_elem == 0 || (!idx_out && !wq_out)) return 0;
    if ((((uintptr_t)w | (uintptr_t)idx_out | (uintptr_t)wq_out) & 15) != 0) return -22;
    const float scale = (float)((1 << bits) - 1);
    const int hi = (1 << bits) - 1, grid = lc_ew_grid(n_elem);
    cudaStream_t st = (cudaStream_t)stream;
    if (!idx_out || idx_bytes == 4)
        lc_quant_affine_kernel<int><<<grid, 256, 0, st>>>(w, n_elem, scale, hi, (int *)idx_out, wq_out);
    else if (idx_bytes == 2)
        lc_quant_affine_kernel<unsigned short><<<grid, 256, 0, st>>>(w, n_elem, scale, hi, (unsigned short *)idx_out, wq_out);
    else
        lc_quant_affine_kernel<unsigned char><<<grid, 256, 0, st>>>(w, n_elem, scale, hi, (unsigned char *)idx_out, wq_out);
    LC_LAUNCHED();
    return 0;
}

int lc_quantize_affine(const float *w, int64_t n_elem, int bits, int32_t *idx_out, float *wq_out, void *stream)
{
    return lc_quantize_affine_t(w, n_elem, bits, idx_out, 4, wq_out, stream);
}

int lc_dequantize_affine_t(const void *idx, int idx_bytes, int64_t n_elem, int bits, float *w_out, void *stream)
{
    if (n_elem < 0 || bits < 1 || bits > 24 || !idx || !w_out) return -22;
    if (idx_bytes != 1 && idx_bytes != 2 && idx_bytes != 4) return -22;
    if (n_elem == 0) return 0;
    if ((((uintptr_t)idx | (uintptr_t)w_out) & 15) != 0) return -22;
    const float scale = (float)((1 << bits) - 1);
    const int tab_n = bits <= 12 ? (1 << bits) : 0; // values tabulated per block (lc_dequant_affine_kernel)
    const size_t sm = (size_t)tab_n * 4;
    const int grid = lc_ew_grid(n_elem, sm);
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_bytes == 4) lc_dequant_affine_kernel<int><<<grid, 256, sm, st>>>((const int *)idx, n_elem, scale, tab_n, w_out);
    else if (idx_bytes == 2)
        lc_dequant_affine_kernel<unsigned short><<<grid, 256, sm, st>>>((const unsigned short *)idx, n_elem, scale, tab_n, w_out);
    else lc_dequant_affine_kernel<unsigned char><<<grid, 256, sm, st>>>((const unsigned char *)idx, n_elem, scale, tab_n, w_out);
    LC_LAUNCHED();
    return 0;
}

int lc_dequantize_affine(const int32_t *idx, int64_t n_elem, int bits, float *w_out, void *stream)
{
    return lc_dequantize_affine_t(idx, 4, n_elem, bits, w_out, stream);
}

int lc_quantize_codebook_t(const float *z, int64_t n_elem, const float *codebook, int n, int sorted_ascending,
                           void *idx_out, int idx_bytes, float *deq_out, void *stream)
{
    if (n_elem < 0 || n < 1 || n > 4096 || !z || !codebook || !idx_out || !lc_idx_bytes_ok(idx_bytes, n)) return -22;
    if (n_elem == 0) return 0;
    if ((((uintptr_t)z | (uintptr_t)idx_out | (uintptr_t)deq_out) & 15) != 0) return -22;
    const int so = sorted_ascending ? 1 : 0;
    int rep = 32; // copies of the table in shared memory (one per bank while they fit 32 KB)
    while (rep > 1 && (size_t)n * rep * 4 > 32 * 1024) rep >>= 1;
    const size_t sm = (size_t)n * rep * 4;
    const int grid = lc_ew_grid(n_elem, sm, LC_K2_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    if (so && n >= 2 && n <= 1024) { // near-uniform tables (any linspace): the no-lookup kernel; others return at once
        int rep_u = 8; // (its table copies only serve the exact path of the values near a decision boundary, and deq_out)
        while (rep_u > 1 && (size_t)n * rep_u * 4 > 8 * 1024) rep_u >>= 1;
        const size_t sm_u = (size_t)n * rep_u * 4;
        const int grid_u = lc_ew_grid(n_elem, sm_u, 256);
        if (idx_bytes == 4)
            lc_quant_codebook_uniform_kernel<int><<<grid_u, 256, sm_u, st>>>(z, n_elem, codebook, n, rep_u, (int *)idx_out, deq_out);
        else if (idx_bytes == 2)
            lc_quant_codebook_uniform_kernel<unsigned short><<<grid_u, 256, sm_u, st>>>(z, n_elem, codebook, n, rep_u, (unsigned short *)idx_out, deq_out);
        else
            lc_quant_codebook_uniform_kernel<unsigned char><<<grid_u, 256, sm_u, st>>>(z, n_elem, codebook, n, rep_u, (unsigned char *)idx_out, deq_out);
        LC_LAUNCHED();
    }
    if (idx_bytes == 4)
        lc_quant_codebook_kernel<int><<<grid, LC_K2_THREADS, sm, st>>>(z, n_elem, codebook, n, rep, so, (int *)idx_out, deq_out);
    else if (idx_bytes == 2)
        lc_quant_codebook_kernel<unsigned short><<<grid, LC_K2_THREADS, sm, st>>>(z, n_elem, codebook, n, rep, so, (unsigned short *)idx_out, deq_out);
    else
        lc_quant_codebook_kernel<unsigned char><<<grid, LC_K2_THREADS, sm, st>>>(z, n_elem, codebook, n, rep, so, (unsigned char *)idx_out, deq_out);
    LC_LAUNCHED();
    return 0;
}

int lc_quantize_codebook(const float *z, int64_t n_elem, const float *codebook, int n, int sorted_ascending,
                         int32_t *idx_out, float *deq_out, void *stream)
{
    return lc_quantize_codebook_t(z, n_elem, codebook, n, sorted_ascending, idx_out, 4, deq_out, stream);
}

int lc_dequantize_codebook_t(const void *idx, int idx_bytes, int64_t n_elem, const float *codebook, int n, float *w_out,
                             void *stream)
{
    if (n_elem < 0 || n < 1 || n > 4096 || !idx || !codebook || !w_out) return -22;
    if (idx_bytes != 1 && idx_bytes != 2 && idx_bytes != 4) return -22;
    if (n_elem == 0) return 0;
    if ((((uintptr_t)idx | (uintptr_t)w_out) & 15) != 0) return -22;
    const size_t sm = (size_t)n * 4;
    const int grid = lc_ew_grid(n_elem, sm);
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_bytes == 4) lc_dequant_codebook_kernel<int><<<grid, 256, sm, st>>>((const int *)idx, n_elem, codebook, n, w_out);
    else if (idx_bytes == 2)
        lc_dequant_codebook_kernel<unsigned short><<<grid, 256, sm, st>>>((const unsigned short *)idx, n_elem, codebook, n, w_out);
    else lc_dequant_codebook_kernel<unsigned char><<<grid, 256, sm, st>>>((const unsigned char *)idx, n_elem, codebook, n, w_out);
    LC_LAUNCHED();
    return 0;
}

int lc_dequantize_codebook(const int32_t *idx, int64_t n_elem, const float *codebook, int n, float *w_out, void *stream)
{
    return lc_dequantize_codebook_t(idx, 4, n_elem, codebook, n, w_out, stream);
}

int lc_coder_grid(int B, int imgs, int R, int C, int n_symbols, int has_ctx)
{
    LcCoderCfg cfg;
    if (B < 1 || lc_make_cfg(cfg, imgs, R, C, n_symbols, 0.05, LC_MODE_REPAIRED, has_ctx)) return -22;
    return lc_grid_for(cfg, B);
}

int64_t lc_coder_scratch_bytes(int B, int imgs, int R, int C, int n_symbols, int has_ctx)
{
    LcCoderCfg cfg;
    if (B < 1 || lc_make_cfg(cfg, imgs, R, C, n_symbols, 0.05, LC_MODE_REPAIRED, has_ctx)) return -22;
    return lc_scratch_need(cfg, B);
}

int64_t lc_encode_slot_bytes(int imgs, int R, int C, int n_symbols)
{
    if (imgs < 1 || R < 1 || C < 1 || n_symbols < 2) return -22;
    int lg = 0;
    while ((1 << lg) < n_symbols) lg++;
    // adaptive coding of i.i.d.-like latents costs about log2(n) bits/symbol; allow 1.5x + slack.
    // A stream that still does not fit reports LC_STATUS_OUT_OVERFLOW and can be retried larger.
    const int64_t total = (int64_t)imgs * R * C;
    const int64_t bytes = (total * (lg + 2) * 3 / 2) / 8 + 256;
    return (bytes + 15) & ~(int64_t)15;
}

int lc_encode_batch_t(const void *idx, int idx_bytes, int B, int imgs, int R, int C, int n_symbols, double adaptation_rate,
                      int mode, int has_ctx, void *scratch, int64_t scratch_bytes, uint8_t *slots, int64_t slot_bytes,
                      uint8_t *out_bytes, int64_t out_capacity, int64_t *out_offsets, int32_t *out_nbits, int32_t *status,
                      int32_t *fault_index, int flags, void *stream)
{
    LcCoderCfg cfg;
    if (B < 0 || !idx || !scratch || !slots || !out_nbits || !status || !fault_index) return -22;
    if (idx_bytes != 1 && idx_bytes != 2 && idx_bytes != 4) return -22;
    if (B == 0) return 0;
    int rc = lc_make_cfg(cfg, imgs, R, C, n_symbols, adaptation_rate, mode, has_ctx);
    if (rc) return rc;
    if (slot_bytes < 16 || (slot_bytes & 15) || slot_bytes > 0xfffffff0ll) return -22;
    if (out_bytes && !out_offsets) return -22;
    if ((((uintptr_t)slots | (uintptr_t)out_bytes | (uintptr_t)scratch) & 15) != 0) return -22;
    if (scratch_bytes < lc_scratch_need(cfg, B)) return -12;
    lc_prepare_device();
    cudaStream_t st = (cudaStream_t)stream;
    const LcCodes all_codes(idx, idx_bytes);
    if (lc_use_parallel_encoder(cfg, flags)) {
        const size_t sort_smem = sizeof(LcBlockSort::TempStorage);
        bool sparse_variant = true;
#ifdef LC_DEBUG_VARIANTS
        if (flags & LC_FLAG_DEBUG_ENC_DENSE_PHASE_A) sparse_variant = false;
#endif
        const int tile = B < LC_PAR_TILE ? B : LC_PAR_TILE;
        // per-launch tables (u after the first update, exact cumsum rows of the model after one update)
        double *tables = (double *)((char *)scratch + (((size_t)tile * LC_PAR_STREAM_BYTES + 255) & ~(size_t)255));
        // ... and, for alphabets up to LCS_T2_MAX_N symbols, the models after two visits.  The second table pays
        // for itself once the batch holds a few times n*n/1000 streams.
        char *t2 = (char *)0;
        if (sparse_variant) {
            lc_v2_tables_kernel<<<cfg.n, 32, (size_t)cfg.n * 8, st>>>(cfg, tables);
            LC_LAUNCHED();
            if (cfg.n > LCD_MAX_N && cfg.n <= LCS_T2_MAX_N && (long long)B * 1000 >= (long long)cfg.n * cfg.n) {
                t2 = (char *)tables + ((lcv_tables_bytes(cfg.n) + 255) & ~(uint64_t)255);
                const size_t t2_smem = (size_t)LCS_BLOCK_WARPS * cfg.n * 8;
                lc_t2_kernel<<<lc_num_sms() * 4, 32 * LCS_BLOCK_WARPS, t2_smem, st>>>(cfg, tables, t2);
                LC_LAUNCHED();
            }
        }
        for (int b0 = 0; b0 < B; b0 += LC_PAR_TILE) {
            const int nb = (B - b0) < LC_PAR_TILE ? (B - b0) : LC_PAR_TILE;
            char *ws = (char *)scratch;
            uint32_t *skeys = (uint32_t *)ws;                 ws += (size_t)nb * LC_PAR_MAX_SYMBOLS * 4;
            double *ivs = (double *)ws;                       ws += (size_t)nb * LC_PAR_MAX_SYMBOLS * 16;
            unsigned short *spos = (unsigned short *)ws;      ws += (size_t)nb * LC_PAR_MAX_SYMBOLS * 2;
            unsigned short *glist = (unsigned short *)ws;     ws += (size_t)nb * LC_PAR_MAX_GROUPS * 2;
            int *first_bad = (int *)ws;                       ws += (size_t)nb * 4;
            int *ngroups = (int *)ws;                         ws += (size_t)nb * 4;
            unsigned int *task_counter = (unsigned int *)ws;
            const LcCodes codes = all_codes + (size_t)b0 * cfg.total;
            if (flags & LC_FLAG_ENC_SORT_V1)
                lc_enc_sort_kernel<<<nb, LC_SORT_THREADS, sort_smem, st>>>(cfg, codes, skeys, spos, first_bad, glist, ngroups, ivs,
                                                               sparse_variant ? tables : (const double *)0);
            else
                lc_enc_sort2_kernel<<<nb, LCS2_THREADS, LCS2_SMEM, st>>>(cfg, codes, skeys, spos, first_bad, glist, ngroups, ivs,
                                                                 sparse_variant ? tables : (const double *)0);
            LC_LAUNCHED();
            if (sparse_variant) {
                const size_t sp_smem = (size_t)LCS_BLOCK_WARPS * cfg.n * 8;
                int blocks_per_sm = (int)((200u * 1024u) / (sp_smem + 1024u));
                if (blocks_per_sm > LCS_BLOCKS_PER_SM) blocks_per_sm = LCS_BLOCKS_PER_SM;
                cudaMemsetAsync(task_counter, 0, 4, st);
                lc_enc_phase_a_sparse_kernel<<<lc_num_sms() * blocks_per_sm, 32 * LCS_BLOCK_WARPS, sp_smem, st>>>(
                    cfg, codes, nb, skeys, spos, first_bad, glist, ngroups, ivs, task_counter, tables, t2);
            }
#ifdef LC_DEBUG_VARIANTS
            else
                lc_enc_phase_a_kernel<<<nb, 256, (size_t)8 * cfg.n * 8, st>>>(cfg, codes, nb, skeys, spos, first_bad, ivs);
#endif
            LC_LAUNCHED();
            const size_t b2_smem = (size_t)slot_bytes + 4 + 24 * 4;
            if (cfg.mode == LC_MODE_REPAIRED && b2_smem <= 96 * 1024) {
                // B1: the serial recurrence (leaves its first finish bit in out_nbits); B2: parallel bit placement
                lc_enc_phase_b1_kernel<<<nb, 32, 0, st>>>(cfg, nb, first_bad, ivs, out_nbits + b0);
                LC_LAUNCHED();
                lc_enc_phase_b2_kernel<<<nb, LC_B2_THREADS, b2_smem, st>>>(cfg, nb, first_bad, ivs,
                                                                           slots + (size_t)b0 * slot_bytes, (uint32_t)slot_bytes,
                                                                           out_nbits + b0, status + b0, fault_index + b0);
            } else {
                // verbatim mode (the fault of defect D3 has to surface at its symbol) and slots beyond 96 KB
                lc_enc_phase_b_kernel<<<nb, 32, 0, st>>>(cfg, nb, first_bad, ivs, slots + (size_t)b0 * slot_bytes,
                                                        (uint32_t)slot_bytes, out_nbits + b0, status + b0, fault_index + b0);
            }
            LC_LAUNCHED();
        }
    } else {
        const int grid = lc_grid_for(cfg, B);
        if (cfg.sm_bytes > 48 * 1024)
            cudaFuncSetAttribute(lc_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.sm_bytes);
        lc_encode_kernel<<<grid, 32, cfg.sm_bytes, st>>>(cfg, all_codes, B, slots, (uint32_t)slot_bytes, out_nbits, status,
                                                         fault_index, (char *)scratch);
        LC_LAUNCHED();
    }
    if (out_bytes) {
        lc_scan_sizes_kernel<<<1, 1024, 0, st>>>(out_nbits, status, B, out_capacity, (long long *)out_offsets);
        LC_LAUNCHED();
        lc_compact_kernel<<<B, 128, 0, st>>>(slots, slot_bytes, out_nbits, status, (const long long *)out_offsets, out_bytes);
        LC_LAUNCHED();
    }
    return 0;
}

int lc_encode_batch(const int32_t *idx, int B, int imgs, int R, int C, int n_symbols, double adaptation_rate, int mode,
                    int has_ctx, void *scratch, int64_t scratch_bytes, uint8_t *slots, int64_t slot_bytes,
                    uint8_t *out_bytes, int64_t out_capacity, int64_t *out_offsets, int32_t *out_nbits, int32_t *status,
                    int32_t *fault_index, void *stream)
{
    return lc_encode_batch_t(idx, 4, B, imgs, R, C, n_symbols, adaptation_rate, mode, has_ctx, scratch, scratch_bytes, slots,
                             slot_bytes, out_bytes, out_capacity, out_offsets, out_nbits, status, fault_index, 0, stream);
}

int lc_decode_batch_t(const uint8_t *bytes, const int64_t *offsets, const int32_t *nbits, int B, int imgs, int R, int C,
                      int n_symbols, double adaptation_rate, int mode, int has_ctx, void *scratch, int64_t scratch_bytes,
                      void *idx_out, int idx_bytes, const float *deq_table, float *deq_out, int32_t *status,
                      int32_t *fault_index, int flags, void *stream)
{
    LcCoderCfg cfg;
    if (B < 0 || !bytes || !offsets || !nbits || !scratch || !status || !fault_index) return -22;
    if (!idx_out && !deq_out) return -22;
    if (B == 0) return 0;
    if (deq_out && !deq_table) return -22;
    int rc = lc_make_cfg(cfg, imgs, R, C, n_symbols, adaptation_rate, mode, has_ctx);
    if (rc) return rc;
    if (idx_out && !lc_idx_bytes_ok(idx_bytes, cfg.n)) return -22;
    if ((((uintptr_t)bytes) & 3) != 0 || (((uintptr_t)scratch) & 15) != 0) return -22;
    const int grid = lc_grid_for(cfg, B);
    if (scratch_bytes < lc_scratch_need(cfg, B)) return -12;
    lc_prepare_device();
    cudaStream_t st = (cudaStream_t)stream;
    const LcIdxOut out(idx_out, idx_out ? idx_bytes : 0);
    if (cfg.sm_bytes > 48 * 1024)
        cudaFuncSetAttribute(lc_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.sm_bytes);
    if (lcd_eligible(cfg) && !(flags & (LC_FLAG_DEC_NO_SMALL | LC_FLAG_DEC_REGISTER_MODEL | LC_FLAG_DEC_SERIAL))) {
        // alphabets of up to 16 symbols: dense model per context (global scratch, L2 resident), one warp per stream;
        // anything unusual (corrupt streams) is flagged and redone by the generic kernel
        const size_t sm = lcd_smem_bytes(cfg.C);
        const int gs = lc_small_grid(B);
        // the generic kernel's scratch region follows the tables (the redo pass runs after this kernel, but flagged
        // streams must not find their scratch overwritten by a later launch's tables -- they are separate areas)
        char *tabs = (char *)scratch + (((size_t)grid * cfg.scratch_stride + 255) & ~(size_t)255);
        const long long *offs = (const long long *)offsets;
        switch (cfg.n) {
        case 2: lc_decode_small_kernel<2><<<gs, 32, sm, st>>>(cfg, bytes, offs, nbits, B, out, deq_table, deq_out, status, fault_index, tabs); break;
        case 4: lc_decode_small_kernel<4><<<gs, 32, sm, st>>>(cfg, bytes, offs, nbits, B, out, deq_table, deq_out, status, fault_index, tabs); break;
        case 8: lc_decode_small_kernel<8><<<gs, 32, sm, st>>>(cfg, bytes, offs, nbits, B, out, deq_table, deq_out, status, fault_index, tabs); break;
        default: lc_decode_small_kernel<16><<<gs, 32, sm, st>>>(cfg, bytes, offs, nbits, B, out, deq_table, deq_out, status, fault_index, tabs); break;
        }
        LC_LAUNCHED();
        lc_decode_kernel<<<grid, 32, cfg.sm_bytes, st>>>(cfg, bytes, offs, nbits, B, out, deq_table, deq_out, status,
                                                         fault_index, (char *)scratch, LC_NEEDS_GENERIC);
        LC_LAUNCHED();
        return 0;
    }
    if (lc_use_decoder_v2(cfg, flags)) {
        // decoder/updater warps; streams with a context of more than 32 distinct symbols are flagged
        // and redone from scratch by the generic kernel
        LcV2Cfg vc;
        lcv_cfg_make(cfg, &vc);
        const int g2 = lc_v2_grid(vc, B, flags);
        const int gmax = lc_v2_grid(vc, B, LC_FLAG_DEC_THROUGHPUT_BUILD);
        double *tables = (double *)((char *)scratch + (((size_t)gmax * vc.g_stride + 255) & ~(size_t)255));
        if (vc.sm_bytes > 64 * 1024) return -22;
        lc_v2_tables_kernel<<<cfg.n, 32, (size_t)cfg.n * 8, st>>>(cfg, tables);
        LC_LAUNCHED();
        // records of the models after two visits: the updater's job for a second visit becomes a copy
        char *t2 = (char *)0;
        // (measured neutral on the decode time -- the updater is not the bottleneck -- so only where the table's
        // n^2 updates are small against the batch: >= 4096 streams at n = 256)
        bool v3 = false;
#ifdef LC_DEBUG_VARIANTS
        v3 = (flags & LC_FLAG_DEBUG_DEC_V3) && cfg.total <= (1 << 21);
#endif
        if (!v3 && (long long)B * 16 >= (long long)cfg.n * cfg.n) {
            t2 = (char *)tables + ((lcv_tables_bytes(cfg.n) + 255) & ~(uint64_t)255);
            lc_t2_kernel<<<lc_num_sms() * 4, 32 * LCS_BLOCK_WARPS, (size_t)LCS_BLOCK_WARPS * cfg.n * 8, st>>>(cfg, tables, t2);
            LC_LAUNCHED();
        }
        if (v3) {
#ifdef LC_DEBUG_VARIANTS
            if (idx_bytes != 4 || !idx_out) return -22;
            lc_decode_v3_kernel<<<g2, 32 * LC3_WARPS, vc.sm_bytes, st>>>(cfg, vc, bytes, (const long long *)offsets, nbits,
                                                                         B, (int *)idx_out, deq_table, deq_out, status,
                                                                         fault_index, (char *)scratch, tables);
#endif
        } else {
            const bool w8 = cfg.n == 256 && cfg.C == 512 && cfg.R == 16 && cfg.imgs == 1 && !(flags & LC_FLAG_DEC_GENERIC_SHAPE);
            const bool thr = lc_v2_per_sm(vc, B, flags) > LC_V2_LAT_PER_SM;
            auto kern = thr ? (w8 ? lc_decode_v2_w8_thr_kernel : lc_decode_v2_thr_kernel)
                            : (w8 ? lc_decode_v2_w8_kernel : lc_decode_v2_kernel);
            kern<<<g2, 32 * LCV_WARPS, vc.sm_bytes, st>>>(cfg, vc, bytes, (const long long *)offsets, nbits, B, out,
                                                          deq_table, deq_out, status, fault_index, (char *)scratch, tables,
                                                          t2);
        }
        LC_LAUNCHED();
        lc_decode_kernel<<<grid, 32, cfg.sm_bytes, st>>>(cfg, bytes, (const long long *)offsets, nbits, B, out,
                                                         deq_table, deq_out, status, fault_index, (char *)scratch,
                                                         LC_NEEDS_GENERIC);
        LC_LAUNCHED();
        return 0;
    }
    if (cfg.mode == LC_MODE_REPAIRED && cfg.has_ctx && !(flags & LC_FLAG_DEC_SERIAL)) {
        // register-model kernel; streams it cannot finish (a context with more than 32 distinct symbols) are
        // flagged and redone from scratch by the generic kernel
        if (cfg.sm_bytes > 48 * 1024)
            cudaFuncSetAttribute(lc_fast_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.sm_bytes);
        lc_fast_decode_kernel<<<grid, 32, cfg.sm_bytes, st>>>(cfg, bytes, (const long long *)offsets, nbits, B, out,
                                                              deq_table, deq_out, status, fault_index, (char *)scratch);
        LC_LAUNCHED();
        lc_decode_kernel<<<grid, 32, cfg.sm_bytes, st>>>(cfg, bytes, (const long long *)offsets, nbits, B, out,
                                                         deq_table, deq_out, status, fault_index, (char *)scratch,
                                                         LC_NEEDS_GENERIC);
        LC_LAUNCHED();
        return 0;
    }
    lc_decode_kernel<<<grid, 32, cfg.sm_bytes, st>>>(cfg, bytes, (const long long *)offsets, nbits, B, out, deq_table,
                                                     deq_out, status, fault_index, (char *)scratch, 0);
    LC_LAUNCHED();
    return 0;
}

int lc_decode_batch(const uint8_t *bytes, const int64_t *offsets, const int32_t *nbits, int B, int imgs, int R, int C,
                    int n_symbols, double adaptation_rate, int mode, int has_ctx, void *scratch, int64_t scratch_bytes,
                    int32_t *idx_out, const float *deq_table, float *deq_out, int32_t *status, int32_t *fault_index,
                    void *stream)
{
    if (!idx_out) return -22;
    return lc_decode_batch_t(bytes, offsets, nbits, B, imgs, R, C, n_symbols, adaptation_rate, mode, has_ctx, scratch,
                             scratch_bytes, idx_out, 4, deq_table, deq_out, status, fault_index, 0, stream);
}

#ifdef LC_DEC_PROFILE
// debug builds only (tools/dec_profile.py): read and clear the decoder's cycle counters
int lc_debug_profile(unsigned long long *out64)
{
    static unsigned long long zero[64];
    if (cudaMemcpyFromSymbol(out64, lc_prof_global, sizeof(zero)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(lc_prof_global, zero, sizeof(zero)) != cudaSuccess) return -1;
    return 0;
}
#endif

int lc_model_update(double *vec, int n_symbols, int symbol, double adaptation_rate, void *stream)
{
    LcCoderCfg cfg;
    if (!vec || (((uintptr_t)vec) & 7) != 0) return -22;
    int rc = lc_make_cfg(cfg, 1, 1, 1, n_symbols, adaptation_rate, LC_MODE_REPAIRED, 0);
    if (rc) return rc;
    if (symbol < -n_symbols || symbol >= n_symbols) return -22;
    lc_model_update_kernel<<<1, 32, (size_t)cfg.n * 8, (cudaStream_t)stream>>>(cfg, vec, symbol);
    LC_LAUNCHED();
    return 0;
}

static int lc_stateful_check(LcCoderCfg &cfg, LcStatefulTable &T, int imgs, int R, int C, int n, double rate, int mode,
                             int has_ctx, void *table, int64_t table_bytes)
{
    int rc = lc_make_cfg(cfg, imgs, R, C, n, rate, mode, has_ctx);
    if (rc) return rc;
    if (!table || (((uintptr_t)table) & 255) != 0) return -22;
    if (table_bytes < (int64_t)lc_stateful_bytes(n, cfg.has_ctx)) return -12;
    T.valid = (unsigned char *)table;
    T.counts = (int *)((char *)table + lc_stateful_off_counts(n, cfg.has_ctx));
    T.vecs = (double *)((char *)table + lc_stateful_off_vecs(n, cfg.has_ctx));
    return 0;
}

int64_t lc_stateful_table_bytes(int n_symbols, int has_ctx)
{
    if (n_symbols < 2 || n_symbols > 1024 || (n_symbols & (n_symbols - 1))) return -22;
    return (int64_t)lc_stateful_bytes(n_symbols, has_ctx ? 1 : 0);
}

int64_t lc_stateful_table_offset(int n_symbols, int has_ctx, int which)
{
    if (lc_stateful_table_bytes(n_symbols, has_ctx) < 0) return -22;
    if (which == 0) return 0;
    if (which == 1) return (int64_t)lc_stateful_off_counts(n_symbols, has_ctx ? 1 : 0);
    if (which == 2) return (int64_t)lc_stateful_off_vecs(n_symbols, has_ctx ? 1 : 0);
    return -22;
}

int lc_stateful_encode(const int32_t *idx, int imgs, int R, int C, int n_symbols, double adaptation_rate, int mode,
                       int has_ctx, void *table, int64_t table_bytes, uint8_t *slot, int64_t slot_bytes,
                       int32_t *out_nbits, int32_t *status, int32_t *fault_index, void *stream)
{
    LcCoderCfg cfg;
    LcStatefulTable T;
    if (!idx || !slot || !out_nbits || !status || !fault_index) return -22;
    if (slot_bytes < 16 || (slot_bytes & 3) || slot_bytes > 0xfffffff0ll) return -22;
    int rc = lc_stateful_check(cfg, T, imgs, R, C, n_symbols, adaptation_rate, mode, has_ctx, table, table_bytes);
    if (rc) return rc;
    const size_t smem = (size_t)cfg.n * 8;
    lc_stateful_encode_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(cfg, idx, T, slot, (uint32_t)slot_bytes, out_nbits,
                                                                    status, fault_index);
    LC_LAUNCHED();
    return 0;
}

int lc_stateful_decode(const uint8_t *bytes, int64_t nbytes, int imgs, int R, int C, int n_symbols,
                       double adaptation_rate, int mode, int has_ctx, void *table, int64_t table_bytes, int32_t *idx_out,
                       int32_t *status, int32_t *fault_index, void *stream)
{
    LcCoderCfg cfg;
    LcStatefulTable T;
    if (!bytes || nbytes < 0 || !idx_out || !status || !fault_index || (((uintptr_t)bytes) & 3) != 0) return -22;
    int rc = lc_stateful_check(cfg, T, imgs, R, C, n_symbols, adaptation_rate, mode, has_ctx, table, table_bytes);
    if (rc) return rc;
    const size_t smem = (size_t)cfg.n * 8;
    lc_stateful_decode_kernel<<<1, 32, smem, (cudaStream_t)stream>>>(cfg, bytes, (long long)nbytes, T, idx_out, status,
                                                                    fault_index);
    LC_LAUNCHED();
    return 0;
}

} // extern "C"
