/*
 * TEST INFRASTRUCTURE ONLY -- a minimal CPU emulation of the CUDA SIMT features the coder kernels
 * use, so the warp-cooperative device code in image_compression_2_b200/csrc/ (the .cuh files) can be exercised
 * in the GPU-less build container (tests/test_hostsim.py).  It is never part of the product path.
 *
 * One warp = 32 coroutines (hand-rolled x86-64 stack switch) scheduled round-robin on one OS
 * thread.  Warp collectives (__shfl*_sync, __ballot_sync, __syncwarp) are rendezvous points; the
 * emulator aborts if the 32 lanes do not all reach the SAME collective call site (divergence bug)
 * or if a lane exits while others wait.  run_warp() runs one warp to completion (blocks whose warps
 * do not synchronise with one another are run warp after warp); run_block() schedules all warps of
 * a block round-robin so that they can hand work to one another through shared memory: spin loops
 * call spin_yield(), __syncthreads() is a real block barrier there.
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#if !defined(__x86_64__)
#error "cuda_emul.h: x86-64 only"
#endif

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)

namespace emu {

struct Dim { unsigned x, y, z; };

struct Warp {
    void *lane_sp[32];
    char *lane_stack[32];
    bool done[32];
    void *main_sp;
    int cur;
    int arrived;
    uint64_t gen;
    uint64_t xchg[2][32];
    int site[32];
    void (*fn)(void *);
    void *arg;
    unsigned block, grid, warp, nwarps;
};

extern Warp *g_warp;

struct Block { int nwarps; int arrived; uint64_t gen; uint64_t events; };
extern Block *g_block; /* non-null while run_block() is scheduling cooperating warps */

extern "C" void emu_switch(void **save_sp, void *load_sp);

inline Dim tid() { return Dim{g_warp->warp * 32u + (unsigned)g_warp->cur, 0, 0}; }
inline Dim bid() { return Dim{g_warp->block, 0, 0}; }
inline Dim bdim() { return Dim{32u * g_warp->nwarps, 1, 1}; }
inline Dim gdim() { return Dim{g_warp->grid, 1, 1}; }

inline void yield_to_main() {
    Warp *w = g_warp;
    emu_switch(&w->lane_sp[w->cur], w->main_sp);
}

[[noreturn]] inline void die(const char *msg, int site) {
    fprintf(stderr, "cuda_emul: %s (call site line %d, lane %d)\n", msg, site, g_warp ? g_warp->cur : -1);
    abort();
}

/* rendezvous of all 32 lanes; returns the generation index that was completed */
inline uint64_t rendezvous(int site) {
    Warp *w = g_warp;
    const int me = w->cur;
    const uint64_t g = w->gen;
    w->site[me] = site;
    if (++w->arrived == 32) {
        for (int i = 0; i < 32; i++)
            if (w->site[i] != site) {
                fprintf(stderr, "cuda_emul: sites:");
                for (int q = 0; q < 32; q++) fprintf(stderr, " %d", w->site[q]);
                fprintf(stderr, "\n");
                die("lanes reached different collectives", site);
            }
        w->arrived = 0;
        w->gen = g + 1;
    } else {
        while (w->gen == g) yield_to_main();
    }
    return g;
}

/* a lane waiting in a spin loop for another warp: give the other lanes and warps a turn */
inline void spin_yield() { yield_to_main(); }

/* __syncthreads(): a block barrier under run_block(), a warp rendezvous under run_warp() */
inline void block_barrier(int site) {
    Block *b = g_block;
    if (!b) { (void)rendezvous(site); return; }
    const uint64_t g = b->gen;
    b->events++;
    if (++b->arrived == 32 * b->nwarps) { b->arrived = 0; b->gen = g + 1; }
    else while (b->gen == g) yield_to_main();
}

template <typename T> inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

template <typename T> inline T shfl(T v, int src, int site) {
    Warp *w = g_warp;
    const int me = w->cur;
    w->xchg[w->gen & 1][me] = to_bits(v);
    uint64_t g = rendezvous(site);
    return from_bits<T>(w->xchg[g & 1][src & 31]);
}
template <typename T> inline T shfl_up(T v, unsigned delta, int site) {
    Warp *w = g_warp;
    const int me = w->cur;
    w->xchg[w->gen & 1][me] = to_bits(v);
    uint64_t g = rendezvous(site);
    int src = me - (int)delta;
    return src < 0 ? v : from_bits<T>(w->xchg[g & 1][src]);
}
template <typename T> inline T shfl_down(T v, unsigned delta, int site) {
    Warp *w = g_warp;
    const int me = w->cur;
    w->xchg[w->gen & 1][me] = to_bits(v);
    uint64_t g = rendezvous(site);
    int src = me + (int)delta;
    return src > 31 ? v : from_bits<T>(w->xchg[g & 1][src]);
}
inline unsigned ballot(int pred, int site) {
    Warp *w = g_warp;
    const int me = w->cur;
    w->xchg[w->gen & 1][me] = pred ? 1u : 0u;
    uint64_t g = rendezvous(site);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(w->xchg[g & 1][i] & 1u) << i;
    return m;
}

void run_warp(void (*fn)(void *), void *arg, unsigned block, unsigned grid, unsigned warp = 0, unsigned nwarps = 1);
void run_block(void (*fn)(void *), void *arg, unsigned block, unsigned grid, unsigned nwarps);

} // namespace emu

#define threadIdx (emu::tid())
#define blockIdx (emu::bid())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())

#define __shfl_sync(mask, v, src) emu::shfl((v), (src), __LINE__)
#define __shfl_xor_sync(mask, v, lanemask) emu::shfl((v), (int)((emu::tid().x & 31u) ^ (unsigned)(lanemask)), __LINE__)
#define __shfl_up_sync(mask, v, delta) emu::shfl_up((v), (delta), __LINE__)
#define __shfl_down_sync(mask, v, delta) emu::shfl_down((v), (delta), __LINE__)
#define __ballot_sync(mask, pred) emu::ballot((pred), __LINE__)
#define __syncwarp() ((void)emu::rendezvous(__LINE__))
#define __syncthreads() (emu::block_barrier(__LINE__))
/* a __shared__ variable inside a device function: blocks run one after the other here, so one static instance is the
 * block's (the emulator's warps of a block are coroutines on this thread) */
#define __shared__ static

static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __clzll(long long x) { return x == 0 ? 64 : __builtin_clzll((unsigned long long)x); }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0x7;
        r |= (unsigned)((v >> (8 * sel)) & 0xff) << (8 * i);
    }
    return r;
}
/* IEEE double ops; the translation unit is built with -ffp-contract=off so these never fuse */
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline long long __double2ll_rz(double a) { return (long long)a; }
static inline double __ll2double_rn(long long a) { return (double)a; }
static inline double __int2double_rn(int a) { return (double)a; }
static inline unsigned int atomicAdd(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = o + v; return o; }
static inline unsigned int atomicOr(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = o | v; return o; }
static inline int atomicMin(int *p, int v) { int o = *p; if (v < o) *p = v; return o; }
static inline unsigned int atomicXor(unsigned int *p, unsigned int v) { unsigned int o = *p; *p = o ^ v; return o; }
static inline unsigned int atomicCAS(unsigned int *p, unsigned int c, unsigned int v) { unsigned int o = *p; if (o == c) *p = v; return o; }
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline unsigned int __funnelshift_lc(unsigned int lo, unsigned int hi, unsigned int sh)
{
    if (sh > 32u) sh = 32u;
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return sh == 32u ? lo : (unsigned int)((v << sh) >> 32);
}
static inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int)(((uint64_t)a * b) >> 32); }
static inline long long __double_as_longlong(double d) { long long b; memcpy(&b, &d, 8); return b; }
template <typename T> static inline T __ldcg(const T *p) { return *p; }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
