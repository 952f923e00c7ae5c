/*
 * TEST INFRASTRUCTURE ONLY -- runs the coder's device code (lc_coder.cuh) on the CPU through the
 * SIMT emulator in cuda_emul.h.  Built by tests/hostsim/build.py with g++ -ffp-contract=off.
 * Used by tests/test_hostsim.py to check the warp-cooperative logic without a GPU.
 */
#define LC_HOSTSIM 1
#include "cuda_emul.h"

#include <vector>

#include "../../image_compression_2_b200/csrc/lc_coder.cuh"
#include "../../image_compression_2_b200/csrc/lc_encoder_par.cuh"
#include "../../image_compression_2_b200/csrc/lc_decoder_fast.cuh"
#include "../../image_compression_2_b200/csrc/lc_decoder_v2.cuh"
#include "../../image_compression_2_b200/csrc/lc_decoder_v3.cuh"
#include "../../image_compression_2_b200/csrc/lc_decoder_small.cuh"
#include "../../image_compression_2_b200/csrc/lc_encoder_sparse.cuh"
#include "../../image_compression_2_b200/csrc/lc_encoder_sort.cuh"
#include "../../image_compression_2_b200/csrc/lc_encoder_pack.cuh"
#include <algorithm>
#include <numeric>

namespace emu {
Warp *g_warp = nullptr;
Block *g_block = nullptr;

asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

static void lane_entry()
{
    Warp *w = g_warp;
    w->fn(w->arg);
    w->done[w->cur] = true;
    emu_switch(&w->lane_sp[w->cur], w->main_sp);
    abort(); // a finished lane is never resumed
}

static void warp_setup(Warp &w, void (*fn)(void *), void *arg, unsigned block, unsigned grid, unsigned warp, unsigned nwarps)
{
    static const size_t STACK = 256 * 1024;
    memset(&w, 0, sizeof(w));
    w.fn = fn; w.arg = arg; w.block = block; w.grid = grid; w.warp = warp; w.nwarps = nwarps;
    for (int i = 0; i < 32; i++) {
        w.lane_stack[i] = (char *)aligned_alloc(64, STACK);
        uintptr_t top = ((uintptr_t)w.lane_stack[i] + STACK) & ~(uintptr_t)15;
        uint64_t *sp = (uint64_t *)top;
        *--sp = 0;                      // fake return address of lane_entry
        *--sp = (uint64_t)&lane_entry;  // popped by emu_switch's ret
        for (int r = 0; r < 6; r++) *--sp = 0;
        w.lane_sp[i] = sp;
    }
}

// all warps of one block, scheduled round-robin (one turn per lane) so that warps can wait for one another
void run_block(void (*fn)(void *), void *arg, unsigned block, unsigned grid, unsigned nwarps)
{
    std::vector<Warp> ws(nwarps);
    for (unsigned w = 0; w < nwarps; w++) warp_setup(ws[w], fn, arg, block, grid, w, nwarps);
    Block blk{(int)nwarps, 0, 0, 0};
    Warp *saved = g_warp;
    Block *saved_b = g_block;
    g_block = &blk;
    int alive = 32 * (int)nwarps;
    uint64_t idle = 0;
    while (alive > 0) {
        uint64_t sig0 = blk.events + blk.gen;
        for (unsigned w = 0; w < nwarps; w++) sig0 += ws[w].gen * 64 + (uint64_t)ws[w].arrived;
        const int alive0 = alive;
        for (unsigned w = 0; w < nwarps; w++) {
            g_warp = &ws[w];
            for (int i = 0; i < 32; i++) {
                if (ws[w].done[i]) continue;
                ws[w].cur = i;
                emu_switch(&ws[w].main_sp, ws[w].lane_sp[i]);
                if (ws[w].done[i]) alive--;
            }
        }
        uint64_t sig1 = blk.events + blk.gen;
        for (unsigned w = 0; w < nwarps; w++) sig1 += ws[w].gen * 64 + (uint64_t)ws[w].arrived;
        if (sig1 == sig0 && alive == alive0) { if (++idle > 200000) die("run_block: no progress (deadlock between warps?)", -1); }
        else idle = 0;
    }
    g_warp = saved;
    g_block = saved_b;
    for (unsigned w = 0; w < nwarps; w++)
        for (int i = 0; i < 32; i++) free(ws[w].lane_stack[i]);
}

void run_warp(void (*fn)(void *), void *arg, unsigned block, unsigned grid, unsigned warp, unsigned nwarps)
{
    Warp w;
    warp_setup(w, fn, arg, block, grid, warp, nwarps);
    Warp *saved = g_warp;
    g_warp = &w;
    int alive = 32;
    while (alive > 0) {
        const uint64_t gen0 = w.gen;
        const int arrived0 = w.arrived;
        const int alive0 = alive;
        for (int i = 0; i < 32; i++) {
            if (w.done[i]) continue;
            w.cur = i;
            emu_switch(&w.main_sp, w.lane_sp[i]);
            if (w.done[i]) alive--;
        }
        if (alive > 0 && w.gen == gen0 && w.arrived == arrived0 && alive == alive0)
            die("deadlock: lanes wait at a collective that not all 32 lanes reach", -1);
    }
    g_warp = saved;
    for (int i = 0; i < 32; i++) free(w.lane_stack[i]);
}
} // namespace emu

struct EncArgs {
    LcCoderCfg cfg; const int *codes; int B; unsigned char *out_slots; uint32_t slot_bytes;
    int *nbits, *status, *fault; char *scratch; char *smem;
};
static void enc_body(void *p)
{
    EncArgs *a = (EncArgs *)p;
    lc_encode_block(a->cfg, a->codes, a->B, a->out_slots, a->slot_bytes, a->nbits, a->status, a->fault, a->scratch,
                    a->smem);
}
struct DecArgs {
    LcCoderCfg cfg; const unsigned char *bytes; const long long *offsets; const int *nbits; int B; int *out; const float *deq_table;
    float *deq_out; int *status, *fault; char *scratch; char *smem; int only_flagged;
};
static void dec_body(void *p)
{
    DecArgs *a = (DecArgs *)p;
    lc_decode_block(a->cfg, a->bytes, a->offsets, a->nbits, a->B, a->out, a->deq_table, a->deq_out, a->status, a->fault,
                    a->scratch, a->smem, a->only_flagged);
}
static void decfast_body(void *p)
{
    DecArgs *a = (DecArgs *)p;
    lc_fast_decode_block(a->cfg, a->bytes, a->offsets, a->nbits, a->B, a->out, a->deq_table, a->deq_out, a->status,
                         a->fault, a->scratch, a->smem);
}

static int make_cfg(LcCoderCfg &cfg, int imgs, int R, int C, int n, double rate, int mode, int has_ctx)
{
    memset(&cfg, 0, sizeof(cfg));
    cfg.n = n; cfg.R = R; cfg.C = C; cfg.imgs = imgs; cfg.has_ctx = has_ctx; cfg.mode = mode; cfg.rate = rate;
    return lc_cfg_finalize(&cfg);
}

extern "C" int hostsim_encode(const int *codes, int B, int imgs, int R, int C, int n, double rate, int mode,
                              int has_ctx, unsigned char *out_slots, unsigned slot_bytes, int *nbits, int *status,
                              int *fault, int grid)
{
    EncArgs a;
    int rc = make_cfg(a.cfg, imgs, R, C, n, rate, mode, has_ctx);
    if (rc) return rc;
    std::vector<char> scratch((size_t)grid * a.cfg.scratch_stride + 256);
    std::vector<char> smem(a.cfg.sm_bytes + 64);
    a.codes = codes; a.B = B; a.out_slots = out_slots; a.slot_bytes = slot_bytes;
    a.nbits = nbits; a.status = status; a.fault = fault; a.scratch = scratch.data();
    a.smem = (char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (int b = 0; b < grid; b++) emu::run_warp(enc_body, &a, (unsigned)b, (unsigned)grid);
    return 0;
}

extern "C" int hostsim_decode(const unsigned char *bytes, const long long *offsets, const int *nbits, int B, int imgs, int R, int C,
                              int n, double rate, int mode, int has_ctx, int *out, const float *deq_table,
                              float *deq_out, int *status, int *fault, int grid)
{
    DecArgs a;
    int rc = make_cfg(a.cfg, imgs, R, C, n, rate, mode, has_ctx);
    if (rc) return rc;
    std::vector<char> scratch((size_t)grid * a.cfg.scratch_stride + 256);
    std::vector<char> smem(a.cfg.sm_bytes + 64);
    a.bytes = bytes; a.offsets = offsets; a.nbits = nbits; a.B = B; a.out = out; a.deq_table = deq_table; a.deq_out = deq_out;
    a.status = status; a.fault = fault; a.scratch = scratch.data(); a.only_flagged = 0;
    a.smem = (char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (int b = 0; b < grid; b++) emu::run_warp(dec_body, &a, (unsigned)b, (unsigned)grid);
    return 0;
}

// fast decoder (repaired mode, (left,up) contexts) followed by the generic redo pass, as the host does
extern "C" int hostsim_decode_fast(const unsigned char *bytes, const long long *offsets, const int *nbits, int B,
                                   int imgs, int R, int C, int n, double rate, int *out, const float *deq_table,
                                   float *deq_out, int *status, int *fault, int grid, int *n_redone)
{
    DecArgs a;
    int rc = make_cfg(a.cfg, imgs, R, C, n, rate, LC_MODE_REPAIRED, 1);
    if (rc) return rc;
    std::vector<char> scratch((size_t)grid * a.cfg.scratch_stride + 256);
    std::vector<char> smem(a.cfg.sm_bytes + 64);
    a.bytes = bytes; a.offsets = offsets; a.nbits = nbits; a.B = B; a.out = out; a.deq_table = deq_table; a.deq_out = deq_out;
    a.status = status; a.fault = fault; a.scratch = scratch.data(); a.only_flagged = 0;
    a.smem = (char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (int b = 0; b < grid; b++) emu::run_warp(decfast_body, &a, (unsigned)b, (unsigned)grid);
    int redo = 0;
    for (int b = 0; b < B; b++) redo += status[b] == LC_NEEDS_GENERIC;
    *n_redone = redo;
    a.only_flagged = LC_NEEDS_GENERIC;
    for (int b = 0; b < grid; b++) emu::run_warp(dec_body, &a, (unsigned)b, (unsigned)grid);
    return 0;
}

// small-alphabet decoder (dense shared-memory model, n <= 16) followed by the generic redo pass, as the host does
static void decsmall_body(void *p)
{
    DecArgs *a = (DecArgs *)p;
    const LcIdxOut out(a->out);
    switch (a->cfg.n) {
    case 2: lcd_decode_block<2>(a->cfg, a->bytes, a->offsets, a->nbits, a->B, out, a->deq_table, a->deq_out, a->status, a->fault, a->scratch, a->smem); break;
    case 4: lcd_decode_block<4>(a->cfg, a->bytes, a->offsets, a->nbits, a->B, out, a->deq_table, a->deq_out, a->status, a->fault, a->scratch, a->smem); break;
    case 8: lcd_decode_block<8>(a->cfg, a->bytes, a->offsets, a->nbits, a->B, out, a->deq_table, a->deq_out, a->status, a->fault, a->scratch, a->smem); break;
    default: lcd_decode_block<16>(a->cfg, a->bytes, a->offsets, a->nbits, a->B, out, a->deq_table, a->deq_out, a->status, a->fault, a->scratch, a->smem); break;
    }
}
extern "C" int hostsim_decode_small(const unsigned char *bytes, const long long *offsets, const int *nbits, int B,
                                    int imgs, int R, int C, int n, double rate, int *out, const float *deq_table,
                                    float *deq_out, int *status, int *fault, int grid, int *n_redone)
{
    DecArgs a;
    int rc = make_cfg(a.cfg, imgs, R, C, n, rate, LC_MODE_REPAIRED, 1);
    if (rc) return rc;
    if (!lcd_eligible(a.cfg)) return -22;
    std::vector<char> scratch((size_t)grid * (a.cfg.scratch_stride > lcd_tab_bytes(n) ? a.cfg.scratch_stride : lcd_tab_bytes(n)) + 256);
    std::vector<char> smem_small(lcd_smem_bytes(C) + 64), smem(a.cfg.sm_bytes + 64);
    a.bytes = bytes; a.offsets = offsets; a.nbits = nbits; a.B = B; a.out = out; a.deq_table = deq_table; a.deq_out = deq_out;
    a.status = status; a.fault = fault; a.scratch = scratch.data(); a.only_flagged = 0;
    a.smem = (char *)(((uintptr_t)smem_small.data() + 15) & ~(uintptr_t)15);
    for (int b = 0; b < grid; b++) emu::run_warp(decsmall_body, &a, (unsigned)b, (unsigned)grid);
    int redo = 0;
    for (int b = 0; b < B; b++) redo += status[b] == LC_NEEDS_GENERIC;
    *n_redone = redo;
    a.only_flagged = LC_NEEDS_GENERIC;
    a.smem = (char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    for (int b = 0; b < grid; b++) emu::run_warp(dec_body, &a, (unsigned)b, (unsigned)grid);
    return 0;
}

// decoder v2 (decoder warp + updater warps per block) followed by the generic redo pass, as the host does
struct DecV2Args { DecArgs d; LcV2Cfg vc; double *tables; char *scratch2; char *t2; };
static void decv2_tables_body(void *p)
{
    DecV2Args *a = (DecV2Args *)p;
    lcv_tables_block(a->d.cfg, a->tables, a->d.smem);
}
static void decv2_t2_body(void *p)
{
    DecV2Args *a = (DecV2Args *)p;
    lcv_t2_block(a->d.cfg, a->tables, a->t2, a->d.smem);
}
static void decv2_body(void *p)
{
    DecV2Args *a = (DecV2Args *)p;
    if (a->d.cfg.n == 256 && a->d.cfg.C == 512 && a->d.cfg.R == 16 && a->d.cfg.imgs == 1)
        lcv_decode_block<256, 512, 16>(a->d.cfg, a->vc, a->d.bytes, a->d.offsets, a->d.nbits, a->d.B, a->d.out, a->d.deq_table,
                                       a->d.deq_out, a->d.status, a->d.fault, a->scratch2, a->tables, a->t2, a->d.smem);
    else
    lcv_decode_block<0, 0, 0>(a->d.cfg, a->vc, a->d.bytes, a->d.offsets, a->d.nbits, a->d.B, a->d.out, a->d.deq_table, a->d.deq_out,
                     a->d.status, a->d.fault, a->scratch2, a->tables, a->t2, a->d.smem);
}
static void decv3_body(void *p)
{
    DecV2Args *a = (DecV2Args *)p;
    lc3_decode_block(a->d.cfg, a->vc, a->d.bytes, a->d.offsets, a->d.nbits, a->d.B, a->d.out, a->d.deq_table, a->d.deq_out,
                     a->d.status, a->d.fault, a->scratch2, a->tables, a->d.smem);
}
static int g_decoder_version = 2;
extern "C" void hostsim_set_decoder_version(int v) { g_decoder_version = v; }
extern "C" int hostsim_decode_v2(const unsigned char *bytes, const long long *offsets, const int *nbits, int B,
                                 int imgs, int R, int C, int n, double rate, int *out, const float *deq_table,
                                 float *deq_out, int *status, int *fault, int grid, int *n_redone)
{
    DecV2Args a;
    int rc = make_cfg(a.d.cfg, imgs, R, C, n, rate, LC_MODE_REPAIRED, 1);
    if (rc) return rc;
    if (!lcv_eligible(a.d.cfg)) return -22;
    lcv_cfg_make(a.d.cfg, &a.vc);
    std::vector<char> scratch((size_t)grid * a.d.cfg.scratch_stride + 256);
    std::vector<char> scratch2((size_t)grid * a.vc.g_stride + 256, (char)0xAB); // never cleared on the GPU either
    std::vector<double> tables(lcv_tables_bytes(n) / 8 + 8, -777.0);
    std::vector<char> smem(std::max((size_t)a.vc.sm_bytes, std::max((size_t)a.d.cfg.sm_bytes, (size_t)(n + 64) * 8)) + 64);
    a.d.bytes = bytes; a.d.offsets = offsets; a.d.nbits = nbits; a.d.B = B; a.d.out = out; a.d.deq_table = deq_table;
    a.d.deq_out = deq_out; a.d.status = status; a.d.fault = fault; a.d.scratch = scratch.data(); a.d.only_flagged = 0;
    a.d.smem = (char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    a.tables = tables.data();
    a.scratch2 = (char *)(((uintptr_t)scratch2.data() + 255) & ~(uintptr_t)255);
    std::vector<char> t2rec;
    a.t2 = (char *)0;
    for (int b = 0; b < 3; b++) emu::run_warp(decv2_tables_body, &a, (unsigned)b, 3u);
    if (g_decoder_version == 2 && n <= 64) { // small alphabets also exercise the two-visit records (updater copy path)
        t2rec.assign((size_t)n * n * 64 + 64, (char)0xCD);
        a.t2 = t2rec.data();
        for (int b = 0; b < 2; b++)
            for (int w = 0; w < 2; w++) emu::run_warp(decv2_t2_body, &a, (unsigned)b, 2u, (unsigned)w, 2u);
    }
    if (g_decoder_version == 3)
        for (int b = 0; b < grid; b++) emu::run_block(decv3_body, &a, (unsigned)b, (unsigned)grid, LC3_WARPS);
    else
        for (int b = 0; b < grid; b++) emu::run_block(decv2_body, &a, (unsigned)b, (unsigned)grid, LCV_WARPS);
    int redo = 0;
    for (int b = 0; b < B; b++) redo += status[b] == LC_NEEDS_GENERIC;
    *n_redone = redo;
    a.d.only_flagged = LC_NEEDS_GENERIC;
    for (int b = 0; b < grid; b++) emu::run_warp(dec_body, &a.d, (unsigned)b, (unsigned)grid);
    return 0;
}

extern "C" void hostsim_stats(long long *out, int reset)
{
    out[0] = g_lc_stats.syms; out[1] = g_lc_stats.fresh; out[2] = g_lc_stats.search_fail;
    out[3] = g_lc_stats.interval_fail; out[4] = g_lc_stats.enc_interval_fail;
    if (reset) memset(&g_lc_stats, 0, sizeof(g_lc_stats));
}

// ---- parallel encoder: phase S restated on the host AND (sparse variant) run as the real kernel body through the
// emulator (lc_encoder_sort.cuh, eight warps per stream), the two compared array by array; phases A and B emulated
struct ParAArgs { LcCoderCfg cfg; const int *codes; int B; const uint32_t *skeys; const unsigned short *spos;
                  const int *first_bad; double *ivs; char *smem; unsigned short *glist; int *ngroups; unsigned int *task_counter;
                  double *tables; char *t2; };
static void para_body(void *p)
{
    ParAArgs *a = (ParAArgs *)p;
    lc_enc_phase_a_block(a->cfg, a->codes, a->B, a->skeys, a->spos, a->first_bad, a->ivs, a->smem);
}
struct ParBArgs { LcCoderCfg cfg; int B; const int *first_bad; const double *ivs; unsigned char *out_slots;
                  uint32_t slot_bytes; int *nbits, *status, *fault; };
static void para_tables_body(void *p)
{
    ParAArgs *a = (ParAArgs *)p;
    lcv_tables_block(a->cfg, a->tables, a->smem);
}
static void para_lanes_body(void *p)
{
    ParAArgs *a = (ParAArgs *)p;
    lc_enc_phase_a_sparse_block(a->cfg, a->codes, a->B, a->skeys, a->spos, a->first_bad, a->glist, a->ngroups, a->ivs,
                                a->task_counter, a->tables, a->t2, a->smem);
}
static void para_t2_body(void *p)
{
    ParAArgs *a = (ParAArgs *)p;
    lcv_t2_block(a->cfg, a->tables, a->t2, a->smem);
}
static void glist_body(void *p)
{   // phase S part: group list + first-visit intervals (on the GPU this is done by lc_enc_sort_kernel)
    ParAArgs *a = (ParAArgs *)p;
    const int lane = (int)(threadIdx.x & 31);
    const double u0 = 1.0 / (double)a->cfg.n;
    for (int b = (int)blockIdx.x; b < a->B; b += (int)gridDim.x) {
        const size_t o = (size_t)b * LC_PAR_MAX_SYMBOLS;
        const int fb = a->first_bad[b];
        const int total = fb < a->cfg.total ? fb : a->cfg.total;
        const int ng = lc_enc_group_list3_warp(lane, a->codes + (size_t)b * a->cfg.total, a->skeys + o, a->spos + o, total,
                                               a->cfg.n, u0, a->tables + 64, a->ivs + 2 * o,
                                               a->glist + (size_t)b * LC_PAR_MAX_GROUPS);
        if (lane == 0) a->ngroups[b] = ng;
    }
}
// the real phase S (lc_encoder_sort.cuh: radix sort from warp ballots, 8 warps per stream) under the emulator
struct Sort2Args { LcCoderCfg cfg; const int *codes; uint32_t *skeys; unsigned short *spos; int *first_bad;
                   unsigned short *glist; int *ngroups; double *ivs; const double *tables; char *smem; };
static void sort2_body(void *p)
{
    Sort2Args *a = (Sort2Args *)p;
    lcs2_block(a->cfg, LcCodes(a->codes), a->skeys, a->spos, a->first_bad, a->glist, a->ngroups, a->ivs, a->tables, a->smem);
}
static void parb_body(void *p)
{
    ParBArgs *a = (ParBArgs *)p;
    lc_enc_phase_b_block(a->cfg, a->B, a->first_bad, a->ivs, a->out_slots, a->slot_bytes, a->nbits, a->status,
                         a->fault);
}
struct ParB2Args { ParBArgs b; char *smem; };
static void parb1_body(void *p)
{
    ParB2Args *a = (ParB2Args *)p;
    alignas(16) static thread_local char ring[LC_B1_SMEM];
    lc_enc_phase_b1_block(a->b.cfg, a->b.B, a->b.first_bad, const_cast<double *>(a->b.ivs), a->b.nbits, ring);
}
static void parb2_body(void *p)
{
    ParB2Args *a = (ParB2Args *)p;
    lc_enc_phase_b2_block(a->b.cfg, a->b.B, a->b.first_bad, a->b.ivs, a->b.out_slots, a->b.slot_bytes, a->b.nbits,
                          a->b.status, a->b.fault, a->smem);
}

extern "C" int hostsim_encode_par(const int *codes, int B, int imgs, int R, int C, int n, double rate, int mode,
                                  unsigned char *out_slots, unsigned slot_bytes, int *nbits, int *status, int *fault,
                                  int grid, int nwarps)
{
    LcCoderCfg cfg;
    int rc = make_cfg(cfg, imgs, R, C, n, rate, mode, 1);
    if (rc) return rc;
    if (cfg.total > LC_PAR_MAX_SYMBOLS) return -22;
    const int total = cfg.total, RC = R * C;
    std::vector<uint32_t> skeys((size_t)B * LC_PAR_MAX_SYMBOLS, LC_PAR_KEY_PAD);
    std::vector<unsigned short> spos((size_t)B * LC_PAR_MAX_SYMBOLS, 0);
    std::vector<int> first_bad(B);
    for (int b = 0; b < B; b++) {
        const int *c = codes + (size_t)b * total;
        int fb = total;
        for (int p = 0; p < total; p++) if (c[p] < 0 || c[p] >= n) { fb = p; break; }
        first_bad[b] = fb;
        std::vector<uint32_t> keys(LC_PAR_MAX_SYMBOLS, LC_PAR_KEY_PAD);
        for (int p = 0; p < fb; p++) {
            const int q = p % RC, cc = q % C, rr = q / C;
            const int left = cc > 0 ? c[p - 1] : -1, up = rr > 0 ? c[p - C] : -1;
            keys[p] = (uint32_t)(left + 1) * (uint32_t)(n + 1) + (uint32_t)(up + 1);
        }
        std::vector<int> order(LC_PAR_MAX_SYMBOLS);
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return keys[x] < keys[y]; });
        for (int j = 0; j < LC_PAR_MAX_SYMBOLS; j++) {
            skeys[(size_t)b * LC_PAR_MAX_SYMBOLS + j] = keys[order[j]];
            spos[(size_t)b * LC_PAR_MAX_SYMBOLS + j] = (unsigned short)order[j];
        }
    }
    std::vector<double> ivs((size_t)B * LC_PAR_MAX_SYMBOLS * 2, -1.0);
    std::vector<char> smem(std::max((size_t)nwarps * n * 8, std::max((size_t)(n + 64) * 8, (size_t)2 * n * 8)) + 64);
    std::vector<unsigned short> glist((size_t)B * LC_PAR_MAX_GROUPS, 0);
    std::vector<int> ngroups(B, 0);
    unsigned int task_counter = 0;
    std::vector<double> tables(lcv_tables_bytes(n) / 8 + 8, -777.0);
    ParAArgs a{cfg, codes, B, skeys.data(), spos.data(), first_bad.data(), ivs.data(),
               (char *)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15), glist.data(), ngroups.data(), &task_counter,
               tables.data(), (char *)0};
    std::vector<char> t2;
    if (nwarps == 0 && n <= 64) { // small alphabets also exercise the records of the models after two visits
        t2.assign((size_t)n * n * 64, (char)0xCD);
        a.t2 = t2.data();
    }
    if (nwarps == 0) { // sparse variant: second visits from the table, one warp per context visited three times or more
        for (int b = 0; b < 3; b++) emu::run_warp(para_tables_body, &a, (unsigned)b, 3u);
        for (int b = 0; b < grid; b++) emu::run_warp(glist_body, &a, (unsigned)b, (unsigned)grid);
        {   // ... and the kernel that does all of phase S on the GPU must leave exactly the same arrays
            std::vector<uint32_t> skeys2((size_t)B * LC_PAR_MAX_SYMBOLS, 0u);
            std::vector<unsigned short> spos2((size_t)B * LC_PAR_MAX_SYMBOLS, 0), glist2((size_t)B * LC_PAR_MAX_GROUPS, 0);
            std::vector<int> first_bad2(B, -1), ngroups2(B, -1);
            std::vector<double> ivs2((size_t)B * LC_PAR_MAX_SYMBOLS * 2, -1.0);
            std::vector<char> smem_s(LCS2_SMEM + 64);
            Sort2Args sa{cfg, codes, skeys2.data(), spos2.data(), first_bad2.data(), glist2.data(), ngroups2.data(), ivs2.data(),
                         tables.data(), (char *)(((uintptr_t)smem_s.data() + 15) & ~(uintptr_t)15)};
            for (int b = 0; b < B; b++) emu::run_block(sort2_body, &sa, (unsigned)b, (unsigned)B, LCS2_WARPS);
            if (skeys2 != skeys || spos2 != spos || first_bad2 != first_bad || ngroups2 != ngroups) return -99;
            for (int b = 0; b < B; b++)
                for (int g = 0; g < ngroups[b]; g++)
                    if (glist2[(size_t)b * LC_PAR_MAX_GROUPS + g] != glist[(size_t)b * LC_PAR_MAX_GROUPS + g]) return -98;
            if (memcmp(ivs2.data(), ivs.data(), ivs.size() * sizeof(double)) != 0) return -97;
        }
        if (a.t2)
            for (int b = 0; b < 2; b++)
                for (int w = 0; w < 2; w++) emu::run_warp(para_t2_body, &a, (unsigned)b, 2u, (unsigned)w, 2u);
        for (int b = 0; b < grid; b++)
            for (int w = 0; w < 2; w++) emu::run_warp(para_lanes_body, &a, (unsigned)b, (unsigned)grid, (unsigned)w, 2u);
    } else
    for (int b = 0; b < grid; b++)
        for (int w = 0; w < nwarps; w++) emu::run_warp(para_body, &a, (unsigned)b, (unsigned)grid, (unsigned)w, (unsigned)nwarps);
    ParBArgs bb{cfg, B, first_bad.data(), ivs.data(), out_slots, slot_bytes, nbits, status, fault};
    if (mode == LC_MODE_REPAIRED && nwarps != 1) { // the two-kernel phase B, as the host launches it in this mode
        std::vector<char> smem2((size_t)slot_bytes + 4 + 24 * 4 + 64);
        ParB2Args b2{bb, (char *)(((uintptr_t)smem2.data() + 15) & ~(uintptr_t)15)};
        for (int b = 0; b < grid; b++) emu::run_warp(parb1_body, &b2, (unsigned)b, (unsigned)grid);
        for (int b = 0; b < grid; b++) emu::run_block(parb2_body, &b2, (unsigned)b, (unsigned)grid, LC_B2_THREADS / 32);
    } else
    for (int b = 0; b < grid; b++) emu::run_warp(parb_body, &bb, (unsigned)b, (unsigned)grid);
    return 0;
}
