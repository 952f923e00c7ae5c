// Phase A of the parallel encoder on SPARSE models, one warp per context group (replaces the dense-column
// lane-per-group variant).
//
// Phase A evolves, for every context that is visited more than once, the float64 model exactly as
// ContextModel.update_model does (cabac_compression.py:119-144) and writes the exact np.cumsum bounds
// (cum[s], cum[s+1]) (:346-347) of every visit.  The dense-column variant kept an [n][32] float64 tile per
// warp (64 KB at n = 256: three warps per SM, 19 % issue utilisation, 4.2 ms on the benchmark, latency bound
// by each lane's np.cumsum walk).  A lane-per-group variant on sparse records was tried and measured SLOWER
// (5.0 ms, 395 warp instructions per symbol): every lane's loops have their own trip counts, and the warp
// executes their union.  This version splits the work by visit number instead:
//   * first visits: closed form i/n (phase S);
//   * second visits (24 % of the benchmark's symbols): the model after one update depends only on the first
//     symbol, so the bounds are read from the per-launch table of exact cumsums (lcv_tables_block), also in
//     phase S -- no model is ever built for a context that is visited exactly twice;
//   * third and later visits: one warp per context keeps (u, sorted [(sym, val)]) in registers (lane j = entry
//     j), exactly like the decoder's updater warps (lcf_update / lcf_exact_at); the only shared memory is the
//     n-double image the pairwise sum is taken from, so 16+ warps fit on an SM and the kernel is issue bound.
//     A context with more than 32 distinct symbols continues on the dense image (lc_dense_update).
// Every float64 operation is the reference's operation on the same operands in the same order.
#pragma once
#include "lc_encoder_par.cuh"
#include "lc_decoder_v2.cuh"
#include "lc_decoder_small.cuh"

#define LCS_TASK_GROUPS 16
#define LCS_T2_MAX_N 256 // the table of models after two visits (lcv_t2_block) is built for alphabets up to this size

// Work list of phase A for one stream, and the intervals that need no model: glist[g] = sorted index of the first
// visit of the g-th context visited at least three times; first visits get i/n, second visits the table row of
// their first symbol.  (On the GPU phase S -- lc_enc_sort_kernel -- produces the same from the sorted keys it holds
// in registers; this warp version feeds the CPU emulator.)
__device__ __forceinline__ int lc_enc_group_list3_warp(int lane, LcCodes codes, const uint32_t *skeys,
                                                       const unsigned short *spos, int total, int n, double u0,
                                                       const double *cum1, double *ivs, unsigned short *glist)
{
    int ngroups = 0;
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        const bool valid = j < total;
        const uint32_t kj = valid ? skeys[j] : 0u;
        const bool head = valid && (j == 0 || skeys[j - 1] != kj);
        const bool second = valid && !head && (j == 1 || skeys[j - 2] != kj);
        if (head) {
            const int p = spos[j];
            const int s = codes[p];
            ivs[2 * p] = LC_DMUL((double)s, u0);
            ivs[2 * p + 1] = LC_DMUL((double)(s + 1), u0);
        }
        if (second) {
            const int p = spos[j];
            const int s = codes[p], s1 = codes[spos[j - 1]];
            const double *row = cum1 + (size_t)s1 * (n + 1);
            ivs[2 * p] = row[s];
            ivs[2 * p + 1] = row[s + 1];
        }
        const bool three = head && (j + 2 < total) && (skeys[j + 1] == kj) && (skeys[j + 2] == kj);
        const unsigned m = __ballot_sync(LC_FULL_MASK, three);
        if (three) glist[ngroups + __popc(m & ((1u << lane) - 1u))] = (unsigned short)j;
        ngroups += __popc(m);
    }
    return ngroups;
}

// Persistent warps pull tasks (stream, chunk of LCS_TASK_GROUPS contexts) from *task_counter.
// smem: one dense image (n doubles) per warp.  tables: u1tab[32] | ru1tab[32] | cum1[n][n+1].
__device__ __forceinline__ void lc_enc_phase_a_sparse_block(const LcCoderCfg &cfg, LcCodes codes_all, int B,
                                                            const uint32_t *skeys_all, const unsigned short *spos_all,
                                                            const int *first_bad, const unsigned short *glist_all,
                                                            const int *ngroups_all, double *ivs_all,
                                                            unsigned int *task_counter, const double *tables,
                                                            const char *t2, char *smem)
{
    const int warp = (int)(threadIdx.x >> 5);
    LcFast F;
    F.n = cfg.n; F.C = cfg.C; F.R = cfg.R; F.total = cfg.total; F.lane = (int)(threadIdx.x & 31);
    F.rate = cfg.rate; F.delta = cfg.delta; F.u0 = LC_DDIV(1.0, (double)cfg.n);
    F.delta_v = 0.0; F.tmargin = 0.0;
    F.P1 = LC_DADD(F.u0, LC_DMUL(F.rate, LC_DSUB(1.0, F.u0)));
    F.slot_cap = 0; F.slot_shift = 0; F.pool_bytes = 0; F.pool_top = 0;
    F.pw_len = cfg.pw_len; F.pw_steps = cfg.pw_steps; F.pw_chains = cfg.pw_chains;
    F.slots = (unsigned long long *)0; F.pool = (char *)0;
    F.dense = (double *)smem + (size_t)warp * cfg.n;
    F.u1tab = const_cast<double *>(tables); F.rows = (unsigned short *)0;
    F.k = 0; F.u = F.u0; F.my_sym = 0x7fffffff; F.my_val = 0.0;
    LcWarp W;
    lc_warp_init(W, cfg, smem, (char *)0);
    W.dense = F.dense;
    const unsigned chunks_per_stream = LC_PAR_MAX_GROUPS / LCS_TASK_GROUPS;
    const unsigned n_tasks = (unsigned)B * chunks_per_stream;
    for (;;) {
        unsigned task = 0;
        if (F.lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(LC_FULL_MASK, task, 0);
        if (task >= n_tasks) break;
        const int sidx = (int)(task / chunks_per_stream);
        const int g_lo = (int)(task % chunks_per_stream) * LCS_TASK_GROUPS;
        const int ngroups = ngroups_all[sidx];
        if (g_lo >= ngroups) continue;
        const int g_hi = (g_lo + LCS_TASK_GROUPS < ngroups) ? g_lo + LCS_TASK_GROUPS : ngroups;
        const size_t o = (size_t)sidx * LC_PAR_MAX_SYMBOLS;
        const LcCodes codes = codes_all + (size_t)sidx * cfg.total;
        const uint32_t *skeys = skeys_all + o;
        const unsigned short *spos = spos_all + o;
        const unsigned short *glist = glist_all + (size_t)sidx * LC_PAR_MAX_GROUPS;
        double *ivs = ivs_all + 2 * o;
        const int fb = first_bad[sidx];
        const int total = fb < cfg.total ? fb : cfg.total;
        if (cfg.n <= LCD_MAX_N) { // small alphabets: dense vector on the lanes (lc_decoder_small.cuh)
            for (int g = g_lo; g < g_hi; g++) {
                const int j0 = glist[g];
                switch (cfg.n) {
                case 2: lcd_enc_group<2>(codes, skeys, spos, j0, total, cfg.rate, F.dense, ivs, F.lane); break;
                case 4: lcd_enc_group<4>(codes, skeys, spos, j0, total, cfg.rate, F.dense, ivs, F.lane); break;
                case 8: lcd_enc_group<8>(codes, skeys, spos, j0, total, cfg.rate, F.dense, ivs, F.lane); break;
                default: lcd_enc_group<16>(codes, skeys, spos, j0, total, cfg.rate, F.dense, ivs, F.lane); break;
                }
                __syncwarp();
            }
            continue;
        }
        for (int g = g_lo; g < g_hi; g++) {
            const int j0 = glist[g];
            const uint32_t key = skeys[j0];
            // model after the first two visits: from the per-launch table, or computed (the first update is closed form)
            const int sa = codes[spos[j0]], sb = codes[spos[j0 + 1]];
            if (t2) lcv_record_load(F, t2 + ((size_t)sa * F.n + sb) * 64);
            else {
                lcf_state_first(F, sa);
                lcf_update(F, sb); // k <= 2: cannot overflow
            }
            bool dense_mode = false;
            for (int t = j0 + 2;; t++) {
                const int p = spos[t];
                const int s = codes[p];
                const bool last = (t + 1 >= total) || (skeys[t + 1] != key);
                if (!dense_mode) {
                    LcInterval iv;
                    lcf_exact_at(F, s, iv);
                    if (F.lane == 0) { ivs[2 * p] = iv.clo; ivs[2 * p + 1] = iv.chi; }
                } else {
                    const double T = lc_dense_prefix(W.dense, s);
                    if (F.lane == 0) { ivs[2 * p] = T; ivs[2 * p + 1] = LC_DADD(T, W.dense[s]); }
                }
                if (last) break; // the update after the last visit is never read
                if (!dense_mode && !lcf_update(F, s)) {
                    // more than 32 distinct symbols: go on with the dense image of the (unchanged) model
                    for (int i = F.lane; i < F.n; i += 32) F.dense[i] = F.u;
                    __syncwarp();
                    if (F.lane < F.k) F.dense[F.my_sym] = F.my_val;
                    __syncwarp();
                    dense_mode = true;
                    lc_dense_update(W, s);
                } else if (dense_mode) lc_dense_update(W, s);
            }
            __syncwarp();
        }
    }
}
