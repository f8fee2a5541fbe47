// j40-b200: pass groups -- HF coefficient entropy decode into token lists (device side, also compiled for the CPU
// kernel-logic tests).
//
// Replaces the reference's
//   j40__hf_coeffs    j40.h:6888-7004  (non-zero counts, coefficient contexts, the symbol loop)
//   j40__pass_group   j40.h:7007-7036  (preset selector; the modular extra channels are the host's / another kernel's)
// for every pass of a frame (j40.h:7847-7854: one section per pass and group; the reference adds the passes'
// coefficients up, j40.h:6989, which the tile back-end does from the per-pass token lists).
//
// Two kernels:
//   hf_prep_body   one warp per group, parallel: the group's varblocks in the raster order of their top-left cells
//                  (the order j40__hf_coeffs visits them in), each with everything the decoder needs resolved:
//                  cell, size class, the three block contexts (BlockContext() of j40.h:6927-6957). The serial
//                  decoder then reads one 16-byte record per varblock instead of scanning the block map and
//                  chasing four dependent tables.
//   HfLane         one *thread* per (pass, group): the lanes of a warp decode 32 sections side by side. The loop
//                  is warp-uniform (all lanes iterate until the last one is done) and has two phases separated by
//                  warp barriers, so that the lanes provably reconverge: (R) lanes whose current channel is
//                  exhausted start the next one -- next varblock record, predicted non-zero count, its symbol --
//                  and (C) every lane reads one coefficient symbol. Phase C is the tight path (~2/3 of all issue
//                  slots); phase R runs for the few lanes that need it.
// Tokens carry the scan index, not the coefficient position: the order table is applied by the tile back-end, where
// a table lookup costs nothing (one thread per token) -- here it would sit on the serial chain.
#pragma once
#include "j40b_vardct.h"

namespace j40b {

struct alignas(16) HfVb {
    uint32_t voff_cell;  // varblock index in the LF group (20 bits) | x8 << 20 | y8 << 25 (cell inside the group)
    uint8_t log_first;   // log2(size) - 6: the scan starts at 1 << log_first
    uint8_t order_idx;
    uint8_t log_w8;      // log2 of the width in cells
    uint8_t pad0;
    uint8_t bctx[3];     // block context per channel in decoding order (Y, X, B)
    uint8_t pad1;
    uint32_t pad2;
};

struct HfPrepWork {
    const DFrame *f;
    const uint8_t *arena;
    const DLfGroup *g;
    DGroup *grp;            // pass 0's record of the group (geometry; receives nvb)
    const uint32_t *lf_err;
};

// CoeffFreqContext / CoeffNumNonzeroContext of the format, pre-multiplied by 2 (j40.h:6935-6947)
J40B_HD J40B_INLINE int coeff_freq_ctx2(int k) { // k in [1, 64)
    return k < 16 ? 2 * (k - 1) : k < 32 ? 30 + 2 * ((k - 16) >> 1) : 46 + 2 * ((k - 32) >> 2);
}
J40B_HD J40B_INLINE int coeff_nnz_ctx2(int q) { // q in [0, 64)
    return q < 2 ? 0 : q == 2 ? 62 : q < 5 ? 124 : q < 9 ? 186 : q < 13 ? 246 : q < 21 ? 304 : q < 33 ? 360 : 412;
}

template <class Sync>
J40B_HD inline void hf_prep_body(const HfPrepWork &w, int lane, int nlanes, Sync sync) {
    if (*w.lf_err) return;
    const DFrame &f = *w.f;
    const DLfGroup &g = *w.g;
    DGroup &grp = *w.grp;
    const int gw8 = ceil_div(grp.gw, 8), gh8 = ceil_div(grp.gh, 8), ncell = gw8 * gh8;
    const int qf_count = f.nb_qf_thr + 1;
    const int lfidx_size = (f.nb_lf_thr[0] + 1) * (f.nb_lf_thr[1] + 1) * (f.nb_lf_thr[2] + 1);
    const int bctxc = 13 * qf_count * lfidx_size;
    const uint8_t *block_ctx_map = w.arena + f.block_ctx_map_off;
    const int w8 = g.width8;
    const int32_t *blocks = g.blocks + grp.gy8 * w8 + grp.gx8;
    const uint8_t *lfidx_map = g.lfidx + grp.gy8 * w8 + grp.gx8;
    HfVb *out = grp.vbs;
    int count = 0;
    for (int base = 0; base < ncell; base += nlanes) {
        const int cell = base + lane;
        int x8 = 0, y8 = 0;
        int32_t b = 0;
        if (cell < ncell) { y8 = cell / gw8; x8 = cell - y8 * gw8; b = blocks[y8 * w8 + x8]; }
        const bool top_left = (b >> 20) >= 2;
        int idx = count;
#if defined(__CUDA_ARCH__)
        const uint32_t m = __ballot_sync(0xffffffffu, top_left);
        idx += __popc(m & ((1u << lane) - 1));
        count += __popc(m);
#else
        count += top_left ? 1 : 0;
#endif
        if (top_left) {
            const int32_t voff = b & 0xfffff;
            const DctSelectInfo d = dct_select_info((b >> 20) - 2);
            const int bctx0 = (d.order_idx * qf_count + g.varblocks[voff].qfidx) * lfidx_size + lfidx_map[y8 * w8 + x8];
            HfVb r;
            r.voff_cell = (uint32_t) voff | (uint32_t) x8 << 20 | (uint32_t) y8 << 25;
            r.log_first = (uint8_t) (d.log_rows + d.log_columns - 6);
            r.order_idx = (uint8_t) d.order_idx;
            r.log_w8 = (uint8_t) (d.log_columns - 3);
            r.pad0 = r.pad1 = 0; r.pad2 = 0;
            for (int c = 0; c < 3; ++c) r.bctx[c] = block_ctx_map[bctx0 + bctxc * c];
            out[idx] = r;
        }
    }
    if (lane == 0) grp.nvb = count;
    sync();
}

// per-lane decoder of one pass-group section
struct HfShared { // what the lanes of a block share (shared memory on the device)
    const uint16_t *ctx_lut; // [128]: [q] = coeff_nnz_ctx2(q), [64 + k] = coeff_freq_ctx2(k); may be null
};

template <int MODE> // MODE 1: rANS without LZ77 (state seeded by the caller), 0: generic
struct HfLane {
    BitReader br;
    ErrSlot es;
    CodeCtx cc;
    CodeState cs;
    const uint64_t *ans_tables; // MODE 1: the clusters' alias tables, contiguous in cluster order
    int32_t las;                // MODE 1: log2 of the table length
    const uint16_t *ctx_lut;
    const HfVb *vbs;
    DToken *tokens;
    uint32_t *vb_tok;           // this pass's {first, count} pairs: [3][n8][2]
    uint8_t *colbuf;            // [3][32]: quantised non-zero count of the latest varblock covering each column
    int32_t n8, nvb, vb_i, ctxoff, nb_block_ctx;
    uint32_t tok, tok_end, first_tok;
    uint32_t rec_voff_cell, rec_bctx; // current record: packed cell, the three block contexts
    int32_t log_first, log_w8;
    int32_t c_yxb, nz, i, prev, cctx, c;
    bool done;

    J40B_HD J40B_INLINE int32_t symbol(int32_t ctx) {
        if (MODE == 1) {
            const uint32_t ci = cc.cluster_map[ctx];
            const uint64_t e = ans_tables[((size_t) ci << las) + ((cs.ans_state & 0xfff) >> cc.log_bucket)];
            const HybridCfg cfg = cc.clusters[ci].cfg;
            const int32_t token = ans_symbol_entry(br, cs.ans_state, cc.log_bucket, e);
            return hybrid_int(br, es, token, cfg);
        }
        return code(br, es, cc, cs, ctx, 0);
    }

    J40B_HD J40B_INLINE void fail(uint32_t code_) { es.set(br, code_); done = true; }

    // phase R, one round: the next channel (and varblock) of the section; reads its non-zero count.
    // Leaves nz > 0 if coefficients follow.
    J40B_HD J40B_INLINE void next_channel() {
        if (++c_yxb == 3) {
            c_yxb = 0;
            if (++vb_i >= nvb) { done = true; return; }
            const HfVb r = vbs[vb_i];
            rec_voff_cell = r.voff_cell;
            rec_bctx = (uint32_t) r.bctx[0] | (uint32_t) r.bctx[1] << 8 | (uint32_t) r.bctx[2] << 16;
            log_first = r.log_first;
            log_w8 = r.log_w8;
        }
        c = c_yxb == 0 ? 1 : c_yxb == 1 ? 0 : 2;
        const int x8 = (int) (rec_voff_cell >> 20) & 31, y8 = (int) (rec_voff_cell >> 25) & 31;
        const int bctx = (int) (rec_bctx >> (8 * c_yxb)) & 0xff;
        uint8_t *col = colbuf + c * 32;
        // j40.h:6959-6961; column x8 was last covered by the block above, column x8 - 1 by the block to the left
        const int top = col[x8], left = col[x8 > 0 ? x8 - 1 : 0];
        const int pred = x8 > 0 ? (y8 > 0 ? (left + top + 1) >> 1 : left) : (y8 > 0 ? top : 32);
        const int ctx = ctxoff + bctx + (pred < 8 ? pred : 4 + pred / 2) * nb_block_ctx;
        cctx = ctxoff + 458 * bctx + 37 * nb_block_ctx;
        const int32_t v = symbol(ctx);
        if (es.err) { done = true; return; }
        if (!((uint32_t) v <= (63u << log_first))) { fail(E_COEF); return; }
        nz = v;
        const uint8_t qnz = (uint8_t) ((nz + (1 << log_first) - 1) >> log_first);
        for (int j = 0; j < (1 << log_w8); ++j) col[x8 + j] = qnz;
        prev = nz <= (1 << (log_first + 2)); // size / 16 (j40.h:6979)
        i = 1 << log_first;
        first_tok = tok;
    }

    // phase C: one coefficient symbol (j40.h:6981-6991)
    J40B_HD J40B_INLINE void coefficient() {
        const int q = (nz + (1 << log_first) - 1) >> log_first, k = i >> log_first;
        const int ctx = cctx + prev + (ctx_lut ? (int) ctx_lut[q] + (int) ctx_lut[64 + k] : coeff_nnz_ctx2(q) + coeff_freq_ctx2(k));
        const int32_t v = symbol(ctx);
        if (es.err) { done = true; return; }
        if (v) {
            const int32_t val = unpack_signed(v);
            const bool wide = token_needs_wide(val);
            if (tok + (wide ? 3u : 1u) > tok_end) { es.set_raw(E_TOKV); done = true; return; }
            tokens[tok++] = DToken::make((uint32_t) i, wide ? TOKEN_WIDE : val);
            if (wide) {
                tokens[tok++] = DToken::make(0, val & 0xffff);
                tokens[tok++] = DToken::make(0, (int32_t) ((uint32_t) val >> 16));
            }
        }
        prev = v != 0;
        nz -= prev;
        ++i;
        if (nz == 0) {
            uint32_t *slot = vb_tok + ((size_t) c * n8 + (rec_voff_cell & 0xfffff)) * 2;
            slot[0] = first_tok;
            slot[1] = tok - first_tok;
        } else if (i >= (64 << log_first)) fail(E_COEF);
    }
};

} // namespace j40b
