// j40-b200: kernel bodies. Each *_body function is what one CUDA warp or thread block executes for one
// work item; the __global__ wrappers live in j40b_cuda.cu. The CPU kernel-logic tests (tests/hostemu) call
// the same bodies with a single "thread" (lane 0 of 1 / tid 0 of 1).
//
// VarDCT frame, per step (all images of the batch in each launch):
//   lf_decode1   warp per LF group   serial: LF image (3 modular channels)              [latency-bound]
//   lf_post      block per LF group  parallel: inverse RCT, dequantise, LF indices, adaptive smoothing
//   lf_decode2   warp per LF group   serial: HF metadata (4 modular channels) + varblock placement
//   lf_llf       block per LF group  parallel: LLF coefficients (forward DCT of LF patches)
//   hf_group     thread per group    serial per lane: HF coefficient entropy decode -> token lists
//   back_tile    block per 64x64 tile  parallel: dequant, CfL, IDCT, XYB->sRGB, RGBA8 store
//   back_generic persistent blocks   varblocks larger than 64x64 / straddling tiles
// Modular frame: modular (warp per group) -> render (thread per pixel).
#pragma once
#include "j40b_hf.h"
#include "j40b_modlane.h"

namespace j40b {

// a modular header with a tree of its own inside an LF-group section (j40.h:3827-3835), as the host read it
struct LfLocal {
    int32_t present;
    uint32_t host_err;      // the header / tree / code spec is broken: reported when the decoder gets there
    uint32_t tree_off, spec_off;
    int32_t uses_wp;
    uint64_t start_bit;     // first bit of the channel data
    ModImage hdr;           // weighted-predictor parameters, transforms, dist_mult (channel geometry is the decoder's)
};

// one work item per LF group / group; all pointers are device pointers
struct LfWork {
    const DFrame *f;
    const uint8_t *arena;   // per-image table arena
    const uint8_t *cs;      // linearised codestream
    DLfGroup *g;
    uint32_t *err;          // this section's error word
    float *llf_scratch;     // [2048] for LF patches larger than 8x8 cells, or null
    ModLaneScratch *lane_scratch; // per-stream record of the lane-per-stream decoder
    int16_t *ring;          // this lane's slot of its warp's row ring (LaneEnv), or null
    int32_t *wring;
    int32_t ring_w, ring_lstride;
    LfLocal local[2];       // LF image / HF metadata: local trees (filled in by the host after an E_LTRE round)
};

struct HfWork { // one per (pass, group)
    const DFrame *f;
    const uint8_t *arena;
    const uint8_t *cs;
    DLfGroup *g;
    DGroup *grp;            // this pass's record of the group
    const DGroup *geo;      // pass 0's record (varblock list)
    DToken *tokens;         // image token array
    const uint32_t *lf_err; // error word of the LF group this group depends on
    uint32_t *err;
};

struct BackWork {
    const DFrame *f;
    const uint8_t *arena;
    const DLfGroup *g;
    const DGroup *grp;
    const DToken *tokens;
    const uint32_t *lf_err, *hf_err; // hf_err[pass * hf_err_stride]: the group's pass sections
    int32_t hf_err_stride;
    uint8_t *rgba;
    int32_t rgba_stride;
    float *big_scratch;     // [4 * 65536] for varblocks larger than 64x64, or null
};

// diagnostics (j40b_batch_debug_dump): the dense coefficient planes of an LF group in the reference's layout
// (gg->coeffs[c][coeffoff + pos], j40.h:6614), rebuilt from the token lists
struct DumpWork {
    const DFrame *f;
    const DLfGroup *g;
    const DToken *tokens;
    float *out;       // [3][width8 * height8 * 64]
    int32_t dequant;  // 0: the decoded integers (j40.h:6989); 1: after j40__dequant_hf (j40.h:7053-7097)
};

struct ModWork { // one modular sub-bitstream: a pass group of a modular frame, or the global channels
    const DFrame *f;
    const uint8_t *arena;
    const uint8_t *cs;
    uint32_t sec_off, sec_size;
    uint64_t sec_start_bit;
    int32_t sidx;
    int32_t header_parsed;  // 1: `m` was filled by the host (global image, or a header with a local tree), 0: parse it here
    int32_t is_global;      // the frame's global image (transforms are applied by the render step)
    int32_t check_end;      // single-section frame: the sub-bitstream must end the section exactly
    uint32_t tree_off, spec_off; // MA tree and code spec of this sub-bitstream in the arena (the global ones or local ones)
    int32_t tree_uses_wp;
    uint32_t preset_err;    // non-zero: the host already found this sub-bitstream's header broken
    ModImage m;             // channel views into the frame planes
    int32_t *wp_scratch;    // [2 * width * 5] or null
    int32_t *lz_window;     // or null
    uint32_t lz_mask;
    uint32_t *err;
    ModLaneScratch *lane_scratch; // per-stream record of the lane-per-stream decoder
    int16_t *ring;          // this lane's slot of its warp's row ring (LaneEnv), or null
    int32_t *wring;
    int32_t ring_w, ring_lstride;
    // extra channels of a VarDCT pass group (j40.h:7024-7033): the sub-bitstream starts where the coefficient decoder of
    // the same section stopped, and only if that one succeeded (both report through the section's error word)
    const uint64_t *start_bit_src; // or null: sec_start_bit
    const uint32_t *skip_if;       // or null
};

struct RenderWork { // modular frames: inverse global transforms + interleave to RGBA8
    const DFrame *f;
    int16_t *plane[MOD_MAX_CH]; // full-frame planes, stride = width
    // a delta palette (the last transform of the list) is undone ahead of the render step by palette_delta_body:
    int16_t *dplane[MOD_MAX_CH]; // its num_c restored channels (null: the frame has none)
    int32_t *dwp[MOD_MAX_CH];    // weighted-predictor error rows [2][width][5] per restored channel (d_pred == 6)
    const uint32_t *any_err;    // non-zero: skip
    uint8_t *rgba;
    int32_t rgba_stride;
};

} // namespace j40b
#include "j40b_backtile.h" // needs BackWork
namespace j40b {

enum { PTREE_CAP = 192 };

// per-warp scratch of the serial decoders (shared memory on the device)
// (the pruned tree of the current channel lives in the stream's global scratch record, ModLaneScratch::ptree: it is
// only read when the channel is set up, or walked through L1 by the fallback decoder)
struct alignas(16) WarpScratch {
    ModImage m;
    int32_t info[8];
};

// cooperative staging of a code spec's blob into `dst` (cap bytes); returns true if it fits
J40B_HD inline bool stage_spec_blob(const uint8_t *arena, uint32_t spec_off, uint8_t *dst, uint32_t cap, int tid, int nth) {
    if (!spec_off) return false;
    const DCodeSpec *spec = (const DCodeSpec *) (arena + spec_off);
    uint32_t lo = spec->blob_lo, hi = spec->blob_hi;
    if (!dst || hi - lo > cap) return false;
    const uint32_t *src = (const uint32_t *) (arena + lo); // lo is 16-byte aligned, hi 4-byte aligned
    uint32_t *d32 = (uint32_t *) dst;
    for (uint32_t i = (uint32_t) tid; i < (hi - lo + 3) / 4; i += (uint32_t) nth) d32[i] = src[i];
    return true;
}

J40B_HD inline void fill_div24(int32_t *div24, int tid, int nth) {
    for (int i = tid; i < 64; i += nth) div24[i] = (int32_t) (0x1000000u / (uint32_t) (i + 1));
}

J40B_HD inline void init_code_ctx(CodeCtx &cc, const uint8_t *arena, uint32_t spec_off, const uint8_t *spec_copy, const uint8_t *copy_arena) {
    if (spec_copy && copy_arena == arena) cc.init_from_copy(spec_copy, ((const DCodeSpec *) (arena + spec_off))->blob_lo, spec_off);
    else cc.init(arena, spec_off);
}

// the four channels of the HF metadata image (j40.h:6766-6772)
J40B_HD J40B_INLINE void hf_meta_channels(const DLfGroup &g, int32_t nvb, ModImage &m) {
    m.num_channels = 4;
    m.ch[0].px = g.xfromy; m.ch[0].w = g.width64; m.ch[0].h = g.height64; m.ch[0].stride = g.width64;
    m.ch[1].px = g.bfromy; m.ch[1].w = g.width64; m.ch[1].h = g.height64; m.ch[1].stride = g.width64;
    m.ch[2].px = g.blockinfo; m.ch[2].w = nvb; m.ch[2].h = 2; m.ch[2].stride = nvb;
    m.ch[3].px = g.sharpness; m.ch[3].w = g.width8; m.ch[3].h = g.height8; m.ch[3].stride = g.width8;
    for (int c = 0; c < 4; ++c) m.ch[c].hshift = m.ch[c].vshift = 0;
}

// The modular header of stage `stage` (0 LF image, 1 HF metadata) of an LF group; `m` holds the channel geometry.
// A header naming a tree of its own cannot be finished here (trees and code specs are parsed by the host): the
// first time round the decoder records where the header starts and reports the internal code E_LTRE; the host reads
// header, tree and code spec there (Batch::resolve_local_trees) and the batch is decoded again with w.local[stage].
// Returns false if the stage cannot go on (es.err is set).
J40B_HD inline bool lf_stage_header(BitReader &br, ErrSlot &es, const LfWork &w, int stage, ModImage &m, const DTreeNode *&tree,
                                    uint32_t &spec_off, int32_t &uses_wp, bool writer) {
    const DFrame &f = *w.f;
    DLfGroup &g = *w.g;
    const LfLocal &lo = w.local[stage];
    tree = (const DTreeNode *) (w.arena + f.global_tree_off);
    spec_off = f.global_spec_off;
    uses_wp = f.global_tree_uses_wp;
    if (lo.present) {
        if (lo.host_err) { es.set_raw(lo.host_err); return false; }
        m.wp = lo.hdr.wp;
        m.nb_transforms = lo.hdr.nb_transforms;
        for (int t = 0; t < MOD_MAX_TRANSFORMS; ++t) m.tr[t] = lo.hdr.tr[t];
        m.dist_mult = lo.hdr.dist_mult;
        m.nb_meta_channels = 0;
        br.init(w.cs + g.sec_off, g.sec_size, lo.start_bit);
        tree = (const DTreeNode *) (w.arena + lo.tree_off);
        spec_off = lo.spec_off;
        uses_wp = lo.uses_wp;
        return true;
    }
    const uint64_t at = br.bits_consumed();
    int local = 0;
    modular_header(br, es, f.have_global_tree != 0, m, &local);
    if (es.err) return false;
    if (local) {
        if (writer) { g.ltree_bit = at; g.ltree_stage = stage; }
        es.set_raw(E_LTRE);
        return false;
    }
    return true;
}

// inverse transforms of the HF metadata image (an RCT over xfromy/bfromy/blockinfo is possible when their sizes coincide;
// none can involve the sharpness channel: its height never equals the block-info channel's 2 rows together with the
// colour-correlation maps' height) and the varblock placement (j40.h:6585-6720)
template <class Sync>
J40B_HD inline void lf_place_tail(const LfWork &w, const ModImage &m, ErrSlot &es, const BitReader &br, const ModSmem &ms,
                                  int lane, int nlanes, Sync sync) {
    const DFrame &f = *w.f;
    DLfGroup &g = *w.g;
    for (int t = m.nb_transforms - 1; t >= 0; --t) {
        ModImage one = m;
        one.nb_transforms = 1;
        one.tr[0] = m.tr[t];
        inverse_transforms(one, lane, nlanes);
        sync();
    }
    // the weighted predictor's shared-memory row and the reference-property rows behind it are free now: 8 words per
    // row of cells serve as the occupancy bitmap of the placement (256 * 8 words; error row 256 * 5, property rows 256 * 2,
    // sample rows 256 * 1.5 words follow each other in the warp's slice, see carve_warp_slice)
    if (ms.rows && ms.cap >= 256) place_varblocks_warp(f, g, es, br, (uint32_t *) ms.wp, lane, nlanes, sync);
    else if (lane == 0) place_varblocks(f, g, es, br);
}

// ---------------------------------------------------------------------------------------------
// The two serial stages of an LF group, one *channel* and one decoder class per kernel (modular_channel_prep / _run in
// j40b_modular.h): stage 0 = the LF image (3 channels, j40.h:6739-6757), stage 1 = the HF metadata image (4 channels)
// followed by the varblock placement (j40.h:6766-6777, 6585-6720). For every channel the executor launches the four
// class kernels one after the other; each LF group is decoded by the one whose class its channel has and left alone by
// the others. Between kernels the decoder state travels through the LF group record (bit position, rANS state, LZ77
// counters) and the modular image header through the group's scratch record. Why: a kernel holding every variant
// needs 168 / 236 registers per thread, the class kernels 64 ... 128, and the register file the long-lived serial
// warps occupy is what keeps tile and coefficient blocks of other batches off the SMs ("what bounds a step").
// `split` (multi-section frames): the varblock placement does not wait for the sharpness channel -- 65 536 samples per
// big LF group that j40 never uses (j40.h:6769-6772) -- but follows the block-info channel; the executor decodes the
// sharpness channel on a side stream next to the LLF / coefficient / tile stages. Errors keep the reference's order:
// whatever the sharpness decode raises (running out of data included) precedes a placement error.
template <int K, class Sync>
J40B_HD inline void lf_chan_body(const LfWork &w, int stage, int c, WarpScratch &ws, const ModSmem &ms, const int32_t *div24,
                                 int lane, int nlanes, Sync sync, bool split = false) {
    const DFrame &f = *w.f;
    DLfGroup &g = *w.g;
    const bool late = split && stage == 1 && c == 3; // the sharpness channel behind the placement
    if (late ? !g.placed : *w.err != 0) return;
    const int n8 = g.width8 * g.height8;
    const int nch = stage == 0 ? 3 : 4;
    const int32_t sidx = stage == 0 ? 1 + g.idx : 1 + 2 * f.num_lf_groups + g.idx;
    BitReader br;
    ErrSlot es;
    CodeCtx cc;
    CodeState cs;
    es.err = 0;
    cs.init(g.lz_window, (1u << 18) - 1);
    ModImage &m = ws.m; // (every lane runs on identical state and writes identical values to the shared image header)
    const DTreeNode *tree;
    uint32_t spec_off;
    int32_t uses_wp;
    bool go = true;
    if (c == 0) {
        // the stage's prologue; every class kernel of channel 0 runs it (same bits, same values)
        br.init(w.cs + g.sec_off, g.sec_size, stage == 0 ? g.sec_start_bit : g.mid_bit);
        if (stage == 0) {
            const int32_t extra_prec = (int32_t) br.u(2);
            if (lane == 0) g.extra_prec = extra_prec;
            m.num_channels = 3;
            for (int k = 0; k < 3; ++k) {
                m.ch[k].px = g.lfq + (size_t) k * n8;
                m.ch[k].stride = g.width8; m.ch[k].w = g.width8; m.ch[k].h = g.height8;
                m.ch[k].hshift = m.ch[k].vshift = 0;
            }
        } else {
            const int32_t nvb = (int32_t) br.u(ceil_lg32((uint32_t) n8)) + 1;
            if (lane == 0) { g.nb_varblocks = nvb; g.placed = 0; }
            hf_meta_channels(g, nvb, m);
        }
        go = lf_stage_header(br, es, w, stage, m, tree, spec_off, uses_wp, lane == 0);
        sync();
        if (go && lane == 0) w.lane_scratch->m = m;
    } else {
        m = w.lane_scratch->m;
        const LfLocal &lo = w.local[stage];
        const bool local = lo.present && !lo.host_err;
        tree = (const DTreeNode *) (w.arena + (local ? lo.tree_off : f.global_tree_off));
        spec_off = local ? lo.spec_off : f.global_spec_off;
        uses_wp = local ? lo.uses_wp : f.global_tree_uses_wp;
        br.init(w.cs + g.sec_off, g.sec_size, g.chan_bit);
        cs.ans_state = g.chan_ans;
        cs.num_to_copy = g.chan_copy[0]; cs.copy_pos = g.chan_copy[1]; cs.num_decoded = g.chan_copy[2];
    }
    cc.init(w.arena, spec_off);
    sync();
    if (go && !es.err) {
        const int cls = modular_channel_prep(cc, tree, uses_wp != 0, g.wp_scratch, w.lane_scratch->ptree, PTREE_CAP, ms, m, c, sidx, lane, sync);
        if (cls != K && !(cls == MC_NONE && K == MC_REST)) return; // another class kernel's channel
        modular_channel_run<K>(cls, br, es, cc, cs, tree, g.wp_scratch, div24, w.lane_scratch->ptree, ms, m, c, sidx, lane, nlanes, sync);
    }
    if (!es.err && c + 1 < nch) {
        if (lane == 0) {
            g.chan_bit = br.bits_consumed();
            g.chan_ans = cs.ans_state;
            g.chan_copy[0] = cs.num_to_copy; g.chan_copy[1] = cs.copy_pos; g.chan_copy[2] = cs.num_decoded;
        }
        if (!(split && stage == 1 && c == 2)) return;
        // split mode: placement right behind the block-info channel
        sync();
        lf_place_tail(w, m, es, br, ms, lane, nlanes, sync);
        if (lane == 0) {
            g.placed = 1;
            if (es.err) *w.err = es.err;
        }
        return;
    }
    // ---- the stage's epilogue (last channel, or an error)
    if (!es.err) finish_code(br, es, cc, cs);
    if (stage == 0) {
        if (lane == 0) {
            g.mid_bit = br.bits_consumed();
            g.nb_tr1 = imin(m.nb_transforms, MOD_MAX_TRANSFORMS);
            for (int t = 0; t < g.nb_tr1; ++t) g.tr1[t] = m.tr[t];
            if (es.err) *w.err = es.err;
        }
        return;
    }
    if (es.err) { if (lane == 0) *w.err = es.err; return; }
    sync();
    if (!late) lf_place_tail(w, m, es, br, ms, lane, nlanes, sync);
    if (lane == 0) {
        // multi-section frames: the reference drops pad0/excs found at a section's end (they are raised
        // on the per-section state and never copied back, j40.h:7791-7798); running short is still an error
        if (late) { if (br.overrun()) *w.err = E_SHRT; }   // (precedes whatever the placement found)
        else if (!es.err && br.overrun()) es.set_raw(E_SHRT);
        if (!es.err) g.end_bit = br.bits_consumed();
        if (es.err) *w.err = es.err;
    }
}

// LF group, stage 2 (one block): inverse transforms, dequantisation, LF indices, smoothing (j40.h:6544-6583)
template <class Sync>
J40B_HD inline void lf_post_body(const LfWork &w, int tid, int nth, Sync sync) {
    if (*w.err) return;
    const DFrame &f = *w.f;
    DLfGroup &g = *w.g;
    const int n8 = g.width8 * g.height8;
    for (int i = tid; i < n8; i += nth) g.blocks[i] = 0;
    for (int t = g.nb_tr1 - 1; t >= 0; --t) {
        ModImage one;
        one.num_channels = 3;
        for (int c = 0; c < 3; ++c) {
            one.ch[c].px = g.lfq + (size_t) c * n8;
            one.ch[c].stride = g.width8; one.ch[c].w = g.width8; one.ch[c].h = g.height8;
            one.ch[c].hshift = one.ch[c].vshift = 0;
        }
        one.nb_transforms = 1;
        one.tr[0] = g.tr1[t];
        inverse_transforms(one, tid, nth);
        sync();
    }
    lf_dequant(f, g, g.extra_prec, tid, nth);
    sync();
    if (!f.skip_adapt_lf_smooth) lf_smooth(f, g, tid, nth);
}

// LF group, stage 4 (one block): LLF coefficients of every varblock (j40.h:6669-6683)
template <class Sync>
J40B_HD inline void lf_llf_body(const LfWork &w, int tid, int nth, Sync sync) {
    if (*w.err) return;
    DLfGroup &g = *w.g;
    for (int v = tid; v < g.nb_varblocks; v += nth) {
        const DVarblock &vb = g.varblocks[v];
        DctSelectInfo d = dct_select_info(vb.dctsel);
        if (d.log_rows + d.log_columns - 6 <= 6) {
            float scratch[64];
            llf_from_lf(g, vb, scratch, 0, 1, NoSync());
        }
    }
    if (tid == 0 && w.llf_scratch && g.has_big) {
        for (int v = 0; v < g.nb_varblocks; ++v) {
            const DVarblock &vb = g.varblocks[v];
            DctSelectInfo d = dct_select_info(vb.dctsel);
            if (d.log_rows + d.log_columns - 6 > 6) llf_from_lf(g, vb, w.llf_scratch, 0, 1, NoSync());
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One (pass, group) section per *thread*: the lanes of a warp decode 32 sections side by side in a warp-uniform
// two-phase loop (see j40b_hf.h). `spec_copy` is an optional shared-memory copy of the coefficient code spec of
// (image `copy_arena`, pass `copy_pass`); lanes of other images / passes read their tables from global memory.
// `active`: this lane has a work item. AnyFn(bool) -> bool: warp-wide "any" (identity with a single lane).
template <int MODE>
J40B_HD J40B_INLINE void hf_lane_init(HfLane<MODE> &L, const HfWork &w, const uint8_t *spec_copy, const uint8_t *copy_arena, int copy_pass,
                                     const uint16_t *ctx_lut) {
    const DFrame &f = *w.f;
    DLfGroup &g = *w.g;
    DGroup &grp = *w.grp;
    L.es.err = 0;
    L.done = false;
    const uint64_t start_bit = grp.sec_start_bit == ~0ull ? g.end_bit : grp.sec_start_bit;
    L.br.init(w.cs + grp.sec_off, grp.sec_size, start_bit);
    init_code_ctx(L.cc, w.arena, f.coeff_spec_off[grp.pass], grp.pass == copy_pass ? spec_copy : nullptr, copy_arena);
    L.cs.init(grp.lz_window, (1u << 18) - 1);
    L.ans_tables = (const uint64_t *) (L.cc.arena + L.cc.spec->ans_tables_off);
    L.las = L.cc.spec->log_alpha_size;
    L.ctx_lut = ctx_lut;
    L.vbs = w.geo->vbs;
    L.nvb = w.geo->nvb;
    L.tokens = w.tokens;
    L.n8 = g.width8 * g.height8;
    L.vb_tok = g.vb_tok + (size_t) grp.pass * 6 * (size_t) L.n8;
    L.colbuf = grp.nonzeros;
    L.nb_block_ctx = f.nb_block_ctx;
    L.tok = grp.tok_first;
    L.tok_end = grp.tok_first + grp.tok_cap;
    L.first_tok = L.tok;
    L.vb_i = -1; L.c_yxb = 2; L.nz = 0; L.i = 0; L.prev = 0; L.cctx = 0; L.c = 0;
    L.rec_voff_cell = L.rec_bctx = 0; L.log_first = L.log_w8 = 0;
    const int32_t preset = (int32_t) L.br.u(ceil_lg32((uint32_t) f.num_hf_presets));
    L.ctxoff = 495 * f.nb_block_ctx * preset;
    // contexts beyond the code spec: the reference reads out of bounds here; treat as corrupt
    if (preset >= f.num_hf_presets) L.fail(E_COEF);
    // MODE 1: every group reads at least one symbol (its first varblock's non-zero count): seed the rANS state now,
    // so that the symbol loop need not test for it
    else if (MODE == 1) ans_seed(L.br, L.cs.ans_state);
}

template <int MODE>
J40B_HD J40B_INLINE void hf_lane_finish(HfLane<MODE> &L, const HfWork &w) {
    DGroup &grp = *w.grp;
    grp.tok_used = L.tok - grp.tok_first;
    if (!L.es.err) finish_code(L.br, L.es, L.cc, L.cs);
    grp.end_bit = L.br.bits_consumed(); // where the section's extra-channel sub-bitstream starts, if it has one
    if (!L.es.err) {
        if (grp.sec_start_bit == ~0ull) { uint32_t e = L.br.finish(); if (e) L.es.set_raw(e); } // single-section frame: real check
        else if (L.br.overrun()) L.es.set_raw(E_SHRT); // see lf_decode2_body
    }
    if (L.es.err) *w.err = L.es.err;
}

template <int MODE, class AnyFn, class Sync>
J40B_HD inline void hf_lanes_run(const HfWork *w, bool active, const uint8_t *spec_copy, const uint8_t *copy_arena, int copy_pass,
                                 const uint16_t *ctx_lut, AnyFn any, Sync sync) {
    HfLane<MODE> L;
    L.done = true;
    if (active) hf_lane_init(L, *w, spec_copy, copy_arena, copy_pass, ctx_lut);
    for (uint32_t iter = 0;; ++iter) {
        if (!any(!L.done)) break;
        sync();
        // phase R: lanes whose channel is exhausted (or that have not started) open the next one. With 32 lanes some
        // lane needs it in every other iteration, and it costs more than a coefficient step, so it is batched: every
        // fourth iteration, or at once when no lane has a coefficient to read; a lane waits 1.5 iterations on average
        // per channel (2 % more iterations) and phase R runs half as often.
        const bool want = !L.done && L.nz == 0;
        if ((iter & 3) == 0 || !any(!L.done && L.nz != 0)) {
            if (want) { do L.next_channel(); while (!L.done && L.nz == 0); }
            sync();
        }
        // phase C: one coefficient symbol per lane
        if (!L.done && L.nz != 0) L.coefficient();
    }
    if (active) hf_lane_finish(L, *w);
}

// whether the section of work item `w` can take the MODE 1 path (the lanes of a warp must agree: the caller votes)
J40B_HD J40B_INLINE bool hf_is_plain_ans(const HfWork &w) {
    const DCodeSpec *spec = (const DCodeSpec *) (w.arena + w.f->coeff_spec_off[w.grp->pass]);
    return !spec->use_prefix_code && !spec->lz77_enabled;
}

// ---------------------------------------------------------------------------------------------
// generic path: varblocks the tile kernel leaves out (larger than 64x64 or straddling tiles), one at a time
// with all threads and buffers in w.big_scratch (4 * 65536 floats of global memory)
template <class Sync>
J40B_HD inline void back_generic_body(const BackWork &w, int tid, int nth, Sync sync) {
    if (*w.lf_err) return;
    for (int p = 0; p < w.f->num_passes; ++p) if (w.hf_err[(size_t) p * w.hf_err_stride]) return;
    if (!w.g->has_big) return;
    const DFrame &f = *w.f;
    const DLfGroup &g = *w.g;
    const DGroup &grp = *w.grp;
    const int gw8 = ceil_div(grp.gw, 8), gh8 = ceil_div(grp.gh, 8);
    for (int y8 = 0; y8 < gh8; ++y8) for (int x8 = 0; x8 < gw8; ++x8) {
        int32_t voff = g.blocks[(y8 + grp.gy8) * g.width8 + x8 + grp.gx8];
        if ((voff >> 20) < 2) continue;
        voff &= 0xfffff;
        const DVarblock &vb = g.varblocks[voff];
        if (!(vb.pad & 1)) continue;
        float *buf = w.big_scratch;
        varblock_to_pixels(f, w.arena, g, vb, voff, w.tokens, buf, buf + 65536, buf + 2 * 65536, buf + 3 * 65536,
                           w.rgba, w.rgba_stride, tid, nth, sync);
    }
}

// ---------------------------------------------------------------------------------------------
// one warp per modular sub-bitstream
template <class Sync>
J40B_HD inline void modular_body(ModWork &w, WarpScratch &ws, const ModSmem &ms, const int32_t *div24,
                                 const uint8_t *spec_copy, const uint8_t *copy_arena, int lane, int nlanes, Sync sync) {
    const DFrame &f = *w.f;
    if (w.preset_err) { if (lane == 0) *w.err = w.preset_err; return; }
    if (w.skip_if && *w.skip_if) return;
    BitReader br;
    ErrSlot es;
    CodeCtx cc;
    CodeState cs;
    es.err = 0;
    br.init(w.cs + w.sec_off, w.sec_size, w.start_bit_src ? *w.start_bit_src : w.sec_start_bit);
    init_code_ctx(cc, w.arena, w.spec_off, spec_copy, copy_arena);
    cs.init(w.lz_window, w.lz_mask);
    const DTreeNode *tree = (const DTreeNode *) (w.arena + w.tree_off);
    ModImage &m = ws.m;
    m = w.m;
    if (!w.header_parsed) modular_header(br, es, f.have_global_tree != 0, m);
    sync();
    for (int c = 0; c < m.num_channels && !es.err; ++c) {
        modular_channel_warp(br, es, cc, cs, tree, w.tree_uses_wp != 0, w.wp_scratch, div24, w.lane_scratch->ptree, PTREE_CAP, ms, m, c, w.sidx, lane, nlanes, sync);
    }
    if (!es.err) finish_code(br, es, cc, cs);
    if (!es.err) {
        if (w.check_end) { uint32_t e = br.finish(); if (e) es.set_raw(e); } // global image of a single-section frame
        else if (br.overrun()) es.set_raw(E_SHRT); // see lf_decode2_body
    }
    if (es.err && lane == 0) *w.err = es.err;
    sync();
    if (es.err || w.is_global) return; // global transforms are applied by the render step
    for (int t = m.nb_transforms - 1; t >= 0; --t) {
        ModImage one = m;
        one.nb_transforms = 1;
        one.tr[0] = m.tr[t];
        inverse_transforms(one, lane, nlanes);
        sync();
    }
}

// ---------------------------------------------------------------------------------------------
// Lane-per-stream variants of the three serial decoders (j40b_modlane.h): every thread of a warp owns one work item.
// `active`: this lane has one. `props`/`pstride`: the lane's 16 property slots. AnyFn / Sync as in hf_lanes_run.
template <int MODE, class AnyFn, class Sync>
J40B_HD J40B_INLINE void mod_lane_loop(ModLane<MODE> &L, AnyFn any, Sync sync) {
    for (;;) {
        if (!any(!L.done)) break;
        sync();
        if (!L.done && L.need_setup) L.setup();   // next channel: rare, lanes of one geometry get here together
        sync();
        if (!L.done) L.sample();
    }
}

// code spec of stage `stage` of an LF group: the global one, or the local one once the host has read it
J40B_HD J40B_INLINE uint32_t lf_stage_spec_off(const LfWork &w, int stage) {
    return w.local[stage].present && !w.local[stage].host_err ? w.local[stage].spec_off : w.f->global_spec_off;
}

J40B_HD J40B_INLINE bool spec_is_plain_ans(const uint8_t *arena, uint32_t spec_off) {
    if (!spec_off) return false;
    const DCodeSpec *spec = (const DCodeSpec *) (arena + spec_off);
    return !spec->use_prefix_code && !spec->lz77_enabled;
}

// LF group, stage 1: LfQuant, the 3-channel modular LF image (j40.h:6739-6757); same hand-over as lf_decode1_body
template <int MODE, class AnyFn, class Sync>
J40B_HD inline void lf_decode1_lanes(const LfWork *wp, bool active, LaneEnv env, AnyFn any, Sync sync) {
    ModLane<MODE> L;
    L.done = true; L.need_setup = false; L.es.err = 0;
    int32_t extra_prec = 0;
    if (active) {
        const LfWork &w = *wp;
        const DFrame &f = *w.f;
        DLfGroup &g = *w.g;
        const int n8 = g.width8 * g.height8;
        L.sc = w.lane_scratch; L.env = env;
        L.env.ring = w.ring; L.env.wring = w.wring; L.env.ring_w = w.ring_w; // (env.lstride is the launch's: w.ring_lstride)
        L.br.init(w.cs + g.sec_off, g.sec_size, g.sec_start_bit);
        extra_prec = (int32_t) L.br.u(2);
        ModImage &m = L.sc->m;
        m.num_channels = 3;
        for (int c = 0; c < 3; ++c) {
            m.ch[c].px = g.lfq + (size_t) c * n8;
            m.ch[c].stride = g.width8; m.ch[c].w = g.width8; m.ch[c].h = g.height8;
            m.ch[c].hshift = m.ch[c].vshift = 0;
        }
        const DTreeNode *tree;
        uint32_t spec_off;
        int32_t uses_wp;
        if (lf_stage_header(L.br, L.es, w, 0, m, tree, spec_off, uses_wp, true))
            L.begin(w.arena, spec_off, tree, uses_wp, 1 + g.idx, g.wp_scratch, g.lz_window, (1u << 18) - 1);
    }
    mod_lane_loop(L, any, sync);
    if (active) {
        const LfWork &w = *wp;
        DLfGroup &g = *w.g;
        const ModImage &m = L.sc->m;
        if (!L.es.err) finish_code(L.br, L.es, L.cc, L.cs);
        g.extra_prec = extra_prec;
        g.mid_bit = L.br.bits_consumed();
        g.nb_tr1 = imin(m.nb_transforms, MOD_MAX_TRANSFORMS);
        for (int t = 0; t < g.nb_tr1; ++t) g.tr1[t] = m.tr[t];
        if (L.es.err) *w.err = L.es.err;
    }
}

// LF group, stage 3a: entropy decode of the HF metadata image; transforms and varblock placement follow in lf_place_body
template <int MODE, class AnyFn, class Sync>
J40B_HD inline void lf_decode2_lanes(const LfWork *wp, bool active, LaneEnv env, AnyFn any, Sync sync) {
    ModLane<MODE> L;
    L.done = true; L.need_setup = false; L.es.err = 0;
    if (active) {
        const LfWork &w = *wp;
        const DFrame &f = *w.f;
        DLfGroup &g = *w.g;
        const int n8 = g.width8 * g.height8;
        L.sc = w.lane_scratch; L.env = env;
        L.env.ring = w.ring; L.env.wring = w.wring; L.env.ring_w = w.ring_w; // (env.lstride is the launch's: w.ring_lstride)
        L.br.init(w.cs + g.sec_off, g.sec_size, g.mid_bit);
        const int32_t nvb = (int32_t) L.br.u(ceil_lg32((uint32_t) n8)) + 1;
        g.nb_varblocks = nvb;
        ModImage &m = L.sc->m;
        hf_meta_channels(g, nvb, m);
        const DTreeNode *tree;
        uint32_t spec_off;
        int32_t uses_wp;
        if (lf_stage_header(L.br, L.es, w, 1, m, tree, spec_off, uses_wp, true))
            L.begin(w.arena, spec_off, tree, uses_wp, 1 + 2 * f.num_lf_groups + g.idx, g.wp_scratch, g.lz_window, (1u << 18) - 1);
    }
    mod_lane_loop(L, any, sync);
    if (active) {
        const LfWork &w = *wp;
        DLfGroup &g = *w.g;
        const ModImage &m = L.sc->m;
        if (!L.es.err) finish_code(L.br, L.es, L.cc, L.cs);
        g.nb_tr2 = imin(m.nb_transforms, MOD_MAX_TRANSFORMS);
        for (int t = 0; t < g.nb_tr2; ++t) g.tr2[t] = m.tr[t];
        g.end_bit = L.br.bits_consumed();
        g.meta_overrun = L.br.overrun() ? 1 : 0;
        if (L.es.err) *w.err = L.es.err;
    }
}

// LF group, stage 3b (one warp): inverse transforms of the HF metadata image, varblock placement (j40.h:6585-6720)
template <class Sync>
J40B_HD inline void lf_place_body(const LfWork &w, uint32_t *bitmap /* [height8 * 8] or null */, int lane, int nlanes, Sync sync) {
    if (*w.err) return;
    const DFrame &f = *w.f;
    DLfGroup &g = *w.g;
    ErrSlot es;
    es.err = 0;
    BitReader br; // stands for the decode stage's reader where errors are classified: only "ran past the end" matters
    br.base = nullptr; br.size = 0; br.buf = 0; br.nbits = 0; br.pos = g.meta_overrun ? 1u : 0u;
    ModImage m;
    hf_meta_channels(g, g.nb_varblocks, m);
    m.nb_transforms = g.nb_tr2;
    for (int t = 0; t < g.nb_tr2; ++t) m.tr[t] = g.tr2[t];
    // (an RCT over xfromy/bfromy/blockinfo is possible when their sizes coincide)
    for (int t = m.nb_transforms - 1; t >= 0; --t) {
        ModImage one = m;
        one.nb_transforms = 1;
        one.tr[0] = m.tr[t];
        inverse_transforms(one, lane, nlanes);
        sync();
    }
    if (bitmap) place_varblocks_warp(f, g, es, br, bitmap, lane, nlanes, sync);
    else if (lane == 0) place_varblocks(f, g, es, br);
    if (lane == 0) {
        // multi-section frames: the reference drops pad0/excs found at a section's end; running short is still an error
        if (!es.err && g.meta_overrun) es.set_raw(E_SHRT);
        if (es.err) *w.err = es.err;
    }
}

// one modular sub-bitstream per lane (same hand-over as modular_body)
template <int MODE, class AnyFn, class Sync>
J40B_HD inline void modular_lanes(ModWork *wp, bool active, LaneEnv env, AnyFn any, Sync sync) {
    ModLane<MODE> L;
    L.done = true; L.need_setup = false; L.es.err = 0;
    bool preset = false;
    if (active) {
        ModWork &w = *wp;
        const DFrame &f = *w.f;
        if (w.preset_err) { *w.err = w.preset_err; preset = true; }
        else if (w.skip_if && *w.skip_if) preset = true;
        else {
            L.sc = w.lane_scratch; L.env = env;
        L.env.ring = w.ring; L.env.wring = w.wring; L.env.ring_w = w.ring_w; // (env.lstride is the launch's: w.ring_lstride)
            L.br.init(w.cs + w.sec_off, w.sec_size, w.start_bit_src ? *w.start_bit_src : w.sec_start_bit);
            L.sc->m = w.m;
            if (!w.header_parsed) modular_header(L.br, L.es, f.have_global_tree != 0, L.sc->m);
            if (!L.es.err) L.begin(w.arena, w.spec_off, (const DTreeNode *) (w.arena + w.tree_off), w.tree_uses_wp, w.sidx, w.wp_scratch, w.lz_window, w.lz_mask);
        }
    }
    mod_lane_loop(L, any, sync);
    if (active && !preset) {
        ModWork &w = *wp;
        const ModImage &m = L.sc->m;
        if (!L.es.err) finish_code(L.br, L.es, L.cc, L.cs);
        if (!L.es.err) {
            if (w.check_end) { uint32_t e = L.br.finish(); if (e) L.es.set_raw(e); }
            else if (L.br.overrun()) L.es.set_raw(E_SHRT);
        }
        if (L.es.err) *w.err = L.es.err;
        else if (!w.is_global) { // global transforms are applied by the render step
            for (int t = m.nb_transforms - 1; t >= 0; --t) {
                ModImage one = m;
                one.nb_transforms = 1;
                one.tr[0] = m.tr[t];
                inverse_transforms(one, 0, 1);
            }
        }
    }
}

// diagnostics: varblock `voff` of w.g into the dense planes; same arithmetic as the tile back-end's scatter phase
template <class Sync>
J40B_HD inline void dump_coeffs_body(const DumpWork &w, int voff, int tid, int nth, Sync sync) {
    const DFrame &f = *w.f;
    const DLfGroup &g = *w.g;
    if (voff >= g.nb_varblocks) return;
    const DVarblock vb = g.varblocks[voff];
    const DctSelectInfo d = dct_select_info(vb.dctsel);
    const int size = 1 << (d.log_rows + d.log_columns);
    const size_t n8 = (size_t) g.width8 * g.height8;
    float *coef[3] = {w.out + vb.coeffoff, w.out + n8 * 64 + vb.coeffoff, w.out + 2 * n8 * 64 + vb.coeffoff};
    for (int c = 0; c < 3; ++c) for (int i = tid; i < size; i += nth) coef[c][i] = 0.0f;
    sync();
    for (int pass = 0; pass < f.num_passes; ++pass) {
        for (int c = 0; c < 3; ++c) {
            const uint32_t *slot = g.vb_tok + (((size_t) pass * 3 + c) * n8 + voff) * 2;
            const int32_t *order = f.order[pass][d.order_idx][c];
            for (uint32_t k = tid; k < slot[1]; k += nth) {
                const DToken *t = w.tokens + slot[0] + k;
                if (t->is_ext()) continue;
                const int32_t pos = order[t->pos()];
                coef[c][pos] = J40B_FADD(coef[c][pos], (float) token_value(t));
            }
        }
        sync();
    }
    if (!w.dequant) return;
    const float *dq = f.dq[d.param_idx];
    float mult[3];
    mult[1] = J40B_FMUL(J40B_FDIV(65536.0f, (float) f.global_scale), vb.hfmul_inv);
    mult[0] = J40B_FMUL(mult[1], f.x_qm_mult);
    mult[2] = J40B_FMUL(mult[1], f.b_qm_mult);
    for (int c = 0; c < 3; ++c) for (int i = tid; i < size; i += nth) {
        float v = coef[c][i];
        if (-1.0f <= v && v <= 1.0f) v = J40B_FMUL(v, f.quant_bias[c]);
        else v = J40B_FSUB(v, J40B_FDIV(f.quant_bias_num, v));
        coef[c][i] = J40B_FMUL(v, J40B_FDIV(mult[c], dq[i * 3 + c]));
    }
}

// one thread per pixel: global inverse RCTs (j40.h:8209) + j40__render_to_u8x4_rgba (j40.h:7910-7957)
// the 72 hard-coded palette deltas (j40.h:4275-4290): entry k serves index 2k as is and 2k + 1 negated
J40B_HD J40B_INLINE int32_t palette_delta(int32_t idx /* 0..142 */, int c) {
    const int16_t T[72][3] = {
        {0, 0, 0}, {4, 4, 4}, {11, 0, 0}, {0, 0, -13}, {0, -12, 0}, {-10, -10, -10},
        {-18, -18, -18}, {-27, -27, -27}, {-18, -18, 0}, {0, 0, -32}, {-32, 0, 0}, {-37, -37, -37},
        {0, -32, -32}, {24, 24, 45}, {50, 50, 50}, {-45, -24, -24}, {-24, -45, -45}, {0, -24, -24},
        {-34, -34, 0}, {-24, 0, -24}, {-45, -45, -24}, {64, 64, 64}, {-32, 0, -32}, {0, -32, 0},
        {-32, 0, 32}, {-24, -45, -24}, {45, 24, 45}, {24, -24, -45}, {-45, -24, 24}, {80, 80, 80},
        {64, 0, 0}, {0, 0, -64}, {0, -64, -64}, {-24, -24, 45}, {96, 96, 96}, {64, 64, 0},
        {45, -24, -24}, {34, -34, 0}, {112, 112, 112}, {24, -45, -45}, {45, 45, -24}, {0, -32, 32},
        {24, -24, 45}, {0, 96, 96}, {45, -24, 24}, {24, -45, -24}, {-24, -45, 24}, {0, -64, 0},
        {96, 0, 0}, {128, 128, 128}, {64, 0, 64}, {144, 144, 144}, {96, 96, 0}, {-36, -36, 36},
        {45, -24, -45}, {45, -45, -24}, {0, 0, -96}, {0, 128, 128}, {0, 96, 0}, {45, 24, -45},
        {-128, 0, 0}, {24, -45, 24}, {-45, 24, -45}, {64, 0, -64}, {64, -64, -64}, {96, 0, 96},
        {45, -45, 24}, {24, 45, -45}, {64, 64, -64}, {128, 128, 0}, {0, 0, -128}, {-24, 45, -45},
    };
    // the reference's table has 144 rows, (x), (-x) alternating, and is entered at idx + 1
    const int32_t v = T[(idx + 1) >> 1][c];
    return ((idx + 1) & 1) ? -v : v;
}

// one colour of a palette entry (j40__inverse_palette without prediction, j40.h:4447-4470; 16-bit buffers)
J40B_HD J40B_INLINE int16_t palette_value(const ModTransform &t, const int16_t *pal /* [num_c][nb_colours] */, int32_t idx_in, int i, int bpp) {
    int16_t idx = (int16_t) idx_in, val;
    if (idx < 0) {
        if (i < 3) {
            idx = (int16_t) (~idx % 143);
            val = (int16_t) palette_delta(idx, i);
            if (bpp > 8) val = (int16_t) (val * (1 << (imin(bpp, 24) - 8))); // (the reference shifts; same bits, defined for negatives)
        } else {
            val = 0;
        }
    } else if (idx < t.nb_colours) {
        val = pal[(size_t) i * (size_t) t.nb_colours + idx];
    } else {
        idx = (int16_t) (idx - t.nb_colours);
        if (idx < 64) {
            val = (int16_t) ((i < 3 ? idx >> (2 * i) : 0) * (((int32_t) 1 << bpp) - 1) / 4 + ((int32_t) 1 << imax(0, bpp - 3)));
        } else {
            val = (int16_t) (idx - 64);
            for (int j = 0; j < i; ++j) val = (int16_t) (val / 5);
            val = (int16_t) ((val % 5) * ((1 << bpp) - 1) / 4);
        }
    }
    return val;
}

// Inverse of a palette transform with delta entries (j40__inverse_palette with use_pred, j40.h:4416-4480): restored
// channel `i`, a serial scan in raster order -- entries below nb_deltas are added to a prediction from the restored
// neighbours. One thread per restored channel. (A rare feature: not worth a wavefront.)
J40B_HD inline void palette_delta_body(const RenderWork &w, int i, const int32_t *div24) {
    const DFrame &f = *w.f;
    const ModTransform &tr = f.global_tr[f.nb_global_transforms - 1];
    if (i >= tr.num_c) return;
    const int32_t width = f.width, height = f.height;
    const int16_t *table = w.plane[0], *index = w.plane[tr.begin_c + 1];
    int16_t *out = w.dplane[i];
    const bool use_wp = tr.d_pred == 6;
    WPState wp;
    wp.errors = w.dwp[i];
    wp.width = width;
    for (int k = 0; k < 5; ++k) wp.pred[k] = 0;
    wp.trueerrw = wp.trueerrn = wp.trueerrnw = wp.trueerrne = 0;
    if (use_wp) for (int32_t k = 0; k < width * 10; ++k) wp.errors[k] = 0;
    for (int32_t y = 0; y < height; ++y) {
        int16_t *row = out + (size_t) y * (size_t) width;
        for (int32_t x = 0; x < width; ++x) {
            const int32_t idx = index[(size_t) y * (size_t) width + x];
            const bool is_delta = idx < tr.nb_deltas;
            int32_t val = palette_value(tr, table, idx, i, f.bpp);
            // neighbours of the restored channel (j40.h:3965-4009)
            const int16_t *p = row + x;
            const int32_t pw = x > 0 ? p[-1] : y > 0 ? p[-width] : 0;
            const int32_t pn = y > 0 ? p[-width] : pw;
            const int32_t pnw = x > 0 && y > 0 ? p[-1 - width] : pw;
            const int32_t pne = x + 1 < width && y > 0 ? p[1 - width] : pn;
            const int32_t pnn = y > 1 ? p[-2 * width] : pn;
            const int32_t pnee = x + 2 < width && y > 0 ? p[2 - width] : pne;
            const int32_t pww = x > 1 ? p[-2] : pw;
            if (use_wp) wp_before_predict(wp, f.global_wp, div24, x, y, pw, pn, pnw, pne, pnn);
            if (is_delta) {
                bool bad = false;
                val = (int16_t) (val + mod_predict(tr.d_pred, pw, pn, pnw, pne, pnn, pww, pnee, wp.pred[4], &bad));
            }
            if (use_wp) wp_after_predict(wp, x, y, val);
            row[x] = (int16_t) val;
        }
    }
}

J40B_HD inline void render_px(const RenderWork &w, int x, int y) {
    const DFrame &f = *w.f;
    int16_t v[MOD_MAX_CH + 4];
    const size_t o = (size_t) y * (size_t) f.width + (size_t) x;
    // the coded channel list at this pixel; palette (meta) channels in front are tables, not samples
    int n = f.num_channels, meta = f.nb_meta_channels;
    for (int c = 0; c < n; ++c) v[c] = c < meta ? (int16_t) 0 : w.plane[c][o];
    for (int t = f.nb_global_transforms - 1; t >= 0; --t) {
        const ModTransform &tr = f.global_tr[t];
        if (tr.kind == 0) {
            int b = tr.begin_c;
            inverse_rct_px(tr.type, v[b], v[b + 1], v[b + 2]);
        } else {
            // every palette transform put its table in front of the list, so the one being undone (transforms are
            // undone last to first) is the first meta channel still there
            const int16_t *table = w.plane[f.nb_meta_channels - meta];
            const int first = tr.begin_c + 1, last = tr.begin_c + tr.num_c;
            const int32_t idx = v[first];
            for (int k = n - 1; k > first; --k) v[k + (last - first)] = v[k]; // channels behind the index channel
            n += last - first;
            for (int i = 0; i < tr.num_c; ++i) v[first + i] = tr.nb_deltas > 0 ? w.dplane[i][o] : palette_value(tr, table, idx, i, f.bpp);
            for (int k = 0; k + 1 < n; ++k) v[k] = v[k + 1]; // drop the palette channel itself
            --n; --meta;
        }
    }
    const int32_t maxpixel = (1 << f.bpp) - 1, half = 1 << (f.bpp - 1);
    uint8_t *out = w.rgba + (size_t) y * (size_t) w.rgba_stride + (size_t) x * 4;
    for (int i = 0; i < 4; ++i) {
        int32_t p = i < 3 ? v[i] : (f.alpha_channel >= 0 ? v[f.alpha_channel] : maxpixel);
        p = imin(imax(0, p), maxpixel);
        out[i] = (uint8_t) ((p * 255 + half) / maxpixel);
    }
}

} // namespace j40b
