// TEST INFRASTRUCTURE ONLY: CPU emulation of the CUDA kernels' *logic* for the `-m "not gpu"` test suite.
// The container that runs the CPU tests has no GPU; this backend executes the very same kernel bodies
// (j40_b200/csrc/j40b_exec.h) one work item after another with a single "thread", so that parser,
// context modelling, bit-exact float code paths and the batch pipeline can be checked against the oracle
// before the real kernels run on a B200. It is NOT part of the product: libj40b200.so has no CPU path and
// the j40_b200 package never loads this library.
#include "../../j40_b200/csrc/j40b_pipeline.h"
#include <stdlib.h>
#include <string.h>

using namespace j40b;

static int g_last_tree_lanes = 1;

struct HostEmuBackend {
    void *dev_alloc(size_t n) { return calloc(n ? n : 1, 1); }
    void dev_free(void *p) { free(p); }
    void *host_alloc(size_t n) { return malloc(n ? n : 1); }
    void host_free(void *p) { free(p); }
    void h2d(void *d, const void *s, size_t n) { memcpy(d, s, n); }
    void d2h(void *d, const void *s, size_t n) { memcpy(d, s, n); }
    void d2h_async(void *d, const void *s, size_t n) { memcpy(d, s, n); }
    void dev_memset(void *d, int v, size_t n) { memset(d, v, n); }
    void sync() {}
    int lane_stride() const { return 1; } // work items per slot of the lane decoders' interleaved buffers
    // shared memory of one warp of the serial decoders
    struct WarpMem {
        WarpScratch ws;
        std::vector<int32_t> wp; // error row [cap][5] followed by the reference-property rows, as in carve_warp_slice
        std::vector<int16_t> rows;
        SimtLane tab[SIMT_LANES];
        SimtLeaf leaves[SIMT_LANES];
        int32_t div24[64];
        ModSmem ms;
        explicit WarpMem(int cap) : wp((size_t) cap * (5 + SIMT_REF_SLOTS) + 1), rows((size_t) cap * 3 + 1) {
            ms.wp = wp.data(); ms.rows = cap ? rows.data() : nullptr; ms.refp = wp.data() + (size_t) cap * 5; ms.tab = tab; ms.leaves = leaves; ms.info = ws.info; ms.cap = cap; ms.tabs = nullptr; ms.tabs_entries = 0; ms.lanes = getenv("HOSTEMU_SIMT_LANES") ? atoi(getenv("HOSTEMU_SIMT_LANES")) : SIMT_LANES;
            fill_div24(div24, 0, 1);
        }
    };
    // HOSTEMU_ROW_CAP overrides the row-path width limit (0 = always take the plain path)
    static int row_cap(int dflt) { const char *e = getenv("HOSTEMU_ROW_CAP"); return e ? atoi(e) : dflt; }
    // HOSTEMU_LANE=1: the lane-per-stream decoders (j40b_modlane.h) instead of the warp-per-stream ones
    static bool lane_mode() { const char *e = getenv("HOSTEMU_LANE"); return e && atoi(e) != 0; }
    // like the device in split mode: the sharpness channel is decoded behind everything else (join_side)
    const LfWork *pending_w = nullptr;
    int pending_n = 0;
    void join_side() {
        if (!pending_w) return;
        WarpMem wm(row_cap(256));
        for (int i = 0; i < pending_n; ++i) lf_chan_body<MC_WP>(pending_w[i], 1, 3, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), true);
        for (int i = 0; i < pending_n; ++i) lf_chan_body<MC_GRAD>(pending_w[i], 1, 3, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), true);
        for (int i = 0; i < pending_n; ++i) lf_chan_body<MC_WIDE>(pending_w[i], 1, 3, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), true);
        for (int i = 0; i < pending_n; ++i) lf_chan_body<MC_GEN>(pending_w[i], 1, 3, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), true);
        for (int i = 0; i < pending_n; ++i) lf_chan_body<MC_REST>(pending_w[i], 1, 3, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), true);
        pending_w = nullptr;
    }
    void launch_lf(const LfWork *w, int n, size_t, bool split, int tree_lanes) {
        g_last_tree_lanes = tree_lanes; // (what the device would size its lane groups by)
        std::vector<uint8_t> copy(40 * 1024);
        WarpMem wm(row_cap(256));
        if (lane_mode()) {
            int32_t props[16], nodes[LANE_NODE_CAP * 4];
            std::vector<uint32_t> bitmap(256 * 8);
            auto any = [](bool p) { return p; };
            for (int i = 0; i < n; ++i) {
                LaneEnv env;
                env.div24 = wm.div24; env.props = props; env.nodes = (i & 2) ? nodes : nullptr; env.lstride = 1;
                env.ring = nullptr; env.wring = nullptr; env.ring_w = 0;
                LfWork ww = w[i];
                if (i & 1) ww.ring = nullptr; // every other stream without the row ring (the path of channels wider than it)
                if (spec_is_plain_ans(w[i].arena, lf_stage_spec_off(w[i], 0))) lf_decode1_lanes<1>(&ww, true, env, any, NoSync());
                else lf_decode1_lanes<0>(&ww, true, env, any, NoSync());
                lf_post_body(w[i], 0, 1, NoSync());
                if (!*w[i].err) {
                    if (spec_is_plain_ans(w[i].arena, lf_stage_spec_off(w[i], 1))) lf_decode2_lanes<1>(&ww, true, env, any, NoSync());
                    else lf_decode2_lanes<0>(&ww, true, env, any, NoSync());
                }
                lf_place_body(w[i], (i & 1) ? bitmap.data() : nullptr, 0, 1, NoSync());
                lf_llf_body(w[i], 0, 1, NoSync());
            }
            return;
        }
        // like the device: per stage and channel the four class kernels in a row, every LF group in each
        if (getenv("HOSTEMU_NO_SPLIT")) split = false;
        auto stage = [&](int st, int c0, int c1, bool sp) {
            for (int c = c0; c < c1; ++c) {
                for (int i = 0; i < n; ++i) lf_chan_body<MC_WP>(w[i], st, c, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), sp);
                for (int i = 0; i < n; ++i) lf_chan_body<MC_GRAD>(w[i], st, c, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), sp);
                for (int i = 0; i < n; ++i) lf_chan_body<MC_WIDE>(w[i], st, c, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), sp);
                for (int i = 0; i < n; ++i) lf_chan_body<MC_GEN>(w[i], st, c, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), sp);
                for (int i = 0; i < n; ++i) lf_chan_body<MC_REST>(w[i], st, c, wm.ws, wm.ms, wm.div24, 0, 1, NoSync(), sp);
            }
        };
        stage(0, 0, 3, false);
        for (int i = 0; i < n; ++i) lf_post_body(w[i], 0, 1, NoSync());
        if (split) { stage(1, 0, 3, true); pending_w = w; pending_n = n; }
        else stage(1, 0, 4, false);
        for (int i = 0; i < n; ++i) lf_llf_body(w[i], 0, 1, NoSync());
    }
    // (the device pads the work list to whole blocks per image and pass; a small block here so that the padding shows up)
    int hf_block_lanes(int) const { return getenv("HOSTEMU_HF_BLOCK") ? atoi(getenv("HOSTEMU_HF_BLOCK")) : 8; }
    void launch_hf(const HfPrepWork *pw, int ngroups, const HfWork *w, int n, size_t, int) {
        for (int i = 0; i < ngroups; ++i) hf_prep_body(pw[i], 0, 1, NoSync());
        std::vector<uint8_t> copy(40 * 1024);
        auto any = [](bool p) { return p; };
        for (int i = 0; i < n; ++i) {
            if (!w[i].grp || *w[i].lf_err) continue;
            // like the device kernel: every other "warp" gets a staged copy of the code spec tables
            const int pass = w[i].grp->pass;
            bool staged = (i & 1) && stage_spec_blob(w[i].arena, w[i].f->coeff_spec_off[pass], copy.data(), (uint32_t) copy.size(), 0, 1);
            uint16_t lut[128];
            for (int k = 0; k < 64; ++k) { lut[k] = (uint16_t) coeff_nnz_ctx2(k); lut[64 + k] = (uint16_t) (k ? coeff_freq_ctx2(k) : 0); }
            const uint8_t *sc = staged ? copy.data() : nullptr;
            if (hf_is_plain_ans(w[i])) hf_lanes_run<1>(&w[i], true, sc, w[i].arena, pass, (i & 2) ? lut : nullptr, any, NoSync());
            else hf_lanes_run<0>(&w[i], true, sc, w[i].arena, pass, (i & 2) ? lut : nullptr, any, NoSync());
        }
    }
    void launch_back(const BackWork *w, int n) {
        std::vector<float> coef(3 * TILE_CH), big(4 * 65536);
        for (int i = 0; i < n; ++i) {
            for (int t = 0; t < 16; ++t) { TileShared ts; ts.tables_staged = 0; back_tile_body(w[i], t & 3, t >> 2, coef.data(), ts, 0, 1, NoSync()); }
            BackWork bw = w[i];
            bw.big_scratch = big.data();
            back_generic_body(bw, 0, 1, NoSync());
        }
    }
    void launch_mod(ModWork *w, int n, size_t, int) {
        std::vector<uint8_t> copy(40 * 1024);
        WarpMem wm(row_cap(1024));
        if (lane_mode()) {
            int32_t props[16], nodes[LANE_NODE_CAP * 4];
            auto any = [](bool p) { return p; };
            for (int i = 0; i < n; ++i) {
                LaneEnv env;
                env.div24 = wm.div24; env.props = props; env.nodes = (i & 1) ? nodes : nullptr; env.lstride = 1;
                env.ring = nullptr; env.wring = nullptr; env.ring_w = 0;
                ModWork ww = w[i];
                if (i & 2) ww.ring = nullptr;
                if (spec_is_plain_ans(w[i].arena, w[i].spec_off)) modular_lanes<1>(&ww, true, env, any, NoSync());
                else modular_lanes<0>(&ww, true, env, any, NoSync());
            }
            return;
        }
        for (int i = 0; i < n; ++i) {
            bool staged = !(i & 1) && stage_spec_blob(w[i].arena, w[i].spec_off, copy.data(), (uint32_t) copy.size(), 0, 1);
            modular_body(w[i], wm.ws, wm.ms, wm.div24, staged ? copy.data() : nullptr, w[i].arena, 0, 1, NoSync());
        }
    }
    void launch_dump(const DumpWork &w, int n) { for (int v = 0; v < n; ++v) dump_coeffs_body(w, v, 0, 1, NoSync()); }
    void mark_modular(int) {}
    void launch_palette_delta(const RenderWork *w, int num_c) { for (int i = 0; i < num_c; ++i) palette_delta_body(*w, i, h_div24_table.v); }
    void launch_render(const RenderWork *w, int width, int height) {
        for (int y = 0; y < height; ++y) for (int x = 0; x < width; ++x) render_px(*w, x, y);
    }
};

extern "C" {

// decodes one image; returns the error code (0 = ok). *out is malloc'ed (stride*height bytes).
__attribute__((visibility("default"))) uint32_t hostemu_decode(const uint8_t *data, size_t size, uint8_t **out, int32_t *w, int32_t *h, int32_t *stride) {
    HostEmuBackend be;
    *out = nullptr;
    *w = *h = *stride = 0;
    uint32_t err = 0;
    Batch<HostEmuBackend> b(be);
    b.add(data, size);
    if (b.plans[0]->err) return b.plans[0]->err;
    for (int attempt = 0; attempt < 6; ++attempt) {
        b.upload();
        b.execute();
        b.collect_errors();
        err = b.results[0].err;
        // internal conditions, handled like j40b_batch_wait does: a token arena that was too small, local MA trees
        if (err == E_TOKV && !b.full_token_cap) { b.full_token_cap = true; continue; }
        if (err == E_LTRE && b.resolve_local_trees()) continue;
        if (!err) {
            const ImageResult &r = b.results[0];
            *out = (uint8_t *) malloc((size_t) r.stride * (size_t) r.height);
            b.download_pixels(0, *out);
            *w = r.width; *h = r.height; *stride = r.stride;
        }
        break;
    }
    return err;
}

__attribute__((visibility("default"))) void hostemu_free(void *p) { free(p); }
// most inner nodes / leaves of any LF-group channel's pruned MA tree in the last decode (1: no LF groups)
__attribute__((visibility("default"))) int hostemu_last_tree_lanes() { return g_last_tree_lanes; }

// decodes one image and returns intermediate array `what` of LF group `lfg` (Batch::debug_dump); bytes written or 0
__attribute__((visibility("default"))) size_t hostemu_dump(const uint8_t *data, size_t size, int lfg, int what, void *dst, size_t cap) {
    HostEmuBackend be;
    Batch<HostEmuBackend> b(be);
    b.full_token_cap = true;
    b.add(data, size);
    if (b.plans[0]->err) return 0;
    b.upload();
    b.execute();
    b.collect_errors();
    return b.debug_dump(0, (size_t) lfg, what, dst, cap);
}

// table helpers exposed for unit tests
__attribute__((visibility("default"))) int hostemu_dq_matrix(int idx, float *out, int cap) {
    std::vector<float> m = compute_dq_matrix_default(idx);
    if ((int) m.size() > cap) return 0;
    memcpy(out, m.data(), m.size() * 4);
    return (int) m.size() / 3;
}
__attribute__((visibility("default"))) int hostemu_natural_order(int log_rows, int log_columns, int32_t *out) {
    std::vector<int32_t> o = compute_natural_order(log_rows, log_columns);
    memcpy(out, o.data(), o.size() * 4);
    return (int) o.size();
}
__attribute__((visibility("default"))) void hostemu_srgb_thresholds(int bpp, float *thr) { compute_srgb_thresholds(bpp, thr); }
__attribute__((visibility("default"))) int hostemu_srgb_lookup(const float *thr, float v) { return srgb_u8_from_linear(thr, v); }
// the tile kernel's two-step table search, on the library's global tables
__attribute__((visibility("default"))) void hostemu_srgb_lut_lookup(const float *v, int n, uint8_t *out) {
    const GlobalTables &gt = GlobalTables::get();
    for (int i = 0; i < n; ++i) out[i] = (uint8_t) srgb_u8_lut(gt.srgb_thr, gt.srgb_lut, v[i]);
}
__attribute__((visibility("default"))) void hostemu_inverse_transform(int dctsel, float *buf) {
    DctSelectInfo d = dct_select_info(dctsel);
    if (is_special_8x8(dctsel)) { inverse_special(dctsel, buf); return; }
    std::vector<float> scratch((size_t) 1 << (d.log_rows + d.log_columns));
    inverse_dct2d(buf, scratch.data(), d.log_rows, d.log_columns, 0, 1, NoSync());
}
__attribute__((visibility("default"))) void hostemu_forward_llf(float *buf, int log_rows, int log_columns) {
    std::vector<float> scratch((size_t) 1 << (log_rows + log_columns));
    forward_dct2d_llf(buf, scratch.data(), log_rows, log_columns, 0, 1, NoSync());
}

}
