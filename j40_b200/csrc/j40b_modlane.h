// j40-b200: modular sub-bitstreams, one per *thread* (device side, also compiled for the CPU kernel-logic tests).
//
// Same arithmetic as j40b_modular.h (j40.h:3965-4229: neighbours, weighted predictor, MA-tree walk, predictors, symbol
// read), organised for throughput instead of latency. The warp-per-stream decoder of j40b_modular.h spends 32 lanes
// and ~20 KB of shared memory on one serial stream; a batch has hundreds of such streams (LF groups, modular groups),
// and what bounds a pipelined step is the issue slots and the shared memory / L1 they take from the other kernels,
// not the latency of one stream. Here the lanes of a warp decode 32 streams side by side:
//   * the decoder is a per-lane state machine (position c, y, x; sliding neighbour and error windows in registers),
//     stepped by a warp-uniform loop with a warp barrier per sample, so that lanes of the same geometry stay converged;
//   * sample rows are read and written in the channel planes themselves (global memory, L1 / L2 hits: every address is
//     known a sample ahead), the weighted predictor's error rows live in the stream's global scratch, the pruned MA
//     tree in a per-stream scratch record; nothing but the 16 property values of the current sample is in shared memory;
//   * the properties of a sample are computed once (14 cheap expressions) and the tree walk indexes them, which keeps
//     the walk free of a per-property switch that would diverge between lanes.
// The executor picks this decoder when a launch has enough streams to fill warps (CudaBackend::launch_lf / launch_mod);
// single images keep the warp-per-stream decoder (lower latency).
#pragma once
#include "j40b_modular.h"

namespace j40b {

enum { LANE_PTREE_CAP = 192 };

// per-stream scratch in global memory
struct alignas(16) ModLaneScratch {
    DTreeNode ptree[LANE_PTREE_CAP];
    ModImage m;
};

template <int MODE> // MODE 1: rANS without LZ77 (see code_cluster); 0: generic
struct ModLane {
    BitReader br;
    ErrSlot es;
    CodeCtx cc;
    CodeState cs;
    const uint64_t *ans_tables;
    int32_t las;
    ModLaneScratch *sc;
    const DTreeNode *full_tree, *tree; // the sub-bitstream's tree; the one walked for the current channel
    int32_t *wp_scratch;               // [2][width][5] or null
    const int32_t *div24;              // shared 64-entry divisor table
    int32_t *props;                    // this lane's 16 property slots, `pstride` words apart
    int32_t pstride;
    int32_t sidx, tree_uses_wp;
    // current channel
    int32_t c, x, y, width, height, stride, nref, dist_mult;
    int16_t *px;
    bool uses_wp, need_setup, done;
    int32_t prev, prev2, n_ww, n_w, n_c, n_e;
    // weighted predictor windows (index 4 = signed true error)
    int32_t e_n[5], e_nw[5], e_ne[5], e_w[5], e_ww[4];
    WPParams wpp;

    J40B_HD J40B_INLINE int32_t symbol(int32_t ctx) {
        if (MODE == 1) {
            const uint32_t ci = cc.cluster_map[ctx];
            const uint64_t e = ans_tables[((size_t) ci << las) + ((cs.ans_state & 0xfff) >> cc.log_bucket)];
            const HybridCfg cfg = cc.clusters[ci].cfg;
            const int32_t token = ans_symbol_entry(br, cs.ans_state, cc.log_bucket, e);
            return hybrid_int(br, es, token, cfg);
        }
        return code(br, es, cc, cs, ctx, dist_mult);
    }

    // (image `sc->m` must be filled in; call begin() after the sub-bitstream's header has been read)
    J40B_HD void begin(const uint8_t *arena, uint32_t spec_off, const DTreeNode *tree_, int32_t tree_uses_wp_, int32_t sidx_,
                       int32_t *wp_scratch_, int32_t *lz_window, uint32_t lz_mask) {
        cc.init(arena, spec_off);
        cs.init(lz_window, lz_mask);
        ans_tables = (const uint64_t *) (arena + cc.spec->ans_tables_off);
        las = cc.spec->log_alpha_size;
        full_tree = tree_;
        tree_uses_wp = tree_uses_wp_;
        sidx = sidx_;
        wp_scratch = wp_scratch_;
        c = -1;
        need_setup = true;
        done = false;
    }

    // opens the next non-empty channel (j40.h:4127-4165); done when there is none
    J40B_HD void setup() {
        const ModImage &m = sc->m;
        for (;;) {
            if (++c >= m.num_channels) { done = true; return; }
            if (m.ch[c].w > 0 && m.ch[c].h > 0) break;
        }
        const ModChannel &ch = m.ch[c];
        width = ch.w; height = ch.h; stride = ch.stride; px = ch.px;
        dist_mult = m.dist_mult;
        wpp = m.wp;
        bool wp_ = tree_uses_wp != 0;
        const int n = prune_tree(full_tree, c, sidx, sc->ptree, LANE_PTREE_CAP, &wp_);
        if (n > 0) tree = sc->ptree; else { tree = full_tree; wp_ = tree_uses_wp != 0; }
        uses_wp = wp_;
        if (uses_wp && !wp_scratch) { es.set_raw(E_MEM); done = true; return; } // (the executor sizes it from the host's view of the tree)
        if (uses_wp) for (int32_t i = 0; i < width * 10; ++i) wp_scratch[i] = 0;
        nref = 0;
        for (int32_t i = c - 1; i >= 0; --i) {
            const ModChannel &r = m.ch[i];
            if (ch.w == r.w && ch.h == r.h && ch.hshift == r.hshift && ch.vshift == r.vshift) ++nref;
        }
        x = 0; y = 0;
        need_setup = false;
        if (MODE == 1 && cs.ans_state == 0) ans_seed(br, cs.ans_state); // the first sample reads a symbol
        row_start();
    }

    J40B_HD J40B_INLINE void row_start() {
        const int16_t *nrow = px + (size_t) (y > 0 ? y - 1 : 0) * (size_t) stride;
        prev = prev2 = 0;
        n_ww = n_w = 0;
        n_c = y > 0 ? nrow[0] : 0;
        n_e = y > 0 && width > 1 ? nrow[1] : n_c;
        if (uses_wp) {
            const int32_t *nerr = wp_scratch + (size_t) ((y & 1) ? 0 : width) * 5;
            for (int i = 0; i < 5; ++i) {
                e_w[i] = 0;
                if (i < 4) e_ww[i] = 0;
                e_n[i] = y > 0 ? nerr[i] : 0;
                e_nw[i] = e_n[i];
                e_ne[i] = y > 0 && width > 1 ? nerr[5 + i] : e_n[i];
            }
        }
    }

    // property `prop` >= 16 of the current sample: a reference channel's value (j40.h:4204-4216)
    J40B_HD int32_t ref_property(int32_t prop, bool *bad) const {
        const ModImage &m = sc->m;
        const ModChannel &ch = m.ch[c];
        int32_t want = (prop - 16) / 4;
        if (want >= nref) { *bad = true; return 0; }
        int32_t ri = -1;
        for (int32_t i = c - 1; i >= 0; --i) {
            const ModChannel &r = m.ch[i];
            if (ch.w != r.w || ch.h != r.h || ch.hshift != r.hshift || ch.vshift != r.vshift) continue;
            if (want-- == 0) { ri = i; break; }
        }
        const ModChannel &r = m.ch[ri];
        const int16_t *rp = r.px + (size_t) y * (size_t) r.stride + x;
        int32_t val = rp[0];
        if (prop & 2) {
            int32_t rw = x > 0 ? rp[-1] : 0;
            int32_t rn = y > 0 ? rp[-r.stride] : rw;
            int32_t rnw = x > 0 && y > 0 ? rp[-1 - r.stride] : rw;
            val -= mod_gradient(rw, rn, rnw);
        }
        if (prop & 1) val = iabs(val);
        return val;
    }

    // one sample (j40.h:4167-4229)
    J40B_HD J40B_INLINE void sample() {
        const int16_t *nrow = px + (size_t) (y > 0 ? y - 1 : 0) * (size_t) stride;
        const int32_t n_ee = y > 0 && x + 2 < width ? nrow[x + 2] : n_e;
        const int32_t pw = x > 0 ? prev : n_c;     // (n_c is 0 in the first row)
        const int32_t pn = y > 0 ? n_c : pw;
        const int32_t pnw = x > 0 && y > 0 ? n_w : pw;
        const int32_t pne = y > 0 ? n_e : pn;      // n_e already equals n_c at the right edge
        const int32_t pnn = y > 1 ? px[(size_t) (y - 2) * (size_t) stride + x] : pn;
        const int32_t pnee = y > 0 ? n_ee : pne;
        const int32_t pww = x > 1 ? prev2 : pw;
        const int32_t pnww = x > 1 && y > 0 ? n_ww : pww;
        int32_t wp_pred[5] = {0, 0, 0, 0, 0};
        int32_t maxerr = 0;
        int32_t ne_next[5] = {0, 0, 0, 0, 0};
        if (uses_wp) {
            // next sample's north-east errors: the address is known now, the values are needed a sample later
            const int32_t *nerr = wp_scratch + (size_t) ((y & 1) ? 0 : width) * 5;
            const bool have = y > 0 && x + 2 < width;
            for (int i = 0; i < 5; ++i) ne_next[i] = have ? nerr[(size_t) (x + 2) * 5 + i] : e_ne[i];
            // j40.h:4011-4072
            const int32_t te_w = e_w[4], te_n = e_n[4], te_nw = e_nw[4], te_ne = e_ne[4];
            wp_pred[0] = (pw + pne - pn) * 8;
            wp_pred[1] = pn * 8 - (((te_w + te_n + te_ne) * wpp.p1) >> 5);
            wp_pred[2] = pw * 8 - (((te_w + te_n + te_nw) * wpp.p2) >> 5);
            wp_pred[3] = pn * 8 - ((te_nw * wpp.p3[0] + te_n * wpp.p3[1] + te_ne * wpp.p3[2] +
                                    (pnn - pn) * 8 * wpp.p3[3] + (pnw - pw) * 8 * wpp.p3[4]) >> 5);
            int32_t w[4];
            for (int i = 0; i < 4; ++i) {
                int32_t errsum = e_n[i] + e_w[i] + e_nw[i] + e_ww[i] + e_ne[i] + (x + 1 < width ? 0 : e_w[i]);
                int32_t shift = imax(floor_lg32((uint32_t) errsum + 1) - 5, 0);
                w[i] = (int32_t) (4 + (((int64_t) wpp.w[i] * div24[errsum >> shift]) >> shift));
            }
            int32_t logw = floor_lg32((uint32_t) (w[0] + w[1] + w[2] + w[3])) - 4;
            int32_t wsum = 0, sum = 0;
            for (int i = 0; i < 4; ++i) {
                w[i] >>= logw;
                wsum += w[i];
                sum += wp_pred[i] * w[i];
            }
            wp_pred[4] = (int32_t) ((((int64_t) sum + (wsum >> 1) - 1) * div24[wsum - 1]) >> 24);
            if (((te_n ^ te_w) | (te_n ^ te_nw)) <= 0) {
                int32_t lo = imin(pw, imin(pn, pne)) * 8;
                int32_t hi = imax(pw, imax(pn, pne)) * 8;
                wp_pred[4] = imin(imax(lo, wp_pred[4]), hi);
            }
            maxerr = te_w;
            if (iabs(maxerr) < iabs(te_n)) maxerr = te_n;
            if (iabs(maxerr) < iabs(te_nw)) maxerr = te_nw;
            if (iabs(maxerr) < iabs(te_ne)) maxerr = te_ne;
        }
        // ---- MA tree: the sample's properties once, then the walk indexes them (j40.h:4178-4219)
        const DTreeNode *n = tree;
        DTreeNode node = *n;
        if (node.a < 0) {
            int32_t *p = props;
            const int32_t ps = pstride;
            p[2 * ps] = y; p[3 * ps] = x; p[4 * ps] = iabs(pn); p[5 * ps] = iabs(pw); p[6 * ps] = pn; p[7 * ps] = pw;
            p[8 * ps] = x > 0 ? pw - (pww + pnw - pnww) : pw;
            p[9 * ps] = pw + pn - pnw; p[10 * ps] = pw - pnw; p[11 * ps] = pnw - pn; p[12 * ps] = pn - pne;
            p[13 * ps] = pn - pnn; p[14 * ps] = pw - pww; p[15 * ps] = maxerr;
            p[0] = c; p[1 * ps] = sidx;
            do {
                const int32_t prop = -1 - node.a;
                int32_t val;
                if (prop < 16) val = p[prop * ps];
                else {
                    bool bad = false;
                    val = ref_property(prop, &bad);
                    if (bad) { es.set(br, E_TREC); done = true; return; }
                }
                node = tree[val > node.b ? node.c : node.d];
            } while (node.a < 0);
        }
        int32_t val = symbol(node.a);
        val = unpack_signed(val) * node.d + node.c;
        bool bad = false;
        val += mod_predict(node.b, pw, pn, pnw, pne, pnn, pww, pnee, wp_pred[4], &bad);
        if (bad) es.set(br, E_PRED);
        if (es.err) { done = true; return; }
        if ((uint32_t) (val + 32768) > 65535u) { es.set(br, E_POVF); done = true; return; }
        px[(size_t) y * (size_t) stride + x] = (int16_t) val;
        prev2 = prev;
        prev = val;
        n_ww = n_w; n_w = n_c; n_c = n_e; n_e = n_ee;
        if (uses_wp) {
            // j40.h:4103-4111, then slide the error windows
            int32_t *err = wp_scratch + ((size_t) ((y & 1) ? width : 0) + (size_t) x) * 5;
            const int32_t v8 = val * 8;
            for (int i = 0; i < 5; ++i) {
                const int32_t e = i < 4 ? (iabs(wp_pred[i] - v8) + 3) >> 3 : wp_pred[4] - v8;
                err[i] = e;
                if (i < 4) e_ww[i] = e_w[i];
                e_w[i] = e;
                e_nw[i] = e_n[i];
                e_n[i] = e_ne[i];
                e_ne[i] = ne_next[i];
            }
        }
        if (++x == width) {
            x = 0;
            if (++y == height) need_setup = true;
            else row_start();
        }
    }
};

} // namespace j40b
