// j40-b200: modular sub-bitstreams, one per *thread* (device side, also compiled for the CPU kernel-logic tests).
//
// Same arithmetic as j40b_modular.h (j40.h:3965-4229: neighbours, weighted predictor, MA-tree walk, predictors, symbol
// read), organised for throughput instead of latency. The warp-per-stream decoder of j40b_modular.h spends 32 lanes
// and ~20 KB of shared memory on one serial stream; a batch has hundreds of such streams (LF groups, modular groups),
// and what bounds a pipelined step is the issue slots and the shared memory / L1 they take from the other kernels,
// not the latency of one stream. Here the lanes of a warp decode 32 streams side by side:
//   * the decoder is a per-lane state machine (position c, y, x; sliding neighbour and error windows in registers),
//     stepped by a warp-uniform loop with a warp barrier per sample, so that lanes of the same geometry stay converged;
//   * everything a lane re-reads lives in a lane-interleaved layout, element e of lane l at [e][l]: the last three
//     sample rows and the weighted predictor's two error rows in a warp-private ring in global memory (lanes at the
//     same x touch one cache line per access instead of 32), the pruned MA tree of the current channel and the 16
//     property values of the current sample in shared memory (bank = lane: conflict-free whatever node or property a
//     lane is at). Finished samples also go to the channel plane, which later stages and reference properties read;
//   * leaves of the compiled tree carry their cluster and hybrid-integer configuration, so that a symbol costs one
//     dependent global load (the alias-table entry) instead of three (cluster map, cluster record, table).
// The executor picks this decoder when a launch has enough streams to fill warps (CudaBackend::launch_lf / launch_mod);
// single images keep the warp-per-stream decoder (lower latency).
#pragma once
#include "j40b_modular.h"

namespace j40b {

enum { LANE_PTREE_CAP = 192, LANE_NODE_CAP = 64 };

// per-stream scratch in global memory
struct alignas(16) ModLaneScratch {
    DTreeNode ptree[LANE_PTREE_CAP];
    ModImage m;
};

// what a lane shares with its warp: tables in shared memory, the warp's ring in global memory. All pointers are already
// offset to this lane's slot; consecutive elements are `lstride` words (or samples) apart.
struct LaneEnv {
    const int32_t *div24;  // 64-entry divisor table of the weighted predictor
    int32_t *props;        // [16] property values of the current sample
    int32_t *nodes;        // [LANE_NODE_CAP * 4] compiled tree of the current channel (DTreeNode words)
    int16_t *ring;         // [3][ring_w] sample rows (null: none)
    int32_t *wring;        // [2][ring_w][5] weighted-predictor error rows
    int32_t ring_w;        // widest channel the ring takes; wider ones use the planes / the stream's own error rows
    int32_t lstride;       // lanes per warp slot (32 on the device, 1 on the CPU)
};

// leaf word b of a compiled tree: predictor | cluster << 8 | split_exp << 16 | msb_in_token << 20 | lsb_in_token << 24
J40B_HD J40B_INLINE HybridCfg lane_leaf_cfg(uint32_t b) {
    HybridCfg c;
    c.split_exp = (int8_t) ((b >> 16) & 15); c.msb_in_token = (int8_t) ((b >> 20) & 15); c.lsb_in_token = (int8_t) ((b >> 24) & 15); c.pad = 0;
    // the largest token whose value fits 32 bits (as the host computes it for the cluster, j40b_host.cc hybrid_cfg)
    c.max_token = (1 << c.split_exp) + ((30 - c.split_exp) << (c.lsb_in_token + c.msb_in_token)) - 1;
    return c;
}

template <int MODE> // MODE 1: rANS without LZ77 (see code_cluster); 0: generic
struct ModLane {
    BitReader br;
    ErrSlot es;
    CodeCtx cc;
    CodeState cs;
    const uint64_t *ans_tables;
    int32_t las;
    LaneEnv env;
    ModLaneScratch *sc;
    const DTreeNode *full_tree, *tree; // the sub-bitstream's tree; the one walked for the current channel
    int32_t *wp_scratch;               // [2][width][5] of this stream, or null
    int32_t sidx, tree_uses_wp;
    // current channel
    int32_t c, x, y, width, height, stride, nref, dist_mult;
    int16_t *px;
    const int16_t *ref0;               // nearest reference channel (properties 16..19), or null
    int32_t ref0_stride;
    bool uses_wp, need_setup, done, ringed, tree_smem, fused;
    int32_t prev, prev2, n_ww, n_w, n_c, n_e;
    // weighted predictor windows (index 4 = signed true error)
    int32_t e_n[5], e_nw[5], e_ne[5], e_w[5], e_ww[4];
    WPParams wpp;

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

    J40B_HD J40B_INLINE int16_t *ring_row(int32_t yy) const { return env.ring + (size_t) ((yy % 3) * env.ring_w) * (size_t) env.lstride; }
    J40B_HD J40B_INLINE int32_t *err_row(int32_t yy) const {
        return ringed ? env.wring + (size_t) ((yy & 1) * env.ring_w) * 5 * (size_t) env.lstride : wp_scratch + (size_t) ((yy & 1) ? width : 0) * 5;
    }
    J40B_HD J40B_INLINE int32_t estride() const { return ringed ? env.lstride : 1; }

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
        ringed = env.ring && width <= env.ring_w;
        bool wp_ = tree_uses_wp != 0;
        const int n = prune_tree(full_tree, c, sidx, sc->ptree, LANE_PTREE_CAP, &wp_);
        fused = false;
        tree_smem = false;
        if (n > 0) {
            tree = sc->ptree;
            if (MODE == 1) { // leaves resolved to their cluster
                for (int i = 0; i < n; ++i) {
                    DTreeNode &t = sc->ptree[i];
                    if (t.a < 0) continue;
                    const uint32_t ci = cc.cluster_map[t.a];
                    const HybridCfg cfg = cc.clusters[ci].cfg;
                    t.b = (int32_t) ((uint32_t) (t.b & 0xff) | ci << 8 | (uint32_t) (cfg.split_exp & 15) << 16 | (uint32_t) (cfg.msb_in_token & 15) << 20 |
                                     (uint32_t) (cfg.lsb_in_token & 15) << 24);
                }
                fused = true;
            }
            if (env.nodes && n <= LANE_NODE_CAP) {
                for (int i = 0; i < n; ++i) {
                    const DTreeNode t = sc->ptree[i];
                    int32_t *d = env.nodes + (size_t) (i * 4) * (size_t) env.lstride;
                    d[0] = t.a; d[env.lstride] = t.b; d[2 * env.lstride] = t.c; d[3 * env.lstride] = t.d;
                }
                tree_smem = true;
            }
        } else { tree = full_tree; wp_ = tree_uses_wp != 0; }
        uses_wp = wp_;
        if (uses_wp && !ringed && !wp_scratch) { es.set_raw(E_MEM); done = true; return; } // (the executor sizes it from the host's view of the tree)
        if (uses_wp) {
            int32_t *e0 = err_row(0), *e1 = err_row(1);
            const int32_t es_ = estride();
            for (int32_t i = 0; i < width * 5; ++i) { e0[(size_t) i * es_] = 0; e1[(size_t) i * es_] = 0; }
        }
        nref = 0;
        ref0 = nullptr; ref0_stride = 0;
        for (int32_t i = c - 1; i >= 0; --i) {
            const ModChannel &r = m.ch[i];
            if (ch.w == r.w && ch.h == r.h && ch.hshift == r.hshift && ch.vshift == r.vshift) {
                if (nref++ == 0) { ref0 = r.px; ref0_stride = r.stride; }
            }
        }
        x = 0; y = 0;
        need_setup = false;
        if (MODE == 1 && cs.ans_state == 0) ans_seed(br, cs.ans_state); // the first sample reads a symbol
        row_start();
    }

    J40B_HD J40B_INLINE int32_t north(int32_t yy, int32_t xx) const { // sample (xx, yy) of an earlier row
        return ringed ? ring_row(yy)[(size_t) xx * env.lstride] : px[(size_t) yy * (size_t) stride + xx];
    }

    J40B_HD J40B_INLINE void row_start() {
        prev = prev2 = 0;
        n_ww = n_w = 0;
        n_c = y > 0 ? north(y - 1, 0) : 0;
        n_e = y > 0 && width > 1 ? north(y - 1, 1) : n_c;
        if (uses_wp) {
            const int32_t *nerr = err_row(y + 1);
            const int32_t es_ = estride();
            for (int i = 0; i < 5; ++i) {
                e_w[i] = 0;
                if (i < 4) e_ww[i] = 0;
                e_n[i] = y > 0 ? nerr[(size_t) i * es_] : 0;
                e_nw[i] = e_n[i];
                e_ne[i] = y > 0 && width > 1 ? nerr[(size_t) (5 + i) * es_] : e_n[i];
            }
        }
    }

    // property `prop` >= 16 of the current sample: a reference channel's value (j40.h:4204-4216)
    J40B_HD int32_t ref_property(int32_t prop, bool *bad) const {
        int32_t want = (prop - 16) / 4;
        if (want >= nref) { *bad = true; return 0; }
        const int16_t *rpx = ref0;
        int32_t rstride = ref0_stride;
        if (want > 0) { // farther reference channels: rare, looked up in the channel list
            const ModImage &m = sc->m;
            const ModChannel &ch = m.ch[c];
            for (int32_t i = c - 1; i >= 0; --i) {
                const ModChannel &r = m.ch[i];
                if (ch.w != r.w || ch.h != r.h || ch.hshift != r.hshift || ch.vshift != r.vshift) continue;
                if (want-- == 0) { rpx = r.px; rstride = r.stride; break; }
            }
        }
        const int16_t *rp = rpx + (size_t) y * (size_t) rstride + x;
        int32_t val = rp[0];
        if (prop & 2) {
            int32_t rw = x > 0 ? rp[-1] : 0;
            int32_t rn = y > 0 ? rp[-rstride] : rw;
            int32_t rnw = x > 0 && y > 0 ? rp[-1 - rstride] : rw;
            val -= mod_gradient(rw, rn, rnw);
        }
        if (prop & 1) val = iabs(val);
        return val;
    }

    J40B_HD J40B_INLINE DTreeNode node_at(int32_t i) const {
        if (tree_smem) {
            const int32_t *d = env.nodes + (size_t) (i * 4) * (size_t) env.lstride;
            DTreeNode t;
            t.a = d[0]; t.b = d[env.lstride]; t.c = d[2 * env.lstride]; t.d = d[3 * env.lstride];
            return t;
        }
        return tree[i];
    }

    // one sample (j40.h:4167-4229)
    J40B_HD J40B_INLINE void sample() {
        const int32_t ls = env.lstride;
        const int32_t n_ee = y > 0 && x + 2 < width ? north(y - 1, x + 2) : n_e;
        const int32_t pw = x > 0 ? prev : n_c;     // (n_c is 0 in the first row)
        const int32_t pn = y > 0 ? n_c : pw;
        const int32_t pnw = x > 0 && y > 0 ? n_w : pw;
        const int32_t pne = y > 0 ? n_e : pn;      // n_e already equals n_c at the right edge
        const int32_t pnn = y > 1 ? north(y - 2, x) : pn;
        const int32_t pnee = y > 0 ? n_ee : pne;
        const int32_t pww = x > 1 ? prev2 : pw;
        const int32_t pnww = x > 1 && y > 0 ? n_ww : pww;
        int32_t wp_pred[5] = {0, 0, 0, 0, 0};
        int32_t maxerr = 0;
        int32_t ne_next[5] = {0, 0, 0, 0, 0};
        if (uses_wp) {
            // next sample's north-east errors: the address is known now, the values are needed a sample later
            const int32_t *nerr = err_row(y + 1);
            const int32_t es_ = estride();
            const bool have = y > 0 && x + 2 < width;
            for (int i = 0; i < 5; ++i) ne_next[i] = have ? nerr[((size_t) (x + 2) * 5 + i) * es_] : e_ne[i];
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
                w[i] = (int32_t) (4 + (((int64_t) wpp.w[i] * env.div24[errsum >> shift]) >> shift));
            }
            int32_t logw = floor_lg32((uint32_t) (w[0] + w[1] + w[2] + w[3])) - 4;
            int32_t wsum = 0, sum = 0;
            for (int i = 0; i < 4; ++i) {
                w[i] >>= logw;
                wsum += w[i];
                sum += wp_pred[i] * w[i];
            }
            wp_pred[4] = (int32_t) ((((int64_t) sum + (wsum >> 1) - 1) * env.div24[wsum - 1]) >> 24);
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
        DTreeNode node = node_at(0);
        if (node.a < 0) {
            int32_t *p = env.props;
            p[2 * ls] = y; p[3 * ls] = x; p[4 * ls] = iabs(pn); p[5 * ls] = iabs(pw); p[6 * ls] = pn; p[7 * ls] = pw;
            p[8 * ls] = x > 0 ? pw - (pww + pnw - pnww) : pw;
            p[9 * ls] = pw + pn - pnw; p[10 * ls] = pw - pnw; p[11 * ls] = pnw - pn; p[12 * ls] = pn - pne;
            p[13 * ls] = pn - pnn; p[14 * ls] = pw - pww; p[15 * ls] = maxerr;
            p[0] = c; p[ls] = sidx;
            do {
                const int32_t prop = -1 - node.a;
                int32_t val;
                if (prop < 16) val = p[prop * ls];
                else {
                    bool bad = false;
                    val = ref_property(prop, &bad);
                    if (bad) { es.set(br, E_TREC); done = true; return; }
                }
                node = node_at(val > node.b ? node.c : node.d);
            } while (node.a < 0);
        }
        int32_t val, predictor = node.b;
        if (MODE == 1 && fused) {
            const uint32_t b = (uint32_t) node.b;
            predictor = (int32_t) (b & 0xff);
            const uint64_t e = ans_tables[((size_t) ((b >> 8) & 0xff) << las) + ((cs.ans_state & 0xfff) >> cc.log_bucket)];
            const int32_t token = ans_symbol_entry(br, cs.ans_state, cc.log_bucket, e);
            val = hybrid_int(br, es, token, lane_leaf_cfg(b));
        } else if (MODE == 1) {
            const uint32_t ci = cc.cluster_map[node.a];
            const uint64_t e = ans_tables[((size_t) ci << las) + ((cs.ans_state & 0xfff) >> cc.log_bucket)];
            const HybridCfg cfg = cc.clusters[ci].cfg;
            const int32_t token = ans_symbol_entry(br, cs.ans_state, cc.log_bucket, e);
            val = hybrid_int(br, es, token, cfg);
        } else {
            val = code(br, es, cc, cs, node.a, dist_mult);
        }
        val = unpack_signed(val) * node.d + node.c;
        bool bad = false;
        val += mod_predict(predictor, pw, pn, pnw, pne, pnn, pww, pnee, wp_pred[4], &bad);
        if (bad) es.set(br, E_PRED);
        if (es.err) { done = true; return; }
        if ((uint32_t) (val + 32768) > 65535u) { es.set(br, E_POVF); done = true; return; }
        px[(size_t) y * (size_t) stride + x] = (int16_t) val;
        if (ringed) ring_row(y)[(size_t) x * ls] = (int16_t) val;
        prev2 = prev;
        prev = val;
        n_ww = n_w; n_w = n_c; n_c = n_e; n_e = n_ee;
        if (uses_wp) {
            // j40.h:4103-4111, then slide the error windows
            int32_t *err = err_row(y);
            const int32_t es_ = estride();
            const int32_t v8 = val * 8;
            for (int i = 0; i < 5; ++i) {
                const int32_t e = i < 4 ? (iabs(wp_pred[i] - v8) + 3) >> 3 : wp_pred[4] - v8;
                err[((size_t) x * 5 + i) * es_] = e;
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
