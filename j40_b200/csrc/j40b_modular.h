// j40-b200: modular sub-bitstream decoding (device side, also compiled for the CPU kernel-logic tests).
//
// Replaces the reference's
//   modular header      j40.h:3717-3845  (j40__modular_header; transforms limited to RCT on the device)
//   per-sample decode   j40.h:3965-4229  (init_neighbors, weighted predictor, predictors, tree walk)
//   inverse RCT         j40.h:4318-4400  (j40__inverse_rct16)
// Samples are 16-bit (the reference's Main-profile level-5 limits force modular_16bit_buffers,
// j40.h:1167, 3158), intermediates 32-bit, exactly as the reference's P=16 instantiation.
#pragma once
#include "j40b_entropy.h"

namespace j40b {

enum { MOD_MAX_CH = 8, MOD_MAX_TRANSFORMS = 8 };

struct ModChannel {
    int16_t *px;     // top-left sample of this (sub)image
    int32_t stride;  // in samples
    int32_t w, h;
    int32_t hshift, vshift;
};

// RCT: begin_c, type. Palette (kind 1; only in the global image's header, see modular_header): begin_c, num_c,
// nb_colours, nb_deltas, d_pred (j40.h:3762-3792)
struct ModTransform { int32_t begin_c, type; int32_t kind, num_c, nb_colours, nb_deltas, d_pred; };

struct ModImage {
    int32_t num_channels;
    ModChannel ch[MOD_MAX_CH];
    WPParams wp;
    int32_t nb_transforms;
    ModTransform tr[MOD_MAX_TRANSFORMS];
    int32_t dist_mult;
    int32_t nb_meta_channels; // palette channels at the front of the list (global image only)
};

// floor(2^24 / (i + 1)), i in [0, 64): the divisor table of the weighted predictor (j40.h:3905)
struct Div24Table {
    int32_t v[64];
    constexpr Div24Table() : v() { for (int i = 0; i < 64; ++i) v[i] = (int32_t) (0x1000000u / (uint32_t) (i + 1)); }
};
static constexpr Div24Table h_div24_table{};

struct WPState {
    int32_t *errors; // [2][width][5]
    int32_t width;
    int32_t pred[5];
    int32_t trueerrw, trueerrn, trueerrnw, trueerrne;
};

J40B_HD J40B_INLINE int32_t mod_gradient(int32_t w, int32_t n, int32_t nw) {
    int32_t lo = imin(w, n), hi = imax(w, n);
    return imin(imax(lo, w + n - nw), hi);
}

// j40.h:4011-4072. `div24` = the 64-entry divisor table (shared memory on the device).
J40B_HD J40B_INLINE void wp_before_predict(WPState &s, const WPParams &pr, const int32_t *div24, int32_t x, int32_t y,
                                           int32_t pw, int32_t pn, int32_t pnw, int32_t pne, int32_t pnn) {
    const int32_t *err = s.errors + (size_t) ((y & 1) ? s.width : 0) * 5;
    const int32_t *nerr = s.errors + (size_t) ((y & 1) ? 0 : s.width) * 5;
    int32_t ew[5], en[5], enw[5], ene[5], eww[4], ew2[4];
    for (int i = 0; i < 5; ++i) ew[i] = x > 0 ? err[(size_t) (x - 1) * 5 + i] : 0;
    for (int i = 0; i < 5; ++i) en[i] = y > 0 ? nerr[(size_t) x * 5 + i] : 0;
    for (int i = 0; i < 5; ++i) enw[i] = x > 0 && y > 0 ? nerr[(size_t) (x - 1) * 5 + i] : en[i];
    for (int i = 0; i < 5; ++i) ene[i] = x + 1 < s.width && y > 0 ? nerr[(size_t) (x + 1) * 5 + i] : en[i];
    for (int i = 0; i < 4; ++i) eww[i] = x > 1 ? err[(size_t) (x - 2) * 5 + i] : 0;
    for (int i = 0; i < 4; ++i) ew2[i] = x + 1 < s.width ? 0 : ew[i];
    s.trueerrw = ew[4];
    s.trueerrn = en[4];
    s.trueerrnw = enw[4];
    s.trueerrne = ene[4];
    s.pred[0] = (pw + pne - pn) * 8;
    s.pred[1] = pn * 8 - (((s.trueerrw + s.trueerrn + s.trueerrne) * pr.p1) >> 5);
    s.pred[2] = pw * 8 - (((s.trueerrw + s.trueerrn + s.trueerrnw) * pr.p2) >> 5);
    s.pred[3] = pn * 8 - ((s.trueerrnw * pr.p3[0] + s.trueerrn * pr.p3[1] + s.trueerrne * pr.p3[2] +
                           (pnn - pn) * 8 * pr.p3[3] + (pnw - pw) * 8 * pr.p3[4]) >> 5);
    int32_t w[4];
    for (int i = 0; i < 4; ++i) {
        int32_t errsum = en[i] + ew[i] + enw[i] + eww[i] + ene[i] + ew2[i];
        int32_t shift = imax(floor_lg32((uint32_t) errsum + 1) - 5, 0);
        w[i] = (int32_t) (4 + (((int64_t) pr.w[i] * div24[errsum >> shift]) >> shift));
    }
    int32_t logw = floor_lg32((uint32_t) (w[0] + w[1] + w[2] + w[3])) - 4;
    int32_t wsum = 0, sum = 0;
    for (int i = 0; i < 4; ++i) {
        w[i] >>= logw;
        wsum += w[i];
        sum += s.pred[i] * w[i];
    }
    s.pred[4] = (int32_t) ((((int64_t) sum + (wsum >> 1) - 1) * div24[wsum - 1]) >> 24);
    if (((s.trueerrn ^ s.trueerrw) | (s.trueerrn ^ s.trueerrnw)) <= 0) {
        int32_t lo = imin(pw, imin(pn, pne)) * 8;
        int32_t hi = imax(pw, imax(pn, pne)) * 8;
        s.pred[4] = imin(imax(lo, s.pred[4]), hi);
    }
}

// j40.h:4103-4111
J40B_HD J40B_INLINE void wp_after_predict(WPState &s, int32_t x, int32_t y, int32_t val) {
    int32_t *err = s.errors + ((size_t) ((y & 1) ? s.width : 0) + (size_t) x) * 5;
    for (int i = 0; i < 4; ++i) err[i] = (iabs(s.pred[i] - val * 8) + 3) >> 3;
    err[4] = s.pred[4] - val * 8;
}

// Copies the part of `tree` that channel `cidx` of stream `sidx` can reach: branches on the two
// static properties (0 = channel index, 1 = stream index, j40.h:4181-4182) are resolved now.
// Returns the number of nodes written (0 if `cap` is too small: use the full tree then) and whether
// any reachable node needs the weighted predictor (property 15 or predictor 6, j40.h:4142-4154).
J40B_HD inline int prune_tree(const DTreeNode *tree, int32_t cidx, int32_t sidx, DTreeNode *out, int cap, bool *uses_wp) {
    // breadth-first copy; out[i].c / .d first hold source indices, patched once the children are placed
    int n = 0;
    *uses_wp = false;
    if (cap < 1) return 0;
    int32_t src = 0;
    for (;;) { // resolve static branches at the root
        const DTreeNode &t = tree[src];
        if (t.a < 0 && (-1 - t.a) <= 1) { int32_t v = (-1 - t.a) == 0 ? cidx : sidx; src = v > t.b ? t.c : t.d; } else break;
    }
    out[n++] = tree[src];
    for (int i = 0; i < n; ++i) {
        DTreeNode &o = out[i];
        if (o.a >= 0) { if (o.b == 6) *uses_wp = true; continue; }
        if (-1 - o.a == 15) *uses_wp = true;
        int32_t child[2] = {o.c, o.d};
        for (int k = 0; k < 2; ++k) {
            int32_t s2 = child[k];
            for (;;) {
                const DTreeNode &t = tree[s2];
                if (t.a < 0 && (-1 - t.a) <= 1) { int32_t v = (-1 - t.a) == 0 ? cidx : sidx; s2 = v > t.b ? t.c : t.d; } else break;
            }
            if (n >= cap) return 0;
            out[n] = tree[s2];
            child[k] = n++;
        }
        o.c = child[0];
        o.d = child[1];
    }
    return n;
}

// Decodes channel `cidx` of `m` (j40.h:4127-4245). `wp_scratch` must hold 2*width*5 int32 when the
// (pruned) tree needs the weighted predictor; `tree` may be the pruned copy or the full tree.
template <bool USE_WP>
J40B_HD inline void modular_channel_t(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs,
                                      const DTreeNode *tree, int32_t *wp_scratch, const int32_t *div24,
                                      const ModImage &m, int32_t cidx, int32_t sidx) {
    const ModChannel &c = m.ch[cidx];
    const int32_t width = c.w, height = c.h, stride = c.stride;
    const int32_t dist_mult = m.dist_mult;
    const WPParams wpp = m.wp;
    WPState wp;
    wp.errors = wp_scratch;
    wp.width = width;
    for (int i = 0; i < 5; ++i) wp.pred[i] = 0;
    wp.trueerrw = wp.trueerrn = wp.trueerrnw = wp.trueerrne = 0;
    if (USE_WP) for (int32_t i = 0; i < width * 2 * 5; ++i) wp_scratch[i] = 0;
    // reference channels for properties >= 16: earlier channels of identical geometry, nearest first
    int32_t refcmap[MOD_MAX_CH], nref = 0;
    for (int32_t i = cidx - 1; i >= 0; --i) {
        const ModChannel &r = m.ch[i];
        if (c.w != r.w || c.h != r.h || c.hshift != r.hshift || c.vshift != r.vshift) continue;
        refcmap[nref++] = i;
    }
    for (int32_t y = 0; y < height; ++y) {
        int16_t *row = c.px + (size_t) y * (size_t) stride;
        int32_t prev = 0, prev2 = 0; // the two samples just decoded in this row (kept out of memory)
        for (int32_t x = 0; x < width; ++x) {
            const int16_t *p = row + x;
            int32_t pw = x > 0 ? prev : y > 0 ? p[-stride] : 0;
            int32_t pn = y > 0 ? p[-stride] : pw;
            int32_t pnw = x > 0 && y > 0 ? p[-1 - stride] : pw;
            int32_t pne = x + 1 < width && y > 0 ? p[1 - stride] : pn;
            int32_t pnn = y > 1 ? p[-2 * stride] : pn;
            int32_t pnee = x + 2 < width && y > 0 ? p[2 - stride] : pne;
            int32_t pww = x > 1 ? prev2 : pw;
            int32_t pnww = x > 1 && y > 0 ? p[-2 - stride] : pww;
            if (USE_WP) wp_before_predict(wp, wpp, div24, x, y, pw, pn, pnw, pne, pnn);

            const DTreeNode *n = tree;
            while (n->a < 0) {
                int32_t prop = -1 - n->a, val;
                switch (prop) {
                case 0: val = cidx; break;
                case 1: val = sidx; break;
                case 2: val = y; break;
                case 3: val = x; break;
                case 4: val = iabs(pn); break;
                case 5: val = iabs(pw); break;
                case 6: val = pn; break;
                case 7: val = pw; break;
                case 8: val = x > 0 ? pw - (pww + pnw - pnww) : pw; break;
                case 9: val = pw + pn - pnw; break;
                case 10: val = pw - pnw; break;
                case 11: val = pnw - pn; break;
                case 12: val = pn - pne; break;
                case 13: val = pn - pnn; break;
                case 14: val = pw - pww; break;
                case 15:
                    val = wp.trueerrw;
                    if (iabs(val) < iabs(wp.trueerrn)) val = wp.trueerrn;
                    if (iabs(val) < iabs(wp.trueerrnw)) val = wp.trueerrnw;
                    if (iabs(val) < iabs(wp.trueerrne)) val = wp.trueerrne;
                    break;
                default: {
                    int32_t refcidx = (prop - 16) / 4;
                    if (refcidx >= nref) { es.set(br, E_TREC); return; }
                    const ModChannel &r = m.ch[refcmap[refcidx]];
                    const int16_t *rp = r.px + (size_t) y * (size_t) r.stride + x;
                    val = rp[0];
                    if (prop & 2) {
                        int32_t rw = x > 0 ? rp[-1] : 0;
                        int32_t rn = y > 0 ? rp[-r.stride] : rw;
                        int32_t rnw = x > 0 && y > 0 ? rp[-1 - r.stride] : rw;
                        val -= mod_gradient(rw, rn, rnw);
                    }
                    if (prop & 1) val = iabs(val);
                    break;
                }
                }
                n = tree + (val > n->b ? n->c : n->d);
            }

            int32_t val = code(br, es, cc, cs, n->a, dist_mult);
            val = unpack_signed(val) * n->d + n->c;
            int32_t pred;
            switch (n->b) {
            case 0: pred = 0; break;
            case 1: pred = pw; break;
            case 2: pred = pn; break;
            case 3: pred = (pw + pn) / 2; break;
            case 4: pred = iabs(pn - pnw) < iabs(pw - pnw) ? pw : pn; break;
            case 5: pred = mod_gradient(pw, pn, pnw); break;
            case 6: pred = (wp.pred[4] + 3) >> 3; break;
            case 7: pred = pne; break;
            case 8: pred = pnw; break;
            case 9: pred = pww; break;
            case 10: pred = (pw + pnw) / 2; break;
            case 11: pred = (pn + pnw) / 2; break;
            case 12: pred = (pn + pne) / 2; break;
            case 13: pred = (6 * pn - 2 * pnn + 7 * pw + pww + pnee + 3 * pne + 8) / 16; break;
            default: es.set(br, E_PRED); return;
            }
            val += pred;
            if (es.err) return;
            if (val < -32768 || val > 32767) { es.set(br, E_POVF); return; }
            row[x] = (int16_t) val;
            prev2 = prev;
            prev = val;
            if (USE_WP) wp_after_predict(wp, x, y, val);
        }
    }
}

// `ptree`/`ptree_cap`: scratch for the pruned tree (shared memory on the device); 0 = use the full tree
J40B_HD inline void modular_channel(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs,
                                    const DTreeNode *tree, bool tree_uses_wp, int32_t *wp_scratch, const int32_t *div24,
                                    DTreeNode *ptree, int ptree_cap,
                                    const ModImage &m, int32_t cidx, int32_t sidx) {
    const ModChannel &c = m.ch[cidx];
    if (c.w <= 0 || c.h <= 0) return;
    bool uses_wp = tree_uses_wp;
    const DTreeNode *t = tree;
    if (ptree_cap > 0 && prune_tree(tree, cidx, sidx, ptree, ptree_cap, &uses_wp) > 0) t = ptree;
    else uses_wp = tree_uses_wp;
    if (uses_wp) modular_channel_t<true>(br, es, cc, cs, t, wp_scratch, div24, m, cidx, sidx);
    else modular_channel_t<false>(br, es, cc, cs, t, wp_scratch, div24, m, cidx, sidx);
}

// ---------------------------------------------------------------------------------------------
// Warp variant ("SIMT-uniform" decoding). The serial decoder is executed by ALL lanes of a warp in lockstep on
// identical state (bit reader, ANS state, neighbours): redundant uniform work costs no extra issue slots, and it
// lets the lanes split the parts of a sample that are parallel:
//   * the MA tree is evaluated without walking it: lane j computes the decision of inner node j (every
//     property is a small linear form of the neighbours, see SimtLane), one ballot collects the decision bits
//     and a second one finds the leaf whose path matches (lane j also owns leaf j);
//   * the weighted predictor's four sub-predictor weights / error updates are computed by lanes 0..3 (+4 for
//     the signed true error), gathered with shuffles;
//   * at the start of each row all lanes compute the row's reference-channel properties (they depend on
//     finished channels only) and the sample rows / error rows live in shared memory.
// On the CPU (kernel-logic tests) the same code runs with one "lane" and loops over the lane-indexed parts.
// Same integer arithmetic as modular_channel_t (j40.h:3965-4229).
enum { SIMT_LANES = 32, SIMT_REF_SLOTS = 2 }; // (trees with more distinct reference properties are walked instead)

struct SimtLane {       // lane j: decision node j and leaf j of the compiled (pruned) tree
    int32_t c[7];       // coefficients of pn, pw, pnw, pne, pnn, pww, pnww
    int32_t cx, cy, ce; // coefficients of x, y and the weighted predictor's max true error (property 15)
    int32_t refslot;    // >= 0: the value is refp[refslot][x] (property >= 16)
    int32_t flags;      // bit 0: absolute value; bit 1: property 8 (is pw at x == 0); bit 2: node present
    int32_t thr;        // decision: value > thr
    uint32_t care, want; // leaf j is selected iff (decisions & care) == want
    int32_t leaf_node;  // index of leaf j in the pruned tree, -1 if none
};

struct alignas(16) SimtLeaf { // leaf j with its context already resolved to a cluster
    DCluster cl;
    int32_t predictor, offset, multiplier, pad;
};

struct ModSmem {
    int16_t *rows;   // [3][cap]  ring of the last three sample rows (null: no SIMT path)
    int32_t *wp;     // [cap][5] weighted predictor error row (updated in place, see modular_channel_simt); refp follows
    int32_t *refp;   // [SIMT_REF_SLOTS][cap] reference-channel property values of the current row
    SimtLane *tab;   // [SIMT_LANES]
    SimtLeaf *leaves; // [SIMT_LANES]
    int32_t *info;   // [8] scratch flags shared by the lanes
    int32_t cap;     // widest channel the SIMT path can take
    uint64_t *tabs;  // room for copies of the leaves' alias tables (device; null: read them through L1)
    int32_t tabs_entries;
    int32_t lanes;   // lanes that decode this stream together (32, or 16 / 8 when a warp is shared by 2 / 4 streams): the
                     // compiled tree may have that many inner nodes and leaves
};

// Compiles the pruned tree into per-lane decision forms. Returns false if it does not fit (more than 32 inner
// nodes or leaves, more than SIMT_REF_SLOTS distinct reference properties, a reference property without its
// channel): the caller then walks the tree instead. refprops[s] = property number of slot s.
J40B_HD inline bool simt_compile_tree(const DTreeNode *t, int n, int nref, const CodeCtx &cc, SimtLane *tab, SimtLeaf *leaves,
                                      int32_t *refprops, int32_t *nslots, int max_lanes = SIMT_LANES) {
    int inner_of[192], ni = 0, nl = 0, ns = 0;
    if (n <= 0 || n > 192) return false;
    for (int i = 0; i < SIMT_LANES; ++i) {
        SimtLane &L = tab[i];
        for (int k = 0; k < 7; ++k) L.c[k] = 0;
        L.cx = L.cy = L.ce = 0; L.refslot = -1; L.flags = 0; L.thr = 0; L.care = 0; L.want = 0xffffffffu; L.leaf_node = -1;
    }
    for (int i = 0; i < n; ++i) {
        inner_of[i] = -1;
        if (t[i].a >= 0) { if (++nl > max_lanes) return false; continue; }
        if (ni >= max_lanes) return false;
        inner_of[i] = ni;
        SimtLane &L = tab[ni++];
        const int prop = -1 - t[i].a;
        L.flags = 4;
        L.thr = t[i].b;
        // neighbours: 0 pn, 1 pw, 2 pnw, 3 pne, 4 pnn, 5 pww, 6 pnww
        switch (prop) {
        case 2: L.cy = 1; break;
        case 3: L.cx = 1; break;
        case 4: L.c[0] = 1; L.flags |= 1; break;
        case 5: L.c[1] = 1; L.flags |= 1; break;
        case 6: L.c[0] = 1; break;
        case 7: L.c[1] = 1; break;
        case 8: L.c[1] = 1; L.c[5] = -1; L.c[2] = -1; L.c[6] = 1; L.flags |= 2; break;
        case 9: L.c[1] = 1; L.c[0] = 1; L.c[2] = -1; break;
        case 10: L.c[1] = 1; L.c[2] = -1; break;
        case 11: L.c[2] = 1; L.c[0] = -1; break;
        case 12: L.c[0] = 1; L.c[3] = -1; break;
        case 13: L.c[0] = 1; L.c[4] = -1; break;
        case 14: L.c[1] = 1; L.c[5] = -1; break;
        case 15: L.ce = 1; break;
        default: {
            if (prop < 16 || (prop - 16) / 4 >= nref) return false; // (static properties 0/1 were pruned away)
            int slot = -1;
            for (int k = 0; k < ns; ++k) if (refprops[k] == prop) slot = k;
            if (slot < 0) { if (ns >= SIMT_REF_SLOTS) return false; refprops[ns] = prop; slot = ns++; }
            L.refslot = slot;
            break;
        }
        }
    }
    // leaves: path masks by walking down from the root with an explicit stack
    int32_t stack_node[64]; uint32_t stack_care[64], stack_want[64];
    int sp = 0, leaf = 0;
    stack_node[sp] = 0; stack_care[sp] = 0; stack_want[sp] = 0; ++sp;
    while (sp > 0) {
        --sp;
        int i = stack_node[sp]; uint32_t care = stack_care[sp], want = stack_want[sp];
        if (t[i].a >= 0) {
            tab[leaf].care = care; tab[leaf].want = want; tab[leaf].leaf_node = i;
            SimtLeaf &lf = leaves[leaf];
            lf.cl = cc.clusters[cc.cluster_map[t[i].a]];
            lf.predictor = t[i].b; lf.offset = t[i].c; lf.multiplier = t[i].d; lf.pad = 0;
            ++leaf;
            continue;
        }
        if (sp + 2 > 64) return false;
        uint32_t bit = 1u << inner_of[i];
        stack_node[sp] = t[i].c; stack_care[sp] = care | bit; stack_want[sp] = want | bit; ++sp; // value > threshold
        stack_node[sp] = t[i].d; stack_care[sp] = care | bit; stack_want[sp] = want; ++sp;
    }
    *nslots = ns;
    return true;
}

// `xs` = x relative to the start of the current segment of the row (refp holds one segment).
// INTERIOR: x >= 2, so the x == 0 special case of property 8 cannot apply. The sum is written as independent
// pairs (integer addition wraps, so the grouping does not change the value) to keep the dependent chain short,
// and the reference-property value is loaded unconditionally and selected (lanes differ in `refslot`; a branch
// here would diverge on every sample).
template <bool INTERIOR>
J40B_HD J40B_INLINE bool simt_decision(const SimtLane &L, int32_t x, int32_t xs, int32_t y, int32_t pn, int32_t pw, int32_t pnw, int32_t pne,
                                       int32_t pnn, int32_t pww, int32_t pnww, int32_t maxerr, const int32_t *refp, int32_t cap) {
    const int32_t s0 = L.c[0] * pn + L.c[1] * pw, s1 = L.c[2] * pnw + L.c[3] * pne, s2 = L.c[4] * pnn + L.c[5] * pww;
    const int32_t s3 = L.c[6] * pnww + L.cx * x, s4 = L.cy * y + L.ce * maxerr;
    int32_t v = ((s0 + s1) + (s2 + s3)) + s4;
    const int32_t rv = refp[(size_t) (L.refslot >= 0 ? L.refslot : 0) * cap + xs];
    if (L.refslot >= 0) v = rv;
    if (!INTERIOR && (L.flags & 2) && x == 0) v = pw;
    if (L.flags & 1) v = iabs(v);
    return (L.flags & 4) && v > L.thr;
}

// index (into ModSmem::leaves) of the leaf the current sample falls into
template <bool INTERIOR, bool FULL>
J40B_HD J40B_INLINE int32_t simt_tree_leaf(const SimtLane *tab, const SimtLane &mine, int32_t x, int32_t xs, int32_t y, int32_t pn, int32_t pw,
                                           int32_t pnw, int32_t pne, int32_t pnn, int32_t pww, int32_t pnww, int32_t maxerr,
                                           const int32_t *refp, int32_t cap, uint32_t gmask, int32_t gshift) {
#ifdef __CUDA_ARCH__
    // A warp shared by several streams (!FULL): the lanes of a group only ever need each other, and they are always
    // together -- every branch of this decoder is uniform within a group -- so they vote and shuffle among whoever is
    // converged with them right now (__activemask(): a plain VOTE / SHFL; naming the group's mask instead makes the
    // compiler guard every collective with a MATCH.ANY convergence check, 5 instructions each). The groups execute as one
    // instruction stream while their control flow agrees, and one after the other where it does not.
    const uint32_t am = FULL ? 0xffffffffu : __activemask();
    const uint32_t dec = (__ballot_sync(am, simt_decision<INTERIOR>(mine, x, xs, y, pn, pw, pnw, pne, pnn, pww, pnww, maxerr, refp, cap)) & gmask) >> gshift;
    const uint32_t hit = (__ballot_sync(am, mine.leaf_node >= 0 && (dec & mine.care) == mine.want) & gmask) >> gshift;
    return __ffs((int) hit) - 1;
#else
    uint32_t dec = 0;
    for (int j = 0; j < SIMT_LANES; ++j) dec |= (uint32_t) simt_decision<INTERIOR>(tab[j], x, xs, y, pn, pw, pnw, pne, pnn, pww, pnww, maxerr, refp, cap) << j;
    for (int j = 0; j < SIMT_LANES; ++j) if (tab[j].leaf_node >= 0 && (dec & tab[j].care) == tab[j].want) return j;
    return -1;
#endif
}

// Weighted predictor with the lane-indexed parts split out. NL = lane-indexed values this thread holds:
// 1 on the device (lane i = min(lane, 4) holds index i), 5 on the CPU (all of them).
#ifdef __CUDA_ARCH__
enum { WP_NL = 1 };
#else
enum { WP_NL = 5 };
#endif
struct WpSimt {
    int32_t e_n[WP_NL], e_nw[WP_NL], e_ne[WP_NL], e_w[WP_NL], e_ww[WP_NL]; // error window, index i per lane
    int32_t te_w, te_n, te_nw, te_ne;                                       // true errors (index 4), uniform
    int32_t pred[5];
};

// everything one sample of the SIMT path reads and updates (lives in registers; pointers are shared memory)
struct SimtCtx {
    const SimtLane *tab; const SimtLeaf *leaves; const int32_t *refp; const int32_t *div24;
    int32_t cap, width, y, seg0, dist_mult, my_i;
    uint32_t gmask; int32_t gshift; // the lanes that decode this stream: their bits in the warp, the first one's number
    // device, per lane: the alias table of "my" leaf's cluster (any valid table for lanes without a leaf), read one
    // sample ahead of the tree decision; masks selecting this lane's weighted sub-predictor without branches
    const uint64_t *my_table;
    int32_t m0, m1, m2, m4;
    int16_t *cur; const int16_t *nrow, *nnrow;
    int32_t *err; const int32_t *nerr;
    int32_t prev, prev2, n_ww, n_w, n_c, n_e; // the two samples to the left; sliding window over the row above
    WPParams wpp;
    WpSimt wp;
};

J40B_HD J40B_INLINE int32_t mod_predict(int32_t predictor, int32_t pw, int32_t pn, int32_t pnw, int32_t pne, int32_t pnn,
                                        int32_t pww, int32_t pnee, int32_t wp_pred, bool *bad) {
    switch (predictor) {
    case 0: return 0;
    case 1: return pw;
    case 2: return pn;
    case 3: return (pw + pn) / 2;
    case 4: return iabs(pn - pnw) < iabs(pw - pnw) ? pw : pn;
    case 5: return mod_gradient(pw, pn, pnw);
    case 6: return (wp_pred + 3) >> 3;
    case 7: return pne;
    case 8: return pnw;
    case 9: return pww;
    case 10: return (pw + pnw) / 2;
    case 11: return (pn + pnw) / 2;
    case 12: return (pn + pne) / 2;
    case 13: return (6 * pn - 2 * pnn + 7 * pw + pww + pnee + 3 * pne + 8) / 16;
    default: *bad = true; return 0;
    }
}

// One sample. PRED >= 0: every leaf of the compiled tree uses that predictor (no dispatch); MODE: see
// code_cluster; INTERIOR: y >= 2 and 2 <= x < width - 2, so no neighbour falls off the image.
// Returns false on error (uniform across the lanes).
template <bool USE_WP, int PRED, int MODE, bool INTERIOR, bool FULL>
J40B_HD J40B_INLINE bool simt_sample(SimtCtx &S, const SimtLane &mine, BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs, int32_t x) {
    const int32_t y = S.y, width = S.width;
    const int32_t n_ee = INTERIOR ? S.nrow[x + 2] : (y > 0 && x + 2 < width ? S.nrow[x + 2] : S.n_e);
    const int32_t pw = INTERIOR ? S.prev : (x > 0 ? S.prev : S.n_c);   // (n_c is 0 in the first row)
    const int32_t pn = INTERIOR ? S.n_c : (y > 0 ? S.n_c : pw);
    const int32_t pnw = INTERIOR ? S.n_w : (x > 0 && y > 0 ? S.n_w : pw);
    const int32_t pne = INTERIOR ? S.n_e : (y > 0 ? S.n_e : pn);        // n_e already equals n_c at the right edge
    const int32_t pnn = INTERIOR ? S.nnrow[x] : (y > 1 ? S.nnrow[x] : pn);
    const int32_t pnee = INTERIOR ? n_ee : (y > 0 ? n_ee : pne);
    const int32_t pww = INTERIOR ? S.prev2 : (x > 1 ? S.prev2 : pw);
    const int32_t pnww = INTERIOR ? S.n_ww : (x > 1 && y > 0 ? S.n_ww : pww);
    WpSimt &wp = S.wp;
    int32_t maxerr = 0;
#ifdef __CUDA_ARCH__
    // MODE 1 (rANS, no LZ77): the bucket of the next symbol depends on the state alone, which cluster's table it
    // is looked up in depends on the leaf. Every leaf lane fetches the entry of its own cluster now; the load
    // completes under the predictor and tree arithmetic, and the lane of the chosen leaf hands it over by shuffle.
    uint64_t e_mine = 0;
    if (MODE == 1) e_mine = S.my_table[(cs.ans_state & 0xfff) >> cc.log_bucket];
#endif
    if (USE_WP) {
        // j40.h:4011-4072
        const WPParams &wpp = S.wpp;
        wp.pred[0] = (pw + pne - pn) * 8;
        wp.pred[1] = pn * 8 - (((wp.te_w + wp.te_n + wp.te_ne) * wpp.p1) >> 5);
        wp.pred[2] = pw * 8 - (((wp.te_w + wp.te_n + wp.te_nw) * wpp.p2) >> 5);
        wp.pred[3] = pn * 8 - ((wp.te_nw * wpp.p3[0] + wp.te_n * wpp.p3[1] + wp.te_ne * wpp.p3[2] +
                                (pnn - pn) * 8 * wpp.p3[3] + (pnw - pw) * 8 * wpp.p3[4]) >> 5);
        int32_t w[4];
#ifdef __CUDA_ARCH__
        {
            const int i = S.my_i & 3;
            int32_t errsum = wp.e_n[0] + wp.e_w[0] + wp.e_nw[0] + wp.e_ww[0] + wp.e_ne[0] + (INTERIOR || x + 1 < width ? 0 : wp.e_w[0]);
            int32_t shift = imax(floor_lg32((uint32_t) errsum + 1) - 5, 0);
            int32_t wi = (int32_t) (4 + (((int64_t) wpp.w[i] * S.div24[errsum >> shift]) >> shift));
            const uint32_t am = FULL ? 0xffffffffu : __activemask();
            for (int k = 0; k < 4; ++k) w[k] = __shfl_sync(am, wi, S.gshift + k);
        }
#else
        for (int i = 0; i < 4; ++i) {
            int32_t errsum = wp.e_n[i] + wp.e_w[i] + wp.e_nw[i] + wp.e_ww[i] + wp.e_ne[i] + (INTERIOR || x + 1 < width ? 0 : wp.e_w[i]);
            int32_t shift = imax(floor_lg32((uint32_t) errsum + 1) - 5, 0);
            w[i] = (int32_t) (4 + (((int64_t) wpp.w[i] * S.div24[errsum >> shift]) >> shift));
        }
#endif
        int32_t logw = floor_lg32((uint32_t) (w[0] + w[1] + w[2] + w[3])) - 4;
        int32_t wsum = 0, sum = 0;
        for (int i = 0; i < 4; ++i) {
            w[i] >>= logw;
            wsum += w[i];
            sum += wp.pred[i] * w[i];
        }
        wp.pred[4] = (int32_t) ((((int64_t) sum + (wsum >> 1) - 1) * S.div24[wsum - 1]) >> 24);
        if (((wp.te_n ^ wp.te_w) | (wp.te_n ^ wp.te_nw)) <= 0) {
            int32_t lo = imin(pw, imin(pn, pne)) * 8;
            int32_t hi = imax(pw, imax(pn, pne)) * 8;
            wp.pred[4] = imin(imax(lo, wp.pred[4]), hi);
        }
        maxerr = wp.te_w;
        if (iabs(maxerr) < iabs(wp.te_n)) maxerr = wp.te_n;
        if (iabs(maxerr) < iabs(wp.te_nw)) maxerr = wp.te_nw;
        if (iabs(maxerr) < iabs(wp.te_ne)) maxerr = wp.te_ne;
    }

    const int32_t li = simt_tree_leaf<INTERIOR, FULL>(S.tab, mine, x, x - S.seg0, y, pn, pw, pnw, pne, pnn, pww, pnww, maxerr, S.refp, S.cap, S.gmask, S.gshift);
    const SimtLeaf leaf = S.leaves[li];
    int32_t val;
#ifdef __CUDA_ARCH__
    if (MODE == 1) {
        const uint32_t am = FULL ? 0xffffffffu : __activemask();
        const uint32_t elo = __shfl_sync(am, (uint32_t) e_mine, S.gshift + li), ehi = __shfl_sync(am, (uint32_t) (e_mine >> 32), S.gshift + li);
        val = ans_symbol_entry(br, cs.ans_state, cc.log_bucket, (uint64_t) ehi << 32 | elo);
        val = hybrid_int(br, es, val, leaf.cl.cfg);
        if (es.err) val = 0;
    } else
#endif
    if (MODE == 1 || !code_copy(cs, val)) val = code_cluster<false, MODE>(br, es, cc, cs, leaf.cl, S.dist_mult);
    val = unpack_signed(val) * leaf.multiplier + leaf.offset;
    bool bad = false;
    val += mod_predict(PRED >= 0 ? PRED : leaf.predictor, pw, pn, pnw, pne, pnn, pww, pnee, USE_WP ? wp.pred[4] : 0, &bad);
    if (bad) es.set(br, E_PRED);
    if (es.err) return false; // uniform: every lane sees the same error
    if ((uint32_t) (val + 32768) > 65535u) { es.set(br, E_POVF); return false; }
    S.cur[x] = (int16_t) val;
    S.prev2 = S.prev;
    S.prev = val;
    S.n_ww = S.n_w; S.n_w = S.n_c; S.n_c = S.n_e; S.n_e = n_ee;
    if (USE_WP) {
        // j40.h:4103-4111, then slide the error windows
        const int32_t v8 = val * 8;
        const int32_t te = wp.pred[4] - v8;
        for (int k = 0; k < WP_NL; ++k) {
            const int i = WP_NL == 1 ? S.my_i : k;
#ifdef __CUDA_ARCH__
            // this lane's sub-predictor by masks (lane constants) instead of a three-way branch on the lane index
            const int32_t p3 = wp.pred[3];
            const int32_t psel = p3 ^ ((wp.pred[0] ^ p3) & S.m0) ^ ((wp.pred[1] ^ p3) & S.m1) ^ ((wp.pred[2] ^ p3) & S.m2);
            const int32_t esub = (iabs(psel - v8) + 3) >> 3;
            const int32_t e = (esub & ~S.m4) | (te & S.m4);
#else
            const int32_t psel = i == 0 ? wp.pred[0] : i == 1 ? wp.pred[1] : i == 2 ? wp.pred[2] : wp.pred[3];
            const int32_t e = i < 4 ? (iabs(psel - v8) + 3) >> 3 : te;
#endif
            S.err[(size_t) x * 5 + i] = e;
            wp.e_ww[k] = wp.e_w[k];
            wp.e_w[k] = e;
            wp.e_nw[k] = wp.e_n[k];
            wp.e_n[k] = wp.e_ne[k];
            wp.e_ne[k] = INTERIOR || (y > 0 && x + 2 < width) ? S.nerr[(size_t) (x + 2) * 5 + i] : wp.e_ne[k];
        }
        wp.te_w = te;
        wp.te_nw = wp.te_n;
        wp.te_n = wp.te_ne;
        wp.te_ne = INTERIOR || (y > 0 && x + 2 < width) ? S.nerr[(size_t) (x + 2) * 5 + 4] : wp.te_ne;
    }
    return true;
}

// WIDE: the channel is wider than the shared-memory rows (HF metadata's block-info channel: 2 x up to 65536).
// Sample and error rows then stay in global memory (read through L1) and the reference properties are
// computed one segment of `cap` samples at a time.
template <bool USE_WP, int PRED, int MODE, bool WIDE, class Sync>
J40B_HD inline void modular_channel_simt(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs,
                                         const ModSmem &ms, int32_t *wp_scratch, const int32_t *refprops, int nslots,
                                         const int32_t *div24, const ModImage &m, int32_t cidx,
                                         int lane, int nlanes, Sync sync) {
    const ModChannel &c = m.ch[cidx];
    const int32_t width = c.w, height = c.h, stride = c.stride, cap = ms.cap;
    int32_t refcmap[MOD_MAX_CH], nref = 0;
    for (int32_t i = cidx - 1; i >= 0; --i) {
        const ModChannel &r = m.ch[i];
        if (c.w != r.w || c.h != r.h || c.hshift != r.hshift || c.vshift != r.vshift) continue;
        refcmap[nref++] = i;
    }
    const SimtLane mine = ms.tab[lane & (SIMT_LANES - 1)];
    SimtCtx S;
    S.tab = ms.tab; S.leaves = ms.leaves; S.refp = ms.refp; S.div24 = div24;
    S.cap = cap; S.width = width; S.dist_mult = m.dist_mult;
    S.gmask = sync.mask(); S.gshift = sync.shift();
    S.my_i = lane < 4 ? lane : 4; // this lane's weighted-predictor index (device)
    S.m0 = -(int32_t) (S.my_i == 0); S.m1 = -(int32_t) (S.my_i == 1); S.m2 = -(int32_t) (S.my_i == 2); S.m4 = -(int32_t) (S.my_i == 4);
    S.my_table = (const uint64_t *) (cc.arena + ms.leaves[mine.leaf_node >= 0 ? (lane & (SIMT_LANES - 1)) : 0].cl.table_off);
#ifdef __CUDA_ARCH__
    if (MODE == 1 && ms.tabs) {
        // The alias tables of this channel's leaves (leaf j's in slot j, as far as they fit) into the stream's slice of shared
        // memory: one entry is fetched per sample on the critical path, and next to a dozen other streams and the tile
        // kernels of other batches an SM has too little L1 left to keep the tables of all of them (measured: hit rate 98 %
        // with one stream per SM, 46-55 % with four)
        const int tlen = 1 << (12 - cc.log_bucket);
        for (int j = 0; j < ms.lanes && ms.tab[j].leaf_node >= 0 && (j + 1) * tlen <= ms.tabs_entries; ++j) {
            const uint64_t *src = (const uint64_t *) (cc.arena + ms.leaves[j].cl.table_off);
            for (int e = lane; e < tlen; e += nlanes) ms.tabs[j * tlen + e] = src[e];
        }
        sync();
        const int mj = mine.leaf_node >= 0 ? (lane & (SIMT_LANES - 1)) : 0;
        if ((mj + 1) * tlen <= ms.tabs_entries) S.my_table = ms.tabs + mj * tlen;
    }
#endif
    S.wpp = m.wp;
    WpSimt &wp = S.wp;
    // the first call below always reads a symbol (or continues an LZ77 copy that an earlier channel began):
    // seed the rANS state here so that the per-sample path does not have to test for it
    if (!cc.prefix && cs.ans_state == 0 && cs.num_to_copy <= 0) ans_seed(br, cs.ans_state);
    for (int32_t y = 0; y < height; ++y) {
        // ---- all lanes: the finished previous row goes to global memory, coalesced
        if (!WIDE && y > 0) {
            const int16_t *done = ms.rows + (size_t) ((y + 2) % 3) * cap;
            int16_t *gdst = c.px + (size_t) (y - 1) * (size_t) stride;
            for (int32_t x = lane; x < width; x += nlanes) gdst[x] = done[x];
        }
        S.y = y;
        if (WIDE) {
            S.cur = c.px + (size_t) y * (size_t) stride;
            S.nrow = S.cur - (y > 0 ? stride : 0);
            S.nnrow = S.cur - (y > 1 ? 2 * stride : 0);
            S.err = wp_scratch + (size_t) ((y & 1) ? width : 0) * 5;
            S.nerr = wp_scratch + (size_t) ((y & 1) ? 0 : width) * 5;
        } else {
            S.cur = ms.rows + (size_t) (y % 3) * cap;
            S.nrow = ms.rows + (size_t) ((y + 2) % 3) * cap;
            S.nnrow = ms.rows + (size_t) ((y + 1) % 3) * cap;
            // one row serves as both the previous and the current row: sample x writes its errors at x and has read
            // the previous row's up to x + 2 by then (halves the weighted predictor's shared memory)
            S.err = ms.wp;
            S.nerr = ms.wp;
        }
        for (int32_t seg0 = 0; seg0 < width; seg0 += cap) {
            const int32_t seg1 = seg0 + cap < width ? seg0 + cap : width;
            // ---- all lanes: this segment's reference-channel property values
            if (seg0 > 0) sync();
            for (int sl = 0; sl < nslots; ++sl) {
                const int prop = refprops[sl];
                const ModChannel &r = m.ch[refcmap[(prop - 16) / 4]];
                const int16_t *rrow = r.px + (size_t) y * (size_t) r.stride;
                int32_t *dst = ms.refp + (size_t) sl * cap;
                for (int32_t x = seg0 + lane; x < seg1; x += nlanes) {
                    int32_t val = rrow[x];
                    if (prop & 2) {
                        int32_t rw = x > 0 ? rrow[x - 1] : 0;
                        int32_t rn = y > 0 ? rrow[x - r.stride] : rw;
                        int32_t rnw = x > 0 && y > 0 ? rrow[x - 1 - r.stride] : rw;
                        val -= mod_gradient(rw, rn, rnw);
                    }
                    if (prop & 1) val = iabs(val);
                    dst[x - seg0] = val;
                }
            }
            sync(); // also orders the previous row's sample / error stores before this row's loads
            S.seg0 = seg0;
            if (seg0 == 0) {
                S.prev = S.prev2 = 0;
                S.n_ww = S.n_w = 0;
                S.n_c = y > 0 ? S.nrow[0] : 0;
                S.n_e = y > 0 && width > 1 ? S.nrow[1] : S.n_c;
                if (USE_WP) {
                    for (int k = 0; k < WP_NL; ++k) {
                        const int i = WP_NL == 1 ? S.my_i : k;
                        wp.e_w[k] = wp.e_ww[k] = 0;
                        wp.e_n[k] = y > 0 ? S.nerr[i] : 0;
                        wp.e_nw[k] = wp.e_n[k];
                        wp.e_ne[k] = y > 0 && width > 1 ? S.nerr[5 + i] : wp.e_n[k];
                    }
                    wp.te_w = 0;
                    wp.te_n = y > 0 ? S.nerr[4] : 0;
                    wp.te_nw = wp.te_n;
                    wp.te_ne = y > 0 && width > 1 ? S.nerr[5 + 4] : wp.te_n;
                }
            }
            int32_t x = seg0;
            if (y >= 2 && width >= 5) {
                const int32_t e0 = seg1 < 2 ? seg1 : 2, e1 = seg1 < width - 2 ? seg1 : width - 2;
                for (; x < e0; ++x) if (!simt_sample<USE_WP, PRED, MODE, false, Sync::kFull>(S, mine, br, es, cc, cs, x)) return;
                for (; x < e1; ++x) if (!simt_sample<USE_WP, PRED, MODE, true, Sync::kFull>(S, mine, br, es, cc, cs, x)) return;
            }
            for (; x < seg1; ++x) if (!simt_sample<USE_WP, PRED, MODE, false, Sync::kFull>(S, mine, br, es, cc, cs, x)) return;
        }
    }
    sync();
    if (!WIDE) {
        const int16_t *done = ms.rows + (size_t) ((height - 1) % 3) * cap;
        int16_t *gdst = c.px + (size_t) (height - 1) * (size_t) stride;
        for (int32_t x = lane; x < width; x += nlanes) gdst[x] = done[x];
    }
    sync();
}

// One channel of a sub-bitstream, warp-cooperative, in two steps so that kernels can be specialised by decoder variant
// (a kernel that holds all variants needs the registers of the greediest one plus what the allocator loses on the
// dispatch: 168 / 236 registers for the LF-group kernels, against 64 ... 124 for a single variant):
//   modular_channel_prep   lane 0 prunes the tree for (channel, stream) and compiles it; returns the class of decoder
//                          the channel needs (uniform across the lanes)
//   modular_channel_run<K> runs it if the class is K (K = MC_ANY: whatever it is)
// Classes: MC_WP  weighted predictor in every leaf, rANS without LZ77, rows in shared memory (the LF image's luma)
//          MC_GRAD gradient predictor in every leaf, no weighted-predictor property, same coding (its chroma, typically)
//          MC_WIDE wider than the shared-memory rows, no weighted predictor (the block-info channel of the HF metadata)
//          MC_GEN  any predictors, no weighted predictor, rANS without LZ77, rows in shared memory (chroma with mixed
//                  predictors, the sharpness and colour-correlation maps)
//          MC_REST everything else (weighted predictor among others, prefix codes, LZ77, trees that do not compile)
enum { MC_NONE = 0, MC_WP = 1, MC_GRAD = 2, MC_WIDE = 3, MC_REST = 4, MC_GEN = 5, MC_ANY = 6 };

template <class Sync>
J40B_HD inline int modular_channel_prep(const CodeCtx &cc, const DTreeNode *tree, bool tree_uses_wp, int32_t *wp_scratch,
                                        DTreeNode *ptree, int ptree_cap, const ModSmem &ms,
                                        const ModImage &m, int32_t cidx, int32_t sidx, int lane, Sync sync) {
    const ModChannel &c = m.ch[cidx];
    if (c.w <= 0 || c.h <= 0) return MC_NONE;
    sync(); // the previous channel's readers of ptree / tab are done
    if (lane == 0) {
        bool uses_wp = tree_uses_wp;
        int n = ptree_cap > 0 ? prune_tree(tree, cidx, sidx, ptree, ptree_cap, &uses_wp) : 0;
        if (n == 0) uses_wp = tree_uses_wp;
        int nref = 0;
        for (int32_t i = cidx - 1; i >= 0; --i) {
            const ModChannel &r = m.ch[i];
            if (c.w == r.w && c.h == r.h && c.hshift == r.hshift && c.vshift == r.vshift) ++nref;
        }
        int32_t nslots = 0;
        bool simt = ms.rows && n > 0 && (c.w <= ms.cap || !uses_wp || wp_scratch) && simt_compile_tree(ptree, n, nref, cc, ms.tab, ms.leaves, ms.info + 4, &nslots, ms.lanes);
        // fast variants: rANS without LZ77 and one predictor (gradient or weighted) shared by all leaves
        int variant = 0;
        if (simt && !cc.prefix && !cc.lz77) {
            int common = -2;
            for (int i = 0; i < n; ++i) if (ptree[i].a >= 0) common = common == -2 || common == ptree[i].b ? ptree[i].b : -1;
            if (common == 5) variant = 1;
            else if (common == 6) variant = 2;
        }
        ms.info[0] = uses_wp;
        ms.info[1] = nslots;
        ms.info[2] = simt ? 1 + variant : 0;
        ms.info[3] = n;
    }
    sync();
    const bool uses_wp = ms.info[0] != 0;
    if (!ms.info[2]) return MC_REST;
    const int variant = ms.info[2] - 1;
    if (c.w > ms.cap) return uses_wp ? MC_REST : MC_WIDE;
    if (variant == 2) return MC_WP;
    if (variant == 1 && !uses_wp) return MC_GRAD;
    if (!uses_wp && !cc.prefix && !cc.lz77) return MC_GEN;
    return MC_REST;
}

// `cls`: what modular_channel_prep returned. Does nothing unless K is that class (or MC_ANY).
template <int K, class Sync>
J40B_HD inline void modular_channel_run(int cls, BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs,
                                        const DTreeNode *tree, int32_t *wp_scratch, const int32_t *div24,
                                        const DTreeNode *ptree, const ModSmem &ms,
                                        const ModImage &m, int32_t cidx, int32_t sidx, int lane, int nlanes, Sync sync) {
    if (cls == MC_NONE || (K != MC_ANY && K != cls)) return;
    const ModChannel &c = m.ch[cidx];
    const bool uses_wp = ms.info[0] != 0;
    const int n = ms.info[3];
    int32_t refprops[SIMT_REF_SLOTS];
    for (int k = 0; k < SIMT_REF_SLOTS; ++k) refprops[k] = ms.info[4 + k];
    const int ns = ms.info[1];
    if ((K == MC_WP || K == MC_ANY) && cls == MC_WP) {
        modular_channel_simt<true, 6, 1, false>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
    } else if ((K == MC_GRAD || K == MC_ANY) && cls == MC_GRAD) {
        modular_channel_simt<false, 5, 1, false>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
    } else if ((K == MC_GEN || K == MC_ANY) && cls == MC_GEN) {
        modular_channel_simt<false, -1, 1, false>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
    } else if ((K == MC_WIDE || K == MC_ANY) && cls == MC_WIDE) {
        if (!cc.prefix && !cc.lz77) modular_channel_simt<false, -1, 1, true>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
        else modular_channel_simt<false, -1, 0, true>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
    } else if ((K == MC_REST || K == MC_ANY) && cls == MC_REST) {
        if (ms.info[2]) {
            const int variant = ms.info[2] - 1;
            // (specialising the generic-predictor variants for rANS without LZ77 as well was measured: no gain)
            if (c.w > ms.cap) modular_channel_simt<true, -1, 0, true>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
            else if (variant == 1) modular_channel_simt<true, 5, 1, false>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
            else if (uses_wp) modular_channel_simt<true, -1, 0, false>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
            else modular_channel_simt<false, -1, 0, false>(br, es, cc, cs, ms, wp_scratch, refprops, ns, div24, m, cidx, lane, nlanes, sync);
        } else {
            // trees that do not compile: walked, redundantly by all lanes so that their state stays in step
            const DTreeNode *t = n > 0 ? ptree : tree;
            if (uses_wp) modular_channel_t<true>(br, es, cc, cs, t, wp_scratch, div24, m, cidx, sidx);
            else modular_channel_t<false>(br, es, cc, cs, t, wp_scratch, div24, m, cidx, sidx);
            sync();
        }
    }
}

// Entry point for a warp that takes whatever the channel needs; every lane calls it with identical arguments and
// identical decoder state, and leaves it with identical state.
template <class Sync>
J40B_HD inline void modular_channel_warp(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs,
                                         const DTreeNode *tree, bool tree_uses_wp, int32_t *wp_scratch, const int32_t *div24,
                                         DTreeNode *ptree, int ptree_cap, const ModSmem &ms,
                                         const ModImage &m, int32_t cidx, int32_t sidx, int lane, int nlanes, Sync sync) {
    const int cls = modular_channel_prep(cc, tree, tree_uses_wp, wp_scratch, ptree, ptree_cap, ms, m, cidx, sidx, lane, sync);
    modular_channel_run<MC_ANY>(cls, br, es, cc, cs, tree, wp_scratch, div24, ptree, ms, m, cidx, sidx, lane, nlanes, sync);
}

// ModularHeader as far as the device understands it (RCT transforms only). Mirrors the checks of
// j40.h:3729-3759, 3816. A tree local to the sub-bitstream follows the header in the stream; reading it is the
// host's job (`local_tree` non-null: set to 1 and left to the caller; null: the device path, which only gets
// here for sub-bitstreams whose header the host could not reach -- LF groups of VarDCT frames -- rejects it).
// `allow_palette`: palette transforms change the channel list (one meta channel nb_colours x num_c in front, the
// num_c channels replaced by one index channel); the executor handles that for the global image only.
J40B_HD inline void modular_header(BitReader &br, ErrSlot &es, bool have_global_tree, ModImage &m, int *local_tree = nullptr,
                                   bool allow_palette = false) {
    m.nb_meta_channels = 0;
    int use_global_tree = (int) br.u(1);
    if (use_global_tree && !have_global_tree) { es.set(br, E_MTRE); return; }
    int default_wp = (int) br.u(1);
    m.wp.p1 = default_wp ? 16 : (int8_t) br.u(5);
    m.wp.p2 = default_wp ? 10 : (int8_t) br.u(5);
    for (int i = 0; i < 5; ++i) m.wp.p3[i] = default_wp ? (int8_t) (7 * (i < 3)) : (int8_t) br.u(5);
    for (int i = 0; i < 4; ++i) m.wp.w[i] = default_wp ? (int8_t) (12 + (i < 1)) : (int8_t) br.u(4);
    m.nb_transforms = (int32_t) br.u32(0, 0, 1, 0, 2, 4, 18, 8);
    if (m.nb_transforms > MOD_MAX_TRANSFORMS) { m.nb_transforms = 0; es.set(br, E_XLIM); return; } // (callers copy tr[0..nb_transforms))
    for (int i = 0; i < m.nb_transforms; ++i) {
        uint32_t id = br.u(2);
        m.tr[i].kind = 0; m.tr[i].num_c = m.tr[i].nb_colours = m.tr[i].nb_deltas = m.tr[i].d_pred = 0;
        if (id == 0) {
            int32_t begin_c = (int32_t) br.u32(0, 3, 8, 6, 72, 10, 1096, 13);
            int32_t type = (int32_t) br.u32(6, 0, 0, 2, 2, 4, 10, 6);
            m.tr[i].begin_c = begin_c;
            m.tr[i].type = type;
            if (type >= 42) { es.set(br, E_RCTT); return; }
            if (begin_c + 3 > m.num_channels) { es.set(br, E_RCTC); return; }
            if (!(begin_c >= m.nb_meta_channels || begin_c + 3 <= m.nb_meta_channels)) { es.set(br, E_RCTC); return; }
            if (begin_c < m.nb_meta_channels) { es.set(br, E_TODO); return; } // an RCT over palette channels: not built
            const ModChannel &a = m.ch[begin_c];
            for (int k = 1; k < 3; ++k) {
                const ModChannel &b = m.ch[begin_c + k];
                if (a.w != b.w || a.h != b.h || a.hshift != b.hshift || a.vshift != b.vshift) { es.set(br, E_RTCD); return; }
            }
        } else if (id == 3) {
            es.set(br, E_XFM);
            return;
        } else if (id == 1) { // palette (j40.h:3762-3792)
            ModTransform &t = m.tr[i];
            t.kind = 1;
            t.begin_c = (int32_t) br.u32(0, 3, 8, 6, 72, 10, 1096, 13);
            t.num_c = (int32_t) br.u32(1, 0, 3, 0, 4, 0, 1, 13);
            t.nb_colours = (int32_t) br.u32(0, 8, 256, 10, 1280, 12, 5376, 16);
            t.nb_deltas = (int32_t) br.u32(0, 0, 1, 8, 257, 10, 1281, 16);
            t.d_pred = (int32_t) br.u(4);
            t.type = 0;
            const int32_t end_c = t.begin_c + t.num_c;
            if (t.d_pred >= 14) { es.set(br, J40B_4CC('p', 'a', 'l', 'p')); return; }
            if (end_c > m.num_channels) { es.set(br, J40B_4CC('p', 'a', 'l', 'c')); return; }
            if (t.begin_c < m.nb_meta_channels) {
                if (end_c > m.nb_meta_channels) { es.set(br, J40B_4CC('p', 'a', 'l', 'c')); return; }
                es.set(br, E_TODO); // a palette of palette channels: not built
                return;
            }
            for (int k = t.begin_c + 1; k < end_c; ++k) {
                const ModChannel &a = m.ch[t.begin_c], &b = m.ch[k];
                if (a.w != b.w || a.h != b.h || a.hshift != b.hshift || a.vshift != b.vshift) { es.set(br, J40B_4CC('p', 'a', 'l', 'd')); return; }
            }
            // (delta palettes, nb_deltas > 0, predict from the restored neighbours: the executor undoes them with a scan
            // kernel ahead of the render step, and only as the last transform of the list; the host checks that)
            if (!allow_palette || m.num_channels + 2 - t.num_c > MOD_MAX_CH) { es.set(br, E_TODO); return; }
            const ModChannel input = m.ch[t.begin_c];
            ModChannel list[MOD_MAX_CH + 1];
            int n = 0;
            list[n] = input; list[n].w = t.nb_colours; list[n].h = t.num_c; list[n].stride = t.nb_colours; list[n].hshift = 0; list[n].vshift = -1; list[n].px = nullptr; ++n;
            for (int k = 0; k < t.begin_c; ++k) list[n++] = m.ch[k];
            list[n++] = input;
            for (int k = end_c; k < m.num_channels; ++k) list[n++] = m.ch[k];
            m.num_channels = n;
            for (int k = 0; k < n; ++k) m.ch[k] = list[k];
            m.nb_meta_channels += 1;
        } else {
            // squeeze: the reference itself rejects it (j40.h:3812)
            es.set(br, E_TODO);
            return;
        }
    }
    if (!use_global_tree) {
        if (!local_tree) { es.set(br, E_TODO); return; }
        *local_tree = 1;
    }
    // (in a local and stored once: `m` may be shared memory that every lane of a warp writes the same values to, and a
    // running maximum kept there is a read-modify-write between lanes -- compute-sanitizer racecheck, round 2)
    int32_t dist_mult = 0;
    for (int i = m.nb_meta_channels; i < m.num_channels; ++i) dist_mult = imax(dist_mult, m.ch[i].w);
    m.dist_mult = imin(dist_mult, 1 << 21);
}

// one pixel of an inverse RCT (j40.h:4341-4399); v[0..2] in, v'[perm[i]] out
J40B_HD J40B_INLINE void inverse_rct_px(int32_t type, int16_t &c0, int16_t &c1, int16_t &c2) {
    int32_t a = c0, b = c1, c = c2;
    switch (type % 7) {
    case 0: break;
    case 1: c = (int16_t) (c + a); break;
    case 2: c = (int16_t) (b + a); break;
    case 3: b = (int16_t) (b + a); c = (int16_t) (c + a); break;
    case 4: b = (int16_t) (b + (a / 2 + c / 2 + (a & c & 1))); break;
    case 5: b = (int16_t) (b + a + (c >> 1)); c = (int16_t) (c + a); break;
    case 6: {
        int32_t tmp = a - (c >> 1);
        int32_t p1 = c + tmp;
        int32_t p2 = tmp - (b >> 1);
        a = (int16_t) (p2 + b);
        b = (int16_t) p1;
        c = (int16_t) p2;
        break;
    }
    }
    int16_t o[3] = {(int16_t) a, (int16_t) b, (int16_t) c};
    // PERMUTATIONS[type / 7]: result i goes to channel begin_c + perm[i]
    const uint8_t P[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
    int16_t r[3];
    const uint8_t *pp = P[type / 7];
    r[pp[0]] = o[0]; r[pp[1]] = o[1]; r[pp[2]] = o[2];
    c0 = r[0]; c1 = r[1]; c2 = r[2];
}

// applies the image's transforms in reverse order to its channels (parallel over pixels)
J40B_HD inline void inverse_transforms(const ModImage &m, int tid, int nthreads) {
    for (int t = m.nb_transforms - 1; t >= 0; --t) {
        const ModChannel &a = m.ch[m.tr[t].begin_c], &b = m.ch[m.tr[t].begin_c + 1], &c = m.ch[m.tr[t].begin_c + 2];
        int32_t n = a.w * a.h;
        for (int32_t i = tid; i < n; i += nthreads) {
            int32_t y = i / a.w, x = i - y * a.w;
            inverse_rct_px(m.tr[t].type, a.px[(size_t) y * a.stride + x], b.px[(size_t) y * b.stride + x], c.px[(size_t) y * c.stride + x]);
        }
        // NOTE: successive transforms touching the same channels need a barrier between them;
        // callers with nb_transforms > 1 invoke this per transform (see kernels)
    }
}

} // namespace j40b
