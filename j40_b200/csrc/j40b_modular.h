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

struct ModTransform { int32_t begin_c, type; };

struct ModImage {
    int32_t num_channels;
    ModChannel ch[MOD_MAX_CH];
    WPParams wp;
    int32_t nb_transforms;
    ModTransform tr[MOD_MAX_TRANSFORMS];
    int32_t dist_mult;
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
// Warp-cooperative variant: lane 0 is the (inherently serial) decoder, the other lanes keep everything it
// touches per sample in shared memory: a ring of the last three sample rows, the weighted predictor's two
// error rows, and the current / previous row of up to two reference channels (properties >= 16). The
// per-sample dependency chain then never leaves the SM. Same arithmetic as modular_channel_t.
enum { MOD_STAGED_REFS = 2 };
struct ModSmem {
    int16_t *rows;   // [3][cap]
    int32_t *wp;     // [2][cap][5]
    int16_t *refs;   // [MOD_STAGED_REFS][2][cap]
    int32_t *info;   // [4]: uses_wp, refs staged, use row path, abort
    int32_t cap;     // widest channel the row path can take
};

template <bool USE_WP, class Sync>
J40B_HD inline void modular_channel_rows(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs,
                                         const DTreeNode *tree, const ModSmem &ms, const int32_t *div24,
                                         const ModImage &m, int32_t cidx, int32_t sidx, int nstaged,
                                         int lane, int nlanes, Sync sync) {
    const ModChannel &c = m.ch[cidx];
    const int32_t width = c.w, height = c.h, stride = c.stride, cap = ms.cap;
    const int32_t dist_mult = m.dist_mult;
    const WPParams wpp = m.wp;
    WPState wp;
    wp.errors = ms.wp;
    wp.width = width;
    for (int i = 0; i < 5; ++i) wp.pred[i] = 0;
    wp.trueerrw = wp.trueerrn = wp.trueerrnw = wp.trueerrne = 0;
    int32_t refcmap[MOD_MAX_CH], nref = 0;
    for (int32_t i = cidx - 1; i >= 0; --i) {
        const ModChannel &r = m.ch[i];
        if (c.w != r.w || c.h != r.h || c.hshift != r.hshift || c.vshift != r.vshift) continue;
        refcmap[nref++] = i;
    }
    if (nstaged > nref) nstaged = nref;
    if (USE_WP) for (int32_t i = lane; i < width * 2 * 5; i += nlanes) ms.wp[i] = 0;
    for (int32_t y = 0; y < height; ++y) {
        // all lanes: row y of the staged reference channels (row y-1 is still there from the last round)
        for (int r = 0; r < nstaged; ++r) {
            const ModChannel &rc = m.ch[refcmap[r]];
            const int16_t *src = rc.px + (size_t) y * (size_t) rc.stride;
            int16_t *dst = ms.refs + (size_t) (r * 2 + (y & 1)) * cap;
            for (int32_t x = lane; x < width; x += nlanes) dst[x] = src[x];
        }
        sync();
        if (lane == 0) {
            int16_t *row = c.px + (size_t) y * (size_t) stride;
            int16_t *cur = ms.rows + (size_t) (y % 3) * cap;
            const int16_t *nrow = ms.rows + (size_t) ((y + 2) % 3) * cap, *nnrow = ms.rows + (size_t) ((y + 1) % 3) * cap;
            int32_t prev = 0, prev2 = 0;
            for (int32_t x = 0; x < width; ++x) {
                int32_t pw = x > 0 ? prev : y > 0 ? nrow[x] : 0;
                int32_t pn = y > 0 ? nrow[x] : pw;
                int32_t pnw = x > 0 && y > 0 ? nrow[x - 1] : pw;
                int32_t pne = x + 1 < width && y > 0 ? nrow[x + 1] : pn;
                int32_t pnn = y > 1 ? nnrow[x] : pn;
                int32_t pnee = x + 2 < width && y > 0 ? nrow[x + 2] : pne;
                int32_t pww = x > 1 ? prev2 : pw;
                int32_t pnww = x > 1 && y > 0 ? nrow[x - 2] : pww;
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
                        if (refcidx >= nref) { es.set(br, E_TREC); break; }
                        int32_t rw, rn, rnw;
                        if (refcidx < nstaged) {
                            const int16_t *ry = ms.refs + (size_t) (refcidx * 2 + (y & 1)) * cap;
                            const int16_t *rp = ms.refs + (size_t) (refcidx * 2 + ((y & 1) ^ 1)) * cap;
                            val = ry[x];
                            rw = x > 0 ? ry[x - 1] : 0;
                            rn = y > 0 ? rp[x] : rw;
                            rnw = x > 0 && y > 0 ? rp[x - 1] : rw;
                        } else {
                            const ModChannel &r = m.ch[refcmap[refcidx]];
                            const int16_t *rp = r.px + (size_t) y * (size_t) r.stride + x;
                            val = rp[0];
                            rw = x > 0 ? rp[-1] : 0;
                            rn = y > 0 ? rp[-r.stride] : rw;
                            rnw = x > 0 && y > 0 ? rp[-1 - r.stride] : rw;
                        }
                        if (prop & 2) val -= mod_gradient(rw, rn, rnw);
                        if (prop & 1) val = iabs(val);
                        break;
                    }
                    }
                    if (es.err) break;
                    n = tree + (val > n->b ? n->c : n->d);
                }
                if (es.err) break;

                int32_t val = code(br, es, cc, cs, n->a, dist_mult);
                val = unpack_signed(val) * n->d + n->c;
                int32_t pred = 0;
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
                default: es.set(br, E_PRED); break;
                }
                val += pred;
                if (es.err) break;
                if (val < -32768 || val > 32767) { es.set(br, E_POVF); break; }
                row[x] = (int16_t) val;
                cur[x] = (int16_t) val;
                prev2 = prev;
                prev = val;
                if (USE_WP) wp_after_predict(wp, x, y, val);
            }
            ms.info[3] = es.err ? 1 : 0;
        }
        sync();
        if (ms.info[3]) return;
    }
}

// Entry point for a warp: decides (lane 0) between the row path above and the plain path, then all lanes
// follow. Every lane must call this with the same arguments; only lane 0's br / es / cs are meaningful.
template <class Sync>
J40B_HD inline void modular_channel_warp(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs,
                                         const DTreeNode *tree, bool tree_uses_wp, int32_t *wp_scratch, const int32_t *div24,
                                         DTreeNode *ptree, int ptree_cap, const ModSmem &ms,
                                         const ModImage &m, int32_t cidx, int32_t sidx, int lane, int nlanes, Sync sync) {
    const ModChannel &c = m.ch[cidx];
    if (c.w <= 0 || c.h <= 0) return;
    if (lane == 0) {
        bool uses_wp = tree_uses_wp;
        int n = ptree_cap > 0 ? prune_tree(tree, cidx, sidx, ptree, ptree_cap, &uses_wp) : 0;
        if (n == 0) uses_wp = tree_uses_wp;
        int maxprop = 0;
        const DTreeNode *t = n ? ptree : tree;
        if (n) for (int i = 0; i < n; ++i) { if (t[i].a < 0 && -1 - t[i].a > maxprop) maxprop = -1 - t[i].a; }
        else maxprop = 16 + 4 * MOD_STAGED_REFS - 1; // unknown: stage what we can
        int want = maxprop >= 16 ? (maxprop - 16) / 4 + 1 : 0;
        ms.info[0] = uses_wp;
        ms.info[1] = want < MOD_STAGED_REFS ? want : MOD_STAGED_REFS;
        ms.info[2] = (ms.rows && c.w <= ms.cap && n > 0) ? 1 : 0;
        ms.info[3] = 0;
    }
    sync();
    const bool uses_wp = ms.info[0] != 0;
    const int nstaged = ms.info[1];
    if (ms.info[2]) {
        if (uses_wp) modular_channel_rows<true>(br, es, cc, cs, ptree, ms, div24, m, cidx, sidx, nstaged, lane, nlanes, sync);
        else modular_channel_rows<false>(br, es, cc, cs, ptree, ms, div24, m, cidx, sidx, nstaged, lane, nlanes, sync);
    } else {
        if (lane == 0) {
            // wide channels (or trees too large to prune into shared memory): plain path through global memory
            const DTreeNode *t = tree;
            bool wpf = tree_uses_wp;
            if (ptree_cap > 0 && prune_tree(tree, cidx, sidx, ptree, ptree_cap, &wpf) > 0) t = ptree; else wpf = tree_uses_wp;
            if (wpf) modular_channel_t<true>(br, es, cc, cs, t, wp_scratch, div24, m, cidx, sidx);
            else modular_channel_t<false>(br, es, cc, cs, t, wp_scratch, div24, m, cidx, sidx);
        }
        sync();
    }
}

// ModularHeader as far as the device understands it (global tree only; RCT transforms only).
// Mirrors the checks of j40.h:3729-3759, 3816.
J40B_HD inline void modular_header(BitReader &br, ErrSlot &es, bool have_global_tree, ModImage &m) {
    int use_global_tree = (int) br.u(1);
    if (use_global_tree && !have_global_tree) { es.set(br, E_MTRE); return; }
    int default_wp = (int) br.u(1);
    m.wp.p1 = default_wp ? 16 : (int8_t) br.u(5);
    m.wp.p2 = default_wp ? 10 : (int8_t) br.u(5);
    for (int i = 0; i < 5; ++i) m.wp.p3[i] = default_wp ? (int8_t) (7 * (i < 3)) : (int8_t) br.u(5);
    for (int i = 0; i < 4; ++i) m.wp.w[i] = default_wp ? (int8_t) (12 + (i < 1)) : (int8_t) br.u(4);
    m.nb_transforms = (int32_t) br.u32(0, 0, 1, 0, 2, 4, 18, 8);
    if (m.nb_transforms > MOD_MAX_TRANSFORMS) { es.set(br, E_XLIM); return; }
    for (int i = 0; i < m.nb_transforms; ++i) {
        uint32_t id = br.u(2);
        if (id == 0) {
            int32_t begin_c = (int32_t) br.u32(0, 3, 8, 6, 72, 10, 1096, 13);
            int32_t type = (int32_t) br.u32(6, 0, 0, 2, 2, 4, 10, 6);
            m.tr[i].begin_c = begin_c;
            m.tr[i].type = type;
            if (type >= 42) { es.set(br, E_RCTT); return; }
            if (begin_c + 3 > m.num_channels) { es.set(br, E_RCTC); return; }
            const ModChannel &a = m.ch[begin_c];
            for (int k = 1; k < 3; ++k) {
                const ModChannel &b = m.ch[begin_c + k];
                if (a.w != b.w || a.h != b.h || a.hshift != b.hshift || a.vshift != b.vshift) { es.set(br, E_RTCD); return; }
            }
        } else if (id == 3) {
            es.set(br, E_XFM);
            return;
        } else {
            // palette and squeeze change the channel list; not decoded on the device yet (the
            // reference itself rejects squeeze, j40.h:3812)
            es.set(br, E_TODO);
            return;
        }
    }
    if (!use_global_tree) {
        // a tree local to this sub-bitstream would have to be parsed here; see DESIGN.md (out of
        // scope for the device path in this round)
        es.set(br, E_TODO);
        return;
    }
    m.dist_mult = 0;
    for (int i = 0; i < m.num_channels; ++i) m.dist_mult = imax(m.dist_mult, m.ch[i].w);
    m.dist_mult = imin(m.dist_mult, 1 << 21);
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
