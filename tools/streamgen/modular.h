// JPEG XL test-stream writer -- TEST INFRASTRUCTURE.
// Modular sub-bitstream encoder: meta-adaptive tree description, the 14 predictors incl. the
// self-correcting (weighted) predictor, property evaluation and residual tokenisation, with the
// semantics the reference decoder implements (j40.h:3461-3513,
// 3965-4229).  The encoder must mirror the decoder's integer arithmetic exactly, otherwise the
// contexts (and with the weighted predictor the predictions) diverge.
#pragma once
#include "entropy.h"
#include <memory>

namespace jxlgen {

struct MANode {
    bool leaf = true;
    // branch
    int prop = 0, value = 0;
    int left = -1, right = -1; // indices into MATree::nodes (left taken when property > value)
    // leaf
    int ctx = 0, predictor = 0, offset = 0, mul_log = 0, mul_bits = 0;
    int multiplier() const { return (mul_bits + 1) << mul_log; }
};

struct MATree {
    std::vector<MANode> nodes; // breadth-first order, root at 0
    int num_leaves = 0;

    // builder helpers: construct as a pointer tree, then flatten breadth-first
    struct B {
        bool leaf; int prop, value; std::shared_ptr<B> l, r; int predictor, offset, mul_log, mul_bits;
    };
    typedef std::shared_ptr<B> P;
    static P Leaf(int predictor, int offset = 0, int mul_log = 0, int mul_bits = 0) {
        P p = std::make_shared<B>();
        p->leaf = true; p->predictor = predictor; p->offset = offset; p->mul_log = mul_log; p->mul_bits = mul_bits;
        p->prop = p->value = 0;
        return p;
    }
    static P Branch(int prop, int value, P gt, P le) {
        P p = std::make_shared<B>();
        p->leaf = false; p->prop = prop; p->value = value; p->l = gt; p->r = le;
        p->predictor = p->offset = p->mul_log = p->mul_bits = 0;
        return p;
    }
    void flatten(P root) {
        nodes.clear();
        num_leaves = 0;
        std::vector<P> q{root};
        // first pass: assign indices breadth-first
        for (size_t i = 0; i < q.size(); ++i) {
            if (!q[i]->leaf) { q.push_back(q[i]->l); q.push_back(q[i]->r); }
        }
        nodes.resize(q.size());
        size_t next = 1;
        for (size_t i = 0; i < q.size(); ++i) {
            MANode &n = nodes[i];
            n.leaf = q[i]->leaf;
            if (n.leaf) {
                n.ctx = num_leaves++;
                n.predictor = q[i]->predictor; n.offset = q[i]->offset;
                n.mul_log = q[i]->mul_log; n.mul_bits = q[i]->mul_bits;
            } else {
                n.prop = q[i]->prop; n.value = q[i]->value;
                n.left = (int) next++; n.right = (int) next++;
            }
        }
    }
    bool uses_wp() const {
        for (const MANode &n : nodes) if (n.leaf ? n.predictor == 6 : n.prop == 15) return true;
        return false;
    }
    int max_prop() const {
        int m = 0;
        for (const MANode &n : nodes) if (!n.leaf) m = std::max(m, n.prop);
        return m;
    }
};

inline uint32_t pack_signed(int32_t v) { return v >= 0 ? (uint32_t) v * 2u : (uint32_t) (-(int64_t) v) * 2u - 1u; }

// writes the tree with its own 6-context code spec, then the sample code spec placeholder is the
// caller's business (the sample spec follows the tree in the bitstream)
inline void write_tree(BitWriter &bw, const MATree &t, const EntropyOpts &opts) {
    TokStream ts;
    for (const MANode &n : t.nodes) {
        if (n.leaf) {
            ts.push_back({1, 0, 0});
            ts.push_back({2, (uint32_t) n.predictor, 0});
            ts.push_back({3, pack_signed(n.offset), 0});
            ts.push_back({4, (uint32_t) n.mul_log, 0});
            ts.push_back({5, (uint32_t) n.mul_bits, 0});
        } else {
            ts.push_back({1, (uint32_t) n.prop + 1, 0});
            ts.push_back({0, pack_signed(n.value), 0});
        }
    }
    CodeSpec cs;
    std::vector<const TokStream *> v{&ts};
    EntropyOpts o = opts;
    o.lz77 = false;
    cs.build(6, o, v);
    cs.write(bw);
    cs.encode(bw, ts);
}

struct WPParams {
    int p1 = 16, p2 = 10, p3[5] = {7, 7, 7, 0, 0}, w[4] = {13, 12, 12, 12};
    bool is_default() const {
        return p1 == 16 && p2 == 10 && p3[0] == 7 && p3[1] == 7 && p3[2] == 7 && p3[3] == 0 && p3[4] == 0 &&
               w[0] == 13 && w[1] == 12 && w[2] == 12 && w[3] == 12;
    }
};

struct Channel {
    int w = 0, h = 0, hshift = 0, vshift = 0;
    std::vector<int32_t> px; // row-major, values must fit int16 for the reference's 16-bit buffers
    int32_t at(int x, int y) const { return px[(size_t) y * (size_t) w + (size_t) x]; }
};

// floor(2^24 / (i + 1)), the divisor table of the weighted predictor
inline int32_t div24(int i) { return (int32_t) ((1u << 24) / (uint32_t) (i + 1)); }

// Tokenises the channels of one modular sub-image in decoding order.  `sidx` is the stream index
// property (property 1).  Channels must be listed exactly as the decoder will see them.
class ModularTokenizer {
public:
    const MATree &tree;
    WPParams wp;
    explicit ModularTokenizer(const MATree &t) : tree(t) {}

    void run(const std::vector<Channel> &ch, int64_t sidx, TokStream &out) const {
        for (size_t c = 0; c < ch.size(); ++c) channel(ch, (int) c, sidx, out);
    }

private:
    struct WPState {
        int width = 0;
        bool on = false;
        std::vector<int32_t> err; // [2][width][5]
        int32_t pred[5] = {0, 0, 0, 0, 0};
        int32_t tw = 0, tn = 0, tnw = 0, tne = 0;
    };

    static int32_t grad(int32_t w, int32_t n, int32_t nw) {
        int32_t lo = std::min(w, n), hi = std::max(w, n);
        return std::min(std::max(lo, w + n - nw), hi);
    }
    static int flg(uint32_t x) { return 31 - __builtin_clz(x); }

    void wp_before(WPState &s, int x, int y, int32_t pw, int32_t pn, int32_t pnw, int32_t pne, int32_t pnn) const {
        if (!s.on) return;
        static const int32_t ZERO[5] = {0, 0, 0, 0, 0};
        int32_t *cur = s.err.data() + (size_t) ((y & 1) ? s.width : 0) * 5;
        int32_t *oth = s.err.data() + (size_t) ((y & 1) ? 0 : s.width) * 5;
        const int32_t *ew = x > 0 ? cur + (size_t) (x - 1) * 5 : ZERO;
        const int32_t *en = y > 0 ? oth + (size_t) x * 5 : ZERO;
        const int32_t *enw = x > 0 && y > 0 ? oth + (size_t) (x - 1) * 5 : en;
        const int32_t *ene = x + 1 < s.width && y > 0 ? oth + (size_t) (x + 1) * 5 : en;
        const int32_t *eww = x > 1 ? cur + (size_t) (x - 2) * 5 : ZERO;
        const int32_t *ew2 = x + 1 < s.width ? ZERO : ew;
        s.tw = x > 0 ? cur[(size_t) (x - 1) * 5 + 4] : 0;
        s.tn = y > 0 ? oth[(size_t) x * 5 + 4] : 0;
        s.tnw = x > 0 && y > 0 ? oth[(size_t) (x - 1) * 5 + 4] : s.tn;
        s.tne = x + 1 < s.width && y > 0 ? oth[(size_t) (x + 1) * 5 + 4] : s.tn;
        s.pred[0] = (pw + pne - pn) * 8;
        s.pred[1] = pn * 8 - (((s.tw + s.tn + s.tne) * wp.p1) >> 5);
        s.pred[2] = pw * 8 - (((s.tw + s.tn + s.tnw) * wp.p2) >> 5);
        s.pred[3] = pn * 8 - ((s.tnw * wp.p3[0] + s.tn * wp.p3[1] + s.tne * wp.p3[2] +
                               (pnn - pn) * 8 * wp.p3[3] + (pnw - pw) * 8 * wp.p3[4]) >> 5);
        int32_t wgt[4];
        for (int i = 0; i < 4; ++i) {
            int32_t errsum = en[i] + ew[i] + enw[i] + eww[i] + ene[i] + ew2[i];
            int shift = std::max(flg((uint32_t) errsum + 1) - 5, 0);
            wgt[i] = (int32_t) (4 + (((int64_t) wp.w[i] * div24(errsum >> shift)) >> shift));
        }
        int logw = flg((uint32_t) (wgt[0] + wgt[1] + wgt[2] + wgt[3])) - 4;
        int32_t wsum = 0, sum = 0;
        for (int i = 0; i < 4; ++i) {
            wgt[i] >>= logw;
            wsum += wgt[i];
            sum += s.pred[i] * wgt[i];
        }
        s.pred[4] = (int32_t) ((((int64_t) sum + (wsum >> 1) - 1) * div24(wsum - 1)) >> 24);
        if (((s.tn ^ s.tw) | (s.tn ^ s.tnw)) <= 0) {
            int32_t lo = std::min(pw, std::min(pn, pne)) * 8;
            int32_t hi = std::max(pw, std::max(pn, pne)) * 8;
            s.pred[4] = std::min(std::max(lo, s.pred[4]), hi);
        }
    }

    void wp_after(WPState &s, int x, int y, int32_t val) const {
        if (!s.on) return;
        int32_t *e = s.err.data() + ((size_t) ((y & 1) ? s.width : 0) + (size_t) x) * 5;
        for (int i = 0; i < 4; ++i) e[i] = (std::abs(s.pred[i] - val * 8) + 3) >> 3;
        e[4] = s.pred[4] - val * 8;
    }

    void channel(const std::vector<Channel> &ch, int cidx, int64_t sidx, TokStream &out) const {
        const Channel &c = ch[(size_t) cidx];
        if (c.w == 0 || c.h == 0) return;
        WPState s;
        s.width = c.w;
        s.on = tree.uses_wp();
        if (s.on) s.err.assign((size_t) c.w * 2 * 5, 0);
        std::vector<int> refs;
        for (int i = cidx - 1; i >= 0; --i) {
            const Channel &r = ch[(size_t) i];
            if (r.w != c.w || r.h != c.h || r.hshift != c.hshift || r.vshift != c.vshift) continue;
            refs.push_back(i);
        }
        for (int y = 0; y < c.h; ++y) for (int x = 0; x < c.w; ++x) {
            int32_t pw = x > 0 ? c.at(x - 1, y) : y > 0 ? c.at(x, y - 1) : 0;
            int32_t pn = y > 0 ? c.at(x, y - 1) : pw;
            int32_t pnw = x > 0 && y > 0 ? c.at(x - 1, y - 1) : pw;
            int32_t pne = x + 1 < c.w && y > 0 ? c.at(x + 1, y - 1) : pn;
            int32_t pnn = y > 1 ? c.at(x, y - 2) : pn;
            int32_t pnee = x + 2 < c.w && y > 0 ? c.at(x + 2, y - 1) : pne;
            int32_t pww = x > 1 ? c.at(x - 2, y) : pw;
            int32_t pnww = x > 1 && y > 0 ? c.at(x - 2, y - 1) : pww;
            wp_before(s, x, y, pw, pn, pnw, pne, pnn);
            const MANode *n = &tree.nodes[0];
            while (!n->leaf) {
                int32_t val;
                switch (n->prop) {
                case 0: val = cidx; break;
                case 1: val = (int32_t) sidx; break;
                case 2: val = y; break;
                case 3: val = x; break;
                case 4: val = std::abs(pn); break;
                case 5: val = std::abs(pw); break;
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
                    val = s.tw;
                    if (std::abs(val) < std::abs(s.tn)) val = s.tn;
                    if (std::abs(val) < std::abs(s.tnw)) val = s.tnw;
                    if (std::abs(val) < std::abs(s.tne)) val = s.tne;
                    break;
                default: {
                    int ridx = (n->prop - 16) / 4;
                    JG_CHECK(ridx < (int) refs.size());
                    const Channel &r = ch[(size_t) refs[(size_t) ridx]];
                    val = r.at(x, y);
                    if (n->prop & 2) {
                        int32_t rw = x > 0 ? r.at(x - 1, y) : 0;
                        int32_t rn = y > 0 ? r.at(x, y - 1) : rw;
                        int32_t rnw = x > 0 && y > 0 ? r.at(x - 1, y - 1) : rw;
                        val -= grad(rw, rn, rnw);
                    }
                    if (n->prop & 1) val = std::abs(val);
                    break;
                }
                }
                n = &tree.nodes[(size_t) (val > n->value ? n->left : n->right)];
            }
            int32_t pred;
            switch (n->predictor) {
            case 0: pred = 0; break;
            case 1: pred = pw; break;
            case 2: pred = pn; break;
            case 3: pred = (pw + pn) / 2; break;
            case 4: pred = std::abs(pn - pnw) < std::abs(pw - pnw) ? pw : pn; break;
            case 5: pred = grad(pw, pn, pnw); break;
            case 6: pred = (s.pred[4] + 3) >> 3; break;
            case 7: pred = pne; break;
            case 8: pred = pnw; break;
            case 9: pred = pww; break;
            case 10: pred = (pw + pnw) / 2; break;
            case 11: pred = (pn + pnw) / 2; break;
            case 12: pred = (pn + pne) / 2; break;
            case 13: pred = (6 * pn - 2 * pnn + 7 * pw + pww + pnee + 3 * pne + 8) / 16; break;
            default: throw GenError("bad predictor");
            }
            int32_t v = c.at(x, y);
            JG_CHECK(v >= -32768 && v <= 32767);
            int32_t resid = v - pred - n->offset;
            int32_t mul = n->multiplier();
            JG_CHECK(resid % mul == 0);
            out.push_back({(uint32_t) n->ctx, pack_signed(resid / mul), 0});
            wp_after(s, x, y, v);
        }
    }
};

// Replaces runs of identical token values by LZ77 copies at distance 1 (run-length coding).
// `dist_mult` is the value the decoder will pass to its symbol reader (0 outside modular images,
// otherwise the widest non-meta channel); it decides how "distance 1" is coded (j40.h:2829-2847).
inline void lz77_rle(TokStream &ts, int min_length, int dist_mult, int min_run = 0) {
    TokStream out;
    size_t n = ts.size();
    if (min_run < min_length) min_run = min_length;
    for (size_t i = 0; i < n;) {
        size_t j = i;
        if (i > 0) while (j < n && ts[j].val == ts[i - 1].val && ts[j].kind == 0) ++j;
        size_t run = j - i;
        if (i > 0 && run >= (size_t) min_run) {
            out.push_back({ts[i].ctx, (uint32_t) (run - (size_t) min_length), 1});
            out.push_back({0, (uint32_t) (dist_mult ? 1 : 0), 2});
            i = j;
        } else {
            out.push_back(ts[i]);
            ++i;
        }
    }
    ts.swap(out);
}

// modular header fields (App. E.6) -- transforms limited to RCTs here
struct ModularHeaderOpts {
    bool use_global_tree = true;
    WPParams wp;
    std::vector<std::pair<int, int>> rcts; // (begin_c, type)
    struct Pal { int begin_c, num_c, nb_colours, nb_deltas, d_pred; };
    std::vector<Pal> palettes;             // written before the RCTs
};

inline void write_modular_header_prefix(BitWriter &bw, const ModularHeaderOpts &h) {
    bw.bit(h.use_global_tree);
    bool def = h.wp.is_default();
    bw.bit(def);
    if (!def) {
        bw.put((uint64_t) h.wp.p1, 5);
        bw.put((uint64_t) h.wp.p2, 5);
        for (int i = 0; i < 5; ++i) bw.put((uint64_t) h.wp.p3[i], 5);
        for (int i = 0; i < 4; ++i) bw.put((uint64_t) h.wp.w[i], 4);
    }
    bw.u32((uint32_t) (h.rcts.size() + h.palettes.size()), 0, 0, 1, 0, 2, 4, 18, 8);
    for (auto &pl : h.palettes) {
        bw.put(1, 2); // palette
        bw.u32((uint32_t) pl.begin_c, 0, 3, 8, 6, 72, 10, 1096, 13);
        bw.u32((uint32_t) pl.num_c, 1, 0, 3, 0, 4, 0, 1, 13);
        bw.u32((uint32_t) pl.nb_colours, 0, 8, 256, 10, 1280, 12, 5376, 16);
        bw.u32((uint32_t) pl.nb_deltas, 0, 0, 1, 8, 257, 10, 1281, 16);
        bw.put((uint64_t) pl.d_pred, 4);
    }
    for (auto &r : h.rcts) {
        bw.put(0, 2); // RCT
        bw.u32((uint32_t) r.first, 0, 3, 8, 6, 72, 10, 1096, 13);
        bw.u32((uint32_t) r.second, 6, 0, 0, 2, 2, 4, 10, 6);
    }
}

} // namespace jxlgen
