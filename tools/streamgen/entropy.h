// JPEG XL test-stream writer -- TEST INFRASTRUCTURE.
// Entropy-coding back end: hybrid-uint tokenisation, context clustering, rANS (alias-table
// compatible) and Brotli-style prefix codes (RFC 7932 §3), LZ77 copies,
// plus serialisation of the code spec exactly as the reference *parses* it (App. E.10,
// j40.h:2526-2777).  Written from the format description; shares no code with the decoder.
#pragma once
#include "bits.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <queue>

namespace jxlgen {

struct HybridCfg { int split_exp, msb, lsb; };

inline void hybrid_split(HybridCfg c, uint32_t v, uint32_t &token, int &nbits, uint32_t &bits) {
    uint32_t split = 1u << c.split_exp;
    if (v < split) { token = v; nbits = 0; bits = 0; return; }
    int n = floor_lg(v);
    uint32_t m = v - (1u << n);
    token = split + (uint32_t) (((n - c.split_exp) << (c.msb + c.lsb)) + (int) ((m >> (n - c.msb)) << c.lsb) + (int) (m & ((1u << c.lsb) - 1)));
    nbits = n - c.msb - c.lsb;
    bits = (m >> c.lsb) & ((1u << nbits) - 1);
}

// one decoded integer (or half of an LZ77 copy) as the decoder will see it
struct Tok {
    uint32_t ctx;
    uint32_t val;
    uint8_t kind; // 0 literal; 1 LZ77 length (val = length - min_length); 2 LZ77 distance (val = coded distance)
};
typedef std::vector<Tok> TokStream;

struct EntropyOpts {
    bool use_prefix = false;
    int log_alpha_size = 6;        // ANS only, 5..8
    int max_clusters = 32;         // <= 256
    HybridCfg cfg = {4, 1, 0};
    bool lz77 = false;
    int min_symbol = 224, min_length = 3;
    HybridCfg lz_len_cfg = {0, 0, 0};
    int cluster_map_mode = 0;      // 0 auto, 1 force complex+mtf, 2 force complex no mtf
    int ans_precision_shift = 13;  // `shift` of the general histogram coding
    bool allow_lz = true;          // whether a nested spec may itself declare lz77 (we never do)
};

class CodeSpec {
public:
    int num_ctx = 0;   // contexts the caller uses; with lz77 one more (distance) context follows
    EntropyOpts o;
    std::vector<uint8_t> cmap; // num_ctx (+1 if lz77)
    int nclusters = 0;
    // ANS
    std::vector<std::vector<uint16_t>> D;     // [cluster][1 << las]
    std::vector<std::vector<uint16_t>> slot;  // [cluster][4096]: (symbol, offset) -> slot via cum
    std::vector<std::vector<uint32_t>> cum;   // [cluster][(1 << las) + 1]
    // prefix
    std::vector<std::vector<uint8_t>> plen;   // [cluster][alphabet]
    std::vector<std::vector<uint16_t>> pcode; // LSB-first code words

    int total_ctx() const { return num_ctx + (o.lz77 ? 1 : 0); }

    void symbolize(const Tok &t, int &cluster, uint32_t &sym, int &nb, uint32_t &bits) const {
        if (t.kind == 0) {
            cluster = cmap[t.ctx];
            hybrid_split(o.cfg, t.val, sym, nb, bits);
            if (o.lz77) JG_CHECK((int) sym < o.min_symbol);
        } else if (t.kind == 1) {
            cluster = cmap[t.ctx];
            hybrid_split(o.lz_len_cfg, t.val, sym, nb, bits);
            sym += (uint32_t) o.min_symbol;
        } else {
            cluster = cmap[num_ctx];
            hybrid_split(o.cfg, t.val, sym, nb, bits);
        }
    }

    // ------------------------------------------------------------------------------------
    void build(int num_ctx_, const EntropyOpts &opts, const std::vector<const TokStream *> &streams) {
        num_ctx = num_ctx_;
        o = opts;
        int nctx = total_ctx();
        int alpha_cap = o.use_prefix ? (1 << 15) : (1 << o.log_alpha_size);
        // histograms per context
        std::vector<std::vector<uint32_t>> hist((size_t) nctx);
        cmap.assign((size_t) nctx, 0);
        for (const TokStream *ts : streams) for (const Tok &t : *ts) {
            int cl; uint32_t sym; int nb; uint32_t bits;
            int ctx = t.kind == 2 ? num_ctx : (int) t.ctx;
            JG_CHECK(ctx < nctx);
            // cluster not yet known: symbolize only needs the config
            if (t.kind == 1) { hybrid_split(o.lz_len_cfg, t.val, sym, nb, bits); sym += (uint32_t) o.min_symbol; }
            else hybrid_split(o.cfg, t.val, sym, nb, bits);
            (void) cl;
            if ((int) sym >= alpha_cap) throw GenError("symbol exceeds alphabet; raise log_alpha_size");
            if (hist[(size_t) ctx].size() <= sym) hist[(size_t) ctx].resize(sym + 1, 0);
            hist[(size_t) ctx][sym]++;
        }
        cluster(hist);
        // per-cluster merged histograms
        std::vector<std::vector<uint32_t>> ch((size_t) nclusters);
        for (int c = 0; c < nctx; ++c) {
            auto &dst = ch[cmap[(size_t) c]];
            const auto &src = hist[(size_t) c];
            if (dst.size() < src.size()) dst.resize(src.size(), 0);
            for (size_t s = 0; s < src.size(); ++s) dst[s] += src[s];
        }
        if (o.use_prefix) build_prefix(ch); else build_ans(ch);
    }

    // ------------------------------------------------------------------------------------
    void write(BitWriter &bw) const {
        int nctx = total_ctx();
        bw.bit(o.lz77);
        if (o.lz77) {
            bw.u32((uint32_t) o.min_symbol, 224, 0, 512, 0, 4096, 0, 8, 15);
            bw.u32((uint32_t) o.min_length, 3, 0, 4, 0, 5, 2, 9, 8);
            write_hybrid_cfg(bw, o.lz_len_cfg, 8);
        }
        write_cluster_map(bw, cmap, nctx, nclusters, o.cluster_map_mode);
        bw.bit(o.use_prefix);
        if (o.use_prefix) {
            for (int i = 0; i < nclusters; ++i) write_hybrid_cfg(bw, o.cfg, 15);
            for (int i = 0; i < nclusters; ++i) {
                int count = (int) plen[(size_t) i].size();
                if (count <= 1) bw.bit(0);
                else {
                    bw.bit(1);
                    int n = floor_lg((uint32_t) (count - 1));
                    bw.put((uint64_t) n, 4);
                    bw.put((uint64_t) (count - 1 - (1 << n)), n);
                }
            }
            for (int i = 0; i < nclusters; ++i) write_prefix_tree(bw, plen[(size_t) i]);
        } else {
            bw.put((uint64_t) (o.log_alpha_size - 5), 2);
            for (int i = 0; i < nclusters; ++i) write_hybrid_cfg(bw, o.cfg, o.log_alpha_size);
            for (int i = 0; i < nclusters; ++i) write_ans_table(bw, D[(size_t) i]);
        }
    }

    // encodes one token stream (one entropy-coded sub-bitstream) into bw
    void encode(BitWriter &bw, const TokStream &ts) const {
        size_t n = ts.size();
        std::vector<int> cl(n);
        std::vector<uint32_t> sym(n), bits(n);
        std::vector<uint8_t> nb(n);
        for (size_t i = 0; i < n; ++i) {
            int nbi;
            symbolize(ts[i], cl[i], sym[i], nbi, bits[i]);
            nb[i] = (uint8_t) nbi;
        }
        if (o.use_prefix) {
            for (size_t i = 0; i < n; ++i) {
                const auto &L = plen[(size_t) cl[i]];
                JG_CHECK(sym[i] < L.size());
                int len = L[sym[i]];
                JG_CHECK(len > 0 || L.size() == 1 || count_nonzero(L) == 1);
                if (count_nonzero(L) > 1) bw.put(pcode[(size_t) cl[i]][sym[i]], len);
                bw.put(bits[i], nb[i]);
            }
            return;
        }
        // rANS: encode in reverse, then lay out forward (App. E.11)
        std::vector<uint32_t> word(n);
        std::vector<uint8_t> has_word(n, 0);
        uint32_t state = 0x130000;
        for (size_t k = n; k-- > 0;) {
            const auto &Dc = D[(size_t) cl[k]];
            uint32_t d = Dc[sym[k]];
            JG_CHECK(d > 0);
            if ((state >> 20) >= d) {
                has_word[k] = 1;
                word[k] = state & 0xffff;
                state >>= 16;
            }
            uint32_t q = state / d, r = state % d;
            state = (q << 12) + slot[(size_t) cl[k]][cum[(size_t) cl[k]][sym[k]] + r];
        }
        bw.put(state & 0xffff, 16);
        bw.put(state >> 16, 16);
        for (size_t k = 0; k < n; ++k) {
            if (has_word[k]) bw.put(word[k], 16);
            bw.put(bits[k], nb[k]);
        }
    }

    // approximate cost in bits of a stream under this spec (for statistics only)
    double cost_bits(const TokStream &ts) const {
        double bitsum = 0;
        for (const Tok &t : ts) {
            int cl, nb; uint32_t sym, b;
            symbolize(t, cl, sym, nb, b);
            if (o.use_prefix) bitsum += (count_nonzero(plen[(size_t) cl]) > 1 ? plen[(size_t) cl][sym] : 0) + nb;
            else bitsum += -std::log2((double) D[(size_t) cl][sym] / 4096.0) + nb;
        }
        return bitsum;
    }

    static void write_hybrid_cfg(BitWriter &bw, HybridCfg c, int log_alpha) {
        bw.at_most((uint32_t) c.split_exp, (uint32_t) log_alpha);
        if (c.split_exp != log_alpha) {
            bw.at_most((uint32_t) c.msb, (uint32_t) c.split_exp);
            bw.at_most((uint32_t) c.lsb, (uint32_t) (c.split_exp - c.msb));
        } else {
            JG_CHECK(c.msb == 0 && c.lsb == 0);
        }
    }

private:
    static int count_nonzero(const std::vector<uint8_t> &L) {
        int n = 0;
        for (uint8_t l : L) n += l != 0;
        return n;
    }

    static double entropy_bits(const std::vector<uint32_t> &h) {
        double tot = 0, acc = 0;
        for (uint32_t c : h) tot += c;
        if (tot == 0) return 0;
        for (uint32_t c : h) if (c) acc -= (double) c * std::log2((double) c / tot);
        return acc;
    }

    void cluster(const std::vector<std::vector<uint32_t>> &hist) {
        int nctx = (int) hist.size();
        std::vector<int> order;
        std::vector<uint64_t> tot((size_t) nctx, 0);
        for (int c = 0; c < nctx; ++c) {
            for (uint32_t v : hist[(size_t) c]) tot[(size_t) c] += v;
            if (tot[(size_t) c]) order.push_back(c);
        }
        std::sort(order.begin(), order.end(), [&](int a, int b) { return tot[(size_t) a] != tot[(size_t) b] ? tot[(size_t) a] > tot[(size_t) b] : a < b; });
        std::vector<std::vector<uint32_t>> ch;
        std::vector<double> ce;
        int maxc = std::max(1, std::min(o.max_clusters, 256));
        for (int c : order) {
            const auto &h = hist[(size_t) c];
            double he = entropy_bits(h);
            int best = -1;
            double bestcost = 1e300;
            for (size_t k = 0; k < ch.size(); ++k) {
                std::vector<uint32_t> m = ch[k];
                if (m.size() < h.size()) m.resize(h.size(), 0);
                for (size_t s = 0; s < h.size(); ++s) m[s] += h[s];
                double cost = entropy_bits(m) - ce[k] - he;
                if (cost < bestcost) { bestcost = cost; best = (int) k; }
            }
            double newcost = 40.0 + 6.0 * (double) h.size(); // rough header cost of one more histogram
            if ((int) ch.size() < maxc && (best < 0 || bestcost > newcost)) {
                cmap[(size_t) c] = (uint8_t) ch.size();
                ch.push_back(h);
                ce.push_back(he);
            } else {
                cmap[(size_t) c] = (uint8_t) best;
                auto &m = ch[(size_t) best];
                if (m.size() < h.size()) m.resize(h.size(), 0);
                for (size_t s = 0; s < h.size(); ++s) m[s] += h[s];
                ce[(size_t) best] = entropy_bits(m);
            }
        }
        nclusters = std::max<int>(1, (int) ch.size());
        // unused contexts go to cluster 0 (already 0)
    }

    // ---------------- ANS ----------------
    void build_ans(const std::vector<std::vector<uint32_t>> &ch) {
        int las = o.log_alpha_size, tsize = 1 << las;
        D.assign((size_t) nclusters, std::vector<uint16_t>((size_t) tsize, 0));
        slot.assign((size_t) nclusters, std::vector<uint16_t>(4096, 0));
        cum.assign((size_t) nclusters, std::vector<uint32_t>((size_t) tsize + 1, 0));
        for (int c = 0; c < nclusters; ++c) {
            normalize(ch[(size_t) c], D[(size_t) c]);
            for (int s = 0; s < tsize; ++s) cum[(size_t) c][(size_t) s + 1] = cum[(size_t) c][(size_t) s] + D[(size_t) c][(size_t) s];
            JG_CHECK(cum[(size_t) c][(size_t) tsize] == 4096);
            build_alias_inverse(D[(size_t) c], las, cum[(size_t) c], slot[(size_t) c]);
        }
    }

    void normalize(const std::vector<uint32_t> &h, std::vector<uint16_t> &out) const {
        int tsize = (int) out.size();
        uint64_t tot = 0;
        int nz = 0;
        for (uint32_t v : h) { tot += v; nz += v != 0; }
        std::fill(out.begin(), out.end(), 0);
        if (tot == 0) { out[0] = 4096; return; }
        JG_CHECK((int) h.size() <= tsize);
        if (nz == 1) {
            for (size_t s = 0; s < h.size(); ++s) if (h[s]) out[s] = 4096;
            return;
        }
        // precision actually representable by the histogram coding for a given exponent
        auto quant = [&](int v) {
            if (v < 2) return v;
            int e = floor_lg((uint32_t) v);
            int bitcount = std::min(std::max(0, o.ans_precision_shift - ((12 - e) >> 1)), e);
            int drop = e - bitcount;
            return (v >> drop) << drop;
        };
        std::vector<int> d(h.size(), 0);
        int sum = 0, imax = 0;
        for (size_t s = 0; s < h.size(); ++s) {
            if (!h[s]) continue;
            double x = (double) h[s] * 4096.0 / (double) tot;
            int v = std::max(1, (int) std::floor(x + 0.5));
            v = std::max(1, quant(std::min(v, 4095)));
            d[s] = v;
            sum += v;
            if (h[s] > h[(size_t) imax] || !h[(size_t) imax]) imax = (int) s;
        }
        // the most probable symbol absorbs the rounding error (it is written as the omitted count
        // if it also has the largest log-count; see write_ans_table which re-checks this)
        int rest = 4096 - (sum - d[(size_t) imax]);
        if (rest < 1) {
            // pathological: too many symbols forced to >= 1; shave the larger ones
            while (rest < 1) {
                int j = -1;
                for (size_t s = 0; s < d.size(); ++s) if ((int) s != imax && d[s] > 1 && (j < 0 || d[s] > d[(size_t) j])) j = (int) s;
                JG_CHECK(j >= 0);
                int nv = std::max(1, quant(d[(size_t) j] - 1 > 0 ? d[(size_t) j] / 2 : 1));
                if (nv >= d[(size_t) j]) nv = d[(size_t) j] - 1;
                nv = std::max(1, quant(nv));
                rest += d[(size_t) j] - nv;
                d[(size_t) j] = nv;
            }
        }
        d[(size_t) imax] = rest;
        for (size_t s = 0; s < d.size(); ++s) out[s] = (uint16_t) d[s];
    }

    // builds the decoder's alias table per the format definition, then inverts it:
    // slot[cum[s] + r] = the 12-bit value whose decode yields (symbol s, offset r)
    static void build_alias_inverse(const std::vector<uint16_t> &Dc, int las, const std::vector<uint32_t> &cumc, std::vector<uint16_t> &slotc) {
        int tsize = 1 << las, lbs = 12 - las, bsize = 1 << lbs;
        std::vector<int> cutoff((size_t) tsize), offnext((size_t) tsize, 0), symbol((size_t) tsize, 0);
        int single = -1, nzc = 0;
        for (int i = 0; i < tsize; ++i) if (Dc[(size_t) i]) { ++nzc; single = i; }
        if (nzc == 1) {
            for (int j = 0; j < tsize; ++j) { symbol[(size_t) j] = single; offnext[(size_t) j] = j << lbs; cutoff[(size_t) j] = 0; }
        } else {
            int u = -1, ov = -1;
            for (int i = 0; i < tsize; ++i) {
                int c = Dc[(size_t) i];
                cutoff[(size_t) i] = c;
                if (c > bsize) { offnext[(size_t) i] = ov; ov = i; }
                else if (c < bsize) { offnext[(size_t) i] = u; u = i; }
                else { symbol[(size_t) i] = i; offnext[(size_t) i] = 0; }
            }
            while (ov >= 0) {
                JG_CHECK(u >= 0);
                int by = bsize - cutoff[(size_t) u];
                int tmp = offnext[(size_t) u];
                cutoff[(size_t) ov] -= by;
                symbol[(size_t) u] = ov;
                offnext[(size_t) u] = cutoff[(size_t) ov] - cutoff[(size_t) u];
                u = tmp;
                if (cutoff[(size_t) ov] < bsize) {
                    tmp = offnext[(size_t) ov];
                    offnext[(size_t) ov] = u;
                    u = ov;
                    ov = tmp;
                } else if (cutoff[(size_t) ov] == bsize) {
                    tmp = offnext[(size_t) ov];
                    symbol[(size_t) ov] = ov;
                    offnext[(size_t) ov] = 0;
                    ov = tmp;
                }
            }
            JG_CHECK(u < 0);
        }
        std::vector<uint8_t> seen(4096, 0);
        for (int idx = 0; idx < 4096; ++idx) {
            int i = idx >> lbs, pos = idx & (bsize - 1);
            int s = pos < cutoff[(size_t) i] ? i : symbol[(size_t) i];
            int off = pos < cutoff[(size_t) i] ? 0 : offnext[(size_t) i];
            int r = off + pos;
            JG_CHECK(r >= 0 && r < Dc[(size_t) s]);
            uint32_t k = cumc[(size_t) s] + (uint32_t) r;
            JG_CHECK(!seen[k]);
            seen[k] = 1;
            slotc[k] = (uint16_t) idx;
        }
    }

    void write_ans_table(BitWriter &bw, const std::vector<uint16_t> &Dc) const {
        int tsize = (int) Dc.size();
        std::vector<int> nzs;
        for (int i = 0; i < tsize; ++i) if (Dc[(size_t) i]) nzs.push_back(i);
        if (nzs.size() == 1) { // "one entry" mode
            bw.put(1, 2);
            bw.u8((uint32_t) nzs[0]);
            return;
        }
        if (nzs.size() == 2 && nzs[1] < tsize) {
            bw.put(3, 2);
            bw.u8((uint32_t) nzs[0]);
            bw.u8((uint32_t) nzs[1]);
            bw.put(Dc[(size_t) nzs[0]], 12);
            return;
        }
        { // flat mode if it happens to match exactly
            int alpha = nzs.back() + 1;
            bool flat = (int) nzs.size() == alpha;
            int d = 4096 / alpha, bias = 4096 % alpha;
            for (int i = 0; flat && i < alpha; ++i) flat = Dc[(size_t) i] == (i < bias ? d + 1 : d);
            if (flat) {
                bw.put(2, 2);
                bw.u8((uint32_t) (alpha - 1));
                return;
            }
        }
        // general mode
        bw.put(0, 2);
        int shift = o.ans_precision_shift;
        JG_CHECK(shift >= 0 && shift <= 13);
        { // shift is coded as u(len) + (1 << len) - 1 with a unary-ish len
            int len = floor_lg((uint32_t) shift + 1);
            JG_CHECK(len <= 3);
            // len = u(1) ? u(1) ? u(1) ? 3 : 2 : 1 : 0
            for (int i = 0; i < len; ++i) bw.bit(1);
            if (len < 3) bw.bit(0);
            bw.put((uint64_t) (shift + 1 - (1 << len)), len);
        }
        int alpha = std::max(3, nzs.back() + 1);
        JG_CHECK(alpha <= tsize);
        bw.u8((uint32_t) (alpha - 3));
        // log-count codes; the first symbol with the largest code is the implicit one
        std::vector<int> code((size_t) alpha);
        int maxcode = 0, omit = -1;
        for (int i = 0; i < alpha; ++i) {
            int v = Dc[(size_t) i];
            code[(size_t) i] = v == 0 ? 0 : floor_lg((uint32_t) v) + 1;
            if (code[(size_t) i] > maxcode) { maxcode = code[(size_t) i]; omit = i; }
        }
        JG_CHECK(maxcode <= 12 || true);
        if (maxcode > 12) { // a count of 4096 cannot occur here (>= 3 symbols), 2048..4095 -> code 12
            throw GenError("ans table: log-count too large");
        }
        // RLE plan: runs of identical counts (>= 5 long) become value + RLE(run-1 in 4..258)
        static const uint16_t LC_BITS[14] = {17, 11, 15, 3, 9, 7, 4, 2, 5, 6, 0, 33, 1, 65};
        static const uint8_t LC_LEN[14] = {5, 4, 4, 4, 4, 4, 3, 3, 3, 3, 3, 6, 7, 7};
        struct Item { int code; int rep; };
        std::vector<Item> items; // rep > 0 => RLE item
        for (int i = 0; i < alpha;) {
            items.push_back({code[(size_t) i], 0});
            int j = i + 1;
            // a run may not start at the omitted position (its stored value is -1 in the reader)
            // and may not cover it either
            if (i != omit) {
                while (j < alpha && Dc[(size_t) j] == Dc[(size_t) i] && j != omit) ++j;
            }
            int run = j - (i + 1);
            int consumed = 0;
            while (run - consumed >= 4) {
                int rep = std::min(run - consumed, 4 + 255);
                items.push_back({13, rep});
                consumed += rep;
            }
            i += 1 + consumed;
        }
        for (const Item &it : items) {
            bw.put(LC_BITS[it.code], LC_LEN[it.code]);
            if (it.rep) bw.u8((uint32_t) (it.rep - 4));
        }
        // mantissas
        {
            int n = 0;
            bool omitted = false;
            for (const Item &it : items) {
                if (it.rep) { n += it.rep; continue; }
                int c = it.code;
                if (c == maxcode && !omitted) { omitted = true; JG_CHECK(n == omit); ++n; continue; }
                if (c >= 2) {
                    int e = c - 1;
                    int bitcount = std::min(std::max(0, shift - ((12 - e) >> 1)), e);
                    int v = Dc[(size_t) n];
                    int mant = v - (1 << e);
                    JG_CHECK((mant & ((1 << (e - bitcount)) - 1)) == 0);
                    bw.put((uint64_t) (mant >> (e - bitcount)), bitcount);
                }
                ++n;
            }
        }
    }

    // ---------------- prefix codes ----------------
    void build_prefix(const std::vector<std::vector<uint32_t>> &ch) {
        plen.assign((size_t) nclusters, {});
        pcode.assign((size_t) nclusters, {});
        for (int c = 0; c < nclusters; ++c) {
            std::vector<uint32_t> h = ch[(size_t) c];
            if (h.empty()) h.assign(1, 0);
            std::vector<uint8_t> L(h.size(), 0);
            huffman_lengths(h, L, 15);
            plen[(size_t) c] = L;
            pcode[(size_t) c] = canonical_codes(L);
        }
    }

    static void huffman_lengths(const std::vector<uint32_t> &h, std::vector<uint8_t> &L, int maxlen) {
        std::vector<int> syms;
        for (size_t s = 0; s < h.size(); ++s) if (h[s]) syms.push_back((int) s);
        std::fill(L.begin(), L.end(), 0);
        if (syms.empty()) return;        // alphabet of one (unused) symbol
        if (syms.size() == 1) { L[(size_t) syms[0]] = 1; return; } // written as a 1-symbol simple code (0 bits)
        std::vector<uint64_t> w;
        for (int s : syms) w.push_back(h[(size_t) s]);
        for (int iter = 0;; ++iter) {
            // plain Huffman on w
            struct Node { uint64_t w; int l, r; };
            std::vector<Node> nodes;
            typedef std::pair<uint64_t, int> P;
            std::priority_queue<P, std::vector<P>, std::greater<P>> pq;
            for (size_t i = 0; i < w.size(); ++i) { nodes.push_back({w[i], -1, -1}); pq.push({w[i], (int) i}); }
            while (pq.size() > 1) {
                P a = pq.top(); pq.pop();
                P b = pq.top(); pq.pop();
                nodes.push_back({a.first + b.first, a.second, b.second});
                pq.push({a.first + b.first, (int) nodes.size() - 1});
            }
            std::vector<int> depth(nodes.size(), 0);
            int mx = 0;
            for (int i = (int) nodes.size() - 1; i >= 0; --i) {
                if (nodes[(size_t) i].l >= 0) {
                    depth[(size_t) nodes[(size_t) i].l] = depth[(size_t) i] + 1;
                    depth[(size_t) nodes[(size_t) i].r] = depth[(size_t) i] + 1;
                }
            }
            for (size_t i = 0; i < w.size(); ++i) mx = std::max(mx, depth[i]);
            if (mx <= maxlen) {
                for (size_t i = 0; i < w.size(); ++i) L[(size_t) syms[i]] = (uint8_t) depth[i];
                return;
            }
            // flatten and retry
            uint64_t tot = 0;
            for (uint64_t x : w) tot += x;
            uint64_t floorw = std::max<uint64_t>(1, tot >> (maxlen - 2 - std::min(iter, 8)));
            for (uint64_t &x : w) x = std::max(x, floorw);
        }
    }

    // canonical code per RFC 7932 §3.2 (shorter codes first, ties by symbol), returned bit-reversed
    // for LSB-first emission
    static std::vector<uint16_t> canonical_codes(const std::vector<uint8_t> &L) {
        std::vector<uint16_t> out(L.size(), 0);
        int blcount[17] = {0}, next[17] = {0};
        for (uint8_t l : L) blcount[l]++;
        blcount[0] = 0;
        int code = 0;
        for (int b = 1; b <= 16; ++b) { code = (code + blcount[b - 1]) << 1; next[b] = code; }
        for (size_t s = 0; s < L.size(); ++s) {
            int l = L[s];
            if (!l) continue;
            int c = next[l]++;
            int r = 0;
            for (int i = 0; i < l; ++i) r |= ((c >> i) & 1) << (l - 1 - i);
            out[s] = (uint16_t) r;
        }
        return out;
    }

    void write_prefix_tree(BitWriter &bw, const std::vector<uint8_t> &L) const {
        int size = (int) L.size();
        if (size <= 1) return; // alphabet size 1: zero-bit code, nothing written
        std::vector<int> used;
        for (int s = 0; s < size; ++s) if (L[(size_t) s]) used.push_back(s);
        int symbits = ceil_lg((uint32_t) size);
        if (used.empty()) used.push_back(0);
        if (used.size() <= 4) {
            // simple code (hskip == 1); the code lengths must be one of the fixed shapes
            bool ok = true;
            std::vector<int> syms = used;
            int nsym = (int) syms.size();
            int tree_select = 0;
            if (nsym == 1) {
            } else if (nsym == 2) {
                ok = L[(size_t) syms[0]] == 1 && L[(size_t) syms[1]] == 1;
            } else if (nsym == 3) {
                std::stable_sort(syms.begin(), syms.end(), [&](int a, int b) { return L[(size_t) a] < L[(size_t) b]; });
                ok = L[(size_t) syms[0]] == 1 && L[(size_t) syms[1]] == 2 && L[(size_t) syms[2]] == 2;
            } else {
                std::stable_sort(syms.begin(), syms.end(), [&](int a, int b) { return L[(size_t) a] < L[(size_t) b]; });
                if (L[(size_t) syms[0]] == 2 && L[(size_t) syms[3]] == 2) tree_select = 0;
                else if (L[(size_t) syms[0]] == 1 && L[(size_t) syms[1]] == 2 && L[(size_t) syms[2]] == 3 && L[(size_t) syms[3]] == 3) tree_select = 1;
                else ok = false;
            }
            if (ok) {
                bw.put(1, 2);
                bw.put((uint64_t) (nsym - 1), 2);
                for (int s : syms) bw.put((uint64_t) s, symbits);
                if (nsym == 4) bw.bit(tree_select);
                return;
            }
        }
        // complex code: code-length code over 18 symbols
        // 1. RLE-encode the length sequence (trailing zeros dropped)
        int last = size;
        while (last > 0 && L[(size_t) last - 1] == 0) --last;
        struct CL { int sym; int extra; int nextra; };
        std::vector<CL> seq;
        int prev = 8;
        for (int i = 0; i < last;) {
            int v = L[(size_t) i];
            int j = i;
            while (j < last && L[(size_t) j] == v) ++j;
            int run = j - i;
            if (v == 0) {
                // zeros: code 17 repeats 3..10; chained 17s multiply (avoid chaining: emit separately
                // only when a literal sits in between, so use one 17 of up to 10 then literals/more 17s
                // separated by a literal 0)
                while (run > 0) {
                    if (run >= 3) {
                        int r = std::min(run, 10);
                        seq.push_back({17, r - 3, 3});
                        run -= r;
                        if (run > 0) { seq.push_back({0, 0, 0}); run -= 1; }
                    } else { seq.push_back({0, 0, 0}); run -= 1; }
                }
            } else {
                if (v != prev) { seq.push_back({v, 0, 0}); run -= 1; prev = v; }
                while (run > 0) {
                    if (run >= 3) {
                        int r = std::min(run, 6);
                        seq.push_back({16, r - 3, 2});
                        run -= r;
                        if (run > 0) { seq.push_back({v, 0, 0}); run -= 1; }
                    } else { seq.push_back({v, 0, 0}); run -= 1; }
                }
            }
            i = j;
        }
        // 2. Huffman code (max length 5) over the 18 code-length symbols
        std::vector<uint32_t> h1(18, 0);
        for (const CL &c : seq) h1[(size_t) c.sym]++;
        std::vector<uint8_t> L1(18, 0);
        huffman_lengths(h1, L1, 5);
        int nz1 = 0;
        for (uint8_t l : L1) nz1 += l != 0;
        if (nz1 == 1) {
            // a lone code-length symbol: the reader accepts a single length-4 ... no: it needs the
            // Kraft sum to hit 32; give a dummy second symbol instead
            for (int s = 0; s < 18; ++s) if (!L1[(size_t) s]) { L1[(size_t) s] = 1; break; }
            for (int s = 0; s < 18; ++s) if (h1[(size_t) s]) L1[(size_t) s] = 1;
        }
        std::vector<uint16_t> C1 = canonical_codes(L1);
        static const uint8_t ZIGZAG[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
        static const uint8_t L0_BITS[6] = {0, 7, 3, 2, 1, 15};
        static const uint8_t L0_LEN[6] = {2, 4, 3, 2, 2, 4};
        int hskip = 0;
        if (L1[ZIGZAG[0]] == 0 && L1[ZIGZAG[1]] == 0) {
            hskip = 2;
            if (L1[ZIGZAG[2]] == 0) hskip = 3;
        }
        bw.put((uint64_t) hskip, 2); // 0, 2 or 3 (1 would mean a simple code)
        {
            int total = 0;
            for (int i = hskip; i < 18 && total < 32; ++i) {
                int l = L1[ZIGZAG[i]];
                bw.put(L0_BITS[l], L0_LEN[l]);
                if (l) total += 32 >> l;
            }
            JG_CHECK(total == 32);
        }
        // 3. the lengths themselves
        {
            int total = 0;
            int prevl = 8;
            for (size_t k = 0; k < seq.size() && total < 32768; ++k) {
                const CL &c = seq[k];
                bw.put(C1[(size_t) c.sym], L1[(size_t) c.sym]);
                if (c.nextra) bw.put((uint64_t) c.extra, c.nextra);
                if (c.sym < 16) { if (c.sym) { total += 32768 >> c.sym; prevl = c.sym; } }
                else if (c.sym == 16) total += (32768 >> prevl) * (c.extra + 3);
            }
            JG_CHECK(total == 32768);
        }
    }

    // ---------------- cluster map ----------------
    static void write_cluster_map(BitWriter &bw, const std::vector<uint8_t> &map, int num_dist, int nclusters, int mode) {
        JG_CHECK((int) map.size() == num_dist);
        if (num_dist == 1) return;
        int nbits = ceil_lg((uint32_t) nclusters);
        if (mode == 0 && nbits <= 3 && (nbits * num_dist < 300 || nclusters == 1)) {
            bw.bit(1);
            bw.put((uint64_t) nbits, 2);
            for (int i = 0; i < num_dist; ++i) bw.put(map[(size_t) i], nbits);
            return;
        }
        bw.bit(0);
        bool use_mtf = mode != 2;
        bw.bit(use_mtf);
        std::vector<uint32_t> vals((size_t) num_dist);
        if (use_mtf) {
            uint8_t mtf[256];
            for (int i = 0; i < 256; ++i) mtf[i] = (uint8_t) i;
            for (int i = 0; i < num_dist; ++i) {
                int j = 0;
                while (mtf[j] != map[(size_t) i]) ++j;
                vals[(size_t) i] = (uint32_t) j;
                uint8_t moved = mtf[j];
                for (; j > 0; --j) mtf[j] = mtf[j - 1];
                mtf[0] = moved;
            }
        } else {
            for (int i = 0; i < num_dist; ++i) vals[(size_t) i] = map[(size_t) i];
        }
        TokStream ts;
        for (uint32_t v : vals) ts.push_back({0, v, 0});
        EntropyOpts no;
        no.use_prefix = false;
        no.log_alpha_size = 8;
        no.cfg = {8, 0, 0};
        // with up to 256 cluster ids an 8-bit alphabet always suffices
        no.max_clusters = 1;
        CodeSpec nested;
        std::vector<const TokStream *> v1{&ts};
        nested.build(1, no, v1);
        nested.write(bw);
        nested.encode(bw, ts);
    }
};

} // namespace jxlgen
