// j40-b200: host-side front end (see j40b_host.h). Everything here is bit-twiddling over a few kilobytes per
// image: the codestream container, image metadata, frame header, TOC, LfGlobal and HfGlobal, with the same
// acceptance rules and four-character error codes as the reference (file:line citations below), emitting
// device-ready tables. Compiled with -ffp-contract=off: the float tables must match the reference bit for bit.
#include "j40b_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <mutex>

namespace j40b {

namespace {

#define FAIL(code) do { if (!err) err = overrun() ? (uint32_t) E_SHRT : (uint32_t) (code); return; } while (0)
#define FAILV(code, v) do { if (!err) err = overrun() ? (uint32_t) E_SHRT : (uint32_t) (code); return (v); } while (0)
#define CHECK(cond, code) do { if (err) return; if (!(cond)) FAIL(code); } while (0)
#define CHECKV(cond, code, v) do { if (err) return (v); if (!(cond)) FAILV(code, v); } while (0)
#define E4(s) J40B_4CC((s)[0], (s)[1], (s)[2], (s)[3])

struct Limits { // Main profile, level 5 (j40.h:1165-1170)
    static constexpr int64_t pixels = 1 << 28;
    static constexpr int32_t width = 1 << 18, height = 1 << 18;
    static constexpr uint64_t icc_output_size = 1u << 22;
    static constexpr int32_t bpp = 16, num_extra_channels = 4, nb_transforms = 8, nb_channels_tr = 256, tree_depth = 64;
};

struct Parser {
    BitReader br;
    uint32_t err = 0;
    FramePlan &plan;
    Arena &arena;

    explicit Parser(FramePlan &p) : plan(p), arena(p.arena) {}

    bool overrun() const { return br.overrun(); }
    uint32_t u(int n) { return br.u(n); }
    uint32_t u32(uint32_t o0, int n0, uint32_t o1, int n1, uint32_t o2, int n2, uint32_t o3, int n3) { return br.u32(o0, n0, o1, n1, o2, n2, o3, n3); }
    uint64_t u64() { // j40.h:1966
        uint32_t sel = u(2);
        uint64_t ret = u((int) sel * 4);
        if (sel < 3) ret += 17u >> (8 - sel * 4);
        else for (int shift = 12; shift < 64 && u(1); shift += 8) ret |= (uint64_t) u(shift < 56 ? 8 : 64 - shift) << shift;
        return ret;
    }
    int32_t enum_() { // j40.h:1979
        int32_t v = (int32_t) u32(0, 0, 1, 0, 2, 4, 18, 6);
        if (v >= 31) { if (!err) err = overrun() ? (uint32_t) E_SHRT : E4("enum"); return 0; }
        return v;
    }
    float f16() { // j40.h:1987
        int32_t bits = (int32_t) u(16);
        int32_t biased_exp = (bits >> 10) & 0x1f;
        if (biased_exp == 31) { if (!err) err = overrun() ? (uint32_t) E_SHRT : E4("!fin"); return 0.0f; }
        return (float) (bits >> 15 ? -1 : 1) * ldexpf((float) ((bits & 0x3ff) | (biased_exp > 0 ? 0x400 : 0)), biased_exp - 25);
    }
    int32_t u8v() { if (u(1)) { int n = (int) u(3); return (int32_t) u(n) + (1 << n); } return 0; }
    int32_t at_most(int32_t max) {
        int32_t v = max > 0 ? (int32_t) u(ceil_lg32((uint32_t) max + 1)) : 0;
        if (v > max) { if (!err) err = overrun() ? (uint32_t) E_SHRT : E4("rnge"); return 0; }
        return v;
    }
    void zero_pad() { if (err) return; if (!br.zero_pad_to_byte()) FAIL(E_PAD0); }
    void check_overrun() { if (!err && overrun()) err = E_SHRT; }
    void skip_bits(uint64_t n) { while (n > 0 && !overrun()) { int k = n > 32 ? 32 : (int) n; u(k); n -= (uint64_t) k; } }

    // ------------------------------------------------------------------------------------------
    // entropy code specs (j40.h:2526-2777)

    HybridCfg hybrid_cfg(int log_alpha_size) {
        HybridCfg c;
        c.pad = 0;
        c.split_exp = (int8_t) at_most(log_alpha_size);
        if (c.split_exp != log_alpha_size) {
            c.msb_in_token = (int8_t) at_most(c.split_exp);
            c.lsb_in_token = (int8_t) at_most(c.split_exp - c.msb_in_token);
        } else {
            c.msb_in_token = c.lsb_in_token = 0;
        }
        c.max_token = (1 << c.split_exp) + ((30 - c.split_exp) << (c.lsb_in_token + c.msb_in_token)) - 1;
        return c;
    }

    struct PCode { int32_t sym; int len; uint32_t lsb; }; // one code word, bits in stream (LSB-first) order

    // two-level LUT; returns arena offset and sets root_bits/max_len
    uint32_t build_prefix_lut(const std::vector<PCode> &codes, int max_len, int16_t *root_bits_out) {
        const int ROOT = 9;
        int root_bits = std::min(max_len, ROOT);
        std::vector<uint32_t> lut((size_t) 1 << root_bits, 0);
        // group long codes by their first root_bits bits
        std::vector<int> sub_len((size_t) 1 << root_bits, 0);
        for (const PCode &c : codes) if (c.len > root_bits) {
            uint32_t pre = c.lsb & ((1u << root_bits) - 1);
            sub_len[pre] = std::max(sub_len[pre], c.len - root_bits);
        }
        std::vector<uint32_t> sub_off((size_t) 1 << root_bits, 0);
        for (size_t pre = 0; pre < sub_len.size(); ++pre) if (sub_len[pre]) {
            sub_off[pre] = (uint32_t) lut.size();
            lut.resize(lut.size() + ((size_t) 1 << sub_len[pre]), 0);
            lut[pre] = 0x8000u | ((uint32_t) sub_len[pre] << 5) | (sub_off[pre] << 16);
        }
        for (const PCode &c : codes) {
            if (c.len <= root_bits) {
                for (uint32_t k = c.lsb; k < (1u << root_bits); k += 1u << c.len) lut[k] = ((uint32_t) c.sym << 16) | (uint32_t) c.len;
            } else {
                uint32_t pre = c.lsb & ((1u << root_bits) - 1), rest = c.lsb >> root_bits;
                int sl = sub_len[pre], rl = c.len - root_bits;
                for (uint32_t k = rest; k < (1u << sl); k += 1u << rl) lut[sub_off[pre] + k] = ((uint32_t) c.sym << 16) | (uint32_t) c.len;
            }
        }
        uint32_t off = arena.alloc(lut.size() * 4, 8);
        memcpy(arena.at<uint8_t>(off), lut.data(), lut.size() * 4);
        *root_bits_out = (int16_t) root_bits;
        return off;
    }

    static uint32_t bitrev(uint32_t v, int n) { uint32_t r = 0; for (int i = 0; i < n; ++i) r |= ((v >> i) & 1) << (n - 1 - i); return r; }

    // canonical code (shorter first, then by symbol), words reversed into stream order
    static std::vector<PCode> canonical(const std::vector<int> &len) {
        std::vector<PCode> out;
        int count[17] = {0}, next[17] = {0};
        for (int l : len) count[l]++;
        count[0] = 0;
        int code = 0;
        for (int b = 1; b <= 16; ++b) { code = (code + count[b - 1]) << 1; next[b] = code; }
        for (size_t s = 0; s < len.size(); ++s) if (len[s]) {
            int l = len[s];
            out.push_back({(int32_t) s, l, bitrev((uint32_t) next[l]++, l)});
        }
        return out;
    }

    // RFC 7932 §3 prefix code over an alphabet of l2size symbols (j40.h:2049-2242)
    void prefix_code_tree(int32_t l2size, DCluster &cl) {
        cl.root_bits = 0;
        cl.max_len = 0;
        if (l2size == 1) { // zero-bit code for symbol 0
            std::vector<PCode> one{{0, 0, 0}};
            cl.table_off = build_prefix_lut(one, 0, &cl.root_bits);
            return;
        }
        uint32_t hskip = u(2);
        if (hskip == 1) { // simple codes
            int nsym = (int) u(2) + 1;
            int32_t syms[4] = {0, 0, 0, 0};
            for (int i = 0; i < nsym; ++i) {
                syms[i] = at_most(l2size - 1);
                for (int j = 0; j < i; ++j) CHECK(syms[i] != syms[j], E4("hufd"));
            }
            int tree_select = 0;
            if (nsym == 4) tree_select = (int) u(1);
            if (err) return;
            std::vector<PCode> codes;
            int max_len = 0;
            if (nsym == 1) {
                codes.push_back({syms[0], 0, 0});
            } else if (nsym == 2) {
                std::sort(syms, syms + 2);
                codes.push_back({syms[0], 1, 0}); codes.push_back({syms[1], 1, 1});
                max_len = 1;
            } else if (nsym == 3) {
                std::sort(syms + 1, syms + 3);
                codes.push_back({syms[0], 1, 0}); codes.push_back({syms[1], 2, 1}); codes.push_back({syms[2], 2, 3});
                max_len = 2;
            } else if (!tree_select) {
                // the reference maps the 2-bit LSB-first value k straight to the k-th smallest symbol
                // (j40.h:2093, TEMPLATES[4]); kept as is for parity (DESIGN.md quirk list)
                std::sort(syms, syms + 4);
                for (int k = 0; k < 4; ++k) codes.push_back({syms[k], 2, (uint32_t) k});
                max_len = 2;
            } else {
                std::sort(syms + 2, syms + 4);
                codes.push_back({syms[0], 1, 0}); codes.push_back({syms[1], 2, 1});
                codes.push_back({syms[2], 3, 3}); codes.push_back({syms[3], 3, 7});
                max_len = 3;
            }
            cl.max_len = (int16_t) max_len;
            cl.table_off = build_prefix_lut(codes, max_len, &cl.root_bits);
            return;
        }
        // complex codes: 18 code-length code lengths, coded with a fixed variable-length code
        static const uint8_t ZIGZAG[18] = {1, 2, 3, 4, 0, 5, 17, 6, 16, 7, 8, 9, 10, 11, 12, 13, 14, 15};
        // fixed code: symbol -> (stream bits, length): 0:(00) 4:(10) 3:(01) 2:(110) 1:(1110) 5:(1111), LSB-first
        auto read_l0 = [&]() -> int {
            uint32_t b = br.peek(4);
            int sym, len;
            if ((b & 3) == 0) { sym = 0; len = 2; }
            else if ((b & 3) == 1) { sym = 4; len = 2; }
            else if ((b & 3) == 2) { sym = 3; len = 2; }
            else if ((b & 7) == 3) { sym = 2; len = 3; }
            else if ((b & 15) == 7) { sym = 1; len = 4; }
            else { sym = 5; len = 4; }
            br.skip(len);
            return sym;
        };
        std::vector<int> l1len(18, 0);
        int total = 0, zeros = (int) hskip, i;
        for (i = (int) hskip; i < 18 && total < 32; ++i) {
            int c = read_l0();
            l1len[ZIGZAG[i]] = c;
            if (c) total += 32 >> c; else ++zeros;
        }
        CHECK(total == 32 && zeros != i, E4("hufd"));
        // layer-1 decoding table (5-bit LUT)
        uint32_t l1lut[32];
        {
            std::vector<PCode> c1 = canonical(l1len);
            for (const PCode &c : c1) for (uint32_t k = c.lsb; k < 32; k += 1u << c.len) l1lut[k] = ((uint32_t) c.sym << 8) | (uint32_t) c.len;
        }
        std::vector<int> l2len((size_t) l2size, 0);
        {
            int prev = 8, prev_rep = 0, n = 0;
            total = 0;
            while (n < l2size && total < 32768) {
                uint32_t e = l1lut[br.peek(5)];
                br.skip((int) (e & 0xff));
                int code = (int) (e >> 8);
                if (code < 16) {
                    l2len[(size_t) n++] = code;
                    if (code) { total += 32768 >> code; prev = code; }
                    prev_rep = 0;
                } else if (code == 16) {
                    if (prev_rep < 0) prev_rep = 0;
                    int rep = (prev_rep > 0 ? 4 * prev_rep - 5 : 3) + (int) u(2);
                    CHECK(n + (rep - prev_rep) <= l2size, E4("hufd"));
                    total += (32768 * (rep - prev_rep)) >> prev;
                    for (; prev_rep < rep; ++prev_rep) l2len[(size_t) n++] = prev;
                } else {
                    if (prev_rep > 0) prev_rep = 0;
                    int rep = (prev_rep < 0 ? 8 * prev_rep + 13 : -3) - (int) u(3);
                    CHECK(n + (prev_rep - rep) <= l2size, E4("hufd"));
                    for (; prev_rep > rep; --prev_rep) l2len[(size_t) n++] = 0;
                }
                if (err) return;
                if (overrun()) FAIL(E_SHRT);
            }
            CHECK(total == 32768, E4("hufd"));
        }
        int max_len = 1;
        for (int l : l2len) max_len = std::max(max_len, l);
        cl.max_len = (int16_t) max_len;
        cl.table_off = build_prefix_lut(canonical(l2len), max_len, &cl.root_bits);
    }

    // 12-bit ANS distribution (j40.h:2601-2709)
    void ans_distribution(int log_alpha_size, std::vector<int32_t> &D) {
        const int table_size = 1 << log_alpha_size;
        D.assign((size_t) table_size, 0);
        switch (u(2)) {
        case 1: {
            int32_t v = u8v();
            CHECK(v < table_size, E4("ansd"));
            D[(size_t) v] = 4096;
            break;
        }
        case 3: {
            int32_t v1 = u8v(), v2 = u8v();
            CHECK(v1 != v2 && v1 < table_size && v2 < table_size, E4("ansd"));
            D[(size_t) v1] = (int32_t) u(12);
            D[(size_t) v2] = 4096 - D[(size_t) v1];
            break;
        }
        case 2: {
            int32_t alpha_size = u8v() + 1;
            int32_t d = 4096 / alpha_size, bias = 4096 % alpha_size;
            CHECK(alpha_size <= table_size, E4("ansd"));
            for (int32_t k = 0; k < alpha_size; ++k) D[(size_t) k] = k < bias ? d + 1 : d;
            break;
        }
        default: {
            int len = u(1) ? u(1) ? u(1) ? 3 : 2 : 1 : 0;
            int shift = (int) u(len) + (1 << len) - 1;
            CHECK(shift <= 13, E4("ansd"));
            int alpha_size = u8v() + 3;
            // log-count codes: fixed prefix code, read through a 7-bit peek
            int codes[260], ncodes = 0, omit_log = -1, n = 0;
            for (n = 0; n < alpha_size;) {
                uint32_t b = br.peek(7);
                int sym, l;
                switch (b & 7) {
                case 0: sym = 10; l = 3; break;
                case 2: sym = 7; l = 3; break;
                case 4: sym = 6; l = 3; break;
                case 5: sym = 8; l = 3; break;
                case 6: sym = 9; l = 3; break;
                default:
                    switch (b & 15) {
                    case 3: sym = 3; l = 4; break;
                    case 7: sym = 5; l = 4; break;
                    case 9: sym = 4; l = 4; break;
                    case 11: sym = 1; l = 4; break;
                    case 15: sym = 2; l = 4; break;
                    default: // ...0001 prefix
                        if (b & 16) { sym = 0; l = 5; }
                        else if (b & 32) { sym = 11; l = 6; }
                        else if (b & 64) { sym = 13; l = 7; }
                        else { sym = 12; l = 7; }
                    }
                }
                br.skip(l);
                if (sym < 13) {
                    ++n;
                    codes[ncodes++] = sym;
                    if (omit_log < sym) omit_log = sym;
                } else {
                    int rep = u8v() + 4;
                    n += rep;
                    codes[ncodes++] = -rep;
                }
                if (ncodes >= 259 || overrun()) break;
            }
            CHECK(n == alpha_size && omit_log >= 0, E4("ansd"));
            int omit_pos = -1, total = 0;
            n = 0;
            for (int k = 0; k < ncodes && n < table_size; ++k) {
                int code = codes[k];
                if (code < 0) {
                    int32_t prev = n > 0 ? D[(size_t) n - 1] : 0;
                    CHECK(prev >= 0, E4("ansd"));
                    int rep = std::min(-code, table_size - n);
                    total += prev * rep;
                    while (rep-- > 0) D[(size_t) n++] = prev;
                } else if (code == omit_log) {
                    omit_pos = n;
                    omit_log = -1;
                    D[(size_t) n++] = -1;
                } else if (code < 2) {
                    total += code;
                    D[(size_t) n++] = code;
                } else {
                    --code;
                    int bitcount = std::min(std::max(0, shift - ((12 - code) >> 1)), code);
                    int v = (1 << code) + ((int) u(bitcount) << (code - bitcount));
                    total += v;
                    D[(size_t) n++] = v;
                }
            }
            for (; n < table_size; ++n) D[(size_t) n] = 0;
            CHECK(omit_pos >= 0, E4("ansd"));
            CHECK(total <= 4096, E4("ansd"));
            D[(size_t) omit_pos] = 4096 - total;
            break;
        }
        }
    }

    // alias table per the format's construction (j40.h:2362-2439), packed for the device
    uint32_t build_alias_table(const std::vector<int32_t> &D, int las) {
        const int tsize = 1 << las, lbs = 12 - las, bsize = 1 << lbs;
        std::vector<int> cutoff((size_t) tsize, 0), offnext((size_t) tsize, 0), symbol((size_t) tsize, 0);
        int first = -1, nz = 0;
        for (int i = 0; i < tsize; ++i) if (D[(size_t) i]) { if (first < 0) first = i; ++nz; }
        if (nz == 1) {
            for (int j = 0; j < tsize; ++j) { symbol[(size_t) j] = first; offnext[(size_t) j] = j << lbs; cutoff[(size_t) j] = 0; }
        } else {
            int un = -1, ov = -1;
            for (int i = 0; i < tsize; ++i) {
                int c = D[(size_t) i];
                cutoff[(size_t) i] = c;
                if (c > bsize) { offnext[(size_t) i] = ov; ov = i; }
                else if (c < bsize) { offnext[(size_t) i] = un; un = i; }
                else { symbol[(size_t) i] = i; offnext[(size_t) i] = 0; }
            }
            while (ov >= 0 && un >= 0) {
                int by = bsize - cutoff[(size_t) un];
                int tmp = offnext[(size_t) un];
                cutoff[(size_t) ov] -= by;
                symbol[(size_t) un] = ov;
                offnext[(size_t) un] = cutoff[(size_t) ov] - cutoff[(size_t) un];
                un = tmp;
                if (cutoff[(size_t) ov] < bsize) {
                    tmp = offnext[(size_t) ov];
                    offnext[(size_t) ov] = un;
                    un = ov;
                    ov = tmp;
                } else if (cutoff[(size_t) ov] == bsize) {
                    tmp = offnext[(size_t) ov];
                    symbol[(size_t) ov] = ov;
                    offnext[(size_t) ov] = 0;
                    ov = tmp;
                }
            }
            // distributions not summing to 4096 (possible with a zero "omitted" count) leave buckets
            // unsettled in the reference as well (undefined there); make them harmless here
            for (; un >= 0;) { int t2 = offnext[(size_t) un]; symbol[(size_t) un] = un; offnext[(size_t) un] = 0; un = t2; }
        }
        uint32_t off = arena.alloc((size_t) tsize * 8, 8);
        uint64_t *tab = arena.at<uint64_t>(off);
        for (int i = 0; i < tsize; ++i) {
            uint32_t lo = (uint32_t) (cutoff[(size_t) i] & 0xff) | ((uint32_t) (symbol[(size_t) i] & 0xff) << 8) | ((uint32_t) (offnext[(size_t) i] & 0xfff) << 16);
            uint32_t hi = (uint32_t) (D[(size_t) i] & 0x1fff) | ((uint32_t) (D[(size_t) symbol[(size_t) i]] & 0x1fff) << 16);
            tab[i] = (uint64_t) lo | ((uint64_t) hi << 32);
        }
        return off;
    }

    void cluster_map(int32_t num_dist, int32_t max_allowed, int32_t *num_clusters, std::vector<uint8_t> &map) {
        if (max_allowed > num_dist) max_allowed = num_dist;
        map.assign((size_t) num_dist, 0);
        if (num_dist == 1) { *num_clusters = 1; return; }
        if (u(1)) { // simple
            int nbits = (int) u(2);
            for (int32_t i = 0; i < num_dist; ++i) {
                map[(size_t) i] = (uint8_t) u(nbits);
                CHECK((int32_t) map[(size_t) i] < max_allowed, E4("clst"));
                if (overrun()) FAIL(E_SHRT);
            }
        } else {
            int use_mtf = (int) u(1);
            uint32_t nested = read_code_spec(num_dist <= 2 ? -1 : 1);
            if (err) return;
            CodeCtx cc;
            cc.init(arena.bytes.data(), nested);
            CodeState cs;
            std::vector<int32_t> window;
            if (cc.spec->lz77_enabled) window.resize(1u << 20);
            cs.init(window.empty() ? nullptr : window.data(), (1u << 20) - 1);
            ErrSlot es = {0};
            for (int32_t i = 0; i < num_dist; ++i) {
                int32_t index = code(br, es, cc, cs, 0, 0);
                if (es.err) { err = es.err; return; }
                CHECK(index < max_allowed, E4("clst"));
                map[(size_t) i] = (uint8_t) index;
                if (overrun()) FAIL(E_SHRT);
                // arena may not move while cc points into it: read_code_spec is not called inside the loop
            }
            finish_code(br, es, cc, cs);
            if (es.err) { err = es.err; return; }
            if (use_mtf) {
                uint8_t mtf[256];
                for (int i = 0; i < 256; ++i) mtf[i] = (uint8_t) i;
                for (int32_t i = 0; i < num_dist; ++i) {
                    int j = map[(size_t) i];
                    uint8_t moved = mtf[j];
                    map[(size_t) i] = moved;
                    for (; j > 0; --j) mtf[j] = mtf[j - 1];
                    mtf[0] = moved;
                }
            }
        }
        bool seen[256] = {false};
        for (int32_t i = 0; i < num_dist; ++i) seen[map[(size_t) i]] = true;
        int k = 0;
        while (k < 256 && seen[k]) ++k;
        *num_clusters = k;
        for (; k < 256; ++k) CHECK(!seen[k], E4("clst"));
    }

    // returns the arena offset of a DCodeSpec (j40.h:2711-2777); num_dist < 0 forbids LZ77
    uint32_t read_code_spec(int32_t num_dist) {
        bool allow_lz77 = num_dist > 0;
        num_dist = num_dist < 0 ? -num_dist : num_dist;
        DCodeSpec spec;
        memset(&spec, 0, sizeof(spec));
        spec.lz77_enabled = (int32_t) u(1);
        if (spec.lz77_enabled) {
            CHECKV(allow_lz77, E4("lz77"), 0);
            spec.min_symbol = (int32_t) u32(224, 0, 512, 0, 4096, 0, 8, 15);
            spec.min_length = (int32_t) u32(3, 0, 4, 0, 5, 2, 9, 8);
            spec.lz_len_cfg = hybrid_cfg(8);
            ++num_dist;
        } else {
            spec.min_symbol = spec.min_length = 0x7fffffff;
        }
        if (err) return 0;
        std::vector<uint8_t> map;
        cluster_map(num_dist, 256, &spec.num_clusters, map);
        if (err) return 0;
        std::vector<DCluster> clusters((size_t) spec.num_clusters);
        memset(clusters.data(), 0, clusters.size() * sizeof(DCluster));
        // everything allocated from here on belongs to this spec alone (nested specs came earlier)
        uint32_t blob_lo = (uint32_t) ((arena.bytes.size() + 15) / 16 * 16);
        arena.alloc(0, 16);
        spec.use_prefix_code = (int32_t) u(1);
        if (spec.use_prefix_code) {
            for (auto &c : clusters) c.cfg = hybrid_cfg(15);
            std::vector<int32_t> count((size_t) spec.num_clusters, 1);
            for (auto &cnt : count) {
                if (u(1)) {
                    int n = (int) u(4);
                    cnt = 1 + (1 << n) + (int32_t) u(n);
                    CHECKV(cnt <= (1 << 15), E4("hufd"), 0);
                }
            }
            if (err) return 0;
            for (size_t i = 0; i < clusters.size(); ++i) {
                prefix_code_tree(count[i], clusters[i]);
                if (err) return 0;
                if (overrun()) FAILV(E_SHRT, 0);
            }
        } else {
            spec.log_alpha_size = 5 + (int32_t) u(2);
            for (auto &c : clusters) c.cfg = hybrid_cfg(spec.log_alpha_size);
            if (err) return 0;
            std::vector<int32_t> D;
            for (auto &c : clusters) {
                ans_distribution(spec.log_alpha_size, D);
                if (err) return 0;
                if (overrun()) FAILV(E_SHRT, 0);
                c.table_off = build_alias_table(D, spec.log_alpha_size);
                // the tables of a spec follow each other in the arena (8-byte entries, 8-byte aligned allocations)
                if (&c == &clusters[0]) spec.ans_tables_off = c.table_off;
                if (c.table_off != spec.ans_tables_off + (uint32_t) ((&c - &clusters[0]) << (spec.log_alpha_size + 3))) { fprintf(stderr, "j40_b200: alias tables not contiguous\n"); abort(); }
            }
        }
        if (overrun()) FAILV(E_SHRT, 0);
        spec.num_dist = num_dist;
        spec.blob_lo = blob_lo;
        spec.cluster_map_off = arena.alloc(map.size(), 8);
        memcpy(arena.at<uint8_t>(spec.cluster_map_off), map.data(), map.size());
        spec.clusters_off = arena.alloc(clusters.size() * sizeof(DCluster), 16);
        memcpy(arena.at<uint8_t>(spec.clusters_off), clusters.data(), clusters.size() * sizeof(DCluster));
        uint32_t off = arena.alloc(sizeof(DCodeSpec), 16);
        spec.blob_hi = off + (uint32_t) sizeof(DCodeSpec);
        memcpy(arena.at<uint8_t>(off), &spec, sizeof(spec));
        return off;
    }

    // host-side symbol reader over a parsed spec
    struct HostCode {
        CodeCtx cc;
        CodeState cs;
        std::vector<int32_t> window;
        ErrSlot es = {0};
        void begin(Arena &a, uint32_t spec_off) {
            cc.init(a.bytes.data(), spec_off);
            if (cc.spec->lz77_enabled) window.resize(1u << 20);
            cs.init(window.empty() ? nullptr : window.data(), (1u << 20) - 1);
        }
    };
    int32_t hcode(HostCode &h, int ctx) {
        if (err) return 0;
        int32_t v = code(br, h.es, h.cc, h.cs, ctx, 0);
        if (h.es.err) { err = h.es.err; return 0; }
        if (overrun()) { err = E_SHRT; return 0; }
        return v;
    }
    void hfinish(HostCode &h) {
        if (err) return;
        finish_code(br, h.es, h.cc, h.cs);
        if (h.es.err) err = h.es.err;
    }

    // ------------------------------------------------------------------------------------------
    // MA tree (j40.h:3461-3513); returns offsets of the node array and of the sample code spec

    void read_tree(int32_t max_tree_size, uint32_t *tree_off, uint32_t *spec_off, int32_t *uses_wp) {
        uint32_t tspec = read_code_spec(6);
        if (err) return;
        std::vector<DTreeNode> nodes;
        {
            HostCode h;
            h.begin(arena, tspec);
            int32_t ctx_id = 0, nodes_left = 1, depth = 0, upto = 1;
            *uses_wp = 0;
            while (nodes_left-- > 0) {
                int32_t idx = (int32_t) nodes.size();
                if (idx == upto) {
                    CHECK(++depth <= Limits::tree_depth, E4("tlim"));
                    upto += nodes_left + 1;
                }
                int32_t prop = hcode(h, 1);
                if (err) return;
                DTreeNode n;
                if (prop > 0) {
                    n.a = -prop; // -1 - (prop - 1)
                    n.b = unpack_signed(hcode(h, 0));
                    n.c = idx + (++nodes_left);
                    n.d = idx + (++nodes_left);
                    if (prop - 1 == 15) *uses_wp = 1;
                } else {
                    n.a = ctx_id++;
                    n.b = hcode(h, 2);
                    n.c = unpack_signed(hcode(h, 3));
                    int32_t shift = hcode(h, 4);
                    CHECK(shift < 31, E4("tree"));
                    int32_t val = hcode(h, 5);
                    CHECK(((val + 1) >> (31 - shift)) == 0, E4("tree"));
                    n.d = (val + 1) << shift;
                    if (n.b == 6) *uses_wp = 1;
                }
                if (err) return;
                nodes.push_back(n);
                CHECK((int32_t) nodes.size() + nodes_left <= max_tree_size, E4("tlim"));
            }
            hfinish(h);
            if (err) return;
            int32_t num_leaves = ctx_id;
            *spec_off = read_code_spec(num_leaves);
        }
        if (err) return;
        *tree_off = arena.alloc(nodes.size() * sizeof(DTreeNode), 16);
        memcpy(arena.at<uint8_t>(*tree_off), nodes.data(), nodes.size() * sizeof(DTreeNode));
    }

    // Lehmer-coded permutation (j40.h:5428-5457)
    void read_permutation(HostCode &h, int32_t size, int32_t skip, std::vector<int32_t> &lehmer) {
        lehmer.clear();
        int32_t end = hcode(h, std::min(7, ceil_lg32((uint32_t) size + 1)));
        CHECK(end <= size - skip, E4("perm"));
        int32_t prev = 0;
        for (int32_t i = 0; i < end; ++i) {
            prev = hcode(h, std::min(7, ceil_lg32((uint32_t) prev + 1)));
            CHECK(prev < size - (skip + i), E4("perm"));
            lehmer.push_back(prev);
        }
    }
    template <class T> static void apply_permutation(T *target, const std::vector<int32_t> &lehmer) {
        for (size_t p = 0; p < lehmer.size(); ++p) {
            size_t x = (size_t) lehmer[p];
            T tmp = target[p + x];
            for (size_t k = p + x; k > p; --k) target[k] = target[k - 1];
            target[p] = tmp;
        }
    }

    // ------------------------------------------------------------------------------------------
    // headers

    void size_header(int32_t *w, int32_t *h) { // j40.h:3008
        int div8 = (int) u(1);
        *h = div8 ? (int32_t) (u(5) + 1) * 8 : (int32_t) u32(1, 9, 1, 13, 1, 18, 1, 30);
        switch (u(3)) {
        case 0: *w = div8 ? (int32_t) (u(5) + 1) * 8 : (int32_t) u32(1, 9, 1, 13, 1, 18, 1, 30); break;
        case 1: *w = *h; break;
        case 2: *w = (int32_t) ((uint64_t) *h * 6 / 5); break;
        case 3: *w = (int32_t) ((uint64_t) *h * 4 / 3); break;
        case 4: *w = (int32_t) ((uint64_t) *h * 3 / 2); break;
        case 5: *w = (int32_t) ((uint64_t) *h * 16 / 9); break;
        case 6: *w = (int32_t) ((uint64_t) *h * 5 / 4); break;
        default:
            CHECK(*h < 0x40000000, E4("bigg"));
            *w = *h * 2;
        }
    }
    void bit_depth(int32_t *bpp, int32_t *exp_bits) { // j40.h:3033
        if (u(1)) {
            *bpp = (int32_t) u32(32, 0, 16, 0, 24, 0, 1, 6);
            *exp_bits = (int32_t) u(4) + 1;
            int32_t mant = *bpp - *exp_bits - 1;
            CHECK(2 <= mant && mant <= 23, E4("bpp?"));
            CHECK(2 <= *exp_bits && *exp_bits <= 8, E4("exp?"));
        } else {
            *bpp = (int32_t) u32(8, 0, 10, 0, 12, 0, 1, 6);
            *exp_bits = 0;
            CHECK(1 <= *bpp && *bpp <= 31, E4("bpp?"));
        }
    }
    void name() { // j40.h:3053 (bytes are validated as UTF-8 and dropped)
        int32_t len = (int32_t) u32(0, 0, 0, 4, 16, 5, 48, 10);
        std::vector<uint8_t> buf((size_t) len + 1, 0);
        for (int32_t i = 0; i < len; ++i) { buf[(size_t) i] = (uint8_t) u(8); if (overrun()) FAIL(E_SHRT); }
        for (int32_t i = 0; i < len;) {
            int c = buf[(size_t) i++], cc = buf[(size_t) i];
            c = c < 0x80 ? 0 : c < 0xc2 ? -1 : c < 0xe0 ? 1 :
                c < 0xf0 ? ((c == 0xe0 ? cc >= 0xa0 : c == 0xed ? cc < 0xa0 : 1) ? 2 : -1) :
                c < 0xf5 ? ((c == 0xf0 ? cc >= 0x90 : c == 0xf4 ? cc < 0x90 : 1) ? 3 : -1) : -1;
            CHECK(c >= 0 && i + c < len, E4("name"));
            while (c-- > 0) CHECK((buf[(size_t) i++] & 0xc0) == 0x80, E4("name"));
        }
    }
    void customxy() { u32(0, 19, 0x80000, 19, 0x100000, 20, 0x200000, 21); u32(0, 19, 0x80000, 19, 0x100000, 20, 0x200000, 21); }
    void extensions() { // j40.h:3087
        uint64_t ext = u64();
        uint64_t nbits = 0;
        for (int i = 0; i < 64; ++i) if (ext >> i & 1) {
            uint64_t n = u64();
            if (err) return;
            CHECK(n <= (uint64_t) INT64_MAX && nbits + n <= (uint64_t) INT64_MAX, E4("flen"));
            nbits += n;
        }
        // j40__skip: needs the bits to exist
        uint64_t avail = (uint64_t) br.size * 8 - std::min<uint64_t>(br.bits_consumed(), (uint64_t) br.size * 8);
        CHECK(nbits <= avail, E_SHRT);
        skip_bits(nbits);
    }

    void image_metadata() { // j40.h:3104-3313
        ImageInfo &im = plan.im;
        static const float OPSIN_INV[3][3] = {
            {11.031566901960783f, -9.866943921568629f, -0.16462299647058826f},
            {-3.254147380392157f, 4.418770392156863f, -0.16462299647058826f},
            {-3.6588512862745097f, 2.7129230470588235f, 1.9459282392156863f},
        };
        memcpy(im.opsin_inv_mat, OPSIN_INV, sizeof(OPSIN_INV));
        im.opsin_bias[0] = im.opsin_bias[1] = im.opsin_bias[2] = -0.0037930732552754493f;
        im.quant_bias[0] = 1.0f - 0.05465007330715401f;
        im.quant_bias[1] = 1.0f - 0.07005449891748593f;
        im.quant_bias[2] = 1.0f - 0.049935103337343655f;
        im.quant_bias_num = 0.145f;
        size_header(&im.width, &im.height);
        CHECK(im.width <= Limits::width && im.height <= Limits::height, E4("slim"));
        CHECK((int64_t) im.width * im.height <= Limits::pixels, E4("slim"));
        if (!u(1)) {
            int extra_fields = (int) u(1);
            if (extra_fields) {
                u(3); // orientation
                if (u(1)) { int32_t iw, ih; size_header(&iw, &ih); }
                if (u(1)) FAIL(E_TODO); // preview
                if (u(1)) {
                    im.anim = 1;
                    u32(100, 0, 1000, 0, 1, 10, 1, 30);
                    u32(1, 0, 1001, 0, 1, 8, 1, 10);
                    u32(0, 0, 0, 3, 0, 16, 0, 32);
                    im.anim_have_timecodes = (int) u(1);
                }
            }
            bit_depth(&im.bpp, &im.exp_bits);
            CHECK(im.bpp <= Limits::bpp, E4("fbpp"));
            im.modular_16bit_buffers = (int) u(1);
            CHECK(im.modular_16bit_buffers, E4("fm32"));
            im.num_extra_channels = (int32_t) u32(0, 0, 1, 0, 2, 4, 1, 12);
            CHECK(im.num_extra_channels <= Limits::num_extra_channels, E4("elim"));
            for (int i = 0; i < im.num_extra_channels; ++i) {
                ImageInfo::EC &ec = im.ec[i];
                memset(&ec, 0, sizeof(ec));
                if (u(1)) {
                    ec.type = 0; ec.bpp = 8;
                } else {
                    ec.type = enum_();
                    bit_depth(&ec.bpp, &ec.exp_bits);
                    ec.dim_shift = (int32_t) u32(0, 0, 3, 0, 4, 0, 1, 3);
                    name();
                    switch (ec.type) {
                    case 0: ec.alpha_associated = (int) u(1); break;
                    case 2: f16(); f16(); f16(); f16(); break;
                    case 5: u32(1, 0, 0, 2, 3, 4, 19, 8); break;
                    case 4: FAIL(E4("fblk"));
                    case 1: case 3: case 6: case 15: case 16: break;
                    default: FAIL(E4("ect?"));
                    }
                }
                CHECK(ec.bpp <= Limits::bpp, E4("fbpp"));
                if (err) return;
            }
            im.xyb_encoded = (int) u(1);
            if (!u(1)) { // ColourEncoding
                im.want_icc = (int) u(1);
                int32_t cspace = enum_();
                CHECK(cspace <= 3, E4("csp?"));
                im.cspace_grey = cspace == 1;
                if (!im.want_icc) {
                    if (cspace != 2) {
                        switch (enum_()) {
                        case 1: case 10: case 11: break;
                        case 2: customxy(); break;
                        default: FAIL(E4("wpt?"));
                        }
                        if (cspace != 1) {
                            switch (enum_()) {
                            case 1: case 9: case 11: break;
                            case 2: customxy(); customxy(); customxy(); break;
                            default: FAIL(E4("prm?"));
                            }
                        }
                    }
                    if (u(1)) {
                        int32_t g = (int32_t) u(24);
                        CHECK(g > 0 && g <= 10000000, E4("gama"));
                        if (cspace == 2) CHECK(g == 3333333, E4("gama"));
                    } else {
                        int32_t tf = enum_();
                        CHECK(tf == 1 || tf == 2 || tf == 8 || tf == 13 || tf == 16 || tf == 17 || tf == 18, E4("tfn?"));
                    }
                    int32_t intent = enum_();
                    CHECK(intent >= 0 && intent <= 3, E4("itt?"));
                }
            }
            if (extra_fields) {
                if (!u(1)) { // ToneMapping
                    im.intensity_target = f16();
                    CHECK(im.intensity_target > 0, E4("tone"));
                    float min_nits = f16();
                    CHECK(0 < min_nits && min_nits <= im.intensity_target, E4("tone"));
                    int rel = (int) u(1);
                    float lb = f16();
                    if (rel) CHECK(0 <= lb && lb <= 1, E4("tone")); else CHECK(0 <= lb, E4("tone"));
                }
            }
            extensions();
        }
        if (err) return;
        if (!u(1)) { // !default_m
            if (im.xyb_encoded) {
                for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) im.opsin_inv_mat[i][j] = f16();
                for (int i = 0; i < 3; ++i) im.opsin_bias[i] = f16();
                for (int i = 0; i < 3; ++i) im.quant_bias[i] = f16();
                im.quant_bias_num = f16();
            }
            int cw_mask = (int) u(3);
            CHECK(cw_mask == 0, E_TODO);
        }
        check_overrun();
    }

    void icc() { // j40.h:3351-3383: decoded and dropped
        uint64_t enc_size = u64();
        uint32_t spec = read_code_spec(41);
        if (err) return;
        HostCode h;
        h.begin(arena, spec);
        uint64_t index = 0, output_size = 0;
        { // varint
            int shift = 0;
            for (;;) {
                CHECK(index++ < enc_size, E4("icc?"));
                int32_t b = hcode(h, 0);
                if (err) return;
                output_size |= (uint64_t) (b & 0x7f) << shift;
                if (b < 128) break;
                shift += 7;
                CHECK(shift < 63, E4("vint"));
            }
        }
        CHECK(output_size <= Limits::icc_output_size, E4("plim"));
        CHECK(output_size >= enc_size / 21, E4("icc?"));
        int32_t byte = 0, prev = 0, pprev = 0;
        for (; index < enc_size; ++index) {
            pprev = prev;
            prev = byte;
            int ctx = 0;
            if (index > 128) {
                if (prev < 16) ctx = prev < 2 ? prev + 3 : 5;
                else if (prev > 240) ctx = 6 + (prev == 255);
                else if (97 <= (prev | 32) && (prev | 32) <= 122) ctx = 1;
                else if (prev == 44 || prev == 46 || (48 <= prev && prev < 58)) ctx = 2;
                else ctx = 8;
                if (pprev < 16) ctx += 2 * 8;
                else if (pprev > 240) ctx += 3 * 8;
                else if (97 <= (pprev | 32) && (pprev | 32) <= 122) ctx += 0 * 8;
                else if (pprev == 44 || pprev == 46 || (48 <= pprev && pprev < 58)) ctx += 1 * 8;
                else ctx += 4 * 8;
            }
            byte = hcode(h, ctx);
            if (err) return;
        }
        hfinish(h);
    }

    void frame_header() { // j40.h:5163-5388
        ImageInfo &im = plan.im;
        FrameInfo &f = plan.fh;
        f.width = im.width;
        f.height = im.height;
        zero_pad();
        if (err) return;
        if (!u(1)) {
            int full_frame = 1;
            int32_t x0 = 0, y0 = 0;
            int64_t duration = 0;
            int save_as_ref = 0, blend_mode0 = 0;
            f.type = (int) u(2);
            f.is_modular = (int) u(1);
            uint64_t flags = u64();
            f.has_noise = (int) (flags & 1);
            f.has_patches = (int) (flags >> 1 & 1);
            f.has_splines = (int) (flags >> 4 & 1);
            f.use_lf_frame = (int) (flags >> 5 & 1);
            f.skip_adapt_lf_smooth = (int) (flags >> 7 & 1);
            if (!im.xyb_encoded) f.do_ycbcr = (int) u(1);
            if (!f.use_lf_frame) {
                if (f.do_ycbcr) f.jpeg_upsampling = (int) u(6);
                CHECK(u(2) == 0, E_TODO); // upsampling
                for (int i = 0; i < im.num_extra_channels; ++i) CHECK(u(2) == 0, E_TODO);
            }
            if (f.is_modular) f.group_size_shift = 7 + (int) u(2);
            else if (im.xyb_encoded) { f.x_qm_scale = (int) u(3); f.b_qm_scale = (int) u(3); }
            if (f.type != 2) {
                f.num_passes = (int) u32(1, 0, 2, 0, 3, 0, 4, 3);
                if (f.num_passes > 1) {
                    int8_t log_ds[4];
                    int32_t ppass = 0, num_ds = (int32_t) u32(0, 0, 1, 0, 2, 0, 3, 1);
                    CHECK(num_ds < f.num_passes, E4("pass"));
                    for (int i = 0; i < f.num_passes - 1; ++i) u(2);
                    for (int i = 0; i < num_ds; ++i) {
                        log_ds[i] = (int8_t) u(2);
                        if (i > 0) CHECK(log_ds[i - 1] >= log_ds[i], E4("pass"));
                    }
                    for (int i = 0; i < num_ds; ++i) {
                        int32_t pass = (int32_t) u32(0, 0, 1, 0, 2, 0, 0, 3);
                        CHECK(i > 0 ? ppass < pass && pass < f.num_passes : pass == 0, E4("pass"));
                        ppass = pass;
                    }
                }
            }
            if (f.type == 1) {
                u(2);
            } else if (u(1)) { // have_crop
                if (f.type != 2) {
                    x0 = unpack_signed((int32_t) u32(0, 8, 256, 11, 2304, 14, 18688, 30));
                    y0 = unpack_signed((int32_t) u32(0, 8, 256, 11, 2304, 14, 18688, 30));
                }
                f.width = (int32_t) u32(0, 8, 256, 11, 2304, 14, 18688, 30);
                f.height = (int32_t) u32(0, 8, 256, 11, 2304, 14, 18688, 30);
                CHECK(f.width <= Limits::width && f.height <= Limits::height, E4("slim"));
                CHECK((int64_t) f.width * f.height <= Limits::pixels, E4("slim"));
                full_frame = x0 <= 0 && y0 <= 0 && f.width + x0 >= im.width && f.height + y0 >= im.height;
            }
            if (f.type == 0 || f.type == 3) {
                for (int i = -1; i < im.num_extra_channels; ++i) {
                    int mode = (int) u32(0, 0, 1, 0, 2, 0, 3, 2);
                    if (i < 0) blend_mode0 = mode;
                    if (im.num_extra_channels > 0) {
                        if (mode == 2 || mode == 3) { u32(0, 0, 1, 0, 2, 0, 3, 3); u(1); }
                        else if (mode == 4) u(1);
                    }
                    if (!full_frame || mode != 0) u(2);
                }
                if (im.anim) {
                    uint32_t sel = u(2);
                    duration = sel == 0 ? 0 : sel == 1 ? 1 : sel == 2 ? (int64_t) u(8) : (int64_t) u(32);
                    if (im.anim_have_timecodes) u(32);
                }
                f.is_last = (int) u(1);
            } else {
                f.is_last = 0;
            }
            if (f.type != 1 && !f.is_last) save_as_ref = (int) u(2);
            if (f.type == 2 || (full_frame && (f.type == 0 || f.type == 3) && blend_mode0 == 0 &&
                                (duration == 0 || save_as_ref != 0) && !f.is_last)) {
                u(1); // save_before_ct
            }
            name();
            { // RestorationFilter -- parsed the way the reference does
                int all_default = (int) u(1);
                int gab = all_default ? 1 : (int) u(1);
                if (gab) { if (u(1)) for (int i = 0; i < 6; ++i) f16(); }
                int epf_iters = all_default ? 2 : (int) u(2);
                if (epf_iters) {
                    if (!f.is_modular && u(1)) for (int i = 0; i < 8; ++i) f16();
                    if (u(1)) { for (int i = 0; i < 3; ++i) f16(); u(32); }
                    if (u(1)) { if (!f.is_modular) f16(); f16(); f16(); f16(); }
                    if (epf_iters && f.is_modular) f16();
                }
                if (!all_default) extensions();
            }
            extensions();
        }
        check_overrun();
        if (err) return;
        f.grows = ceil_div(f.height, 1 << f.group_size_shift);
        f.gcolumns = ceil_div(f.width, 1 << f.group_size_shift);
        f.num_groups = (int64_t) f.grows * f.gcolumns;
        f.ggrows = ceil_div(f.height, 8 << f.group_size_shift);
        f.ggcolumns = ceil_div(f.width, 8 << f.group_size_shift);
        f.num_lf_groups = (int64_t) f.ggrows * f.ggcolumns;
    }

    // ------------------------------------------------------------------------------------------
    struct TocEntry { uint64_t off; uint32_t size; };
    uint64_t lf_global_off = 0, hf_global_off = 0;
    uint32_t lf_global_size = 0, hf_global_size = 0;

    void read_toc() { // j40.h:5479-5648
        FrameInfo &f = plan.fh;
        int64_t nsections = f.num_passes == 1 && f.num_groups == 1 ? 1 : 1 + f.num_lf_groups + 1 + f.num_passes * f.num_groups;
        CHECK(nsections <= INT32_MAX, E4("flen"));
        std::vector<int32_t> lehmer;
        bool permuted = u(1) != 0;
        if (permuted) {
            uint32_t spec = read_code_spec(8);
            if (err) return;
            HostCode h;
            h.begin(arena, spec);
            read_permutation(h, (int32_t) nsections, 0, lehmer);
            hfinish(h);
            if (err) return;
        }
        zero_pad();
        if (err) return;
        if (nsections == 1) {
            uint32_t size = u32(0, 10, 1024, 14, 17408, 22, 4211712, 30);
            zero_pad();
            check_overrun();
            if (err) return;
            plan.single_section = true;
            lf_global_off = br.bits_consumed() / 8;
            lf_global_size = size;
            plan.end_codeoff = lf_global_off + size;
            return;
        }
        std::vector<TocEntry> sec((size_t) nsections);
        for (auto &s : sec) { s.size = u32(0, 10, 1024, 14, 17408, 22, 4211712, 30); if (overrun()) FAIL(E_SHRT); }
        zero_pad();
        check_overrun();
        if (err) return;
        uint64_t off = br.bits_consumed() / 8;
        for (auto &s : sec) { s.off = off; off += s.size; }
        plan.end_codeoff = off;
        if (permuted) apply_permutation(sec.data(), lehmer);
        lf_global_off = sec[0].off; lf_global_size = sec[0].size;
        size_t nlf = (size_t) f.num_lf_groups, ng = (size_t) f.num_groups;
        hf_global_off = sec[1 + nlf].off; hf_global_size = sec[1 + nlf].size;
        plan.lfg_sec.resize(nlf);
        CHECK(f.num_passes == 1 || !f.is_modular, E_TODO); // every pass of a modular frame re-codes the same channels (j40.h:7022): not built
        const size_t npg = ng * (size_t) f.num_passes; // pass groups: index pass * ng + group (j40.h:5555-5561)
        plan.pg_sec.resize(npg);
        for (size_t i = 0; i < nlf; ++i) { plan.lfg_sec[i].off = sec[1 + i].off; plan.lfg_sec[i].size = sec[1 + i].size; }
        for (size_t i = 0; i < npg; ++i) { plan.pg_sec[i].off = sec[2 + nlf + i].off; plan.pg_sec[i].size = sec[2 + nlf + i].size; }
        // decoding order of the reference: by codestream offset, except that a group stored before
        // its LF group is decoded right after that LF group
        struct Item { uint64_t off; int kind; size_t idx; };
        std::vector<Item> main_list;
        std::vector<std::vector<Item>> reloc(nlf);
        for (size_t i = 0; i < nlf; ++i) main_list.push_back({plan.lfg_sec[i].off, 0, i});
        for (size_t pg = 0; pg < npg; ++pg) {
            size_t g = pg % ng;
            size_t grow = g / (size_t) f.gcolumns, gcol = g % (size_t) f.gcolumns;
            size_t gg = (grow / 8) * (size_t) f.ggcolumns + gcol / 8;
            if (plan.pg_sec[pg].off > plan.lfg_sec[gg].off) main_list.push_back({plan.pg_sec[pg].off, 1, pg});
            else reloc[gg].push_back({plan.pg_sec[pg].off, 1, pg});
        }
        auto by_off = [](const Item &a, const Item &b) { return a.off < b.off; };
        std::stable_sort(main_list.begin(), main_list.end(), by_off);
        int32_t rank = 0;
        for (const Item &it : main_list) {
            if (it.kind == 0) {
                plan.lfg_sec[it.idx].rank = rank++;
                std::stable_sort(reloc[it.idx].begin(), reloc[it.idx].end(), by_off);
                for (const Item &r : reloc[it.idx]) plan.pg_sec[r.idx].rank = rank++;
            } else {
                plan.pg_sec[it.idx].rank = rank++;
            }
        }
    }

    // ------------------------------------------------------------------------------------------
    void init_global_modular() { // j40.h:3596-3637
        ImageInfo &im = plan.im;
        FrameInfo &f = plan.fh;
        ModImage &m = plan.gmod;
        memset(&m, 0, sizeof(m));
        m.num_channels = im.num_extra_channels;
        if (f.is_modular) m.num_channels += (!f.do_ycbcr && !im.xyb_encoded && im.cspace_grey) ? 1 : 3;
        for (int i = 0; i < im.num_extra_channels; ++i) CHECK(im.ec[i].dim_shift == 0, E_TODO);
        CHECK(m.num_channels <= MOD_MAX_CH, E_TODO);
        for (int i = 0; i < m.num_channels; ++i) { m.ch[i].w = f.width; m.ch[i].h = f.height; m.ch[i].stride = f.width; }
    }

    // ModularHeader on the host, including a local tree (j40.h:3717-3850). `m` holds the channel geometry.
    void host_modular_header(ModImage &m, FramePlan::LocalHeader &lh, bool allow_palette = false) {
        ErrSlot es = {0};
        int local = 0;
        modular_header(br, es, plan.df.have_global_tree != 0, m, &local, allow_palette);
        if (es.err) { err = es.err; return; }
        if (local) {
            int64_t max_tree_size = 1024;
            for (int i = 0; i < m.num_channels; ++i) max_tree_size = std::min<int64_t>(INT32_MAX, max_tree_size + (int64_t) m.ch[i].w * m.ch[i].h);
            max_tree_size = std::min<int64_t>(1 << 20, max_tree_size);
            read_tree((int32_t) max_tree_size, &lh.tree_off, &lh.spec_off, &lh.uses_wp);
            if (err) return;
            lh.present = true;
            lh.hdr = m;
        }
    }

    // Decodes channels [0, count) of a modular image on the host, where a sub-bitstream sits in the middle of what
    // the host has to read anyway: the extra channels of a single-section VarDCT frame (between LfGlobal and
    // HfGlobal) and RAW dequantisation matrices (inside HfGlobal). Runs the device's own channel decoder
    // (modular_channel_t is __host__ __device__); a few thousand to 65536 samples per channel.
    void host_modular_channels(ModImage &m, int count, const FramePlan::LocalHeader &lh, int32_t sidx, std::vector<std::vector<int16_t>> &planes) {
        const DFrame &d = plan.df;
        HostCode hc;
        hc.begin(arena, lh.present ? lh.spec_off : d.global_spec_off);
        const DTreeNode *tree = (const DTreeNode *) (arena.bytes.data() + (lh.present ? lh.tree_off : d.global_tree_off));
        const bool uses_wp = (lh.present ? lh.uses_wp : d.global_tree_uses_wp) != 0;
        planes.assign((size_t) m.num_channels, std::vector<int16_t>());
        int maxw = 1;
        for (int c = 0; c < m.num_channels; ++c) {
            planes[(size_t) c].assign((size_t) std::max(1, m.ch[c].w) * (size_t) std::max(1, m.ch[c].h), 0);
            m.ch[c].px = planes[(size_t) c].data();
            m.ch[c].stride = m.ch[c].w;
            maxw = std::max(maxw, m.ch[c].w);
        }
        std::vector<int32_t> wp_scratch(uses_wp ? (size_t) maxw * 10 : 1);
        for (int c = 0; c < count && !err; ++c) {
            if (m.ch[c].w <= 0 || m.ch[c].h <= 0) continue;
            if (uses_wp) modular_channel_t<true>(br, hc.es, hc.cc, hc.cs, tree, wp_scratch.data(), h_div24_table.v, m, c, sidx);
            else modular_channel_t<false>(br, hc.es, hc.cc, hc.cs, tree, wp_scratch.data(), h_div24_table.v, m, c, sidx);
            if (hc.es.err) err = hc.es.err;
            else if (overrun()) err = E_SHRT;
        }
        hfinish(hc);
    }

    void lf_global() { // j40.h:6257-6340
        ImageInfo &im = plan.im;
        FrameInfo &f = plan.fh;
        DFrame &d = plan.df;
        CHECK(!f.has_patches && !f.has_splines && !f.has_noise, E_TODO);
        d.m_lf_scaled[0] = 1.0f / 4096.0f; d.m_lf_scaled[1] = 1.0f / 512.0f; d.m_lf_scaled[2] = 1.0f / 256.0f;
        if (!u(1)) for (int i = 0; i < 3; ++i) d.m_lf_scaled[i] = f16() / 128.0f;
        d.inv_colour_factor = 1 / 84.0f;
        d.base_corr_x = 0.0f; d.base_corr_b = 1.0f;
        int32_t x_factor_lf = 0, b_factor_lf = 0;
        if (!f.is_modular) {
            d.global_scale = (int32_t) u32(1, 11, 2049, 11, 4097, 12, 8193, 16);
            d.quant_lf = (int32_t) u32(16, 0, 1, 5, 1, 8, 1, 16);
            std::vector<uint8_t> map;
            if (u(1)) {
                static const uint8_t DEF[39] = {
                    0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6, 7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14,
                    7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14,
                };
                map.assign(DEF, DEF + 39);
                d.nb_qf_thr = d.nb_lf_thr[0] = d.nb_lf_thr[1] = d.nb_lf_thr[2] = 0;
                d.nb_block_ctx = 15;
            } else {
                if (err) return;
                int32_t size = 39;
                for (int i = 0; i < 3; ++i) {
                    d.nb_lf_thr[i] = (int32_t) u(4);
                    for (int j = 0; j < d.nb_lf_thr[i]; ++j) {
                        // 64-bit variant of U32 with a 32-bit last field (j40.h:6292)
                        uint32_t sel = u(2);
                        uint64_t v = sel == 0 ? u(4) : sel == 1 ? (uint64_t) u(8) + 16 : sel == 2 ? (uint64_t) u(16) + 272 : (((uint64_t) u(32) + 65808) & 0xffffffffull);
                        int64_t s = (v & 1) ? -(int64_t) (v / 2 + 1) : (int64_t) (v / 2);
                        d.lf_thr[i][j] = (int32_t) s;
                    }
                    size *= d.nb_lf_thr[i] + 1;
                }
                d.nb_qf_thr = (int32_t) u(4);
                for (int i = 0; i < d.nb_qf_thr; ++i) d.qf_thr[i] = (int32_t) u32(0, 2, 4, 3, 12, 5, 44, 8) + 1;
                size *= d.nb_qf_thr + 1;
                CHECK(size <= 39 * 64, E4("hfbc"));
                cluster_map(size, 16, &d.nb_block_ctx, map);
                if (err) return;
            }
            d.block_ctx_size = (int32_t) map.size();
            d.block_ctx_map_off = arena.alloc(map.size(), 8);
            memcpy(arena.at<uint8_t>(d.block_ctx_map_off), map.data(), map.size());
            if (!u(1)) {
                d.inv_colour_factor = 1.0f / (float) u32(84, 0, 256, 0, 2, 8, 258, 16);
                d.base_corr_x = f16();
                d.base_corr_b = f16();
                x_factor_lf = (int32_t) u(8) - 127;
                b_factor_lf = (int32_t) u(8) - 127;
            }
        }
        d.kx_lf = d.base_corr_x + (float) x_factor_lf * d.inv_colour_factor;
        d.kb_lf = d.base_corr_b + (float) b_factor_lf * d.inv_colour_factor;
        init_global_modular();
        if (err) return;
        if (u(1)) { // global tree
            int64_t px = std::min<int64_t>((int64_t) f.width * f.height, INT32_MAX);
            int64_t t = std::min<int64_t>(px * plan.gmod.num_channels, INT32_MAX) / 16;
            int32_t max_tree_size = (int32_t) std::min<int64_t>(1 << 22, 1024 + t);
            read_tree(max_tree_size, &d.global_tree_off, &d.global_spec_off, &d.global_tree_uses_wp);
            d.have_global_tree = 1;
        }
        if (err) return;
        if (plan.gmod.num_channels > 0) {
            host_modular_header(plan.gmod, plan.gmod_local, f.is_modular);
            if (err) return;
            check_overrun();
            if (err) return;
            if (f.width <= (1 << f.group_size_shift) && f.height <= (1 << f.group_size_shift)) plan.num_gm_channels = plan.gmod.num_channels;
            else plan.num_gm_channels = plan.gmod.nb_meta_channels; // palettes are always coded with the global image
            plan.gmod_has_stream = true;
            plan.gmod_sec.start_bit = br.bits_consumed(); // relative to the reader's base (set by the caller)
        }
    }

    // dequantisation matrix encodings other than the library default (j40.h:4696-4770)
    void read_dq_matrix(int idx, int rows, int columns) {
        int mode = (int) u(3);
        if (mode == 0) return;
        if (mode == 7) { // RAW: a 3-channel modular image of columns x rows weights over a common denominator (j40.h:4705-4743)
            const float denom = f16();
            if (err) return;
            CHECK(!(denom == 0.0f), E4("dqm0")); // j40__surely_nonzero
            const float inv_denom = 1.0f / denom;
            ModImage m;
            memset(&m, 0, sizeof(m));
            m.num_channels = 3;
            for (int c = 0; c < 3; ++c) { m.ch[c].w = columns; m.ch[c].h = rows; m.ch[c].stride = columns; }
            FramePlan::LocalHeader lh;
            host_modular_header(m, lh);
            if (err) return;
            std::vector<std::vector<int16_t>> planes;
            host_modular_channels(m, 3, lh, (int32_t) (1 + 3 * plan.fh.num_lf_groups + idx), planes);
            if (err) return;
            for (int t = m.nb_transforms - 1; t >= 0; --t) {
                ModImage one = m;
                one.nb_transforms = 1;
                one.tr[0] = m.tr[t];
                inverse_transforms(one, 0, 1);
            }
            std::vector<float> mat((size_t) rows * columns * 3);
            for (int c = 0; c < 3; ++c) for (int i = 0; i < rows * columns; ++i) mat[(size_t) i * 3 + (size_t) c] = (float) planes[(size_t) c][(size_t) i] * inv_denom;
            uint32_t off = arena.alloc(mat.size() * 4, 16);
            memcpy(arena.at<uint8_t>(off), mat.data(), mat.size() * 4);
            plan.custom_dq_off[idx] = off;
            return;
        }
        static const int8_t HOW[7][4] = {{0, 0, 0, 0}, {1, 3, 3, 0}, {1, 6, 6, 0}, {1, 2, 2, 1}, {1, 1, 0, 1}, {1, 9, 6, 2}, {1, 0, 0, 1}};
        int nparams = HOW[mode][1], nscaled = HOW[mode][2], ndct = HOW[mode][3];
        if (HOW[mode][0]) CHECK(rows == 8 && columns == 8, E4("dqm?"));
        std::vector<float> params((size_t) (nparams + ndct * 16) * 3, 0.0f); // [j][c]
        int n = 0, m = 0, pidx = nparams;
        for (int c = 0; c < 3; ++c) for (int j = 0; j < nparams; ++j) params[(size_t) j * 3 + (size_t) c] = f16() * (j < nscaled ? 64.0f : 1.0f);
        for (int i = 0; i < ndct; ++i) {
            int k = (int) u(4) + 1;
            if (i == 0) n = k; else m = k;
            for (int c = 0; c < 3; ++c) for (int j = 0; j < k; ++j) params[(size_t) (pidx + j) * 3 + (size_t) c] = f16() * (j == 0 ? 64.0f : 1.0f);
            pidx += k;
        }
        check_overrun();
        if (err) return;
        std::vector<float> mat = compute_dq(idx, mode, n, m, params.data());
        if (mat.empty()) {
            // the reference computes the weights later (j40__load_dq_matrix via j40__prepare_dq_matrices, after the LF
            // groups) and therefore reports anything else wrong with the rest of HfGlobal first: keep parsing and
            // raise "band" at the end of the host parse if nothing else failed. (It also computes them only for
            // transforms the frame uses; an invalid but unused matrix still fails here -- DESIGN.md §8.)
            if (!plan.deferred_err) plan.deferred_err = E4("band");
            return;
        }
        uint32_t off = arena.alloc(mat.size() * 4, 16);
        memcpy(arena.at<uint8_t>(off), mat.data(), mat.size() * 4);
        plan.custom_dq_off[idx] = off;
    }
    static std::vector<float> compute_dq(int idx, int mode, int n, int m, const float *params);

    void hf_global() { // j40.h:6819-6866
        FrameInfo &f = plan.fh;
        DFrame &d = plan.df;
        static const int8_t PARAM_DIMS[17][2] = {{3, 3}, {3, 3}, {3, 3}, {3, 3}, {4, 4}, {5, 5}, {3, 4}, {3, 5}, {4, 5}, {3, 3}, {3, 3},
                                                 {6, 6}, {5, 6}, {7, 7}, {6, 7}, {8, 8}, {7, 8}};
        static const int8_t ORDER_LOG[13][2] = {{3, 3}, {3, 3}, {4, 4}, {5, 5}, {3, 4}, {3, 5}, {4, 5}, {6, 6}, {5, 6}, {7, 7}, {6, 7}, {8, 8}, {7, 8}};
        if (!u(1)) {
            for (int i = 0; i < 17; ++i) {
                read_dq_matrix(i, 1 << PARAM_DIMS[i][0], 1 << PARAM_DIMS[i][1]);
                if (err) return;
            }
        }
        d.num_hf_presets = (int32_t) u(ceil_lg32((uint32_t) f.num_groups)) + 1;
        check_overrun();
        if (err) return;
        // HfPass, once per pass: custom coefficient orders, then the pass's coefficient code spec
        d.num_passes = f.num_passes;
        for (int pass = 0; pass < f.num_passes; ++pass) {
            int32_t used_orders = (int32_t) u32(0x5f, 0, 0x13, 0, 0, 0, 0, 13);
            if (used_orders > 0) {
                uint32_t spec = read_code_spec(8);
                if (err) return;
                HostCode h;
                h.begin(arena, spec);
                const GlobalTables &gt = GlobalTables::get();
                pending_orders.clear();
                for (int j = 0; j < 13; ++j) if (used_orders >> j & 1) {
                    int32_t size = 1 << (ORDER_LOG[j][0] + ORDER_LOG[j][1]);
                    for (int c = 0; c < 3; ++c) {
                        std::vector<int32_t> lehmer;
                        read_permutation(h, size, size / 64, lehmer);
                        if (err) return;
                        if (lehmer.empty()) continue;
                        std::vector<int32_t> order = gt.order[j];
                        apply_permutation(order.data() + size / 64, lehmer);
                        // (the arena may grow here; `h` holds pointers into it, so allocate afterwards)
                        pending_orders.push_back({j, c, std::move(order)});
                    }
                }
                hfinish(h);
                if (err) return;
                for (auto &po : pending_orders) {
                    uint32_t off = arena.alloc(po.order.size() * 4, 16);
                    memcpy(arena.at<uint8_t>(off), po.order.data(), po.order.size() * 4);
                    plan.custom_order_off[pass][po.j][po.c] = off;
                }
            }
            d.coeff_spec_off[pass] = read_code_spec(495 * d.nb_block_ctx * d.num_hf_presets);
            if (err) return;
        }
    }
    struct PendingOrder { int j, c; std::vector<int32_t> order; };
    std::vector<PendingOrder> pending_orders;
};

// ---------------------------------------------------------------------------------------------
// dequantisation matrices (j40.h:4780-4978). Uses libm (powf, hypotf) exactly like the reference.

static const float kLibraryParams[129][3] = {
    // DCT8x8
    {3150.0f, 560.0f, 512.0f}, {0.0f, 0.0f, -2.0f}, {-0.4f, -0.3f, -1.0f}, {-0.4f, -0.3f, 0.0f}, {-0.4f, -0.3f, -1.0f}, {-2.0f, -0.3f, -2.0f},
    // Hornuss
    {280.0f, 60.0f, 18.0f}, {3160.0f, 864.0f, 200.0f}, {3160.0f, 864.0f, 200.0f},
    // DCT2x2
    {3840.0f, 960.0f, 640.0f}, {2560.0f, 640.0f, 320.0f}, {1280.0f, 320.0f, 128.0f}, {640.0f, 180.0f, 64.0f}, {480.0f, 140.0f, 32.0f}, {300.0f, 120.0f, 16.0f},
    // DCT4x4: params + bands
    {1.0f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, {2200.0f, 392.0f, 112.0f}, {0.0f, 0.0f, -0.25f}, {0.0f, 0.0f, -0.25f}, {0.0f, 0.0f, -0.5f},
    // DCT16x16
    {8996.8725711814115328f, 3191.48366296844234752f, 1157.50408145487200256f},
    {-1.3000777393353804f, -0.67424582104194355f, -2.0531423165804414f},
    {-0.49424529824571225f, -0.80745813428471001f, -1.4f},
    {-0.439093774457103443f, -0.44925837484843441f, -0.50687130033378396f},
    {-0.6350101832695744f, -0.35865440981033403f, -0.42708730624733904f},
    {-0.90177264050827612f, -0.31322389111877305f, -1.4856834539296244f},
    {-1.6162099239887414f, -0.37615025315725483f, -4.9209142884401604f},
    // DCT32x32
    {15718.40830982518931456f, 7305.7636810695983104f, 3803.53173721215041536f},
    {-1.025f, -0.8041958212306401f, -3.060733579805728f},
    {-0.98f, -0.7633036457487539f, -2.0413270132490346f},
    {-0.9012f, -0.55660379990111464f, -2.0235650159727417f},
    {-0.4f, -0.49785304658857626f, -0.5495389509954993f},
    {-0.48819395464f, -0.43699592683512467f, -0.4f},
    {-0.421064f, -0.40180866526242109f, -0.4f},
    {-0.27f, -0.27321683125358037f, -0.3f},
    // DCT8x16
    {7240.7734393502f, 1448.15468787004f, 506.854140754517f},
    {-0.7f, -0.5f, -1.4f}, {-0.7f, -0.5f, -0.2f}, {-0.2f, -0.5f, -0.5f}, {-0.2f, -0.2f, -0.5f}, {-0.2f, -0.2f, -1.5f}, {-0.5f, -0.2f, -3.6f},
    // DCT8x32
    {16283.2494710648897f, 5089.15750884921511936f, 3397.77603275308720128f},
    {-1.7812845336559429f, -0.320049391452786891f, -0.321327362693153371f},
    {-1.6309059012653515f, -0.35362849922161446f, -0.34507619223117997f},
    {-1.0382179034313539f, -0.30340000000000003f, -0.70340000000000003f},
    {-0.85f, -0.61f, -0.9f}, {-0.7f, -0.5f, -1.0f}, {-0.9f, -0.5f, -1.0f},
    {-1.2360638576849587f, -0.6f, -1.1754605576265209f},
    // DCT16x32
    {13844.97076442300573f, 4798.964084220744293f, 1807.236946760964614f},
    {-0.97113799999999995f, -0.61125308982767057f, -1.2f},
    {-0.658f, -0.83770786552491361f, -1.2f}, {-0.42026f, -0.79014862079498627f, -0.7f},
    {-0.22712f, -0.2692727459704829f, -0.7f}, {-0.2206f, -0.38272769465388551f, -0.7f},
    {-0.226f, -0.22924222653091453f, -0.4f}, {-0.6f, -0.20719098826199578f, -0.5f},
    // DCT4x8: param + bands
    {1.0f, 1.0f, 1.0f},
    {2198.050556016380522f, 764.3655248643528689f, 527.107573587542228f},
    {-0.96269623020744692f, -0.92630200888366945f, -1.4594385811273854f},
    {-0.76194253026666783f, -0.9675229603596517f, -1.450082094097871593f},
    {-0.6551140670773547f, -0.27845290869168118f, -1.5843722511996204f},
    // AFV: 9 params + 4x8 bands + 4x4 bands
    {3072.0f, 1024.0f, 384.0f}, {3072.0f, 1024.0f, 384.0f}, {256.0f, 50.0f, 12.0f}, {256.0f, 50.0f, 12.0f}, {256.0f, 50.0f, 12.0f},
    {414.0f, 58.0f, 22.0f}, {0.0f, 0.0f, -0.25f}, {0.0f, 0.0f, -0.25f}, {0.0f, 0.0f, -0.25f},
    {2198.050556016380522f, 764.3655248643528689f, 527.107573587542228f},
    {-0.96269623020744692f, -0.92630200888366945f, -1.4594385811273854f},
    {-0.76194253026666783f, -0.9675229603596517f, -1.450082094097871593f},
    {-0.6551140670773547f, -0.27845290869168118f, -1.5843722511996204f},
    {2200.0f, 392.0f, 112.0f}, {0.0f, 0.0f, -0.25f}, {0.0f, 0.0f, -0.25f}, {0.0f, 0.0f, -0.5f},
#define LARGE(m) \
    {m * 23629.073922049845f, m * 8611.3238710010046f, m * 4492.2486445538634f}, \
    {-1.025f, -0.3041958212306401f, -1.2f}, {-0.78f, 0.3633036457487539f, -1.2f}, \
    {-0.65012f, -0.35660379990111464f, -0.8f}, {-0.19041574084286472f, -0.3443074455424403f, -0.7f}, \
    {-0.20819395464f, -0.33699592683512467f, -0.7f}, {-0.421064f, -0.30180866526242109f, -0.4f}, \
    {-0.32733845535848671f, -0.27321683125358037f, -0.5f}
    LARGE(0.9f), LARGE(0.65f), LARGE(1.8f), LARGE(1.3f), LARGE(3.6f), LARGE(2.6f),
#undef LARGE
};
// per parameter set: log rows, log columns, offset into kLibraryParams, default mode, n, m
static const int16_t kDctParams[17][6] = {
    {3, 3, 0, 6, 6, 0}, {3, 3, 6, 1, 0, 0}, {3, 3, 9, 2, 0, 0}, {3, 3, 15, 3, 4, 0}, {4, 4, 21, 6, 7, 0}, {5, 5, 28, 6, 8, 0},
    {3, 4, 36, 6, 7, 0}, {3, 5, 43, 6, 8, 0}, {4, 5, 51, 6, 8, 0}, {3, 3, 59, 4, 4, 0}, {3, 3, 64, 5, 4, 4}, {6, 6, 81, 6, 8, 0},
    {5, 6, 89, 6, 8, 0}, {7, 7, 97, 6, 8, 0}, {6, 7, 105, 6, 8, 0}, {8, 8, 113, 6, 8, 0}, {7, 8, 121, 6, 8, 0},
};

static float dq_interpolate(float pos, int c, const float (*bands)[3], int len) { // j40.h:4780
    if (len == 1) return bands[0][c];
    float scaled_pos = pos * (float) (len - 1);
    int32_t scaled_idx = (int32_t) scaled_pos;
    float frac_idx = scaled_pos - (float) scaled_idx;
    float a = bands[scaled_idx][c], b = bands[scaled_idx + 1][c];
    return a * powf(b / a, frac_idx);
}
static bool dq_bands(const float *params /*[i][3]*/, int n, float (*out)[3]) { // j40.h:4792
    for (int c = 0; c < 3; ++c) {
        out[0][c] = params[c];
        if (!(out[0][c] > 0)) return false;
        for (int i = 1; i < n; ++i) {
            float v = params[i * 3 + c];
            out[i][c] = v > 0 ? out[i - 1][c] * (1.0f + v) : out[i - 1][c] / (1.0f - v);
            if (!(out[i][c] > 0)) return false;
        }
    }
    return true;
}
static void dq_weights(int rows, int columns, const float (*bands)[3], int len, float *out /*[i][3]*/) { // j40.h:4811
    float inv_rows_m1 = 1.0f / (float) (rows - 1), inv_columns_m1 = 1.0f / (float) (columns - 1);
    static const float INV_SQRT2 = 1.0f / 1.414214562373095f;
    for (int c = 0; c < 3; ++c) for (int y = 0; y < rows; ++y) for (int x = 0; x < columns; ++x) {
        float d = hypotf((float) x * inv_columns_m1, (float) y * inv_rows_m1);
        out[(y * columns + x) * 3 + c] = dq_interpolate(d * INV_SQRT2, c, bands, len);
    }
}

std::vector<float> Parser::compute_dq(int idx, int mode, int n, int m, const float *params) { // j40.h:4828-4978
    const int rows = 1 << kDctParams[idx][0], columns = 1 << kDctParams[idx][1];
    std::vector<float> raw((size_t) rows * (size_t) columns * 3, 0.0f);
    float bands[15][3], scratch[64 * 3];
    auto P = [&](int i, int c) { return params[i * 3 + c]; };
    switch (mode) {
    case 6:
        if (n > 15 || !dq_bands(params, n, bands)) return {};
        dq_weights(rows, columns, bands, n, raw.data());
        break;
    case 3:
        if (n > 15 || !dq_bands(params + 2 * 3, n, bands)) return {};
        dq_weights(4, 4, bands, n, scratch);
        for (int c = 0; c < 3; ++c) {
            for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) raw[(size_t) (y * 8 + x) * 3 + (size_t) c] = scratch[((y / 2) * 4 + (x / 2)) * 3 + c];
            raw[1 * 3 + (size_t) c] /= P(0, c);
            raw[8 * 3 + (size_t) c] /= P(0, c);
            raw[9 * 3 + (size_t) c] /= P(1, c);
        }
        break;
    case 2:
        for (int c = 0; c < 3; ++c) {
            static const int8_t MAP[64] = {
                0, 0, 2, 2, 4, 4, 4, 4, 0, 1, 2, 2, 4, 4, 4, 4, 2, 2, 3, 3, 4, 4, 4, 4, 2, 2, 3, 3, 4, 4, 4, 4,
                4, 4, 4, 4, 5, 5, 5, 5, 4, 4, 4, 4, 5, 5, 5, 5, 4, 4, 4, 4, 5, 5, 5, 5, 4, 4, 4, 4, 5, 5, 5, 5,
            };
            for (int i = 0; i < 64; ++i) raw[(size_t) i * 3 + (size_t) c] = P(MAP[i], c);
            raw[(size_t) c] = -1.0f;
        }
        break;
    case 1:
        for (int c = 0; c < 3; ++c) {
            for (int i = 0; i < 64; ++i) raw[(size_t) i * 3 + (size_t) c] = P(0, c);
            raw[(size_t) c] = 1.0f;
            raw[1 * 3 + (size_t) c] = raw[8 * 3 + (size_t) c] = P(1, c);
            raw[9 * 3 + (size_t) c] = P(2, c);
        }
        break;
    case 4:
        if (n > 15 || !dq_bands(params + 1 * 3, n, bands)) return {};
        dq_weights(4, 8, bands, n, scratch);
        for (int c = 0; c < 3; ++c) {
            for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) raw[(size_t) (y * 8 + x) * 3 + (size_t) c] = scratch[((y / 2) * 8 + x) * 3 + c];
            raw[1 * 3 + (size_t) c] /= P(0, c);
        }
        break;
    case 5: {
        if (n > 15 || m > 15) return {};
        if (!dq_bands(params + 9 * 3, n, bands)) return {};
        dq_weights(4, 8, bands, n, scratch);
        if (!dq_bands(params + (9 + n) * 3, m, bands)) return {};
        dq_weights(4, 4, bands, m, scratch + 32 * 3);
        if (!dq_bands(params + 5 * 3, 4, bands)) return {};
        static const float FREQS[12] = {
            0.000000000f, 0.373436417f, 0.320380100f, 0.379332596f, 0.066671353f, 0.259756761f,
            0.530035651f, 0.789731061f, 0.149436598f, 0.559318823f, 0.669198646f, 0.999999917f,
        };
        static const int8_t MAP[64] = {
            60, 32, 62, 33, 48, 34, 49, 35, 0, 1, 2, 3, 4, 5, 6, 7, 61, 36, 63, 37, 50, 38, 51, 39, 8, 9, 10, 11, 12, 13, 14, 15,
            52, 40, 53, 41, 54, 42, 55, 43, 16, 17, 18, 19, 20, 21, 22, 23, 56, 44, 57, 45, 58, 46, 59, 47, 24, 25, 26, 27, 28, 29, 30, 31,
        };
        for (int c = 0; c < 3; ++c) {
            scratch[0 * 3 + c] = P(0, c);
            scratch[32 * 3 + c] = P(1, c);
            for (int i = 0; i < 12; ++i) scratch[(i + 48) * 3 + c] = dq_interpolate(FREQS[i], c, bands, 4);
            scratch[60 * 3 + c] = 1.0f;
            for (int i = 0; i < 3; ++i) scratch[(i + 61) * 3 + c] = P(i + 2, c);
        }
        for (int c = 0; c < 3; ++c) for (int i = 0; i < 64; ++i) raw[(size_t) i * 3 + (size_t) c] = scratch[MAP[i] * 3 + c];
        break;
    }
    default: return {};
    }
    return raw;
}

} // namespace

std::vector<float> compute_dq_matrix_default(int idx) {
    const int16_t *p = kDctParams[idx];
    return Parser::compute_dq(idx, p[3], p[4], p[5], &kLibraryParams[p[2]][0]);
}

// natural (zig-zag) coefficient order for log_rows <= log_columns (j40.h:4980-5030): the LLF corner first in
// raster order, then anti-diagonals of the aspect-scaled grid, alternating direction
std::vector<int32_t> compute_natural_order(int log_rows, int log_columns) {
    const int size = 1 << (log_rows + log_columns), log_slope = log_columns - log_rows;
    const int rows8 = 1 << (log_rows - 3), columns8 = 1 << (log_columns - 3);
    const int R = 1 << log_rows, C = 1 << log_columns, slope = 1 << log_slope;
    std::vector<int32_t> order;
    order.reserve((size_t) size);
    for (int y = 0; y < rows8; ++y) for (int x = 0; x < columns8; ++x) order.push_back(y << log_columns | x);
    for (int key = columns8; (int) order.size() < size; ++key) {
        // cells with x + y * slope == key; x steps by `slope` as y steps by one
        int x0 = key & (slope - 1), y0 = key >> log_slope, x1 = key, y1 = 0;
        if (x1 >= C) { int ex = ceil_div(x1 - (C - 1), slope); x1 -= ex << log_slope; y1 += ex; }
        if (y0 >= R) { int ex = y0 - (R - 1); x0 += ex << log_slope; y0 -= ex; }
        if (key & 1) {
            for (int x = x1, y = y1; x >= x0; x -= slope, ++y) if (y >= rows8 || x >= columns8) order.push_back(y << log_columns | x);
        } else {
            for (int x = x0, y = y0; x <= x1; x += slope, --y) if (y >= rows8 || x >= columns8) order.push_back(y << log_columns | x);
        }
    }
    return order;
}

// The sample quantiser of the reference, v -> (int16)((2^bpp-1) * sRGB(v) + 0.5) followed by the 8-bit
// render step (j40.h:7233-7235, 7950-7952), is a non-decreasing step function of the float v (checked
// exhaustively for bpp = 8). thr[k] is the smallest positive v mapping to >= k+1.
static int quantise_like_reference(float v, int bpp) {
    float s = (v <= 0.0031308f ? 12.92f * v : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f);
    float q = (float) ((1 << bpp) - 1) * s + 0.5f;
    int32_t i16;
    if (!(q == q)) i16 = 0;
    else if (q >= 32767.0f) i16 = 32767;
    else if (q <= -32768.0f) i16 = -32768;
    else i16 = (int32_t) (int16_t) q;
    int32_t maxpixel = (1 << bpp) - 1, half = 1 << (bpp - 1);
    int32_t p = std::min(std::max(0, i16), maxpixel);
    return (p * 255 + half) / maxpixel;
}
void compute_srgb_thresholds(int bpp, float *thr) {
    for (int k = 0; k < 255; ++k) {
        uint32_t lo = 1, hi = 0x7f7fffffu; // positive finite floats, ordered like their bit patterns
        float fhi;
        memcpy(&fhi, &hi, 4);
        if (quantise_like_reference(fhi, bpp) < k + 1) { uint32_t inf = 0x7f800000u; memcpy(&thr[k], &inf, 4); continue; }
        while (lo < hi) {
            uint32_t mid = lo + (hi - lo) / 2;
            float fm;
            memcpy(&fm, &mid, 4);
            if (quantise_like_reference(fm, bpp) >= k + 1) hi = mid; else lo = mid + 1;
        }
        memcpy(&thr[k], &lo, 4);
    }
}

const GlobalTables &GlobalTables::get() {
    static GlobalTables *g = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        GlobalTables *t = new GlobalTables;
        for (int i = 0; i < 17; ++i) t->dq[i] = compute_dq_matrix_default(i);
        static const int8_t ORDER_LOG[13][2] = {{3, 3}, {3, 3}, {4, 4}, {5, 5}, {3, 4}, {3, 5}, {4, 5}, {6, 6}, {5, 6}, {7, 7}, {6, 7}, {8, 8}, {7, 8}};
        for (int i = 0; i < 13; ++i) t->order[i] = compute_natural_order(ORDER_LOG[i][0], ORDER_LOG[i][1]);
        compute_srgb_thresholds(8, t->srgb_thr);
        t->srgb_thr[255] = NAN;
        { // smallest float whose encoded value (255 * sRGB(v) + 0.5, the expression of j40.h:7233-7234) reaches 32768
            uint32_t lo = 1, hi = 0x7f7fffffu;
            while (lo < hi) {
                uint32_t mid = lo + (hi - lo) / 2;
                float fm;
                memcpy(&fm, &mid, 4);
                float s_ = (fm <= 0.0031308f ? 12.92f * fm : 1.055f * powf(fm, 1.0f / 2.4f) - 0.055f);
                if (255.0f * s_ + 0.5f >= 32768.0f) hi = mid; else lo = mid + 1;
            }
            memcpy(&t->srgb_wrap_hi, &lo, 4);
        }
        memset(t->srgb_lut, 0, sizeof(t->srgb_lut));
        for (int b = 0; b <= SRGB_LUT_N; ++b) {
            int n = 0;
            while (n < 255 && t->srgb_thr[n] <= (float) b / (float) SRGB_LUT_N) ++n;
            t->srgb_lut[b] = (uint8_t) n;
            // srgb_u8_lut() takes one step inside a bucket
            if (b > 0 && n - t->srgb_lut[b - 1] > 1) { fprintf(stderr, "j40_b200: sRGB start table too coarse\n"); abort(); }
        }
        g = t;
    });
    return *g;
}

// ---------------------------------------------------------------------------------------------
// container (j40.h:1393-1648): collects the codestream out of jxlc / jxlp boxes

static uint32_t be32(const uint8_t *p) { return ((uint32_t) p[0] << 24) | ((uint32_t) p[1] << 16) | ((uint32_t) p[2] << 8) | p[3]; }

static uint32_t linearise(const uint8_t *data, size_t size, FramePlan &plan) {
    static const uint8_t JXL_BOX[12] = {0, 0, 0, 0x0c, 'J', 'X', 'L', ' ', 0x0d, 0x0a, 0x87, 0x0a};
    static const uint8_t FTYP_BOX[20] = {0, 0, 0, 0x14, 'f', 't', 'y', 'p', 'j', 'x', 'l', ' ', 0, 0, 0, 0, 'j', 'x', 'l', ' '};
    if (size < 2) return E_SHRT;
    if (data[0] == 0xff && data[1] == 0x0a) { plan.cs = data; plan.cs_size = size; return 0; }
    if (!(data[0] == JXL_BOX[0] && data[1] == JXL_BOX[1])) return E4("!jxl");
    if (size < 32) return E_SHRT;
    if (memcmp(data, JXL_BOX, 12) != 0) return E4("!jxl");
    if (memcmp(data + 12, FTYP_BOX, 20) != 0) return E4("ftyp");
    size_t pos = 32;
    bool seen_jxll = false, seen_jxli = false, seen_jxlc = false, seen_jxlp = false, no_more = false;
    while (pos < size) {
        if (size - pos < 8) {
            // a truncated box header. Before the codestream is complete it is a short file. Behind the last
            // codestream box the reference sees it if it lies in the first 64 KiB of the file, which its first
            // buffer fill scans box by box (j40.h:1676), or when it reads a single-section frame through to its
            // end; a multi-section frame is left by a seek and later refills stay inside the codestream box.
            if (!no_more || pos < 65536) return E_SHRT;
            plan.trailing_partial_box = true; // a single-section frame still runs into it at its end (see collect_errors)
            break;
        }
        uint32_t size32 = be32(data + pos), type = be32(data + pos + 4);
        size_t hdr = 8;
        uint64_t payload;
        bool to_eof = false;
        if (size32 == 0) { to_eof = true; payload = size - pos - 8; }
        else if (size32 == 1) {
            if (size - pos < 16) return E_SHRT;
            uint64_t s64 = ((uint64_t) be32(data + pos + 8) << 32) | be32(data + pos + 12);
            if (s64 < 16) return E4("boxx");
            if (s64 > (uint64_t) INT64_MAX) return E4("flen");
            payload = s64 - 16;
            hdr = 16;
        } else {
            if (size32 < 8) return E4("boxx");
            payload = size32 - 8;
        }
        size_t body = pos + hdr;
        size_t avail = (size_t) std::min<uint64_t>(payload, size - body);
        bool codestream_box = false, stop = false;
        size_t skip = 0;
        switch (type) {
        // A misplaced box behind the last codestream box is seen by the reference like a truncated header there: if it
        // lies in the first 64 KiB of the file (first buffer fill, j40.h:1676) or when a single-section frame is read
        // through to its end (collect_errors); a multi-section frame is left by a seek and never gets that far.
#define J40B_DUP_BOX() do { if (!no_more || pos < 65536) return E4("box?"); plan.trailing_box_err = E4("box?"); stop = true; } while (0)
        case 0x6a786c6c: if (seen_jxll) J40B_DUP_BOX(); seen_jxll = true; break;
        case 0x6a786c69: if (seen_jxli) J40B_DUP_BOX(); seen_jxli = true; break;
        case 0x6a786c63:
            if (no_more || seen_jxlp || seen_jxlc) { J40B_DUP_BOX(); break; }
            seen_jxlc = true; no_more = true; codestream_box = true;
            break;
        case 0x6a786c70:
            if (no_more || seen_jxlc) { J40B_DUP_BOX(); break; }
            seen_jxlp = true; codestream_box = true;
            if (payload < 4) return E4("jxlp");
            if (avail < 4) return E_SHRT;
            // the reference treats a CLEAR top bit of the jxlp index as "last box" (j40.h:1550), the
            // opposite of ISO/IEC 18181-2; kept for parity (DESIGN.md quirk list)
            if (!(data[body] >> 7)) no_more = true;
            skip = 4;
            break;
        case 0x62726f62:
            if (!no_more) {
                if (payload <= 4) return E4("brot");
                if (avail < 4) return E_SHRT;
                uint32_t inner = be32(data + body);
                if (inner == 0x62726f62 || (inner >> 8) == 0x6a786c) return E4("brot");
            }
            break;
        default: break;
        }
        if (stop) break;
        if (codestream_box) plan.cs_owned.insert(plan.cs_owned.end(), data + body + skip, data + body + avail);
        if (to_eof) break;
        if (payload > size - body) break; // truncated box: whatever was there has been taken
        pos = body + (size_t) payload;
    }
    if (!seen_jxlc && !seen_jxlp) return E_SHRT;
    plan.cs = plan.cs_owned.data();
    plan.cs_size = plan.cs_owned.size();
    return 0;
}

// ---------------------------------------------------------------------------------------------

void parse_lf_group_local_tree(FramePlan &plan, size_t lfg, int stage, uint64_t start_bit, int32_t nb_varblocks) {
    const FrameInfo &f = plan.fh;
    if (plan.lfg_local.empty()) plan.lfg_local.resize(2 * plan.lfg_sec.size());
    FramePlan::LocalHeader &lh = plan.lfg_local[2 * lfg + (size_t) stage];
    const SectionRef &s = plan.lfg_sec[lfg];
    const int ggx = (int) (lfg % (size_t) f.ggcolumns), ggy = (int) (lfg / (size_t) f.ggcolumns);
    const int w = std::min(2048, f.width - ggx * 2048), h = std::min(2048, f.height - ggy * 2048);
    const int w8 = ceil_div(w, 8), h8 = ceil_div(h, 8), w64 = ceil_div(w, 64), h64 = ceil_div(h, 64);
    ModImage m;
    memset(&m, 0, sizeof(m));
    if (stage == 0) {
        m.num_channels = 3;
        for (int c = 0; c < 3; ++c) { m.ch[c].w = w8; m.ch[c].h = h8; m.ch[c].stride = w8; }
    } else {
        m.num_channels = 4;
        m.ch[0].w = m.ch[1].w = w64; m.ch[0].h = m.ch[1].h = h64;
        m.ch[2].w = nb_varblocks; m.ch[2].h = 2;
        m.ch[3].w = w8; m.ch[3].h = h8;
        for (int c = 0; c < 4; ++c) m.ch[c].stride = m.ch[c].w;
    }
    Parser h_(plan);
    h_.br.init(plan.cs + s.off, s.size, start_bit);
    h_.host_modular_header(m, lh);
    if (!h_.err) h_.check_overrun();
    lh.present = true;
    if (h_.err) { lh.host_err = h_.err; return; }
    lh.start_bit = h_.br.bits_consumed();
}

uint32_t parse_frame(const uint8_t *data, size_t size, FramePlan &plan) {
    memset(&plan.df, 0, sizeof(plan.df));
    plan.err = linearise(data, size, plan);
    if (plan.err) return plan.err;
    Parser p(plan);
    p.br.init(plan.cs, (uint32_t) std::min<size_t>(plan.cs_size, 0xffffffffu));
    // signature
    if (p.u(16) != 0x0aff) return plan.err = p.overrun() ? (uint32_t) E_SHRT : E4("!jxl");
    p.image_metadata();
    if (!p.err && plan.im.want_icc) p.icc();
    if (!p.err) p.frame_header();
    if (!p.err && !plan.fh.is_last) p.err = E_TODO;
    if (!p.err && plan.fh.type != 0) p.err = E_TODO;
    if (!p.err) p.read_toc();
    if (p.err) return plan.err = p.err;

    const ImageInfo &im = plan.im;
    const FrameInfo &f = plan.fh;
    DFrame &d = plan.df;
    d.width = f.width; d.height = f.height;
    d.is_modular = f.is_modular; d.xyb_encoded = im.xyb_encoded; d.bpp = im.bpp;
    d.group_size_shift = f.group_size_shift;
    d.num_groups = (int32_t) f.num_groups; d.num_lf_groups = (int32_t) f.num_lf_groups;
    d.gcolumns = f.gcolumns; d.grows = f.grows; d.ggcolumns = f.ggcolumns; d.ggrows = f.ggrows;
    d.skip_adapt_lf_smooth = f.skip_adapt_lf_smooth;
    static const float QM_SCALE[8] = {1.5625f, 1.25f, 1.0f, 0.8f, 0.64f, 0.512f, 0.4096f, 0.32768f};
    d.x_qm_mult = QM_SCALE[f.x_qm_scale]; d.b_qm_mult = QM_SCALE[f.b_qm_scale];
    for (int i = 0; i < 3; ++i) {
        d.quant_bias[i] = im.quant_bias[i];
        d.opsin_bias[i] = im.opsin_bias[i];
        d.cbrt_opsin_bias[i] = cbrtf(im.opsin_bias[i]);
        for (int j = 0; j < 3; ++j) d.opsin_inv_mat[i * 3 + j] = im.opsin_inv_mat[i][j];
    }
    d.quant_bias_num = im.quant_bias_num;
    d.itscale = 255.0f / im.intensity_target;
    d.alpha_channel = -1;

    // ---- LfGlobal (its own section unless the frame is a single section)
    uint64_t avail = plan.cs_size > p.lf_global_off ? plan.cs_size - p.lf_global_off : 0;
    Parser q(plan);
    q.br.init(plan.cs + std::min<uint64_t>(p.lf_global_off, plan.cs_size), (uint32_t) std::min<uint64_t>(p.lf_global_size, avail));
    q.lf_global();
    if (q.err) return plan.err = q.err;
    if (plan.gmod_has_stream) {
        plan.gmod_sec.off = p.lf_global_off;
        plan.gmod_sec.size = q.br.size;
        plan.gmod_sec.rank = -2;
    }
    if (plan.gmod_has_stream && plan.num_gm_channels == 0) {
        // no globally coded channel: the (empty) entropy stream still has to be closed (j40.h:6337)
        Parser::HostCode hc;
        hc.begin(plan.arena, plan.gmod_local.present ? plan.gmod_local.spec_off : d.global_spec_off);
        q.hfinish(hc);
        if (q.err) return plan.err = q.err;
        plan.gmod_has_stream = false;
    }
    if (!plan.single_section) {
        // section-end padding/excess errors are dropped by the reference in multi-section frames
        // (j40.h:7791-7798); only running short counts
        if (q.br.overrun()) return plan.err = E_SHRT;
        // (with a global modular stream the device finishes the section and reports its error)
        // ---- HfGlobal
        if (f.is_modular) {
            if (p.hf_global_size != 0) return plan.err = E_EXCS;
        } else {
            uint64_t av2 = plan.cs_size > p.hf_global_off ? plan.cs_size - p.hf_global_off : 0;
            Parser h(plan);
            h.br.init(plan.cs + std::min<uint64_t>(p.hf_global_off, plan.cs_size), (uint32_t) std::min<uint64_t>(p.hf_global_size, av2));
            h.hf_global();
            if (!h.err) { h.check_overrun(); }
            if (h.err) return plan.err = h.err;
        }
        for (auto &s : plan.lfg_sec) { uint64_t a = plan.cs_size > s.off ? plan.cs_size - s.off : 0; s.size = (uint32_t) std::min<uint64_t>(s.size, a); s.off = std::min<uint64_t>(s.off, plan.cs_size); }
        for (auto &s : plan.pg_sec) { uint64_t a = plan.cs_size > s.off ? plan.cs_size - s.off : 0; s.size = (uint32_t) std::min<uint64_t>(s.size, a); s.off = std::min<uint64_t>(s.off, plan.cs_size); }
        if (f.is_modular && plan.gmod.num_channels > 0) {
            // pass groups whose modular header (the first thing in the section) names a local tree: header,
            // tree and code spec are read here. A section too short to hold them fails like on the device.
            const int gsize = 1 << f.group_size_shift;
            for (size_t g = 0; g < plan.pg_sec.size(); ++g) {
                const SectionRef &s = plan.pg_sec[g];
                if (s.size == 0 || (plan.cs[s.off] & 1)) continue; // use_global_tree = 1 (or nothing to read: the device reports it)
                if (plan.pg_local.empty()) plan.pg_local.resize(plan.pg_sec.size());
                Parser h(plan);
                h.br.init(plan.cs + s.off, s.size);
                ModImage m = plan.gmod;
                const int gx = (int) (g % (size_t) f.gcolumns) * gsize, gy = (int) (g / (size_t) f.gcolumns) * gsize;
                for (int c = 0; c < m.num_channels; ++c) { m.ch[c].w = std::min(f.width, gx + gsize) - gx; m.ch[c].h = std::min(f.height, gy + gsize) - gy; }
                FramePlan::LocalHeader &lh = plan.pg_local[g];
                h.host_modular_header(m, lh);
                if (!h.err) h.check_overrun();
                if (h.err) {
                    // errors surface in the reference's decoding order: leave this one to the section's slot
                    lh.present = true; lh.host_err = h.err;
                    continue;
                }
                lh.start_bit = h.br.bits_consumed();
            }
        }
    } else {
        // single section: LfGlobal, HfGlobal, LfGroup, PassGroup back to back
        if (!f.is_modular) {
            if (plan.gmod_has_stream) {
                // extra channels of a single-group VarDCT frame: coded right here, between LfGlobal and HfGlobal;
                // the reference decodes and then discards them (j40.h:7869-7870)
                if (plan.gmod.nb_transforms != 0) return plan.err = E_TODO;
                std::vector<std::vector<int16_t>> scratch;
                ModImage m = plan.gmod;
                q.host_modular_channels(m, plan.num_gm_channels, plan.gmod_local, 0, scratch);
                if (q.err) return plan.err = q.err;
                plan.gmod_has_stream = false;
            }
            q.hf_global();
            if (!q.err) q.check_overrun();
            if (q.err) return plan.err = q.err;
            SectionRef s;
            s.off = p.lf_global_off; s.size = q.br.size; s.start_bit = q.br.bits_consumed(); s.rank = 0;
            plan.lfg_sec.assign(1, s);
            s.rank = 1; s.start_bit = ~0ull; // continues where the LF group ends
            plan.pg_sec.assign(1, s);
        } else {
            // all channels were coded globally; LF group and pass group are empty
            plan.lfg_sec.clear();
            plan.pg_sec.clear();
        }
    }
    // extra channels / alpha
    d.num_channels = plan.gmod.num_channels;
    d.num_gm_channels = plan.num_gm_channels;
    d.nb_meta_channels = plan.gmod.nb_meta_channels;
    int restored = plan.gmod.num_channels; // channel count once the global transforms are undone
    for (int i = 0; i < plan.gmod.nb_transforms; ++i) if (plan.gmod.tr[i].kind == 1) restored += plan.gmod.tr[i].num_c - 2;
    d.num_out_channels = restored;
    if (f.is_modular) {
        for (int i = 3; i < restored; ++i) {
            const ImageInfo::EC &ec = im.ec[i - 3];
            if (ec.type == 0) {
                if (!(ec.bpp == im.bpp && ec.exp_bits == im.exp_bits) || ec.alpha_associated) return plan.err = E_TODO;
                d.alpha_channel = i;
                break;
            }
        }
        if (restored < 3) return plan.err = E_TODO; // grey modular frames (reference: assertion)
        if (im.bpp < 8 || im.exp_bits != 0) return plan.err = E_TODO;
    } else {
        if (f.do_ycbcr || im.cspace_grey) return plan.err = E_TODO;
        if (f.use_lf_frame) return plan.err = E_TODO;
        if (f.jpeg_upsampling) return plan.err = E_TODO;
        if (im.bpp < 8 || im.exp_bits != 0) return plan.err = E_TODO;
        if (plan.gmod.num_channels > 0 && !plan.single_section) {
            // VarDCT frame with extra channels (typically alpha): the reference decodes them -- per pass group,
            // behind the HF coefficients (j40.h:7024-7033) -- and then discards them (j40.h:7869-7870; A = 255).
            // Each pass group has its own byte window, so the device simply stops after the coefficients; the
            // price is that corruption confined to the extra channels' data goes unnoticed here.
            if (plan.num_gm_channels != 0 || plan.gmod.nb_transforms != 0) return plan.err = E_TODO;
        }
    }
    // a palette with delta entries is undone by a scan kernel ahead of the per-pixel render step: only where that is the
    // first thing to undo, i.e. the last transform of the list
    for (int i = 0; i + 1 < plan.gmod.nb_transforms; ++i) if (plan.gmod.tr[i].kind == 1 && plan.gmod.tr[i].nb_deltas > 0) return plan.err = E_TODO;
    d.nb_global_transforms = plan.gmod.nb_transforms;
    for (int i = 0; i < plan.gmod.nb_transforms; ++i) d.global_tr[i] = plan.gmod.tr[i];
    d.global_wp = plan.gmod.wp;
    if (plan.deferred_err) return plan.err = plan.deferred_err;
    return 0;
}

} // namespace j40b
