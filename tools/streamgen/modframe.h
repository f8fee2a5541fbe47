// JPEG XL test-stream writer -- TEST INFRASTRUCTURE.
// Modular (lossless, "fjxl-shaped") frame writer: 8-bit RGB(A), reversible colour transform,
// global MA tree, per-group modular sub-bitstreams, prefix codes + LZ77 run-lengths or ANS
// (SURVEY.md §8d C4, App. E.2/E.3/E.6).
#pragma once
#include <atomic>
#include <exception>
#include <mutex>
#include <thread>
#include "modular.h"
#include "synth.h"
#include "vardct.h"

namespace jxlgen {

struct ModularParams {
    int width = 256, height = 256;
    uint64_t seed = 0;
    int group_shift = 8;     // 7..10
    int rct_type = 6;        // 0..41 (type % 7 = transform, type / 7 = permutation); -1 = no RCT
    int tree_preset = 1;     // 0 single gradient leaf; 1 fjxl-like (gradient, contexts by |residual-ish| props); 2 WP tree
    bool use_ans = false;    // fjxl uses prefix codes
    bool lz77 = true;
    bool alpha = false;      // add an 8-bit alpha extra channel
    bool container = false;
    int smooth = 1;          // 0 = raw synthetic photo, 1 = posterise a little so that runs exist
    int max_clusters = 8;
    int palette = 0;         // 1: RGB coded through a palette transform (colours posterised to <= 216), no RCT; a few rows
                             // use the implicit entries (index < 0 and index >= nb_colours)
    int pal_deltas = 0;      // palette: nb_deltas (indices below it are deltas on top of predictor `pal_pred`, j40.h:4402-4490)
    int pal_pred = 5;
    int local_tree = 0;      // 1: odd pass groups (or the global image of a single-group frame) carry a tree and code
                             // spec of their own; 2: all of them do and the frame has no global tree at all
};

class ModularFrameEncoder {
public:
    ModularParams P;
    GenStats stats;
    explicit ModularFrameEncoder(const ModularParams &p) : P(p) {}

    std::vector<uint8_t> encode(const ImageRGB8 &im) {
        JG_CHECK(im.w == P.width && im.h == P.height);
        int W = P.width, H = P.height, nch = P.alpha ? 4 : 3;
        // full-frame planes after the forward transform
        std::vector<std::vector<int32_t>> plane((size_t) nch, std::vector<int32_t>((size_t) W * (size_t) H));
        Rng rng(P.seed * 31 + 5);
        Channel pal; // palette transform: meta channel nb_colours x 3 (j40.h:3762-3792)
        if (P.palette) {
            nch = P.alpha ? 2 : 1;
            plane.assign((size_t) nch, std::vector<int32_t>((size_t) W * (size_t) H));
            std::map<uint32_t, int> index;
            std::vector<uint32_t> colours;
            for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
                size_t o = (size_t) y * (size_t) W + (size_t) x;
                uint32_t key = 0;
                for (int c = 0; c < 3; ++c) key = key << 8 | (uint32_t) (im.px[o * 3 + (size_t) c] * 5 / 255 * 51);
                auto it = index.find(key);
                if (it == index.end()) { it = index.insert({key, (int) colours.size()}).first; colours.push_back(key); }
                plane[0][o] = it->second;
                if (P.alpha) plane[1][o] = ((x / 37 + y / 53) % 5 == 0) ? 128 + ((x + y) & 63) : 255;
            }
            int K = (int) colours.size();
            if (H > 24) for (int x = 0; x < W; ++x) for (int r = 0; r < 4; ++r) {
                plane[0][(size_t) (8 + r) * (size_t) W + (size_t) x] = K + (x % 64);        // 4x4x4 cube
                plane[0][(size_t) (12 + r) * (size_t) W + (size_t) x] = K + 64 + (x % 125); // 5x5x5 cube
                plane[0][(size_t) (16 + r) * (size_t) W + (size_t) x] = -1 - (x % 143);     // hard-coded deltas
            }
            pal.w = K; pal.h = 3; pal.vshift = -1;
            pal.px.resize((size_t) K * 3);
            for (int c = 0; c < 3; ++c) for (int k = 0; k < K; ++k) pal.px[(size_t) c * (size_t) K + (size_t) k] = (int32_t) (colours[(size_t) k] >> (8 * (2 - c)) & 255);
        } else
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            size_t o = (size_t) y * (size_t) W + (size_t) x;
            int v[3];
            for (int c = 0; c < 3; ++c) {
                int s = im.px[o * 3 + (size_t) c];
                if (P.smooth) s = (s >> 2) << 2;
                v[c] = s;
            }
            forward_rct(v);
            for (int c = 0; c < 3; ++c) plane[(size_t) c][o] = v[c];
            if (P.alpha) plane[3][o] = ((x / 37 + y / 53) % 5 == 0) ? 128 + ((x + y) & 63) : 255;
        }
        int gsize = 1 << P.group_shift;
        int gcols = (W + gsize - 1) / gsize, grows = (H + gsize - 1) / gsize;
        int num_groups = gcols * grows;
        int lfcols = (W + gsize * 8 - 1) / (gsize * 8), lfrows = (H + gsize * 8 - 1) / (gsize * 8);
        int num_lfg = lfcols * lfrows;
        bool single = num_groups == 1;

        MATree tree = make_tree();
        stats.tree_nodes = (int32_t) tree.nodes.size();
        ModularTokenizer mt(tree);
        std::vector<TokStream> ts((size_t) num_groups);
        std::vector<int> gwidth((size_t) num_groups);
        std::vector<std::vector<Channel>> group_channels((size_t) (P.local_tree ? num_groups : 0));
        // groups are tokenised independently: on all cores for big frames (the 8192x8192 configuration has 1024)
        auto tokenize = [&](int g, ModularTokenizer &tk) {
            int gx = (g % gcols) * gsize, gy = (g / gcols) * gsize;
            int gw = std::min(W, gx + gsize) - gx, gh = std::min(H, gy + gsize) - gy;
            gwidth[(size_t) g] = gw;
            std::vector<Channel> ch((size_t) nch);
            for (int c = 0; c < nch; ++c) {
                ch[(size_t) c].w = gw; ch[(size_t) c].h = gh;
                ch[(size_t) c].px.resize((size_t) gw * (size_t) gh);
                for (int y = 0; y < gh; ++y) for (int x = 0; x < gw; ++x) {
                    ch[(size_t) c].px[(size_t) y * (size_t) gw + (size_t) x] = plane[(size_t) c][(size_t) (gy + y) * (size_t) W + (size_t) (gx + x)];
                }
            }
            int64_t sidx = single ? 0 : 1 + 3 * (int64_t) num_lfg + 17 + g;
            if (single && P.palette) ch.insert(ch.begin(), pal);
            tk.run(ch, sidx, ts[(size_t) g]);
            if (P.local_tree) group_channels[(size_t) g] = ch;
        };
        parallel_groups(num_groups, [&](int g) { ModularTokenizer tk(tree); tokenize(g, tk); });
        for (int g = 0; g < num_groups; ++g) stats.lf_symbols += (int64_t) ts[(size_t) g].size();
        EntropyOpts eo;
        eo.use_prefix = !P.use_ans;
        eo.log_alpha_size = 8;
        eo.cfg = {4, 1, 0};
        eo.max_clusters = P.max_clusters;
        eo.lz77 = P.lz77;
        if (eo.lz77) parallel_groups(num_groups, [&](int g) { lz77_rle(ts[(size_t) g], eo.min_length, gwidth[(size_t) g], 4); });
        TokStream global_ts; // multi-group frames: the palette is coded with the global image in LfGlobal
        if (P.palette && !single) {
            std::vector<Channel> gch{pal};
            mt.run(gch, 0, global_ts);
            if (eo.lz77) lz77_rle(global_ts, eo.min_length, pal.w, 4);
        }
        CodeSpec spec;
        {
            std::vector<const TokStream *> all;
            for (auto &s : ts) all.push_back(&s);
            all.push_back(&global_ts);
            spec.build(tree.num_leaves, eo, all);
        }
        stats.coef_clusters = spec.nclusters;

        // ---- LfGlobal
        BitWriter lfglobal;
        lfglobal.bit(1); // LF dequant defaults (parsed even for modular frames)
        const bool have_global_tree = P.local_tree != 2;
        auto is_local = [&](int g) { return P.local_tree == 2 || (P.local_tree == 1 && (single || (g & 1))); };
        lfglobal.bit(have_global_tree); // global tree present
        EntropyOpts to;
        to.use_prefix = !P.use_ans;
        to.log_alpha_size = 8;
        to.cfg = {4, 1, 0};
        to.max_clusters = 6;
        if (have_global_tree) {
            write_tree(lfglobal, tree, to);
            spec.write(lfglobal);
        }
        ModularHeaderOpts gh;
        if (P.palette) gh.palettes.push_back({0, 3, pal.w, P.pal_deltas, P.pal_pred});
        else if (P.rct_type >= 0) gh.rcts.push_back({0, P.rct_type});
        gh.use_global_tree = !(single ? is_local(0) : !have_global_tree);
        write_modular_header_prefix(lfglobal, gh);
        if (!gh.use_global_tree) { // the same tree, but stored with the sub-bitstream (j40.h:3827-3835)
            write_tree(lfglobal, tree, to);
            spec.write(lfglobal);
        }
        if (single) {
            spec.encode(lfglobal, ts[0]);
        } else {
            spec.encode(lfglobal, global_ts); // (an empty ANS stream still carries its final state)
        }

        BitWriter out;
        out.put(0xff, 8); out.put(0x0a, 8);
        VarDCTEncoder::write_size_header(out, W, H);
        // ImageMetadata
        out.bit(0);        // !all_default
        out.bit(0);        // extra_fields
        out.bit(0);        // integer samples
        out.u32(8, 8, 0, 10, 0, 12, 0, 1, 6);
        out.bit(1);        // modular_16bit_buffers
        out.u32(P.alpha ? 1 : 0, 0, 0, 1, 0, 2, 4, 1, 12);
        if (P.alpha) out.bit(1); // d_alpha
        out.bit(0);        // xyb_encoded
        out.bit(1);        // ColourEncoding all_default
        out.u64(0);        // extensions
        out.bit(1);        // default_m
        out.pad();
        // FrameHeader
        out.bit(0);
        out.put(0, 2);     // regular
        out.bit(1);        // modular
        out.u64(0);        // flags
        out.bit(0);        // do_ycbcr
        out.put(0, 2);     // log_upsampling
        if (P.alpha) out.put(0, 2); // ec upsampling
        out.put((uint64_t) (P.group_shift - 7), 2);
        out.u32(1, 1, 0, 2, 0, 3, 0, 4, 3); // passes
        out.bit(0);        // have_crop
        out.u32(0, 0, 0, 1, 0, 2, 0, 3, 2); // blend mode (colour)
        if (P.alpha) out.u32(0, 0, 0, 1, 0, 2, 0, 3, 2); // blend mode (alpha)
        out.bit(1);        // is_last
        out.u32(0, 0, 0, 0, 4, 16, 5, 48, 10); // name
        out.bit(0);        // restoration all_default = 0 (SURVEY B-1)
        out.bit(0);        // gab off
        out.put(0, 2);     // epf_iters 0
        out.u64(0);        // restoration extensions
        out.u64(0);        // frame extensions
        // TOC
        if (single) {
            lfglobal.pad();
            out.bit(0); out.pad();
            out.u32((uint32_t) lfglobal.bytes.size(), 0, 10, 1024, 14, 17408, 22, 4211712, 30);
            out.pad();
            out.append_bytes(lfglobal.bytes);
            stats.sections = 1;
        } else {
            std::vector<BitWriter> secs((size_t) (2 + num_lfg + num_groups));
            secs[0] = lfglobal;
            parallel_groups(num_groups, [&](int g) {
                BitWriter &bw = secs[(size_t) (2 + num_lfg + g)];
                ModularHeaderOpts mh;
                if (is_local(g)) {
                    // a different tree than the global one (single gradient leaf for every fourth group), with a
                    // code spec built from this group's symbols alone
                    MATree lt = tree;
                    if ((g & 3) == 3) { lt = MATree(); lt.flatten(MATree::Leaf(5)); }
                    ModularTokenizer lmt(lt);
                    TokStream lts;
                    lmt.run(group_channels[(size_t) g], 1 + 3 * (int64_t) num_lfg + 17 + g, lts);
                    if (eo.lz77) lz77_rle(lts, eo.min_length, gwidth[(size_t) g], 4);
                    CodeSpec ls;
                    std::vector<const TokStream *> one{&lts};
                    ls.build(lt.num_leaves, eo, one);
                    mh.use_global_tree = false;
                    write_modular_header_prefix(bw, mh);
                    write_tree(bw, lt, to);
                    ls.write(bw);
                    ls.encode(bw, lts);
                    return;
                }
                write_modular_header_prefix(bw, mh);
                spec.encode(bw, ts[(size_t) g]);
            });
            out.bit(0); out.pad();
            for (auto &s : secs) { s.pad(); out.u32((uint32_t) s.bytes.size(), 0, 10, 1024, 14, 17408, 22, 4211712, 30); }
            out.pad();
            for (auto &s : secs) out.append_bytes(s.bytes);
            stats.sections = (int64_t) secs.size();
        }
        std::vector<uint8_t> code = out.bytes;
        if (P.container) code = VarDCTEncoder::wrap_container(code, false);
        stats.bytes = (int64_t) code.size();
        return code;
    }

private:
    template <class F>
    static void parallel_groups(int n, F fn) {
        unsigned th = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 32u);
        if (n < 64 || th <= 1) { for (int g = 0; g < n; ++g) fn(g); return; }
        std::atomic<int> next(0);
        std::vector<std::thread> pool;
        std::exception_ptr err;
        std::mutex mu;
        auto work = [&] { try { for (int g; (g = next.fetch_add(1)) < n;) fn(g); } catch (...) { std::lock_guard<std::mutex> l(mu); err = std::current_exception(); } };
        for (unsigned t = 1; t < th; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
        if (err) std::rethrow_exception(err);
    }
    static int fl_avg(int a, int b) { return (a + b) >> 1; }

    void forward_rct(int v[3]) const {
        if (P.rct_type < 0) return;
        static const uint8_t PERM[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
        // the decoder computes c' from c and then stores c'[i] into channel PERM[type/7][i];
        // so the transformed triple we need is defined by: inverse(t)[i] == v[PERM[i]]
        int perm = P.rct_type / 7, op = P.rct_type % 7;
        int a = v[PERM[perm][0]], b = v[PERM[perm][1]], c = v[PERM[perm][2]]; // desired outputs of the inverse
        int t0, t1, t2;
        switch (op) {
        case 0: t0 = a; t1 = b; t2 = c; break;
        case 1: t0 = a; t1 = b; t2 = c - a; break;
        case 2: t0 = a; t1 = b; t2 = c; /* out2 = c1 + c0 overwrites c2 */ t2 = 0; v[PERM[perm][2]] = a + b; break;
        case 3: t0 = a; t1 = b - a; t2 = c - a; break;
        case 4: t0 = a; t2 = c; t1 = b - fl_avg(a, c); break;
        case 5: t0 = a; t2 = c - a; t1 = b - a - (t2 >> 1); break;
        case 6: { // YCgCo: a = R-like, b = G-like, c = B-like
            int co = a - c;
            int tmp = c + (co >> 1);
            int cg = b - tmp;
            int yy = tmp + (cg >> 1);
            t0 = yy; t1 = co; t2 = cg;
            break;
        }
        default: throw GenError("bad rct");
        }
        v[0] = t0; v[1] = t1; v[2] = t2;
    }

    MATree make_tree() const {
        typedef MATree M;
        MATree t;
        if (P.tree_preset == 0) { t.flatten(M::Leaf(5)); return t; }
        if (P.tree_preset == 1) {
            // contexts by channel and by local gradient magnitude; gradient predictor everywhere
            auto sub = [&]() {
                return M::Branch(10, 8, M::Leaf(5), M::Branch(10, -9, M::Branch(10, 1, M::Leaf(5), M::Branch(10, -2, M::Leaf(5), M::Leaf(5))), M::Leaf(5)));
            };
            t.flatten(M::Branch(0, 0, M::Branch(0, 2, M::Leaf(1), sub()), sub()));
            return t;
        }
        auto subw = [&]() {
            return M::Branch(15, 12, M::Leaf(6), M::Branch(15, -13, M::Branch(15, 2, M::Leaf(6), M::Branch(15, -3, M::Leaf(6), M::Leaf(6))), M::Leaf(6)));
        };
        t.flatten(M::Branch(0, 0, M::Branch(16, 0, subw(), M::Leaf(5)), subw()));
        return t;
    }
};

} // namespace jxlgen
