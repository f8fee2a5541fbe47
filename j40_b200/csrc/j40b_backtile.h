// j40-b200: tile-parallel back half of the VarDCT path (dequantisation, chroma-from-luma, LLF insertion,
// inverse transforms, XYB -> sRGB, RGBA8 store) for varblocks up to 64x64 that lie inside one 64x64-pixel
// tile -- the overwhelmingly common case. One thread block owns one tile: the coefficients of all its
// varblocks live in shared memory (3 x 4096 floats), every 1-D inverse DCT runs in the registers of one
// thread, and pixels leave as 256-byte row segments. Larger or tile-straddling varblocks are left to the
// generic path (varblock_to_pixels in j40b_vardct.h).
//
// The arithmetic is the reference's, operation for operation (see j40b_vardct.h for the citations); what
// changes is who computes what and where the numbers sit:
//   * j40__inverse_dct2d runs IDCT(columns) -> transpose -> IDCT(rows) over whole arrays; here each 1-D
//     transform is an independent unit working in place on its own address set, so no transposes and no
//     barriers inside a pass are needed:
//       square/tall blocks keep the stored [u][v] layout: pass A transforms along u (stride R), pass B along
//       v (contiguous), leaving samples as [x][y];
//       wide blocks keep [v][u]: pass A transforms along u (contiguous), pass B along v (stride C),
//       leaving samples as [y][x].
#pragma once
#include "j40b_vardct.h"

namespace j40b {

// 1-D inverse DCT of N points in place, the recursion of j40__inverse_dct_core (j40.h:5802-5841) unrolled
template <int N> struct Idct1D {
    J40B_HD static J40B_INLINE void run(float *v) {
        float a[N / 2], b[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) a[i] = v[2 * i];
        b[0] = J40B_FMUL(J40B_SQRT2, v[1]);
#pragma unroll
        for (int i = 1; i < N / 2; ++i) b[i] = J40B_FADD(v[2 * i - 1], v[2 * i + 1]);
        Idct1D<N / 2>::run(a);
        Idct1D<N / 2>::run(b);
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            float t = J40B_FMUL(b[i], J40B_HALF_SECANT(N / 2 + i));
            v[i] = J40B_FADD(a[i], t);
            v[N - 1 - i] = J40B_FSUB(a[i], t);
        }
    }
};
template <> struct Idct1D<2> {
    J40B_HD static J40B_INLINE void run(float *v) {
        float x = v[0], y = v[1];
        v[0] = J40B_FADD(x, y);
        v[1] = J40B_FSUB(x, y);
    }
};
template <> struct Idct1D<1> { J40B_HD static J40B_INLINE void run(float *) {} };

// Shared-memory layout of one varblock's coefficients: the stored index i = a * M + b (a = major, b = minor,
// M = 1 << mlog = the longer side, so a < M) lives at a * M + (b ^ a). Both IDCT passes then touch 32
// different banks from the 32 lanes of a warp (lanes differ in the fixed index of a 1-D transform), instead of
// the 4 banks of the plain layout when each lane walks 8 consecutive floats.
J40B_HD J40B_INLINE int tile_swz(int i, int mlog) {
    const int a = i >> mlog;
    return i ^ (a & ((1 << mlog) - 1));
}

// 1-D transform number f of a pass over a block: ALONG_MINOR walks b with a = f fixed, else walks a with b = f
template <int N, bool ALONG_MINOR>
J40B_HD J40B_INLINE void idct_swz(float *blk, int f, int mlog) {
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = ALONG_MINOR ? blk[(f << mlog) + (i ^ f)] : blk[(i << mlog) + (f ^ i)];
    Idct1D<N>::run(v);
#pragma unroll
    for (int i = 0; i < N; ++i) { if (ALONG_MINOR) blk[(f << mlog) + (i ^ f)] = v[i]; else blk[(i << mlog) + (f ^ i)] = v[i]; }
}

template <bool ALONG_MINOR>
J40B_HD J40B_INLINE void idct_swz_dispatch(float *blk, int f, int mlog, int log_n) {
    switch (log_n) {
    case 3: idct_swz<8, ALONG_MINOR>(blk, f, mlog); break;
    case 4: idct_swz<16, ALONG_MINOR>(blk, f, mlog); break;
    case 5: idct_swz<32, ALONG_MINOR>(blk, f, mlog); break;
    case 6: idct_swz<64, ALONG_MINOR>(blk, f, mlog); break;
    }
}

// floats per channel buffer: 4096 coefficients + the per-varblock skew (8 floats per varblock and 8 more per
// started row of eight, see chunk_off) that spreads neighbouring varblocks over the banks
enum { TILE_CH = 4672 };

struct TileVb {
    int32_t voff;        // varblock index inside the LF group
    int32_t coeffoff;
    uint16_t chunk_off;  // offset (floats) of this varblock's coefficients in each channel's buffer
    uint16_t log_off;    // the same without the bank skew: position in the tile's raster order of 64-float chunks
    uint8_t dctsel, log_rows, log_cols, param_idx;
    uint8_t cx, cy;      // top-left cell inside the tile
    uint8_t special;
    uint8_t mlog;        // log2 of the swizzled layout's row length (longer side); 0 = plain layout (special 8x8)
    float m[3];          // dequantisation multipliers mult[c] of j40__dequant_hf
    float kx_hf, kb_hf;
};

struct TileShared {
    int32_t nvb;
    TileVb vb[64];
    int32_t cell_voff[64];  // per cell: varblock index if the cell is a top-left handled here, else -1
    uint8_t cell_size64[64];
    uint8_t cover[64];      // tile cell -> index into vb[], 0xff = not handled here (generic path / outside)
    uint8_t chunk_vb[64];   // 64-float chunk -> index into vb[]
    float thr[255];
    uint8_t lut[SRGB_LUT_BYTES];
};

// tile (tx, ty) of group w.grp; coef = 3 * TILE_CH floats of shared memory
template <class Sync>
J40B_HD inline void back_tile_body(const BackWork &w, int tx, int ty, float *coef, TileShared &ts, int tid, int nth, Sync sync) {
    if (*w.lf_err || *w.hf_err) return;
    const DFrame &f = *w.f;
    const DLfGroup &g = *w.g;
    const DGroup &grp = *w.grp;
    const int gw8 = ceil_div(grp.gw, 8), gh8 = ceil_div(grp.gh, 8);
    if (tx * 8 >= gw8 || ty * 8 >= gh8) return;
    const int n8 = g.width8 * g.height8;
    float *coefx = coef, *coefy = coef + TILE_CH, *coefb = coef + 2 * TILE_CH;

    // ---- 0. which varblocks start in this tile (one cell per thread), thresholds, zeroed coefficients
    for (int c = tid; c < 64; c += nth) {
        int cx = c & 7, cy = c >> 3;
        int x8 = tx * 8 + cx, y8 = ty * 8 + cy;
        int32_t voff = -1;
        int size64 = 0;
        if (x8 < gw8 && y8 < gh8) {
            int32_t b = g.blocks[(y8 + grp.gy8) * g.width8 + x8 + grp.gx8];
            if ((b >> 20) >= 2 && !(g.varblocks[b & 0xfffff].pad & 1)) {
                voff = b & 0xfffff;
                DctSelectInfo d = dct_select_info((b >> 20) - 2);
                size64 = 1 << (d.log_rows + d.log_columns - 6);
            }
        }
        ts.cell_voff[c] = voff;
        ts.cell_size64[c] = (uint8_t) size64;
        ts.cover[c] = 0xff;
    }
    for (int i = tid; i < 255; i += nth) ts.thr[i] = f.srgb_thr[i];
    for (int i = tid; i < SRGB_LUT_BYTES / 4; i += nth) ((uint32_t *) ts.lut)[i] = ((const uint32_t *) f.srgb_lut)[i];
    for (int i = tid; i < 3 * TILE_CH; i += nth) coef[i] = 0.0f;
    sync();
    // ---- 0b. compact them in raster order (each top-left cell computes its own rank and chunk offset)
    {
        const float gs = J40B_FDIV(65536.0f, (float) f.global_scale);
        for (int c = tid; c < 64; c += nth) {
            if (c == 63) {
                int n = 0;
                for (int k = 0; k < 64; ++k) n += ts.cell_voff[k] >= 0;
                ts.nvb = n;
            }
            const int32_t voff = ts.cell_voff[c];
            if (voff < 0) continue;
            int rank = 0, off64 = 0;
            for (int k = 0; k < c; ++k) { rank += ts.cell_voff[k] >= 0; off64 += ts.cell_size64[k]; }
            const DVarblock vb = g.varblocks[voff];
            DctSelectInfo d = dct_select_info(vb.dctsel);
            TileVb &t = ts.vb[rank];
            t.voff = voff;
            t.coeffoff = vb.coeffoff;
            t.log_off = (uint16_t) (off64 * 64);
            t.chunk_off = (uint16_t) (off64 * 64 + 8 * rank + 8 * (rank >> 3));
            t.dctsel = vb.dctsel; t.log_rows = (uint8_t) d.log_rows; t.log_cols = (uint8_t) d.log_columns; t.param_idx = (uint8_t) d.param_idx;
            t.cx = (uint8_t) (c & 7); t.cy = (uint8_t) (c >> 3);
            t.special = is_special_8x8(vb.dctsel) ? 1 : 0;
            t.mlog = t.special ? 0 : (uint8_t) (d.log_rows > d.log_columns ? d.log_rows : d.log_columns);
            t.m[1] = J40B_FMUL(gs, vb.hfmul_inv);
            t.m[0] = J40B_FMUL(t.m[1], f.x_qm_mult);
            t.m[2] = J40B_FMUL(t.m[1], f.b_qm_mult);
            t.kx_hf = J40B_FADD(f.base_corr_x, J40B_FMUL(f.inv_colour_factor, (float) g.xfromy[(vb.y8 / 8) * g.width64 + (vb.x8 / 8)]));
            t.kb_hf = J40B_FADD(f.base_corr_b, J40B_FMUL(f.inv_colour_factor, (float) g.bfromy[(vb.y8 / 8) * g.width64 + (vb.x8 / 8)]));
            for (int k = 0; k < ts.cell_size64[c]; ++k) ts.chunk_vb[off64 + k] = (uint8_t) rank;
            for (int i = 0; i < (1 << (d.log_rows - 3)); ++i) for (int j = 0; j < (1 << (d.log_columns - 3)); ++j) {
                ts.cover[((c >> 3) + i) * 8 + (c & 7) + j] = (uint8_t) rank;
            }
        }
    }
    sync();
    const int nvb = ts.nvb;
    if (nvb == 0) return;

    // ---- 1. scatter + dequantise the decoded coefficients (one warp-sized stripe of threads per
    // varblock-channel; j40.h:7078-7094). Only non-zero coefficients have tokens (one per position: a single
    // pass), and a zero coefficient dequantises to +0 whatever its weight, so the dense loop of the reference
    // reduces to the token list: q = |c| <= 1 ? c * bias : c - bias_num / c; q *= mult / weight.
    {
        const float qbn = f.quant_bias_num;
        const int lanes = nth < 32 ? nth : 32, groups = nth / lanes;
        const int lane = tid % lanes, grp_id = tid / lanes;
        for (int pair = grp_id; pair < nvb * 3; pair += groups) {
            int v = pair / 3, c = pair - v * 3;
            const TileVb &t = ts.vb[v];
            uint32_t first = g.vb_tok[((size_t) c * n8 + t.voff) * 2 + 0], cnt = g.vb_tok[((size_t) c * n8 + t.voff) * 2 + 1];
            float *dst = coef + c * TILE_CH + t.chunk_off;
            const int mlog = t.mlog;
            const float qb = f.quant_bias[c], m = t.m[c];
            const float *dq = f.dq[t.param_idx] + c;
            for (uint32_t k = (uint32_t) lane; k < cnt; k += (uint32_t) lanes) {
                DToken tk = w.tokens[first + k];
                float q = (float) tk.val;
                q = (-1.0f <= q && q <= 1.0f) ? J40B_FMUL(q, qb) : J40B_FSUB(q, J40B_FDIV(qbn, q));
                q = J40B_FMUL(q, J40B_FDIV(m, dq[(size_t) tk.pos * 3]));
                dst[tile_swz(tk.pos, mlog)] = q;
            }
        }
    }
    sync();
    // ---- 2. chroma from luma (j40.h:7155-7175): x += y * kx, b += y * kb; a no-op wherever y is zero, so
    // again only the Y tokens are visited
    {
        const int lanes = nth < 32 ? nth : 32, groups = nth / lanes;
        const int lane = tid % lanes, grp_id = tid / lanes;
        for (int v = grp_id; v < nvb; v += groups) {
            const TileVb &t = ts.vb[v];
            uint32_t first = g.vb_tok[((size_t) 1 * n8 + t.voff) * 2 + 0], cnt = g.vb_tok[((size_t) 1 * n8 + t.voff) * 2 + 1];
            const int mlog = t.mlog;
            for (uint32_t k = (uint32_t) lane; k < cnt; k += (uint32_t) lanes) {
                const int p = t.chunk_off + tile_swz(w.tokens[first + k].pos, mlog);
                const float vy = coefy[p];
                coefx[p] = J40B_FADD(coefx[p], J40B_FMUL(vy, t.kx_hf));
                coefb[p] = J40B_FADD(coefb[p], J40B_FMUL(vy, t.kb_hf));
            }
        }
    }
    // (no barrier: the LLF corner below is disjoint from every token position -- the coefficient scan starts
    // behind the LLF coefficients, j40.h:6978, and custom orders leave that prefix in place)
    // ---- 3. LLF corner from the LF image (j40.h:7158-7172)
    {
        const float kx_lf = f.kx_lf, kb_lf = f.kb_lf;
        for (int slot = tid; slot < nvb * 64; slot += nth) {
            const TileVb &t = ts.vb[slot >> 6];
            const int e = slot & 63;
            const int lmin = t.log_rows < t.log_cols ? t.log_rows : t.log_cols, lmax = t.log_rows < t.log_cols ? t.log_cols : t.log_rows;
            const int vh8 = 1 << (lmin - 3), vw8 = 1 << (lmax - 3);
            if (e >= vh8 * vw8) continue;
            int y = e / vw8, x = e - y * vw8;
            int p = t.chunk_off + tile_swz(y * vw8 * 8 + x, t.mlog);
            float l0 = g.llf[(size_t) 0 * n8 + (t.coeffoff >> 6) + e];
            float l1 = g.llf[(size_t) 1 * n8 + (t.coeffoff >> 6) + e];
            float l2 = g.llf[(size_t) 2 * n8 + (t.coeffoff >> 6) + e];
            coefx[p] = J40B_FADD(l0, J40B_FMUL(l1, kx_lf));
            coefy[p] = l1;
            coefb[p] = J40B_FADD(l2, J40B_FMUL(l1, kb_lf));
        }
    }
    sync();
    // ---- 4. pass A: 1-D inverse DCTs along the horizontal frequency u; special 8x8 transforms whole
    // slot = (channel, tile pixel row, tile cell column); active where the cell starts a varblock's row of cells
    for (int slot = tid; slot < 3 * 64 * 8; slot += nth) {
        const int c = slot / 512, r64 = slot & 63, cx = (slot >> 6) & 7;
        const int cell = (r64 >> 3) * 8 + cx;
        const uint8_t vi = ts.cover[cell];
        if (vi == 0xff) continue;
        const TileVb &t = ts.vb[vi];
        if (t.cx != cx) continue; // not the leftmost cell of this varblock
        const int r = r64 - t.cy * 8; // row inside the varblock (vertical frequency index v at this stage)
        float *blk = coef + c * TILE_CH + t.chunk_off;
        if (t.special) {
            if (r == 0) inverse_special(t.dctsel, blk);
            continue;
        }
        if (t.log_cols > t.log_rows) idct_swz_dispatch<true>(blk, r, t.mlog, t.log_cols);    // [v][u]: row v = r, along u
        else idct_swz_dispatch<false>(blk, r, t.mlog, t.log_cols);                          // [u][v]: column v = r, along u
    }
    sync();
    // ---- 5. pass B: 1-D inverse DCTs along the vertical frequency v
    // slot = (channel, tile pixel column, tile cell row); active where the cell starts a varblock's column of cells
    for (int slot = tid; slot < 3 * 64 * 8; slot += nth) {
        const int c = slot / 512, x64 = slot & 63, cy = (slot >> 6) & 7;
        const int cell = cy * 8 + (x64 >> 3);
        const uint8_t vi = ts.cover[cell];
        if (vi == 0xff) continue;
        const TileVb &t = ts.vb[vi];
        if (t.cy != cy || t.special) continue;
        const int x = x64 - t.cx * 8;
        float *blk = coef + c * TILE_CH + t.chunk_off;
        if (t.log_cols > t.log_rows) idct_swz_dispatch<false>(blk, x, t.mlog, t.log_rows);   // [v][x] -> [y][x]: column x, along v
        else idct_swz_dispatch<true>(blk, x, t.mlog, t.log_rows);                           // [x][v] -> [x][y]: row x, along v
    }
    sync();
    // ---- 6. XYB -> sRGB -> RGBA8 (j40.h:7208-7237, 7941-7952), one thread per pixel, row-major
    const int gx0 = g.left + (grp.gx8 + tx * 8) * 8, gy0 = g.top + (grp.gy8 + ty * 8) * 8;
    const int fw = f.width, fh = f.height;
    const float cb0 = f.cbrt_opsin_bias[0], cb1 = f.cbrt_opsin_bias[1], cb2 = f.cbrt_opsin_bias[2];
    const float ob0 = f.opsin_bias[0], ob1 = f.opsin_bias[1], ob2 = f.opsin_bias[2], itscale = f.itscale;
    float om[9];
    for (int i = 0; i < 9; ++i) om[i] = f.opsin_inv_mat[i];
    uint8_t *rgba = w.rgba;
    const size_t rgba_stride = (size_t) w.rgba_stride;
    for (int pix = tid; pix < 4096; pix += nth) {
        const int py = pix >> 6, px = pix & 63;
        const int X = gx0 + px, Y = gy0 + py;
        if (X >= fw || Y >= fh) continue;
        const uint8_t vi = ts.cover[(py >> 3) * 8 + (px >> 3)];
        if (vi == 0xff) continue;
        const TileVb &t = ts.vb[vi];
        const int ly = py - t.cy * 8, lx = px - t.cx * 8;
        const int R = 1 << t.log_rows, C = 1 << t.log_cols;
        // special transforms and wide blocks end as [y][x]; square / tall DCT blocks as [x][y]
        const int idx = t.chunk_off + tile_swz((t.special || t.log_cols > t.log_rows) ? ly * C + lx : lx * R + ly, t.mlog);
        float sx = coefx[idx], sy = coefy[idx], sb = coefb[idx];
        float p0 = J40B_FSUB(J40B_FADD(sy, sx), cb0), p1 = J40B_FSUB(J40B_FSUB(sy, sx), cb1), p2 = J40B_FSUB(sb, cb2);
        float l0 = J40B_FMUL(J40B_FADD(J40B_FMUL(J40B_FMUL(p0, p0), p0), ob0), itscale);
        float l1 = J40B_FMUL(J40B_FADD(J40B_FMUL(J40B_FMUL(p1, p1), p1), ob1), itscale);
        float l2 = J40B_FMUL(J40B_FADD(J40B_FMUL(J40B_FMUL(p2, p2), p2), ob2), itscale);
        uint32_t out = 0xff000000u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = J40B_FADD(J40B_FADD(J40B_FMUL(l0, om[c * 3 + 0]), J40B_FMUL(l1, om[c * 3 + 1])), J40B_FMUL(l2, om[c * 3 + 2]));
            out |= (uint32_t) srgb_u8_lut(ts.thr, ts.lut, v) << (8 * c);
        }
        *(uint32_t *) (rgba + (size_t) Y * rgba_stride + (size_t) X * 4) = out;
    }
}

} // namespace j40b
