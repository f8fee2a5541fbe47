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

// per-phase cycle counters of the tile kernel (diagnostic builds only: make PHASE_CLOCKS=1)
#if defined(__CUDA_ARCH__) && defined(J40B_PHASE_CLOCKS) && defined(J40B_KERN_BACK_TU) // ts.phase: set by k_back_tile
#define J40B_PHASE(i) do { if (tid == 0) { long long now_ = clock64(); atomicAdd(&ts.phase[i], (unsigned long long) (now_ - phase_t0_)); phase_t0_ = now_; } } while (0)
#define J40B_PHASE_BEGIN long long phase_t0_ = clock64()
#else
#define J40B_PHASE(i) do {} while (0)
#define J40B_PHASE_BEGIN do {} while (0)
#endif

// Shared-memory layout of one varblock's coefficients: the stored index i = a * M + b (a = major, b = minor,
// M = 1 << mlog = the longer side, so a < M) lives at a * M + (b ^ a). Both IDCT passes then touch 32
// different banks from the 32 lanes of a warp (lanes differ in the fixed index of a 1-D transform), instead of
// the 4 banks of the plain layout when each lane walks 8 consecutive floats.
J40B_HD J40B_INLINE int tile_swz(int i, int mlog) {
    const int a = i >> mlog;
    return i ^ (a & ((1 << mlog) - 1));
}

// *p += v on a coefficient buffer that other threads of the block may be adding to as well
J40B_HD J40B_INLINE void tile_add(float *p, float v) {
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p = *p + v;
#endif
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
    uint8_t special, order_idx;
    uint8_t mlog;        // log2 of the swizzled layout's row length (longer side); 0 = plain layout (special 8x8)
    uint8_t pad_[3];
    float m[3];          // dequantisation multipliers mult[c] of j40__dequant_hf
    uint32_t first[3];   // first token of channel c (X, Y, B) in the image's token array
    uint16_t cnt[3];     // number of tokens of channel c
};

struct alignas(16) TilePx4 { uint32_t v[4]; };

struct TileShared {
    unsigned long long *phase; // diagnostic builds: global per-phase cycle counters
    int32_t nvb, nchunks;
    TileVb vb[64];
    int32_t cell_voff[64];  // per cell: varblock index if the cell is a top-left handled here, else -1
    uint32_t cell_first[64][3]; // per cell: first token / token count per channel of the varblock starting there
    uint16_t cell_cnt[64][3];
    uint16_t tstart[65];    // per compacted varblock: start of its tokens in the tile's flattened token list
    // 1-D transform tasks of the two IDCT passes, listed per transform type so that the lanes of a warp run the
    // same code: pstart[pass][type][v] = tasks of that type in varblocks before v, [64] = their total.
    // type = 2 * (log2 N - 3) + ALONG_MINOR (see idct_swz)
    uint16_t pstart[2][8][65];
    // ... and the inverse map: the types' task lists of a pass laid end to end in units of 8 tasks (a varblock contributes
    // a multiple of 24), tmap[pass][tbase[pass][type] + (k >> 3)] = index into vb[] of task k of that type
    uint8_t tmap[2][192];
    uint16_t tbase[2][8];
    int32_t tmap_used[2];
    // flattened token list of pass 0: tokmap[m] = index into vb[] of the varblock that holds token 64 m
    uint8_t tokmap[576];
    float kx_hf, kb_hf;     // chroma-from-luma factors of this tile (one 64x64 cell of the XFromY / BFromY maps)
    uint8_t cell_size64[64];
    uint8_t cover[64];      // tile cell -> index into vb[], 0xff = not handled here (generic path / outside)
    uint8_t chunk_vb[64];   // 64-float chunk -> index into vb[]
    // sRGB threshold table and start LUT (padded to multiples of 16 bytes: the persistent kernel stages them with bulk copies)
    alignas(16) float thr[256];
    alignas(16) uint8_t lut[(SRGB_LUT_BYTES + 15) / 16 * 16];
    int32_t tables_staged; // the kernel wrapper has filled thr / lut already (once per persistent block)
};

// number of 1-D transforms of `type` that varblock t contributes to pass 0 (along u) / pass 1 (along v)
J40B_HD J40B_INLINE int tile_task_count(const TileVb &t, int pass, int type) {
    if (t.special) return 0;
    const bool wide = t.log_cols > t.log_rows;
    const int log_n = pass == 0 ? t.log_cols : t.log_rows;
    const int minor = pass == 0 ? (wide ? 1 : 0) : (wide ? 0 : 1);
    if (2 * (log_n - 3) + minor != type) return 0;
    return 3 << (pass == 0 ? t.log_rows : t.log_cols);
}

// exclusive prefix sums of the task counts over the compacted varblocks, one (pass, type) pair per warp
J40B_HD inline void tile_type_scan(TileShared &ts, int nvb, int tid, int nth) {
#if defined(__CUDA_ARCH__)
    if (nth >= 32 && !(nth & 31)) {
        const int lane = tid & 31;
        for (int pair = tid >> 5; pair < 16; pair += nth >> 5) {
            const int pass = pair >> 3, type = pair & 7;
            int carry = 0;
            for (int base = 0; base < 64; base += 32) {
                const int v = base + lane;
                const int cnt = v < nvb ? tile_task_count(ts.vb[v], pass, type) : 0;
                int x = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
                ts.pstart[pass][type][v] = (uint16_t) (carry + x - cnt);
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) ts.pstart[pass][type][64] = (uint16_t) carry;
            if (carry == 0) continue;
            int tb = 0;
            if (lane == 0) {
                tb = atomicAdd(&ts.tmap_used[pass], carry >> 3);
                ts.tbase[pass][type] = (uint16_t) tb;
            }
            tb = __shfl_sync(0xffffffffu, tb, 0);
            __syncwarp();
            for (int v = lane; v < nvb; v += 32) {
                const int cnt = tile_task_count(ts.vb[v], pass, type) >> 3;
                uint8_t *m = ts.tmap[pass] + tb + (ts.pstart[pass][type][v] >> 3);
                for (int e = 0; e < cnt; ++e) m[e] = (uint8_t) v;
            }
        }
        return;
    }
#endif
    for (int pair = tid; pair < 16; pair += nth) {
        const int pass = pair >> 3, type = pair & 7;
        int sum = 0;
        for (int v = 0; v < 64; ++v) {
            ts.pstart[pass][type][v] = (uint16_t) sum;
            if (v < nvb) sum += tile_task_count(ts.vb[v], pass, type);
        }
        ts.pstart[pass][type][64] = (uint16_t) sum;
        if (sum == 0) continue;
#if defined(__CUDA_ARCH__)
        const int tb = atomicAdd(&ts.tmap_used[pass], sum >> 3);
#else
        const int tb = ts.tmap_used[pass];
        ts.tmap_used[pass] += sum >> 3;
#endif
        ts.tbase[pass][type] = (uint16_t) tb;
        for (int v = 0; v < nvb; ++v) {
            const int cnt = tile_task_count(ts.vb[v], pass, type) >> 3;
            uint8_t *m = ts.tmap[pass] + tb + (ts.pstart[pass][type][v] >> 3);
            for (int e = 0; e < cnt; ++e) m[e] = (uint8_t) v;
        }
    }
}

// all 1-D transforms of one type in one pass: task k -> (varblock, channel, transform number)
template <int LOGN, bool MINOR>
J40B_HD J40B_INLINE void tile_pass_type(float *coef, const TileShared &ts, int pass, int nvb, int tid, int nth) {
    const int type = 2 * (LOGN - 3) + (MINOR ? 1 : 0);
    const uint16_t *start = ts.pstart[pass][type];
    const int total = start[64];
    if (total == 0) return;
    const uint8_t *map = ts.tmap[pass] + ts.tbase[pass][type];
    for (int k = tid; k < total; k += nth) {
        const int lo = map[k >> 3];
        const TileVb &t = ts.vb[lo];
        const int j = k - (int) start[lo];
        const int lcnt = pass == 0 ? t.log_rows : t.log_cols;
        idct_swz<(1 << LOGN), MINOR>(coef + (j >> lcnt) * TILE_CH + t.chunk_off, j & ((1 << lcnt) - 1), t.mlog);
    }
}

J40B_HD J40B_INLINE void tile_pass(float *coef, const TileShared &ts, int pass, int nvb, int tid, int nth) {
    tile_pass_type<6, false>(coef, ts, pass, nvb, tid, nth);
    tile_pass_type<6, true>(coef, ts, pass, nvb, tid, nth);
    tile_pass_type<5, false>(coef, ts, pass, nvb, tid, nth);
    tile_pass_type<5, true>(coef, ts, pass, nvb, tid, nth);
    tile_pass_type<4, false>(coef, ts, pass, nvb, tid, nth);
    tile_pass_type<4, true>(coef, ts, pass, nvb, tid, nth);
    tile_pass_type<3, false>(coef, ts, pass, nvb, tid, nth);
    tile_pass_type<3, true>(coef, ts, pass, nvb, tid, nth);
}

// tile (tx, ty) of group w.grp; coef = 3 * TILE_CH floats of shared memory
template <class Sync>
J40B_HD inline void back_tile_body(const BackWork &w, int tx, int ty, float *coef, TileShared &ts, int tid, int nth, Sync sync) {
    if (*w.lf_err) return;
    const DFrame &f = *w.f;
    const int npass = f.num_passes;
    for (int p = 0; p < npass; ++p) if (w.hf_err[(size_t) p * w.hf_err_stride]) return;
    const DLfGroup &g = *w.g;
    const DGroup &grp = *w.grp;
    const int gw8 = ceil_div(grp.gw, 8), gh8 = ceil_div(grp.gh, 8);
    if (tx * 8 >= gw8 || ty * 8 >= gh8) return;
    const int n8 = g.width8 * g.height8;
    float *coefx = coef, *coefy = coef + TILE_CH, *coefb = coef + 2 * TILE_CH;
    J40B_PHASE_BEGIN;

    // ---- 0. which varblocks start in this tile (one cell per thread), thresholds, zeroed coefficients
    for (int c = tid; c < 64; c += nth) {
        int cx = c & 7, cy = c >> 3;
        int x8 = tx * 8 + cx, y8 = ty * 8 + cy;
        int32_t voff = -1;
        int size64 = 0;
        if (x8 < gw8 && y8 < gh8) {
            int32_t b = g.blocks[(y8 + grp.gy8) * g.width8 + x8 + grp.gx8];
            if ((b >> 20) >= 2) {
                // the varblock record and its three token ranges are fetched together (one round trip to L2)
                const int32_t vo = b & 0xfffff;
                const uint16_t pad = g.varblocks[vo].pad;
                const uint64_t r0 = *(const uint64_t *) (g.vb_tok + ((size_t) 0 * n8 + vo) * 2);
                const uint64_t r1 = *(const uint64_t *) (g.vb_tok + ((size_t) 1 * n8 + vo) * 2);
                const uint64_t r2 = *(const uint64_t *) (g.vb_tok + ((size_t) 2 * n8 + vo) * 2);
                if (!(pad & 1)) {
                    voff = vo;
                    DctSelectInfo d = dct_select_info((b >> 20) - 2);
                    size64 = 1 << (d.log_rows + d.log_columns - 6);
                    ts.cell_first[c][0] = (uint32_t) r0; ts.cell_first[c][1] = (uint32_t) r1; ts.cell_first[c][2] = (uint32_t) r2;
                    ts.cell_cnt[c][0] = (uint16_t) (r0 >> 32); ts.cell_cnt[c][1] = (uint16_t) (r1 >> 32); ts.cell_cnt[c][2] = (uint16_t) (r2 >> 32);
                }
            }
        }
        ts.cell_voff[c] = voff;
        ts.cell_size64[c] = (uint8_t) size64;
        ts.cover[c] = 0xff;
    }
    if (tid == 0) {
        // every varblock that starts in this tile lies in the same 64x64 cell of the colour-correlation maps
        const int m64 = ((grp.gy8 >> 3) + ty) * g.width64 + (grp.gx8 >> 3) + tx;
        ts.kx_hf = J40B_FADD(f.base_corr_x, J40B_FMUL(f.inv_colour_factor, (float) g.xfromy[m64]));
        ts.kb_hf = J40B_FADD(f.base_corr_b, J40B_FMUL(f.inv_colour_factor, (float) g.bfromy[m64]));
        ts.tmap_used[0] = ts.tmap_used[1] = 0;
    }
    if (!ts.tables_staged) {
        for (int i = tid; i < 256; i += nth) ts.thr[i] = f.srgb_thr[i];
        for (int i = tid; i < SRGB_LUT_BYTES / 4; i += nth) ((uint32_t *) ts.lut)[i] = ((const uint32_t *) f.srgb_lut)[i];
    }
    for (int i = tid; i < 3 * TILE_CH; i += nth) coef[i] = 0.0f;
    sync();
    J40B_PHASE(0);
    // ---- 0b. compact them in raster order (each top-left cell computes its own rank and chunk offset)
    {
        const float gs = J40B_FDIV(65536.0f, (float) f.global_scale);
        for (int c = tid; c < 64; c += nth) {
            if (c == 63) {
                int n = 0, tt = 0, nch = 0;
                for (int k = 0; k < 64; ++k) if (ts.cell_voff[k] >= 0) { ++n; nch += ts.cell_size64[k]; tt += ts.cell_cnt[k][0] + ts.cell_cnt[k][1] + ts.cell_cnt[k][2]; }
                ts.nvb = n;
                ts.nchunks = nch;
                ts.tstart[n] = (uint16_t) tt;
            }
            const int32_t voff = ts.cell_voff[c];
            if (voff < 0) continue;
            int rank = 0, off64 = 0, tstart = 0;
            for (int k = 0; k < c; ++k) if (ts.cell_voff[k] >= 0) {
                ++rank; off64 += ts.cell_size64[k];
                tstart += ts.cell_cnt[k][0] + ts.cell_cnt[k][1] + ts.cell_cnt[k][2];
            }
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
            t.order_idx = (uint8_t) d.order_idx;
            t.mlog = t.special ? 0 : (uint8_t) (d.log_rows > d.log_columns ? d.log_rows : d.log_columns);
            t.m[1] = J40B_FMUL(gs, vb.hfmul_inv);
            t.m[0] = J40B_FMUL(t.m[1], f.x_qm_mult);
            t.m[2] = J40B_FMUL(t.m[1], f.b_qm_mult);
            for (int k = 0; k < 3; ++k) { t.first[k] = ts.cell_first[c][k]; t.cnt[k] = ts.cell_cnt[c][k]; }
            ts.tstart[rank] = (uint16_t) tstart;
            {
                const int tend = tstart + t.cnt[0] + t.cnt[1] + t.cnt[2];
                for (int m = (tstart + 63) >> 6; (m << 6) < tend; ++m) ts.tokmap[m] = (uint8_t) rank;
            }
            for (int k = 0; k < ts.cell_size64[c]; ++k) ts.chunk_vb[off64 + k] = (uint8_t) rank;
            for (int i = 0; i < (1 << (d.log_rows - 3)); ++i) for (int j = 0; j < (1 << (d.log_columns - 3)); ++j) {
                ts.cover[((c >> 3) + i) * 8 + (c & 7) + j] = (uint8_t) rank;
            }
        }
    }
    sync();
    J40B_PHASE(1);
    const int nvb = ts.nvb;
    if (nvb == 0) return;
    tile_type_scan(ts, nvb, tid, nth); // read by the IDCT passes, behind the barrier that ends the scatter phase

    // ---- 1+2. scatter, dequantise (j40.h:7078-7094) and chroma from luma (j40.h:7155-7175), one thread per
    // token over the tile's flattened token list, so that every global load of the phase (token, then its
    // dequantisation weight) is in flight for 256 tokens at once instead of one varblock-channel per warp.
    // Only non-zero coefficients have tokens (one per position: a single pass), and a zero coefficient
    // dequantises to +0 whatever its weight, so the dense loop of the reference reduces to the token list:
    // q = |c| <= 1 ? c * bias : c - bias_num / c; q *= mult / weight. Chroma from luma, x += y * kx and
    // b += y * kb, is a no-op wherever y is zero: it is applied from the Y tokens. A position of the X (B)
    // buffer thus receives at most two addends on top of the initial +0, its own coefficient and the luma
    // term; float addition is commutative and 0 + a is exact, so adding them in either order (shared-memory
    // atomics, which keep denormals) gives the reference's x + y * kx bit for bit.
    for (int pass = 0; pass < npass; ++pass) {
        if (pass > 0) {
            // the token ranges of this pass (pass 0's were fetched in phase 0), then their prefix sums
            sync();
            for (int c = tid; c < 64; c += nth) {
                const int32_t vo = ts.cell_voff[c];
                if (vo < 0) continue;
                TileVb &t = ts.vb[ts.cover[c]];
                for (int k = 0; k < 3; ++k) {
                    const uint32_t *slot = g.vb_tok + (((size_t) pass * 3 + k) * n8 + vo) * 2;
                    t.first[k] = slot[0];
                    t.cnt[k] = (uint16_t) slot[1];
                }
            }
            sync();
            for (int v = tid; v <= nvb; v += nth) {
                int tt = 0;
                for (int k = 0; k < v; ++k) tt += ts.vb[k].cnt[0] + ts.vb[k].cnt[1] + ts.vb[k].cnt[2];
                ts.tstart[v] = (uint16_t) tt;
            }
            sync();
        }
        const float qbn = f.quant_bias_num;
        const float kx_hf = ts.kx_hf, kb_hf = ts.kb_hf;
        const int total = ts.tstart[nvb];
        const int nblk = (total + 63) >> 6;
        for (int k = tid; k < total; k += nth) {
            // last varblock whose first token is <= k; pass 0 has a map of every 64th token to narrow the search with
            int lo = 0, hi = nvb - 1;
            if (pass == 0) {
                const int m = k >> 6;
                lo = ts.tokmap[m];
                if (m + 1 < nblk) hi = ts.tokmap[m + 1];
            }
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((int) ts.tstart[mid] <= k) lo = mid; else hi = mid - 1; }
            const TileVb &t = ts.vb[lo];
            int j = k - (int) ts.tstart[lo];
            // tokens of a varblock were written in decoding order: Y, X, B
            int c = 1;
            if (j >= (int) t.cnt[1]) { j -= t.cnt[1]; c = 0; if (j >= (int) t.cnt[0]) { j -= t.cnt[0]; c = 2; } }
            const DToken *tk = w.tokens + t.first[c] + (uint32_t) j;
            if (tk->is_ext()) continue; // second or third word of a wide value
            const int32_t pos = f.order[pass][t.order_idx][c][tk->pos()]; // the token carries the scan index
            const int p = t.chunk_off + tile_swz(pos, t.mlog);
            float q = (float) token_value(tk);
            if (npass > 1) {
                // several passes: their values add up first (j40.h:6989; integers, exact in any order); dequantised below
                tile_add(c == 0 ? &coefx[p] : c == 1 ? &coefy[p] : &coefb[p], q);
                continue;
            }
            q = (-1.0f <= q && q <= 1.0f) ? J40B_FMUL(q, f.quant_bias[c]) : J40B_FSUB(q, J40B_FDIV(qbn, q));
            q = J40B_FMUL(q, J40B_FDIV(t.m[c], f.dq[t.param_idx][(size_t) pos * 3 + c]));
            if (c == 1) {
                coefy[p] = q;
                tile_add(&coefx[p], J40B_FMUL(q, kx_hf));
                tile_add(&coefb[p], J40B_FMUL(q, kb_hf));
            } else {
                tile_add(c == 0 ? &coefx[p] : &coefb[p], q);
            }
        }
    }
    if (npass > 1) {
        // dense form of j40__dequant_hf and the chroma-from-luma step over the tile's coefficients (the LLF corner is
        // still zero here and stays zero: 0 * bias * weight)
        sync();
        const float qbn = f.quant_bias_num;
        const float kx_hf = ts.kx_hf, kb_hf = ts.kb_hf;
        for (int e = tid; e < ts.nchunks * 64; e += nth) {
            const TileVb &t = ts.vb[ts.chunk_vb[e >> 6]];
            const int i = e - (int) t.log_off;
            const int p = t.chunk_off + tile_swz(i, t.mlog);
            const float *dq = f.dq[t.param_idx] + (size_t) i * 3;
            float q[3] = {coefx[p], coefy[p], coefb[p]};
            for (int c = 0; c < 3; ++c) {
                q[c] = (-1.0f <= q[c] && q[c] <= 1.0f) ? J40B_FMUL(q[c], f.quant_bias[c]) : J40B_FSUB(q[c], J40B_FDIV(qbn, q[c]));
                q[c] = J40B_FMUL(q[c], J40B_FDIV(t.m[c], dq[c]));
            }
            coefy[p] = q[1];
            coefx[p] = J40B_FADD(q[0], J40B_FMUL(q[1], kx_hf));
            coefb[p] = J40B_FADD(q[2], J40B_FMUL(q[1], kb_hf));
        }
        sync();
    }
    J40B_PHASE(2);
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
    J40B_PHASE(3);
    // ---- 4. pass A: 1-D inverse DCTs along the horizontal frequency u (square / tall blocks keep [u][v]:
    // column v = r, along u; wide blocks keep [v][u]: row v = r, along u); special 8x8 transforms whole,
    // one thread per block and channel, issued first because they are the longest tasks of the pass
    for (int sp = tid; sp < nvb * 3; sp += nth) {
        const TileVb &t = ts.vb[sp / 3];
        if (t.special) inverse_special(t.dctsel, coef + (sp % 3) * TILE_CH + t.chunk_off);
    }
    // ---- 5. pass B: 1-D inverse DCTs along the vertical frequency v ([v][x] -> [y][x]: column x, along v;
    // [x][v] -> [x][y]: row x, along v). One copy of the eight typed transform loops serves both passes (not unrolled:
    // the kernel's code is what its warps wait for most once the instruction count is down -- ncu: `no_instruction`
    // 3.2 stall cycles per issue with a 290 KB kernel, and every tile walks all of its phases)
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int pass = 0; pass < 2; ++pass) {
        tile_pass(coef, ts, pass, nvb, tid, nth);
        sync();
        J40B_PHASE(4 + pass);
    }
    // ---- 6. XYB -> sRGB -> RGBA8 (j40.h:7208-7237, 7941-7952), one thread per pixel, row-major
    const int gx0 = g.left + (grp.gx8 + tx * 8) * 8, gy0 = g.top + (grp.gy8 + ty * 8) * 8;
    const int fw = f.width, fh = f.height;
    const float cb0 = f.cbrt_opsin_bias[0], cb1 = f.cbrt_opsin_bias[1], cb2 = f.cbrt_opsin_bias[2];
    const float ob0 = f.opsin_bias[0], ob1 = f.opsin_bias[1], ob2 = f.opsin_bias[2], itscale = f.itscale;
    float om[9];
    for (int i = 0; i < 9; ++i) om[i] = f.opsin_inv_mat[i];
    const float wrap_hi = f.srgb_wrap_hi;
    uint8_t *rgba = w.rgba;
    const size_t rgba_stride = (size_t) w.rgba_stride;
    // four horizontally adjacent pixels per thread (they share a cell, hence a varblock): one 16-byte store
    for (int q4 = tid; q4 < 1024; q4 += nth) {
        const int py = q4 >> 4, px = (q4 & 15) * 4;
        const int X = gx0 + px, Y = gy0 + py;
        if (X >= fw || Y >= fh) continue;
        const uint8_t vi = ts.cover[(py >> 3) * 8 + (px >> 3)];
        if (vi == 0xff) continue;
        const TileVb &t = ts.vb[vi];
        const int ly = py - t.cy * 8, lx = px - t.cx * 8;
        // special transforms and wide blocks end as [y][x]; square / tall DCT blocks as [x][y]
        const bool rows = t.special || t.log_cols > t.log_rows;
        const int i0 = rows ? (ly << t.log_cols) + lx : (lx << t.log_rows) + ly, di = rows ? 1 : 1 << t.log_rows;
        const int chunk_off = t.chunk_off, mlog = t.mlog;
        TilePx4 o;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int idx = chunk_off + tile_swz(i0 + e * di, mlog);
            float sx = coefx[idx], sy = coefy[idx], sb = coefb[idx];
            float p0 = J40B_FSUB(J40B_FADD(sy, sx), cb0), p1 = J40B_FSUB(J40B_FSUB(sy, sx), cb1), p2 = J40B_FSUB(sb, cb2);
            float l0 = J40B_FMUL(J40B_FADD(J40B_FMUL(J40B_FMUL(p0, p0), p0), ob0), itscale);
            float l1 = J40B_FMUL(J40B_FADD(J40B_FMUL(J40B_FMUL(p1, p1), p1), ob1), itscale);
            float l2 = J40B_FMUL(J40B_FADD(J40B_FMUL(J40B_FMUL(p2, p2), p2), ob2), itscale);
            uint32_t out = 0xff000000u;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = J40B_FADD(J40B_FADD(J40B_FMUL(l0, om[c * 3 + 0]), J40B_FMUL(l1, om[c * 3 + 1])), J40B_FMUL(l2, om[c * 3 + 2]));
                out |= (uint32_t) (srgb_needs_wrap(v, wrap_hi) ? srgb_u8_wrapped(v, 8) : srgb_u8_lut(ts.thr, ts.lut, v)) << (8 * c);
            }
            o.v[e] = out;
        }
        uint8_t *dst = rgba + (size_t) Y * rgba_stride + (size_t) X * 4;
        if (X + 3 < fw) *(TilePx4 *) dst = o; // 16-byte aligned: X is a multiple of 4, the stride one of 32
        else for (int e = 0; X + e < fw; ++e) ((uint32_t *) dst)[e] = o.v[e];
    }
    J40B_PHASE(6);
}

} // namespace j40b
