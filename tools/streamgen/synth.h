// JPEG XL test-stream writer -- TEST INFRASTRUCTURE.
// Deterministic "photo-like" synthetic source images: multi-octave value noise
// with a 1/f-ish spectrum + smooth gradients + hard-edged shapes, 8-bit sRGB.
#pragma once
#include "bits.h"

namespace jxlgen {

struct ImageRGB8 {
    int w = 0, h = 0;
    std::vector<uint8_t> px; // [h][w][3]
};

inline ImageRGB8 synth_photo(int w, int h, uint64_t seed) {
    Rng rng(seed ^ 0xabcdef12345ull);
    ImageRGB8 im;
    im.w = w; im.h = h;
    im.px.assign((size_t) w * (size_t) h * 3, 0);
    std::vector<float> acc((size_t) w * (size_t) h * 3, 0.0f);
    // octaves of bilinearly interpolated random lattices; amplitude proportional to cell size^0.9
    int maxcell = 256;
    while (maxcell > std::max(w, h)) maxcell >>= 1;
    if (maxcell < 2) maxcell = 2;
    double norm = 0;
    for (int cell = maxcell; cell >= 2; cell >>= 1) {
        double amp = std::pow((double) cell, 0.55);
        norm += amp;
        int gw = w / cell + 2, gh = h / cell + 2;
        std::vector<float> lat((size_t) gw * (size_t) gh * 3);
        // colour noise is mostly achromatic with a smaller chromatic part (like natural images)
        for (size_t i = 0; i < (size_t) gw * (size_t) gh; ++i) {
            float l = (float) (rng.uni() - 0.5);
            for (int c = 0; c < 3; ++c) lat[i * 3 + (size_t) c] = l + 0.35f * (float) (rng.uni() - 0.5);
        }
        float inv = 1.0f / (float) cell;
        for (int y = 0; y < h; ++y) {
            int gy = y / cell;
            float fy = (float) (y - gy * cell) * inv;
            fy = fy * fy * (3 - 2 * fy);
            for (int x = 0; x < w; ++x) {
                int gx = x / cell;
                float fx = (float) (x - gx * cell) * inv;
                fx = fx * fx * (3 - 2 * fx);
                const float *p00 = &lat[((size_t) gy * (size_t) gw + (size_t) gx) * 3];
                const float *p01 = p00 + 3, *p10 = p00 + (size_t) gw * 3, *p11 = p10 + 3;
                float *o = &acc[((size_t) y * (size_t) w + (size_t) x) * 3];
                for (int c = 0; c < 3; ++c) {
                    float a = p00[c] + (p01[c] - p00[c]) * fx;
                    float b = p10[c] + (p11[c] - p10[c]) * fx;
                    o[c] += (float) amp * (a + (b - a) * fy);
                }
            }
        }
    }
    // finest-scale sensor-like noise
    float fine = 0.02f;
    // gradients
    float g0[3], gx[3], gy[3];
    for (int c = 0; c < 3; ++c) {
        g0[c] = 0.35f + 0.3f * (float) rng.uni();
        gx[c] = 0.3f * (float) (rng.uni() - 0.5);
        gy[c] = 0.3f * (float) (rng.uni() - 0.5);
    }
    // shapes: rectangles and discs with hard edges
    struct Shape { int kind; float cx, cy, rx, ry; float col[3]; float alpha; };
    std::vector<Shape> shapes;
    int nshapes = 4 + rng.below(5) + (w * h) / (512 * 512);
    for (int i = 0; i < nshapes; ++i) {
        Shape s;
        s.kind = rng.below(2);
        s.cx = (float) rng.uni() * (float) w;
        s.cy = (float) rng.uni() * (float) h;
        s.rx = (0.02f + 0.15f * (float) rng.uni()) * (float) std::max(w, h);
        s.ry = (0.02f + 0.15f * (float) rng.uni()) * (float) std::max(w, h);
        for (int c = 0; c < 3; ++c) s.col[c] = (float) rng.uni();
        s.alpha = 0.5f + 0.5f * (float) rng.uni();
        shapes.push_back(s);
    }
    float scale = 1.6f / (float) norm;
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        float *o = &acc[((size_t) y * (size_t) w + (size_t) x) * 3];
        float v[3];
        for (int c = 0; c < 3; ++c) {
            v[c] = g0[c] + gx[c] * ((float) x / (float) w - 0.5f) + gy[c] * ((float) y / (float) h - 0.5f) + o[c] * scale;
        }
        for (const Shape &s : shapes) {
            float dx = ((float) x - s.cx) / s.rx, dy = ((float) y - s.cy) / s.ry;
            bool in = s.kind == 0 ? (std::fabs(dx) < 1 && std::fabs(dy) < 1) : (dx * dx + dy * dy < 1);
            if (in) for (int c = 0; c < 3; ++c) v[c] = v[c] * (1 - s.alpha) + s.col[c] * s.alpha + 0.25f * o[c] * scale;
        }
        for (int c = 0; c < 3; ++c) {
            float f = v[c] + fine * (float) (rng.uni() - 0.5) * 2.0f;
            int q = (int) std::floor(f * 255.0f + 0.5f);
            im.px[((size_t) y * (size_t) w + (size_t) x) * 3 + (size_t) c] = (uint8_t) std::min(255, std::max(0, q));
        }
    }
    return im;
}

} // namespace jxlgen
