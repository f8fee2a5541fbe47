// JPEG XL test-stream writer -- TEST INFRASTRUCTURE (tests/, bench.py input synthesis).
// Bit-level output primitives, mirroring the *reader* side described in SURVEY.md App. E.1
// (j40.h:1914-2008): LSB-first bit packing, U32/U64/F16/Enum field codings.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <stdexcept>

namespace jxlgen {

struct GenError : std::runtime_error {
    explicit GenError(const std::string &s) : std::runtime_error(s) {}
};
#define JG_CHECK(cond) do { if (!(cond)) throw GenError(std::string("jxlgen check failed: ") + #cond + " @" + __FILE__ + ":" + std::to_string(__LINE__)); } while (0)

inline int floor_lg(uint32_t x) { return 31 - __builtin_clz(x); }
inline int ceil_lg(uint32_t x) { return x > 1 ? 32 - __builtin_clz(x - 1) : 0; }

class BitWriter {
public:
    std::vector<uint8_t> bytes;
    uint64_t acc = 0;
    int nacc = 0;

    void put(uint64_t v, int n) {
        JG_CHECK(n >= 0 && n <= 56);
        if (n == 0) return;
        JG_CHECK(n == 64 || (v >> n) == 0);
        acc |= v << nacc;
        nacc += n;
        while (nacc >= 8) {
            bytes.push_back((uint8_t) acc);
            acc >>= 8;
            nacc -= 8;
        }
    }
    void pad() {
        if (nacc > 0) {
            bytes.push_back((uint8_t) acc);
            acc = 0;
            nacc = 0;
        }
    }
    size_t bitpos() const { return bytes.size() * 8 + (size_t) nacc; }
    // append a byte-aligned blob (this writer must be byte-aligned)
    void append_bytes(const std::vector<uint8_t> &b) {
        JG_CHECK(nacc == 0);
        bytes.insert(bytes.end(), b.begin(), b.end());
    }
    // append all bits of another writer (which need not be aligned)
    void append_bits(const BitWriter &o) {
        for (uint8_t b : o.bytes) put(b, 8);
        if (o.nacc) put(o.acc & ((1ull << o.nacc) - 1), o.nacc);
    }

    void bit(int b) { put((uint64_t) (b ? 1 : 0), 1); }

    // U32(o0,n0,o1,n1,o2,n2,o3,n3): picks the first selector able to represent v
    void u32(uint32_t v, uint32_t o0, int n0, uint32_t o1, int n1, uint32_t o2, int n2, uint32_t o3, int n3, int force_sel = -1) {
        const uint32_t o[4] = {o0, o1, o2, o3};
        const int n[4] = {n0, n1, n2, n3};
        for (int s = 0; s < 4; ++s) {
            if (force_sel >= 0 && s != force_sel) continue;
            if (v >= o[s] && (uint64_t) (v - o[s]) < (1ull << n[s])) {
                put((uint64_t) s, 2);
                put(v - o[s], n[s]);
                return;
            }
        }
        throw GenError("u32: value not representable");
    }
    void u64(uint64_t v) {
        if (v == 0) { put(0, 2); }
        else if (v >= 1 && v <= 16) { put(1, 2); put(v - 1, 4); }
        else if (v >= 17 && v <= 272) { put(2, 2); put(v - 17, 8); }
        else {
            put(3, 2);
            put(v & 0xfff, 12);
            v >>= 12;
            int shift = 12;
            while (v) {
                put(1, 1);
                if (shift < 56) { put(v & 0xff, 8); v >>= 8; shift += 8; }
                else { put(v & 0xf, 4); v = 0; shift = 64; }
            }
            if (shift < 64) put(0, 1);
        }
    }
    void enum_(uint32_t v) { u32(v, 0, 0, 1, 0, 2, 4, 18, 6); }
    // IEEE half; value must be exactly representable (we only ever write "nice" numbers)
    void f16(float f) {
        uint32_t x;
        std::memcpy(&x, &f, 4);
        uint32_t sign = x >> 31, exp = (x >> 23) & 0xff, man = x & 0x7fffff;
        uint32_t h;
        if (exp == 0 && man == 0) h = sign << 15;
        else {
            int e = (int) exp - 127 + 15;
            JG_CHECK(e > 0 && e < 31);      // normal halves only
            JG_CHECK((man & 0x1fff) == 0);  // exactly representable
            h = (sign << 15) | ((uint32_t) e << 10) | (man >> 13);
        }
        put(h, 16);
    }
    // u8() of the ANS histogram coding (j40.h:1994)
    void u8(uint32_t v) {
        if (v == 0) { put(0, 1); return; }
        int n = floor_lg(v);
        put(1, 1);
        put((uint64_t) n, 3);
        put(v - (1u << n), n);
    }
    void at_most(uint32_t v, uint32_t max) {
        JG_CHECK(v <= max);
        if (max > 0) put(v, ceil_lg(max + 1));
    }
};

// round a float to the nearest value exactly representable as a (normal) IEEE half
inline float round_to_f16(float f) {
    if (f == 0.0f) return 0.0f;
    uint32_t x;
    std::memcpy(&x, &f, 4);
    x = (x + 0x1000) & ~0x1fffu;
    float r;
    std::memcpy(&r, &x, 4);
    return r;
}

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    double uni() { return (double) (next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int) (next() % (uint64_t) n); }
    double gauss() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
};

} // namespace jxlgen
