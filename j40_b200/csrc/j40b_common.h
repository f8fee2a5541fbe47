// j40-b200: Blackwell-native JPEG XL group decoder behind the j40 C API.
// Common definitions shared by host parser, CUDA kernels and the CPU kernel-logic emulator used by
// the `-m "not gpu"` unit tests (tests/hostemu; never part of the product library).
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define J40B_HD __host__ __device__
#define J40B_D __device__
#define J40B_INLINE __forceinline__
#define J40B_NOINLINE __noinline__
#else
#define J40B_HD
#define J40B_D
#define J40B_INLINE inline
#define J40B_NOINLINE
#endif

// float arithmetic must round once per operation, in the written order:
// never let the compiler contract a*b+c.  The CUDA translation units are built with -fmad=false,
// host ones with -ffp-contract=off; these wrappers additionally pin it at the call sites that matter.
#if defined(__CUDA_ARCH__)
#define J40B_FMUL(a, b) __fmul_rn((a), (b))
#define J40B_FADD(a, b) __fadd_rn((a), (b))
#define J40B_FSUB(a, b) __fsub_rn((a), (b))
#define J40B_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define J40B_FMUL(a, b) ((a) * (b))
#define J40B_FADD(a, b) ((a) + (b))
#define J40B_FSUB(a, b) ((a) - (b))
#define J40B_FDIV(a, b) ((a) / (b))
#endif

namespace j40b {

// four-character error codes, same packing as the reference (j40.h:482-483)
#define J40B_4CC(a, b, c, d) ((uint32_t) (((uint32_t) (a) << 24) | ((uint32_t) (b) << 16) | ((uint32_t) (c) << 8) | (uint32_t) (d)))
enum : uint32_t {
    E_OK = 0,
    E_SHRT = J40B_4CC('s', 'h', 'r', 't'),
    E_EXCS = J40B_4CC('e', 'x', 'c', 's'),
    E_PAD0 = J40B_4CC('p', 'a', 'd', '0'),
    E_ANS = J40B_4CC('a', 'n', 's', '?'),
    E_COEF = J40B_4CC('c', 'o', 'e', 'f'),
    E_VBLK = J40B_4CC('v', 'b', 'l', 'k'),
    E_DCT = J40B_4CC('d', 'c', 't', '?'),
    E_POVF = J40B_4CC('p', 'o', 'v', 'f'),
    E_PRED = J40B_4CC('p', 'r', 'e', 'd'),
    E_TREC = J40B_4CC('t', 'r', 'e', 'c'),
    E_IOVF = J40B_4CC('i', 'o', 'v', 'f'),
    E_MTRE = J40B_4CC('m', 't', 'r', 'e'),
    E_XLIM = J40B_4CC('x', 'l', 'i', 'm'),
    E_RCTT = J40B_4CC('r', 'c', 't', 't'),
    E_RCTC = J40B_4CC('r', 'c', 't', 'c'),
    E_RTCD = J40B_4CC('r', 't', 'c', 'd'),
    E_XFM = J40B_4CC('x', 'f', 'm', '?'),
    E_TODO = J40B_4CC('T', 'O', 'D', 'O'),
    E_MEM = J40B_4CC('!', 'm', 'e', 'm'),
    E_TOKV = J40B_4CC('t', 'o', 'k', 'v'), // internal: token arena too small, host retries larger
    E_LTRE = J40B_4CC('l', 't', 'r', 'e'), // internal: an LF-group sub-bitstream has a tree of its own; the host reads it, then retries
};

// j40__unpack_signed: (x & 1) ? -(x / 2 + 1) : x / 2 with C's truncating division, also for negative x (a hybrid integer
// that overflowed): x / 2 == (x + (x < 0)) >> 1 and -(q + 1) == ~q
J40B_HD J40B_INLINE int32_t unpack_signed(int32_t x) {
    const int32_t q = (x + (int32_t) ((uint32_t) x >> 31)) >> 1;
    return (x & 1) ? ~q : q;
}
J40B_HD J40B_INLINE int32_t ceil_div(int32_t x, int32_t y) { return (x + y - 1) / y; }
J40B_HD J40B_INLINE int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
J40B_HD J40B_INLINE int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
J40B_HD J40B_INLINE int32_t iabs(int32_t a) { return a < 0 ? -a : a; }
J40B_HD J40B_INLINE int floor_lg32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int) x);
#else
    return 31 - __builtin_clz(x);
#endif
}
J40B_HD J40B_INLINE int ceil_lg32(uint32_t x) { return x > 1 ? floor_lg32(x - 1) + 1 : 0; }

// ---------------------------------------------------------------------------------------------
// entropy-code tables as laid out in the per-image table arena (all offsets in bytes from the arena base)

struct HybridCfg {
    int8_t split_exp, msb_in_token, lsb_in_token, pad;
    int32_t max_token;
};

struct alignas(16) DCluster {
    HybridCfg cfg;
    uint32_t table_off; // ANS: uint64_t[1 << log_alpha_size]; prefix: uint32_t[] two-level LUT
    int16_t root_bits;  // prefix only: index width of the first-level LUT (0 = zero-length code)
    int16_t max_len;    // prefix only
};

struct DCodeSpec {
    int32_t num_dist;
    int32_t lz77_enabled, use_prefix_code;
    int32_t min_symbol, min_length;
    int32_t log_alpha_size;
    int32_t num_clusters;
    HybridCfg lz_len_cfg;
    uint32_t cluster_map_off; // uint8_t[num_dist]
    uint32_t clusters_off;    // DCluster[num_clusters]
    uint32_t blob_lo, blob_hi; // arena byte range holding everything this spec points to (and the spec)
    uint32_t ans_tables_off;   // ANS: the clusters' alias tables, contiguous in cluster order (cluster k at + (k << log_alpha_size) entries)
};

// ANS alias entry: one 64-bit word per bucket (see j40b_entropy.h)
//   bits  0..7   cutoff            bits 32..44  D[bucket index]
//   bits  8..15  alias symbol      bits 48..60  D[alias symbol]
//   bits 16..27  alias offset
// prefix LUT entry (uint32): bits 0..4 code length, bit 15 = pointer to a second-level table
//   (then bits 5..8 = its index width and bits 16..31 = its offset in entries), else bits 16..31 = symbol

struct DTreeNode {
    int32_t a; // branch: -1 - property (negative); leaf: context (>= 0)
    int32_t b; // branch: threshold value;          leaf: predictor
    int32_t c; // branch: index of the "greater" child; leaf: offset
    int32_t d; // branch: index of the other child;     leaf: multiplier
};

struct WPParams { int8_t p1, p2, p3[5], w[4]; int8_t pad; };

} // namespace j40b
