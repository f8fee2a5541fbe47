// j40-b200: section bit reader and entropy decoder (rANS / prefix / hybrid-uint / LZ77).
//
// Replaces, for sections decoded on the device, the reference's
//   bit reader        j40.h:1847-1932, 2011-2016  (j40__always_refill, j40__u, j40__no_more_bytes)
//   prefix decode     j40.h:2244-2273             (j40__match_overflow, j40__prefix_code)
//   hybrid integers   j40.h:2313-2327             (j40__hybrid_int)
//   rANS step         j40.h:2441-2461             (j40__ans_code)
//   symbol reader     j40.h:2804-2876, 2884-2898  (j40__code, j40__finish_and_free_code)
// The same functions (compiled for the host) serve the host-side header parser.
//
// Error model: the reference stops at the first error; running past the end of a section is `shrt`.
// Here reads past the section end deliver zero bits and are detected from the running bit count:
// `check()` turns any later error into `shrt` if the overrun came first (which is what the
// reference would have reported), see DESIGN.md "error parity".
#pragma once
#include "j40b_common.h"

namespace j40b {

struct BitReader {
    const uint8_t *base; // first byte of the section
    uint32_t size;       // section size in bytes
    uint32_t pos;        // next (virtual) byte to load
    uint64_t buf;
    int32_t nbits;

#if defined(__CUDA_ARCH__)
    // Device: aligned 32-bit loads, one word ahead (`nxt` holds the word at `pos`, so that a refill never waits for
    // memory: the load of the following word is issued when the current one is consumed). Bytes past the section end
    // are whatever follows in the codestream buffer (the executor pads it by 16 bytes) up to 8 bytes behind the end
    // and zeros from there on; they are never *accepted*: any consumption past the end is an overrun, which is the
    // only thing the error model looks at. The clamp keeps a decoder that runs on after a truncated section (until
    // its next overrun check) inside the buffer.
    uint32_t nxt;
    J40B_HD J40B_INLINE uint32_t load_word(uint32_t at) const {
        return at < size + 8 ? *(const uint32_t *) (base + at) : 0u;
    }
    J40B_HD void init(const uint8_t *b, uint32_t size_bytes, uint64_t start_bit = 0) {
        base = b;
        size = size_bytes;
        const uint8_t *p = b + (start_bit >> 3);
        uint32_t mis = (uint32_t) ((uintptr_t) p & 3);
        const uint32_t *wp = (const uint32_t *) (p - mis); // (callers start inside the section or right behind it)
        buf = (uint64_t) (*wp >> (8 * mis));
        nbits = 32 - 8 * (int32_t) mis;
        pos = (uint32_t) (start_bit >> 3) + 4 - mis;       // base + pos is 4-byte aligned from here on
        nxt = load_word(pos);
        int skip = (int) (start_bit & 7);
        if (skip) { buf >>= skip; nbits -= skip; }
    }
    J40B_HD J40B_INLINE void refill() { // precondition: nbits <= 32
        buf |= (uint64_t) nxt << nbits;
        nbits += 32;
        pos += 4;
        nxt = load_word(pos);
    }
    J40B_HD J40B_INLINE uint32_t u(int n) { // n in [0, 32]
        if (nbits < 32) refill();
        uint32_t lo = (uint32_t) buf;
        uint32_t v = n >= 32 ? lo : (lo & ((1u << n) - 1));
        buf >>= n;
        nbits -= n;
        return v;
    }
    J40B_HD J40B_INLINE uint32_t peek(int n) {
        if (nbits < 32) refill();
        uint32_t lo = (uint32_t) buf;
        return n >= 32 ? lo : (lo & ((1u << n) - 1));
    }
#else
    J40B_HD void init(const uint8_t *b, uint32_t size_bytes, uint64_t start_bit = 0) {
        base = b;
        size = size_bytes;
        pos = (uint32_t) (start_bit >> 3);
        buf = 0;
        nbits = 0;
        int skip = (int) (start_bit & 7);
        if (skip) { refill(); buf >>= skip; nbits -= skip; }
    }
    J40B_HD J40B_INLINE void refill() {
        // byte-granular like the reference (j40.h:1858); bytes past the end are zeros
        while (nbits <= 56) {
            uint64_t byte = pos < size ? (uint64_t) base[pos] : 0;
            buf |= byte << nbits;
            nbits += 8;
            ++pos;
        }
    }
    J40B_HD J40B_INLINE uint32_t u(int n) { // n in [0, 32]
        if (nbits < n) refill();
        uint32_t v = (uint32_t) (buf & ((1ull << n) - 1));
        buf >>= n;
        nbits -= n;
        return v;
    }
    J40B_HD J40B_INLINE uint32_t peek(int n) { // n <= 32; zero-padded past the end
        if (nbits < n) refill();
        return (uint32_t) (buf & ((1ull << n) - 1));
    }
#endif
    J40B_HD J40B_INLINE void skip(int n) { buf >>= n; nbits -= n; }
    J40B_HD J40B_INLINE uint64_t bits_consumed() const { return (uint64_t) pos * 8 - (uint64_t) nbits; }
    // consumed more bits than the section has: 8 * pos - nbits > 8 * size, i.e. pos - size > floor(nbits / 8) (nbits >= 0;
    // pos and size are below 2^31, so the difference fits 32 bits)
    J40B_HD J40B_INLINE bool overrun() const { return (int32_t) (pos - size) > (nbits >> 3); }
    // j40__zero_pad_to_byte: returns false if a padding bit is set
    J40B_HD bool zero_pad_to_byte() {
        int n = (int) ((8 - (bits_consumed() & 7)) & 7);
        return u(n) == 0;
    }
    // end of a single-section frame (j40__end_of_frame, j40.h:7884-7893): 0 or an error code. Note the reference's
    // naming: a frame that ends *before* its advertised size is "shrt" there (the bytes up to the advertised end
    // would have to be skipped), and one that runs past its size into trailing data is "excs"; running past the
    // end of the file is "shrt" from the bit reader itself, which is all a section-bounded reader can see.
    J40B_HD uint32_t finish() {
        if (overrun()) return E_SHRT;
        if (!zero_pad_to_byte()) return E_PAD0;
        if (overrun()) return E_SHRT;
        if (bits_consumed() != (uint64_t) size * 8) return E_SHRT;
        return 0;
    }
    // U32 field coding (j40.h:1934)
    J40B_HD uint32_t u32(uint32_t o0, int n0, uint32_t o1, int n1, uint32_t o2, int n2, uint32_t o3, int n3) {
        uint32_t sel = u(2);
        switch (sel) {
        case 0: return u(n0) + o0;
        case 1: return u(n1) + o1;
        case 2: return u(n2) + o2;
        default: return u(n3) + o3;
        }
    }
};

// sticky first-error slot with the overrun rule described above
struct ErrSlot {
    uint32_t err;
    J40B_HD void set(const BitReader &br, uint32_t code) {
        if (!err) err = br.overrun() ? (uint32_t) E_SHRT : code;
    }
    J40B_HD void set_raw(uint32_t code) { if (!err) err = code; }
};

struct CodeState {
    uint32_t ans_state; // 0 = not yet initialised
    int32_t num_to_copy, copy_pos, num_decoded;
    int32_t *window;     // >= min(1 << 20, symbols in this sub-bitstream) entries when LZ77 is enabled
    uint32_t window_mask; // capacity - 1 (capacity is a power of two <= 1 << 20)
    J40B_HD void init(int32_t *win, uint32_t mask) {
        ans_state = 0;
        num_to_copy = copy_pos = num_decoded = 0;
        window = win;
        window_mask = mask;
    }
};

J40B_HD J40B_INLINE int32_t hybrid_int(BitReader &br, ErrSlot &es, int32_t token, HybridCfg c) {
    int32_t split = 1 << c.split_exp;
    if (token < split) return token;
    if (token > c.max_token) {
        token = c.max_token;
        es.set(br, E_IOVF);
    }
    int32_t bits_in_token = c.msb_in_token + c.lsb_in_token;
    int32_t midbits = c.split_exp - bits_in_token + ((token - split) >> bits_in_token);
    int32_t mid = (int32_t) br.u(midbits);
    int32_t top = 1 << c.msb_in_token;
    int32_t lo = token & ((1 << c.lsb_in_token) - 1);
    int32_t hi = (token >> c.lsb_in_token) & (top - 1);
    return (int32_t) (((uint32_t) (top | hi) << (midbits + c.lsb_in_token)) | (((uint32_t) mid << c.lsb_in_token) | (uint32_t) lo));
}

// INIT_CHECK = false: the caller has already seeded the state (ans_seed) before its first symbol
J40B_HD J40B_INLINE void ans_seed(BitReader &br, uint32_t &state) {
    state = br.u(16);
    state |= br.u(16) << 16;
}

// the rANS step once the alias entry `e` of bucket (state & 0xfff) >> log_bucket_size is known
J40B_HD J40B_INLINE int32_t ans_symbol_entry(BitReader &br, uint32_t &state, int log_bucket_size, uint64_t e) {
    uint32_t idx = state & 0xfff;
    uint32_t i = idx >> log_bucket_size;
    uint32_t p = idx & ((1u << log_bucket_size) - 1);
    uint32_t lo = (uint32_t) e, hi = (uint32_t) (e >> 32);
    bool own = p < (lo & 0xff);
    uint32_t sym = own ? i : ((lo >> 8) & 0xff);
    uint32_t off = own ? 0u : ((lo >> 16) & 0xfff);
    uint32_t d = own ? (hi & 0x1fff) : ((hi >> 16) & 0x1fff);
    state = d * (state >> 12) + off + p;
    if (state < (1u << 16)) state = (state << 16) | br.u(16);
    return (int32_t) sym;
}

template <bool INIT_CHECK = true>
J40B_HD J40B_INLINE int32_t ans_symbol(BitReader &br, uint32_t &state, int log_bucket_size, const uint64_t *table) {
    if (INIT_CHECK && state == 0) ans_seed(br, state);
    return ans_symbol_entry(br, state, log_bucket_size, table[(state & 0xfff) >> log_bucket_size]);
}

J40B_HD J40B_INLINE int32_t prefix_symbol(BitReader &br, int root_bits, const uint32_t *table) {
    uint32_t bits = br.peek(16);
    uint32_t e = table[bits & ((1u << root_bits) - 1)];
    if (e & 0x8000u) { // second level
        uint32_t sub_bits = (e >> 5) & 15;
        e = table[(e >> 16) + ((bits >> root_bits) & ((1u << sub_bits) - 1))];
    }
    br.skip((int) (e & 31));
    return (int32_t) (e >> 16);
}

struct CodeCtx { // everything a symbol read needs besides the state
    const uint8_t *arena;
    const DCodeSpec *spec;
    const uint8_t *cluster_map;
    const DCluster *clusters;
    // per-symbol fields of the spec, kept in registers
    int32_t min_symbol;   // INT32_MAX when LZ77 is off, so that one compare decides
    int32_t log_bucket;   // ANS: 12 - log_alpha_size
    bool prefix, lz77;
    J40B_HD void init(const uint8_t *arena_, uint32_t spec_off) {
        arena = arena_;
        spec = (const DCodeSpec *) (arena_ + spec_off);
        cluster_map = arena_ + spec->cluster_map_off;
        clusters = (const DCluster *) (arena_ + spec->clusters_off);
        prefix = spec->use_prefix_code != 0;
        lz77 = spec->lz77_enabled != 0;
        min_symbol = lz77 ? spec->min_symbol : 0x7fffffff;
        log_bucket = 12 - spec->log_alpha_size;
    }
    // `copy` holds the bytes [blob_lo, blob_hi) of the arena (cluster map, clusters, tables, spec): all
    // table offsets keep working relative to (copy - blob_lo)
    J40B_HD void init_from_copy(const uint8_t *copy, uint32_t blob_lo, uint32_t spec_off) { init(copy - blob_lo, spec_off); }
};

// MODE 0: read the spec's flags at run time; MODE 1: the caller knows this is rANS without LZ77
template <bool INIT_CHECK = true, int MODE = 0>
J40B_HD J40B_INLINE int32_t cluster_symbol(BitReader &br, const CodeCtx &cc, const DCluster &cl, uint32_t &ans_state) {
    if (MODE == 0 && cc.prefix) {
        return prefix_symbol(br, cl.root_bits, (const uint32_t *) (cc.arena + cl.table_off));
    } else {
        return ans_symbol<INIT_CHECK>(br, ans_state, cc.log_bucket, (const uint64_t *) (cc.arena + cl.table_off));
    }
}

// the reference's special LZ77 distances for 2-D data, {a, b} encoded as (a+7)*16+b (j40.h:2834)
J40B_HD J40B_INLINE int32_t special_distance(int idx) {
    // generated from the format definition: offsets ordered by increasing Euclidean-ish distance
    const uint8_t T[120] = {
        0x71, 0x80, 0x81, 0x61, 0x72, 0x90, 0x82, 0x62, 0x91, 0x51, 0x92, 0x52,
        0x73, 0xa0, 0x83, 0x63, 0xa1, 0x41, 0x93, 0x53, 0xa2, 0x42, 0x74, 0xb0,
        0x84, 0x64, 0xb1, 0x31, 0xa3, 0x43, 0x94, 0x54, 0xb2, 0x32, 0x75, 0xa4,
        0x44, 0xb3, 0x33, 0xc0, 0x85, 0x65, 0xc1, 0x21, 0x95, 0x55, 0xc2, 0x22,
        0xb4, 0x34, 0xa5, 0x45, 0xc3, 0x23, 0x76, 0xd0, 0x86, 0x66, 0xd1, 0x11,
        0x96, 0x56, 0xd2, 0x12, 0xb5, 0x35, 0xc4, 0x24, 0xa6, 0x46, 0xd3, 0x13,
        0x77, 0xe0, 0x87, 0x67, 0xc5, 0x25, 0xe1, 0x01, 0xb6, 0x36, 0xd4, 0x14,
        0x97, 0x57, 0xe2, 0x02, 0xa7, 0x47, 0xe3, 0x03, 0xc6, 0x26, 0xd5, 0x15,
        0xf0, 0xb7, 0x37, 0xe4, 0x04, 0xf1, 0xf2, 0xd6, 0x16, 0xf3, 0xc7, 0x27,
        0xe5, 0x05, 0xf4, 0xd7, 0x17, 0xe6, 0x06, 0xf5, 0xe7, 0x07, 0xf6, 0xf7,
    };
    return (int32_t) T[idx];
}

// one decoded integer (aka DecodeHybridVarLenUint) whose context has already been resolved to cluster `cl`;
// mirrors j40__code incl. its LZ77 quirks
template <bool INIT_CHECK = true, int MODE = 0>
J40B_HD J40B_INLINE int32_t code_cluster(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs, const DCluster &cl, int32_t dist_mult) {
    int32_t token = cluster_symbol<INIT_CHECK, MODE>(br, cc, cl, cs.ans_state);
    if (MODE == 0 && token >= cc.min_symbol) { // only possible when LZ77 is enabled
        const DCodeSpec *spec = cc.spec;
        const DCluster lz = cc.clusters[cc.cluster_map[spec->num_dist - 1]];
        int32_t num_to_copy = hybrid_int(br, es, token - spec->min_symbol, spec->lz_len_cfg) + spec->min_length;
        token = cluster_symbol<INIT_CHECK>(br, cc, lz, cs.ans_state);
        int32_t distance = hybrid_int(br, es, token, lz.cfg);
        if (es.err) return 0;
        if (!dist_mult) {
            ++distance;
        } else if (distance >= 120) {
            distance -= 119;
        } else {
            int32_t special = special_distance(distance);
            distance = imax(1, ((special >> 4) - 7) + dist_mult * (special & 7));
        }
        distance = imin(imin(distance, cs.num_decoded), 1 << 20);
        cs.copy_pos = cs.num_decoded - distance;
        cs.num_to_copy = num_to_copy - 1;
        int32_t v = cs.copy_pos == cs.num_decoded ? 0 : cs.window[(uint32_t) cs.copy_pos & cs.window_mask];
        ++cs.copy_pos;
        cs.window[(uint32_t) cs.num_decoded++ & cs.window_mask] = v;
        return v;
    }
    token = hybrid_int(br, es, token, cl.cfg);
    if (es.err) return 0;
    if (MODE == 0 && cc.lz77) cs.window[(uint32_t) cs.num_decoded++ & cs.window_mask] = token;
    return token;
}

// the LZ77 copy in progress, if any (the part of j40__code that precedes the symbol read)
J40B_HD J40B_INLINE bool code_copy(CodeState &cs, int32_t &v) {
    if (cs.num_to_copy <= 0) return false;
    --cs.num_to_copy;
    v = cs.copy_pos == cs.num_decoded ? 0 : cs.window[(uint32_t) cs.copy_pos & cs.window_mask];
    ++cs.copy_pos;
    cs.window[(uint32_t) cs.num_decoded++ & cs.window_mask] = v;
    return true;
}

J40B_HD J40B_INLINE int32_t code(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs, int32_t ctx, int32_t dist_mult) {
    int32_t v;
    if (code_copy(cs, v)) return v;
    const DCluster cl = cc.clusters[cc.cluster_map[ctx]];
    return code_cluster(br, es, cc, cs, cl, dist_mult);
}

// end of one entropy-coded stream (j40__finish_and_free_code): the rANS state must be back at its seed
J40B_HD J40B_INLINE void finish_code(BitReader &br, ErrSlot &es, const CodeCtx &cc, CodeState &cs) {
    if (!cc.prefix) {
        if (cs.ans_state) {
            if (cs.ans_state != 0x130000u) es.set(br, E_ANS);
        } else {
            uint32_t lo = br.u(16), hi = br.u(16);
            if (lo != 0x0000u || hi != 0x0013u) es.set(br, E_ANS);
        }
    }
}

} // namespace j40b
