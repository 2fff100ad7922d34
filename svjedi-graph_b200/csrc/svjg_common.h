// Shared between the host table builder (tables.cpp) and the kernels (filter.cu):
// device table layout and the hash functions both sides must agree on.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SVJG_HD __host__ __device__ __forceinline__
#else
#define SVJG_HD inline
#endif

namespace svjg {

// ---- token hash: two independent 32-bit multiplicative streams over the name
// taken as little-endian 4-byte words (the last one zero padded), closed with
// the length
struct TokHash {
    uint32_t a, b;
};
SVJG_HD TokHash tok_init() { return TokHash{0x811C9DC5u, 0x2F0B4A67u}; }
SVJG_HD void tok_step(TokHash &h, uint32_t w) {
    uint32_t x = h.a ^ w;
    h.a = ((x << 13) | (x >> 19)) * 0x9E3779B1u;
    h.b = (h.b ^ w) * 0x85EBCA77u + 0x7F4A7C15u;
}
SVJG_HD uint64_t tok_value(const TokHash &h, uint32_t len) {
    return (uint64_t(h.a ^ (len * 0x27D4EB2Fu)) << 32) | h.b;
}

SVJG_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
SVJG_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// link (nL, sL, nR, sR): strands are 1 for '+', 0 for '-'
SVJG_HD uint64_t link_hash(uint64_t tokL, uint32_t sL, uint64_t tokR, uint32_t sR) {
    uint64_t x = tokL * 0x9E3779B97F4A7C15ull + rotl64(tokR, 23) * 0xC2B2AE3D27D4EB4Full +
                 uint64_t(sL * 2 + sR + 1) * 0x165667B19E3779F9ull;
    return mix64(x);
}
SVJG_HD uint64_t alt_hash(uint64_t tok) { return mix64(tok ^ 0xA0761D6478BD642Full); }

// ---- device tables ---------------------------------------------------------
// One 32-byte sector per slot: a probe that lands on the right slot needs one
// DRAM/L2 sector for everything but the name check.
struct LinkSlot {          // open addressing, linear probing, capacity = power of two
    uint64_t hash;         // link_hash of the key
    uint32_t name_off;     // into blob (multiple of 4): left name, zero padded to 4 bytes, then right name, padded
    uint16_t len_l, len_r;
    uint32_t ent_begin;    // entries[ent_begin .. ent_begin + count)
    uint32_t meta;         // count << 4 | poison_key << 3 | sL << 2 | sR << 1 | used
    uint32_t ent0;         // copy of entries[ent_begin] (most keys have one entry)
    uint32_t pad;
};
static_assert(sizeof(LinkSlot) == 32, "LinkSlot must be one sector");

struct AltSlot {
    uint64_t hash;         // alt_hash of the node name
    uint32_t name_off;
    uint32_t name_len;
    int64_t seq_len;
    uint32_t used;
    uint32_t pad;
};
static_assert(sizeof(AltSlot) == 32, "AltSlot must be one sector");

constexpr uint32_t ENTRY_POISON = 0xFFFFFFFFu;  // entry on which the reference raises

struct DevTables {
    const LinkSlot *links;
    const AltSlot *alts;
    const uint8_t *blob;
    const uint32_t *entries;   // 2*sv_index + allele, or ENTRY_POISON
    uint32_t link_mask;        // capacity - 1
    uint32_t alt_mask;         // capacity - 1 (0 capacity is never used: min 2 slots)
    uint32_t num_sv;
    uint32_t pad;
};

}  // namespace svjg
