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

// ---- device tables ---------------------------------------------------------
// Node names are interned: a path node is looked up once (hash probe + byte compare of the name),
// after that a link is the exact integer key (node id, strand, node id, strand).
SVJG_HD uint64_t node_hash(uint64_t tok) { return mix64(tok ^ 0xA0761D6478BD642Full); }
constexpr uint32_t NO_NODE = 0xFFFFFFFFu;

// link (nL, sL, nR, sR): strands are 1 for '+', 0 for '-'; ids < 2^31
SVJG_HD uint64_t link_key(uint32_t idL, uint32_t sL, uint32_t idR, uint32_t sR) {
    return (uint64_t(idL) << 33) | (uint64_t(sL) << 32) | (uint64_t(idR) << 1) | uint64_t(sR);
}
SVJG_HD uint64_t link_hash(uint64_t key) { return mix64(key + 0x9E3779B97F4A7C15ull); }

struct LinkSlot {          // open addressing, linear probing, capacity = power of two; two slots per sector
    uint64_t key;          // link_key()
    uint32_t val;          // count == 1: the entry itself; otherwise entries[val .. val + count)
    uint32_t meta;         // count << 4 | poison_key << 3 | used
};
static_assert(sizeof(LinkSlot) == 16, "LinkSlot is half a sector");

struct NodeSlot {
    uint64_t hash;         // node_hash of the name's token hash
    uint32_t name_off;     // into blob (multiple of 4), zero padded to 4 bytes
    uint32_t name_len;
    int64_t seq_len;       // alt node: length of its GFA sequence (filter-alignments.py:103-113); -1 if the GFA has none
    uint32_t id1;          // node id + 1; 0 = empty slot
    uint32_t pad;
};
static_assert(sizeof(NodeSlot) == 32, "NodeSlot must be one sector");

constexpr uint32_t ENTRY_POISON = 0xFFFFFFFFu;  // entry on which the reference raises

struct DevTables {
    const LinkSlot *links;
    const NodeSlot *nodes;
    const uint8_t *blob;
    const uint32_t *entries;   // 2*sv_index + allele, or ENTRY_POISON
    uint32_t link_mask;        // capacity - 1
    uint32_t node_mask;        // capacity - 1 (0 capacity is never used: min 2 slots)
    uint32_t num_sv;
    uint32_t pad;
};

}  // namespace svjg
