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
// 32-bit mixing only: the probe is made for every link of every multi-node line
SVJG_HD uint32_t link_hash(uint64_t key) {
    uint32_t h = uint32_t(key) * 0x9E3779B1u ^ uint32_t(key >> 32) * 0x85EBCA77u;
    h = (h ^ (h >> 15)) * 0x2C1B3C6Du;
    return h ^ (h >> 13);
}

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

// Plain node names -- chrom:start-end (reference node) or chrom:pos.k (alt node), read from the END of
// the name: 1-9 digits, '-' or '.', 1-9 digits (neither number with a leading zero), ':', and in front
// of that colon a chrom of at most 15 bytes (any bytes) -- have an exact 24-byte key, so the scan
// kernel resolves them from shared memory with one probe and no name bytes are compared.  The last
// byte of the 16-byte chrom field is the chrom's length, so the key determines the name whatever bytes
// the chrom holds.  Any other name is reachable only through the name-hash table above (the exact route).
constexpr uint32_t PN_ALT = 0x80000000u;       // in PNodeSlot::b: chrom:pos.k
constexpr uint32_t PN_NO_LEN = 0xFFFFFFFFu;    // alt node without a usable GFA sequence length
constexpr uint32_t PN_ID_MASK = 0x0FFFFFFFu;   // PNodeSlot::id1: node id + 1
struct PNodeSlot {
    uint64_t c0, c1;       // chrom bytes, little endian, zero padded to 15; byte 15 (top byte of c1) = chrom length
    uint32_t a;            // start / pos
    uint32_t b;            // end, or k | PN_ALT
    uint32_t id1;          // low 28 bits: node id + 1 (same ids as NodeSlot), 0 = empty slot; high 4 bits: the
                           // roles the node has in link keys -- bit 28+s: left node with strand s, bit 30+s:
                           // right node with strand s (s = 1 for '+').  A key (L, sL, R, sR) can exist only if
                           // L has role 28+sL and R has role 30+sR: most reverse-key probes are never made.
    uint32_t alt_len;      // alt node: GFA sequence length (1 .. 2^31-1) or PN_NO_LEN
};
static_assert(sizeof(PNodeSlot) == 32, "PNodeSlot must be one sector");
SVJG_HD uint32_t pnode_hash(uint64_t c0, uint64_t c1, uint32_t a, uint32_t b) {
    uint32_t h = uint32_t(c0) * 0x9E3779B1u;
    h = (h ^ uint32_t(c0 >> 32)) * 0x85EBCA77u;
    h = (h ^ uint32_t(c1)) * 0xC2B2AE3Du;
    h = (h ^ uint32_t(c1 >> 32)) * 0x27D4EB2Fu;
    h = (h ^ a) * 0x165667B1u;
    h = (h ^ b ^ (h >> 15)) * 0x2C1B3C6Du;
    h ^= h >> 13;
    h *= 0x297A2D39u;
    h ^= h >> 16;
    return h;
}

constexpr uint32_t ENTRY_POISON = 0xFFFFFFFFu;  // entry on which the reference raises

struct DevTables {
    const LinkSlot *links;
    const NodeSlot *nodes;
    const uint8_t *blob;
    const uint32_t *entries;   // 2*sv_index + allele, or ENTRY_POISON
    const PNodeSlot *pnodes;
    uint32_t link_mask;        // capacity - 1
    uint32_t node_mask;        // capacity - 1 (0 capacity is never used: min 2 slots)
    uint32_t pnode_mask;
    uint32_t num_sv;
};

}  // namespace svjg
