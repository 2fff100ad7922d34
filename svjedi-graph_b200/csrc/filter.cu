// Kernels 1-3 of the hot path, fused: GAF ingest (newline scan + field split),
// alignment -> allele resolution through the device link hash, and the per-SV
// counter / hit-tuple reduction.  One pass over the GAF bytes.
//
// Reference semantics restated (filter-alignments.py): per-line loop :123-166,
// read_gaf_line :184-198, extract_nodes :351-373, get_aln_links :200-219,
// reverse_link :221-225, check_bkpt_overlap :258-273, get_node_len :343-349.
//
// Layout: every WARP is an independent worker with its own slice of shared memory
// (no block-wide barrier anywhere).  A warp walks 4 KiB tiles of the byte buffer;
// a tile plus 2 KiB of look-ahead (and the 32 bytes in front of it) is staged into
// the warp's window by one TMA bulk copy (cp.async.bulk + mbarrier).  Per tile the
// 32 lanes are re-assigned at the granularity that keeps them busy:
//   A  byte-parallel: every lane classifies 32 bytes with SWAR compares; ballots
//      turn the newline flags into the ordered list of line boundaries;
//   B  line-parallel: one lane per line finds the 12 columns (tab bitmaps of fixed
//      spans, so lanes stay converged), validates the integer columns, walks the
//      path column and, for paths with >= 2 nodes, appends one record per node
//      to the token list (slots handed out by a warp scan);
//   C  token-parallel: one lane per path node hashes the name (4 bytes a step),
//      parses chrom:start-end / looks the alt node up, and stores the record;
//   D  line-parallel over multi-node lines: Tlen/Ts/Te, running sums of the node
//      lengths, the breakpoint-overlap verdict of every link; token-parallel
//      check that no node name can occur inside an earlier one;
//   E  link-parallel: forward and reverse key probes of the link hash; hits are
//      staged in shared memory and written out (one cursor atomic per warp and
//      tile, warp-aggregated counter atomics) while the next tile's bytes fly in.
// Anything that is not of the plain shape (odd integers, odd node names, a name
// that could be a substring of an earlier one, lines longer than the window, ...)
// is handed to the exact per-line routines parse_fields() / general(), which
// follow the reference's string semantics literally.
// A line belongs to the tile its first byte is in.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "svjg_internal.h"

namespace cg = cooperative_groups;
using namespace svjg;

namespace {

constexpr int TILE = 4096;                     // bytes a warp owns per step
constexpr int LOOKAHEAD = 2048;                // staged behind the tile so that lines starting in it are whole
constexpr int HEAD = 32;                       // staged in front of it (the newline that starts the first line)
constexpr int WIN = HEAD + TILE + LOOKAHEAD;   // 6176 = 32 * 193
constexpr int WARPS = 4;                       // per block; warps never synchronise with each other
constexpr int THREADS = WARPS * 32;
constexpr int NPAIRS = WIN / 32;               // 32-byte pairs of 16-byte chunks
constexpr int HEAD_SPAN = 96;                  // bytes searched for the tabs of columns 1-5 on the fast route
constexpr int TAIL_SPAN = 96;                  // bytes searched for the tabs of columns 7-12
constexpr int SPARE = 176;                     // readable bytes behind the window for those fixed-span reads
constexpr int NLCAP = 400;                     // newlines per window (more => some line is shorter than 16 bytes)
constexpr int MCAP = 64;                       // multi-node lines per tile on the token-parallel route
constexpr int TCAP = 192;                      // path nodes per tile on the token-parallel route
constexpr int HCAP = 128;                      // hits staged per tile before the flush
static_assert(WIN % 32 == 0 && WIN + SPARE <= 65536, "window offsets are 16 bit");

// shared memory map of ONE warp (bytes)
constexpr int OFF_WIN = 0;
constexpr int OFF_NL = OFF_WIN + WIN + SPARE;               // newline positions, ascending  u16[NLCAP]
constexpr int OFF_ML = (OFF_NL + NLCAP * 2 + 15) & ~15;     // MLine[MCAP]
constexpr int OFF_TH = (OFF_ML + MCAP * 16 + 15) & ~15;     // token hash   u64[TCAP]
constexpr int OFF_TLEN = OFF_TH + TCAP * 8;                 // node length  i32[TCAP]
constexpr int OFF_TS = OFF_TLEN + TCAP * 4;                 // start value  u32[TCAP]
constexpr int OFF_TB = OFF_TS + TCAP * 4;                   // token begin  u16[TCAP]
constexpr int OFF_TL = OFF_TB + TCAP * 2;                   // token length u16[TCAP]
constexpr int OFF_TLINE = OFF_TL + TCAP * 2;                // MLine index  u16[TCAP]
constexpr int OFF_TF = OFF_TLINE + TCAP * 2;                // flags        u8[TCAP]
constexpr int OFF_HSV = (OFF_TF + TCAP + 15) & ~15;         // staged hits: 2*sv + allele  u32[HCAP]
constexpr int OFF_HML = OFF_HSV + HCAP * 4;                 //              MLine index    u16[HCAP]
constexpr int WARP_SMEM = (OFF_HML + HCAP * 2 + 127) & ~127;
constexpr int SMEM_BYTES = WARP_SMEM * WARPS;

constexpr uint32_t FLAG_EXACT_CHECKS = SVJG_FLAG_EXACT_CHECKS;   // probe links whose overlap test fails too
constexpr uint32_t FLAG_FORCE_GENERAL = SVJG_FLAG_FORCE_GENERAL; // test hook: every multi-node line through general()
constexpr uint32_t COMMA_PATH = 0xFFFFFFFFu;

// token flags
constexpr uint32_t TF_PLUS = 1;     // delimiter in front is '>'
constexpr uint32_t TF_ALT = 2;      // chrom:pos.k (length from the GFA)
constexpr uint32_t TF_PLAIN = 4;    // exactly one ':', digits-digits or digits.<anything>, length > 0
constexpr uint32_t TF_OK = 8;       // overlap verdict of the link (previous node, this node)
// line flags
constexpr uint32_t LF_GENERAL = 1;  // must go through general()
constexpr uint32_t LF_HAS_NL = 2;   // the line ends in a newline (it counts in the hit's length)
constexpr uint32_t LF_SKIP = 4;     // reported as an error: no links

struct MLine {
    uint16_t s, e;        // window offsets of the line's first byte and of its end (newline or end of data)
    uint16_t p6;          // tab that ends the path column
    uint16_t tok0, ntok;  // tokens [tok0, tok0 + ntok); ntok == 0 marks a slot that was given up
    uint16_t flags;
    uint16_t ps;          // first byte of the path column
    uint16_t pad;
};
static_assert(sizeof(MLine) == 16, "MLine is 16 bytes");

struct FilterArgs {
    const uint8_t *gaf;
    uint64_t n;
    uint64_t base;
    int64_t d_over;
    DevTables tb;
    uint32_t *counts;
    uint32_t *hit_sv2, *hit_off, *hit_len;
    uint64_t hit_cap;
    unsigned long long *stats;   // svjg_filter_stats as 8 x u64
    uint32_t n_tiles;
    uint32_t flags;
};

struct Local {
    uint32_t n_rec = 0, n_multi = 0, n_generic = 0;
    uint64_t n_checks = 0;
};

struct SmemSrc {
    typedef uint32_t pos_t;
    const uint8_t *p;
    __device__ __forceinline__ uint32_t operator[](uint32_t i) const { return p[i]; }
};
struct GmemSrc {
    typedef uint64_t pos_t;
    const uint8_t *p;
    __device__ __forceinline__ uint32_t operator[](uint64_t i) const { return __ldg(p + i); }
};

__device__ __forceinline__ bool py_space(uint32_t c) {
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f);
}
__device__ __forceinline__ bool is_delim(uint32_t c) { return (c | 2u) == '>'; }   // '<' = 0x3C, '>' = 0x3E

// ---- SWAR byte classes: the answer is bit 7 of each byte, exact for all 256 byte values.
// (w & 0x7F..) ^ C is zero in its low 7 bits iff the byte's low 7 bits equal C; adding 0x7F carries
// into bit 7 unless they are zero; a byte with bit 7 set never matches (C < 0x80).  3 instructions.
#define SVJG_M7 0x7F7F7F7Fu
#define SVJG_H8 0x80808080u
__device__ __forceinline__ uint32_t lop_and_xor(uint32_t w, uint32_t m, uint32_t c) {      // (w & m) ^ c
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(r) : "r"(w), "r"(m), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop_nor_and(uint32_t t, uint32_t w, uint32_t h) {      // ~t & ~w & h
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(r) : "r"(t), "r"(w), "r"(h));
    return r;
}
__device__ __forceinline__ uint32_t lop_or_and(uint32_t t, uint32_t w, uint32_t h) {       // (t | w) & h
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xA8;" : "=r"(r) : "r"(t), "r"(w), "r"(h));
    return r;
}
__device__ __forceinline__ uint32_t eq_bytes(uint32_t w, uint32_t c4) {                    // c4: four copies of a byte < 0x80
    return lop_nor_and(lop_and_xor(w, SVJG_M7, c4) + SVJG_M7, w, SVJG_H8);
}
__device__ __forceinline__ uint32_t digit_bytes(uint32_t w) {                              // '0'..'9'
    return lop_nor_and(lop_and_xor(w, SVJG_M7, 0x30303030u) + 0x76767676u, w, SVJG_H8);
}
__device__ __forceinline__ uint32_t nondigit_bytes(uint32_t w) {
    return lop_or_and(lop_and_xor(w, SVJG_M7, 0x30303030u) + 0x76767676u, w, SVJG_H8);
}
__device__ __forceinline__ uint32_t delim_bytes(uint32_t w) {                              // '<' 0x3C or '>' 0x3E
    return lop_nor_and(lop_and_xor(w, 0x7D7D7D7Du, 0x3C3C3C3Cu) + SVJG_M7, w, SVJG_H8);
}
// flags of bytes [lo, hi) of a word (byte indices; any ints)
__device__ __forceinline__ uint32_t byte_range(int lo, int hi) {
    uint32_t m = 0xFFFFFFFFu;
    if (lo > 0) m = lo >= 4 ? 0u : (m << (8 * lo));
    if (hi < 4) m = hi <= 0 ? 0u : (m & (0xFFFFFFFFu >> (8 * (4 - hi))));
    return m;
}
// 16 class flags of a 16-byte chunk, bit i = byte i.  dp4a gathers the four bit-7 flags of a
// word: sum(0x80 * weight) = bits << 7.
template <class F>
__device__ __forceinline__ uint32_t mask16(const uint4 &v, F cls) {
    uint32_t lo = __dp4a(cls(v.x), 0x08040201u, 0u);
    lo = __dp4a(cls(v.y), 0x80402010u, lo);
    uint32_t hi = __dp4a(cls(v.z), 0x08040201u, 0u);
    hi = __dp4a(cls(v.w), 0x80402010u, hi);
    return (lo >> 7) | (hi << 1);
}
struct IsNewline {
    __device__ __forceinline__ uint32_t operator()(uint32_t w) const { return eq_bytes(w, 0x0A0A0A0Au); }
};
struct IsTab {
    __device__ __forceinline__ uint32_t operator()(uint32_t w) const { return eq_bytes(w, 0x09090909u); }
};

__device__ __noinline__ void report(const FilterArgs &a, uint32_t code, uint64_t line_off) {
    atomicCAS(a.stats + 4, 0ull, (unsigned long long)code);
    atomicMin(a.stats + 5, (unsigned long long)(a.base + line_off));
}

// Python int(): optional surrounding whitespace, optional sign, decimal digits.
// (Underscore separators and non-ASCII digits are not accepted: documented.)
template <class Src>
__device__ __noinline__ int parse_int(const Src &src, typename Src::pos_t b, typename Src::pos_t e, int64_t &out) {
    while (b < e && py_space(src[b])) ++b;
    while (e > b && py_space(src[e - 1])) --e;
    bool neg = false;
    if (b < e) {
        uint32_t c = src[b];
        if (c == '+' || c == '-') {
            neg = (c == '-');
            ++b;
        }
    }
    if (b >= e) return SVJG_BAD_INT;
    int64_t v = 0;
    int nd = 0;
    for (; b < e; ++b) {
        uint32_t c = src[b] - '0';
        if (c > 9) return SVJG_BAD_INT;
        if (v != 0 || c != 0) ++nd;
        if (nd > 18) return SVJG_BAD_RANGE;
        v = v * 10 + int64_t(c);
    }
    out = neg ? -v : v;
    return 0;
}

// ---------------------------------------------------------------------------
// Rec: one GAF line.  parse_fields()/general() are the exact, string-level
// routines; link()/probe()/emit() are shared with the token-parallel route.
// ---------------------------------------------------------------------------
template <class Src>
struct Rec {
    typedef typename Src::pos_t P;
    const FilterArgs &a;
    Src src;
    P ps, pe;          // path column
    bool angle;
    int64_t tlen, ts, te;
    uint32_t line_off, line_len;
    Local &loc;
    int err;

    __device__ Rec(const FilterArgs &a_, Src s, uint32_t off, uint32_t len, Local &l)
        : a(a_), src(s), ps(0), pe(0), angle(true), tlen(0), ts(0), te(0), line_off(off), line_len(len), loc(l), err(0) {}

    struct Tok {
        P b;
        uint32_t l;
    };

    // read_gaf_line (:184-198) on line.rstrip() (:126): 12 tab-separated columns,
    // 9 of them int().  Returns the number of path tokens for a '<'/'>' path,
    // COMMA_PATH for any other non-empty path.  Sets err where the reference raises.
    __device__ __noinline__ uint32_t parse_fields(P s, P e) {
        while (e > s && py_space(src[e - 1])) --e;
        P pos = s;
        int col = 0;
        int64_t alen = 1;
        uint32_t ntok = 0;
        for (;;) {
            P f = pos;
            int r = 0;
            if (col == 5) {
                // path column: find its end and count tokens on the way (extract_nodes :366-367)
                bool prev_delim = true;
                for (; f < e; ++f) {
                    uint32_t c = src[f];
                    if (c == '\t') break;
                    bool d = is_delim(c);
                    ntok += (prev_delim && !d);
                    prev_delim = d;
                }
                ps = pos;
                pe = f;
            } else {
                uint32_t nd = 0;
                bool plain = true;
                int64_t v = 0;
                for (; f < e; ++f) {
                    uint32_t c = src[f];
                    if (c == '\t') break;
                    uint32_t d = c - '0';
                    plain &= (d <= 9);
                    v = v * 10 + int64_t(d);
                    ++nd;
                }
                bool numeric = (col >= 1 && col <= 3) || (col >= 6 && col <= 11);
                if (numeric) {
                    const bool used = col >= 6 && col <= 8;         // Tlen, Ts, Te: the only values that matter
                    if (!plain || nd == 0 || (nd > 18 && used)) {
                        r = parse_int(src, pos, f, v);
                        // Python's int() has no size limit; only the used columns are held to 18 digits
                        if (r == SVJG_BAD_RANGE && !used) r = 0, v = 1;
                    } else if (nd > 18) {
                        v = 1;                                       // digits only: int() accepts; non-zero unless all '0'
                        bool all0 = true;
                        for (P j = pos; j < f; ++j) all0 &= src[j] == '0';
                        if (all0) v = 0;
                    }
                    if (col == 6) tlen = v;
                    else if (col == 7) ts = v;
                    else if (col == 8) te = v;
                    else if (col == 10) alen = v;
                }
            }
            if (r && !err) err = r;
            ++col;
            if (f >= e || col == 12) break;
            pos = f + 1;
        }
        if (col < 12) err = SVJG_BAD_COLUMNS;
        if (err) return 0;
        if (alen == 0) {                                        // Am / Alen (:196) unless "id:f:" in line (:193)
            bool found = false;
            for (P j = s; j + 5 <= e && !found; ++j)
                found = src[j] == 'i' && src[j + 1] == 'd' && src[j + 2] == ':' && src[j + 3] == 'f' && src[j + 4] == ':';
            if (!found) {
                err = SVJG_BAD_ALEN;
                return 0;
            }
        }
        if (ps == pe) {                                         // p[0] on an empty string (:366)
            err = SVJG_BAD_PATH;
            return 0;
        }
        angle = is_delim(src[ps]);
        return angle ? ntok : COMMA_PATH;
    }

    // Tlen, Ts, Te (columns 7-9) of a line whose columns were validated already
    __device__ void reparse_coords(P e) {
        while (e > pe && py_space(src[e - 1])) --e;
        P pos = pe + 1;
        int64_t *dst[3] = {&tlen, &ts, &te};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            P f = pos;
            while (f < e && src[f] != '\t') ++f;
            parse_int(src, pos, f, *dst[k]);
            pos = f + 1;
        }
        angle = true;
    }

    // tokens of the path column: '<'/'>' separated, or (path not starting with
    // one of those) ','-separated pieces minus their last character
    __device__ bool next_tok(P &cur, Tok &t) const {
        if (angle) {
            while (cur < pe && is_delim(src[cur])) ++cur;
            if (cur >= pe) return false;
            t.b = cur;
            while (cur < pe && !is_delim(src[cur])) ++cur;
            t.l = uint32_t(cur - t.b);
            return true;
        }
        while (cur <= pe) {
            P q = cur;
            while (q < pe && src[q] != ',') ++q;
            P pb = cur;
            uint32_t pl = uint32_t(q - cur);
            cur = q + 1;
            if (pl == 0) continue;
            t.b = pb;
            t.l = pl - 1;
            return true;
        }
        return false;
    }

    __device__ bool tok_eq(const Tok &x, const Tok &y) const {
        if (x.l != y.l) return false;
        if (x.b == y.b) return true;
        for (uint32_t i = 0; i < x.l; ++i)
            if (src[x.b + i] != src[y.b + i]) return false;
        return true;
    }

    __device__ uint64_t tok_hash(const Tok &t) const {
        TokHash h = tok_init();
        uint32_t w = 0;
        for (uint32_t i = 0; i < t.l; ++i) {
            w |= src[t.b + i] << (8 * (i & 3));
            if ((i & 3) == 3) {
                tok_step(h, w);
                w = 0;
            }
        }
        if (t.l & 3) tok_step(h, w);
        return tok_value(h, t.l);
    }

    // name bytes against the table's copy (4-byte aligned, zero padded).  In shared memory the
    // compare is word-wide and never exits early, so all table loads are in flight together.
    __device__ bool names_match(uint32_t off, const Tok &t) const {
        if constexpr (std::is_same<Src, SmemSrc>::value) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(a.tb.blob + off);
            const uint32_t sh = (uint32_t(t.b) & 3u) * 8u;
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(src.p + (uint32_t(t.b) & ~3u));
            const uint32_t nw = (t.l + 3u) >> 2;
            uint32_t cur = wp[0], diff = 0;
#pragma unroll 4
            for (uint32_t k = 0; k < nw; ++k) {
                uint32_t nxt = wp[k + 1];
                uint32_t w = __funnelshift_r(cur, nxt, sh);
                cur = nxt;
                if (k == nw - 1 && (t.l & 3u)) w &= (1u << (8u * (t.l & 3u))) - 1u;
                diff |= w ^ __ldg(q + k);
            }
            return diff == 0;
        } else {
            const uint8_t *q = a.tb.blob + off;
            for (uint32_t i = 0; i < t.l; ++i)
                if (__ldg(q + i) != src[t.b + i]) return false;
            return true;
        }
    }

    __device__ bool probe(uint64_t hl, uint32_t sl, const Tok &tl_, uint64_t hr, uint32_t sr, const Tok &tr_,
                          LinkSlot &out) const {
        uint64_t h = link_hash(hl, sl, hr, sr);
        uint32_t i = uint32_t(h) & a.tb.link_mask;
        for (;;) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(a.tb.links + i);
            uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
            uint32_t meta = hi.y;
            if (!(meta & 1u)) return false;
            uint64_t sh = (uint64_t(lo.y) << 32) | lo.x;
            if (sh == h && ((meta >> 2) & 1u) == sl && ((meta >> 1) & 1u) == sr && (lo.w & 0xFFFFu) == tl_.l &&
                (lo.w >> 16) == tr_.l && names_match(lo.z, tl_) && names_match(lo.z + ((tl_.l + 3u) & ~3u), tr_)) {
                out.hash = sh;
                out.name_off = lo.z;
                out.ent_begin = hi.x;
                out.meta = meta;
                out.ent0 = hi.z;
                return true;
            }
            i = (i + 1) & a.tb.link_mask;
        }
    }

    // alt_node_len[name] (:346); false when the name is not in the GFA
    __device__ bool alt_lookup(uint64_t tokh, const Tok &t, int64_t &len) const {
        uint64_t h = alt_hash(tokh);
        uint32_t i = uint32_t(h) & a.tb.alt_mask;
        for (;;) {
            const uint4 *sp = reinterpret_cast<const uint4 *>(a.tb.alts + i);
            uint4 lo = __ldg(sp), hi = __ldg(sp + 1);
            if (!hi.z) return false;
            uint64_t sh = (uint64_t(lo.y) << 32) | lo.x;
            if (sh == h && lo.w == t.l && names_match(lo.z, t)) {
                len = int64_t((uint64_t(hi.y) << 32) | hi.x);
                return true;
            }
            i = (i + 1) & a.tb.alt_mask;
        }
    }

    // Hits of the token-parallel route are staged in shared memory (flushed once per tile with one
    // cursor atomic for the whole block); everything else appends to the global arrays directly.
    uint32_t *stage_sv = nullptr;
    uint16_t *stage_ml = nullptr;
    uint32_t *stage_n = nullptr;
    uint32_t stage_line = 0;

    __device__ void emit(uint32_t sv2) {
        cg::coalesced_group active = cg::coalesced_threads();
        if (stage_sv) {
            uint32_t base = 0;
            if (active.thread_rank() == 0) base = atomicAdd(stage_n, active.size());
            base = active.shfl(base, 0) + active.thread_rank();
            if (base < HCAP) {
                stage_sv[base] = sv2;
                stage_ml[base] = uint16_t(stage_line);
                return;
            }
        }
        // warp-aggregated: one counter atomic per distinct SV allele among the
        // lanes that are here together, one cursor atomic for all of them
        cg::coalesced_group now = cg::coalesced_threads();
        cg::coalesced_group same = cg::labeled_partition(now, sv2);
        if (same.thread_rank() == 0) atomicAdd(a.counts + sv2, same.size());
        unsigned long long base = 0;
        if (now.thread_rank() == 0) base = atomicAdd(a.stats + 0, (unsigned long long)now.size());
        base = now.shfl(base, 0) + now.thread_rank();
        if (base < a.hit_cap) {
            a.hit_sv2[base] = sv2;
            a.hit_off[base] = line_off;
            a.hit_len[base] = line_len;
        }
    }

    // forward and reverse key of one link (:141-148) and their entries (:150-166).
    // ok_known: the overlap verdict is already there; otherwise it is computed
    // once, on the first key that has entries, by the general overlap().
    __device__ void link(const Tok &A, uint64_t hA, int sA, const Tok &B, uint64_t hB, int sB, bool ok_known, bool ok) {
#pragma unroll 1
        for (int dir = 0; dir < 2; ++dir) {
            LinkSlot s;
            bool hit = dir == 0 ? probe(hA, sA, A, hB, sB, B, s) : probe(hB, !sB, B, hA, !sA, A, s);
            if (!hit) continue;
            if (s.meta & 8u) {
                err = SVJG_BAD_ENTRY;
                return;
            }
            uint32_t cnt = s.meta >> 4;
            if (cnt == 0) continue;
            loc.n_checks += cnt;
            if (!ok_known) {
                ok = overlap(A, B);
                if (err) return;
                ok_known = true;
            }
            if (!ok) continue;
            for (uint32_t k = 0; k < cnt; ++k) {
                uint32_t sv2 = k == 0 ? s.ent0 : __ldg(a.tb.entries + s.ent_begin + k);
                if (sv2 == ENTRY_POISON) {
                    err = SVJG_BAD_ENTRY;
                    return;
                }
                emit(sv2);
            }
        }
    }

    // ---------------------------------------------------- general (exact) routine
    // strand of a token: the byte in front of the FIRST occurrence of its text
    // anywhere in the path (:206).  1 = '+'.
    __device__ int strand(const Tok &t) {
        if (t.l == 0) {
            err = SVJG_BAD_PATH;
            return 0;
        }
        uint32_t c0 = src[t.b];
        P j = ps;
        for (; j < t.b; ++j) {
            if (src[j] != c0) continue;
            uint32_t i = 1;
            while (i < t.l && src[j + i] == src[t.b + i]) ++i;
            if (i == t.l) break;
        }
        if (j == ps) {
            err = SVJG_BAD_PATH;
            return 0;
        }
        return src[j - 1] == '>';
    }

    // get_node_len (:343-349)
    __device__ int64_t node_len(const Tok &t) {
        P end = t.b + t.l, q = end;
        while (q > t.b && src[q - 1] != ':') --q;
        bool dot = false;
        for (P i = q; i < end; ++i) dot |= (src[i] == '.');
        if (dot) {
            int64_t len = 0;
            if (!alt_lookup(tok_hash(t), t, len)) err = SVJG_BAD_ALTNODE;
            return len;
        }
        P d1 = q;
        while (d1 < end && src[d1] != '-') ++d1;
        if (d1 >= end) {
            err = SVJG_BAD_NODENAME;
            return 0;
        }
        P d2 = d1 + 1;
        while (d2 < end && src[d2] != '-') ++d2;
        int64_t s0, e0;
        int r1 = parse_int(src, d1 + 1, d2, e0);
        int r0 = parse_int(src, q, d1, s0);
        if (r0 || r1) {
            err = (r0 == SVJG_BAD_RANGE || r1 == SVJG_BAD_RANGE) ? SVJG_BAD_RANGE : SVJG_BAD_NODENAME;
            return 0;
        }
        return e0 - s0 + 1;
    }

    // check_bkpt_overlap (:258-273): list.index() picks the FIRST equal node
    __device__ bool overlap(const Tok &A, const Tok &B) {
        P cur = ps;
        Tok t;
        int64_t left = 0, right = 0;
        bool found_l = false, found_r = false;
        while (next_tok(cur, t)) {
            bool is_b = !found_r && tok_eq(t, B);
            if (!found_l || found_r || is_b) {
                int64_t len = node_len(t);
                if (err) return false;
                if (!found_l) {
                    left += len;
                    if (tok_eq(t, A)) found_l = true;
                }
                if (is_b) found_r = true;
                if (found_r) right += len;
            }
        }
        return (left - ts >= a.d_over) && (right - (tlen - te - 1) >= a.d_over);
    }

    // extract_nodes / get_aln_links / lookups for any path (:130-166)
    __device__ __noinline__ void general() {
        uint32_t n = 0;
        {
            P cur = ps;
            Tok t;
            while (next_tok(cur, t)) ++n;
        }
        if (n < 2) return;                                       // :133
        if (!angle) loc.n_multi++;
        loc.n_generic++;
        P cur = ps;
        Tok A, B;
        next_tok(cur, A);
        uint64_t hA = tok_hash(A);
        int sA = strand(A);
        if (err) return;
        for (uint32_t i = 1; i < n; ++i) {
            next_tok(cur, B);
            uint64_t hB = tok_hash(B);
            int sB = strand(B);
            if (err) return;
            link(A, hA, sA, B, hB, sB, false, false);
            if (err) return;
            A = B;
            hA = hB;
            sA = sB;
        }
    }
};

// ---- TMA bulk copy + mbarrier helpers ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// the exact per-line route for a line that is inside the window
__device__ __noinline__ void slow_line(const FilterArgs &a, const uint8_t *win, uint32_t s, uint32_t e, uint32_t off,
                                       uint32_t len, Local &loc) {
    Rec<SmemSrc> rec(a, SmemSrc{win}, off, len, loc);
    uint32_t ntok = rec.parse_fields(s, e);
    if (!rec.err && ntok >= 2) {
        if (ntok != COMMA_PATH) loc.n_multi++;
        rec.general();
    }
    if (rec.err) report(a, rec.err, off);
}

// a line that runs past the staged window: exact route on global memory
__device__ __noinline__ void long_line(const FilterArgs &a, uint64_t from, uint32_t off, Local &loc) {
    uint64_t e = from;
    while (e < a.n && __ldg(a.gaf + e) != '\n') ++e;
    uint32_t len = uint32_t(e - off) + (e < a.n ? 1u : 0u);
    Rec<GmemSrc> rec(a, GmemSrc{a.gaf}, off, len, loc);
    uint32_t ntok = rec.parse_fields(uint64_t(off), e);
    if (!rec.err && ntok >= 2) {
        if (ntok != COMMA_PATH) loc.n_multi++;
        rec.general();
    }
    if (rec.err) report(a, rec.err, off);
}

// general() for a line whose columns are known to be valid (path [ps, pe), coordinates after pe)
__device__ __noinline__ void general_line(const FilterArgs &a, const uint8_t *win, uint32_t ps, uint32_t pe, uint32_t e,
                                          uint32_t off, uint32_t len, Local &loc) {
    Rec<SmemSrc> rec(a, SmemSrc{win}, off, len, loc);
    rec.ps = ps;
    rec.pe = pe;
    rec.reparse_coords(e);
    rec.general();
    if (rec.err) report(a, rec.err, off);
}

__device__ __forceinline__ uint32_t lds32(const uint8_t *win, uint32_t a) { return *reinterpret_cast<const uint32_t *>(win + a); }

// number of non-digit bytes in window bytes [lo, hi), hi > lo
__device__ __forceinline__ uint32_t count_nondigits(const uint8_t *win, uint32_t lo, uint32_t hi) {
    const uint32_t a0 = lo & ~3u, a1 = (hi - 1u) & ~3u;
    const uint32_t lom = 0xFFFFFFFFu << (8u * (lo & 3u)), him = 0xFFFFFFFFu >> (8u * (3u - ((hi - 1u) & 3u)));
    uint32_t f = nondigit_bytes(lds32(win, a0)) & lom;
    if (a0 == a1) return __popc(f & him);
    uint32_t n = __popc(f);
    for (uint32_t a = a0 + 4; a < a1; a += 4) n += __popc(nondigit_bytes(lds32(win, a)));
    return n + __popc(nondigit_bytes(lds32(win, a1)) & him);
}

// any ',' in window bytes [lo, hi), hi > lo
__device__ __forceinline__ bool has_comma(const uint8_t *win, uint32_t lo, uint32_t hi) {
    const uint32_t a0 = lo & ~3u, a1 = (hi - 1u) & ~3u;
    const uint32_t lom = 0xFFFFFFFFu << (8u * (lo & 3u)), him = 0xFFFFFFFFu >> (8u * (3u - ((hi - 1u) & 3u)));
    uint32_t f = eq_bytes(lds32(win, a0), 0x2C2C2C2Cu) & lom;
    if (a0 == a1) return (f & him) != 0;
    for (uint32_t a = a0 + 4; a < a1; a += 4) f |= eq_bytes(lds32(win, a), 0x2C2C2C2Cu);
    return (f | (eq_bytes(lds32(win, a1), 0x2C2C2C2Cu) & him)) != 0;
}

// Writes a warp's staged hits: coalesced tuple stores and warp-aggregated counter atomics
// (one per distinct SV allele among the 32 hits the warp holds at a time).
__device__ __forceinline__ void flush_hits(const FilterArgs &a, const uint32_t *h_sv, const uint16_t *h_ml, const MLine *ml,
                                           uint32_t n, uint64_t tile_start, unsigned long long base, int lane) {
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool act = i < n;
        const uint32_t sv2 = act ? h_sv[i] : 0xFFFFFFFFu;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, sv2);
        if (act) {
            if (lane == __ffs(peers) - 1) atomicAdd(a.counts + sv2, uint32_t(__popc(peers)));
            const MLine L = ml[h_ml[i]];
            const unsigned long long k = base + i;
            if (k < a.hit_cap) {
                a.hit_sv2[k] = sv2;
                a.hit_off[k] = uint32_t(tile_start) + (uint32_t(L.s) - HEAD);
                a.hit_len[k] = uint32_t(L.e) - L.s + ((L.flags & LF_HAS_NL) ? 1u : 0u);
            }
        }
    }
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

__global__ void __launch_bounds__(THREADS, 4) filter_kernel(const FilterArgs a) {
    extern __shared__ __align__(128) uint8_t smem_all[];
    __shared__ __align__(8) uint64_t mbars[WARPS];
    __shared__ uint32_t s_hits[WARPS];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint8_t *smem = smem_all + warp * WARP_SMEM;
    uint8_t *win = smem + OFF_WIN;
    uint16_t *nl = reinterpret_cast<uint16_t *>(smem + OFF_NL);
    MLine *ml = reinterpret_cast<MLine *>(smem + OFF_ML);
    uint64_t *t_hash = reinterpret_cast<uint64_t *>(smem + OFF_TH);
    int32_t *t_len = reinterpret_cast<int32_t *>(smem + OFF_TLEN);
    uint32_t *t_sval = reinterpret_cast<uint32_t *>(smem + OFF_TS);
    uint16_t *t_b = reinterpret_cast<uint16_t *>(smem + OFF_TB);
    uint16_t *t_l = reinterpret_cast<uint16_t *>(smem + OFF_TL);
    uint16_t *t_line = reinterpret_cast<uint16_t *>(smem + OFF_TLINE);
    uint8_t *t_flags = smem + OFF_TF;
    uint32_t *h_sv = reinterpret_cast<uint32_t *>(smem + OFF_HSV);
    uint16_t *h_ml = reinterpret_cast<uint16_t *>(smem + OFF_HML);
    uint64_t *mbar = &mbars[warp];
    uint32_t *s_nhits = &s_hits[warp];

    if (lane == 0) {
        mbar_init(mbar, 1);
        *s_nhits = 0;
    }
    __syncwarp();
    uint32_t phase = 0;
    Local loc;
    const bool all_links = a.flags & FLAG_EXACT_CHECKS;
    const uint32_t stop_after = (a.flags >> 8) & 7u;     // profiling hook (SVJG_STOP_AFTER): 1 = A, 2 = B, 3 = C, 4 = D
    uint64_t prev_tile_start = 0;
    const uint32_t n_workers = gridDim.x * WARPS;

    for (uint32_t tile = blockIdx.x * WARPS + warp; tile < a.n_tiles; tile += n_workers) {
        const uint64_t tile_start = uint64_t(tile) * TILE;
        const uint64_t g0 = tile_start ? tile_start - HEAD : 0;
        const uint32_t dst0 = tile_start ? 0 : HEAD;
        uint64_t g1 = tile_start + TILE + LOOKAHEAD;
        if (g1 > a.n) g1 = a.n;
        const bool at_eof = (g1 == a.n);
        const uint32_t nbytes = uint32_t(g1 - g0);
        const uint32_t bulk = nbytes & ~15u;
        const uint32_t valid_end = dst0 + nbytes;

        // the previous tile's hits go out while this tile's bytes come in
        const uint32_t n_staged = min(*s_nhits, uint32_t(HCAP));
        unsigned long long hbase = 0;
        if (lane == 0) {
            if (bulk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbar, bulk);
                bulk_g2s(win + dst0, a.gaf + g0, bulk, mbar);
            }
            if (n_staged) hbase = atomicAdd(a.stats + 0, (unsigned long long)n_staged);
        }
        if (n_staged) {
            hbase = __shfl_sync(0xFFFFFFFFu, hbase, 0);
            flush_hits(a, h_sv, h_ml, ml, n_staged, prev_tile_start, hbase, lane);
        }
        prev_tile_start = tile_start;
        __syncwarp();
        if (lane == 0) *s_nhits = 0;
        for (uint32_t i = bulk + lane; i < nbytes; i += 32) win[dst0 + i] = __ldg(a.gaf + g0 + i);
        if (dst0) win[lane] = lane == HEAD - 1 ? '\n' : 0;               // "newline" in front of byte 0 of the file
        for (uint32_t i = valid_end + lane; i < WIN + SPARE; i += 32) win[i] = 0;
        if (bulk) {
            mbar_wait(mbar, phase);
            phase ^= 1;
        }
        __syncwarp();

        // ---- phase A: ordered list of the newline positions from HEAD-1 on.  A lane takes two
        // adjacent 16-byte chunks; the scan stops behind the tile once the last owned line has its end.
        const uint32_t own_end = min(uint32_t(HEAD + TILE), valid_end);   // lines starting before own_end are ours
        uint32_t n_nl = 0, n_own = 0;
        for (int c0 = 0; c0 < NPAIRS; c0 += 32) {
            if (uint32_t(c0) * 32u >= own_end && n_nl > n_own) break;
            const int c = c0 + lane;
            const uint32_t p0 = uint32_t(c) * 32u;
            uint32_t m = 0;
            if (c < NPAIRS) {
                const uint4 v0 = *reinterpret_cast<const uint4 *>(win + p0);
                const uint4 v1 = *reinterpret_cast<const uint4 *>(win + p0 + 16);
                m = mask16(v0, IsNewline()) | (mask16(v1, IsNewline()) << 16);
                if (c == 0) m &= 0x80000000u;                             // positions before HEAD-1 are not ours to see
            }
            const uint32_t cnt = __popc(m);
            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, cnt != 0);
            if (!bal) continue;
            uint32_t q, total;
            if (__ballot_sync(0xFFFFFFFFu, cnt > 1) == 0) {
                q = __popc(bal & lt_mask);
                total = __popc(bal);
            } else {                                                      // two newlines within 32 bytes (rare)
                const uint32_t incl = warp_incl_scan(cnt, lane);
                q = incl - cnt;
                total = __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
            // newline at p starts an owned line iff p + 1 < own_end
            const int hi = int(own_end) - 1 - int(p0);
            const uint32_t mo = hi >= 32 ? m : (hi <= 0 ? 0u : (m & ((1u << hi) - 1u)));
            n_own += __reduce_add_sync(0xFFFFFFFFu, uint32_t(__popc(mo)));
            uint32_t idx = n_nl + q;
            while (m) {
                if (idx < NLCAP) nl[idx] = uint16_t(p0 + uint32_t(__ffs(m) - 1));
                ++idx;
                m &= m - 1;
            }
            n_nl += total;
        }
        __syncwarp();

        uint32_t n_ml = 0, n_tok = 0;
        if (stop_after == 1) continue;
        if (n_nl > NLCAP) {
            // that many lines in 6 KiB: some line is shorter than 16 bytes and cannot hold 12 columns
            if (lane == 0) report(a, SVJG_BAD_SHORTLINE, tile_start);
        } else {
            // ---- phase B: one lane per line
            bool route_full = false;
            for (uint32_t k0 = 0; k0 < n_own; k0 += 32) {
                const uint32_t k = k0 + lane;
                uint32_t want = 0;                 // path nodes of a line that asks for the token-parallel route
                uint32_t s = 0, e = 0, ps = 0, pe = 0, off = 0, len = 0;
                bool has_nl = false;
                if (k < n_own) {
                    s = uint32_t(nl[k]) + 1u;
                    off = uint32_t(tile_start) + (s - HEAD);
                    loc.n_rec++;
                    bool in_win = true;
                    if (k + 1 < n_nl) {
                        e = nl[k + 1];
                        has_nl = true;
                    } else if (at_eof) {
                        e = valid_end;
                    } else {
                        in_win = false;
                        long_line(a, tile_start + TILE + LOOKAHEAD, off, loc);
                    }
                    if (in_win) {
                        len = e - s + (has_nl ? 1u : 0u);
                        // columns 1-5: tab bitmap of the first HEAD_SPAN bytes (fixed trip count: lanes stay together)
                        bool plain = e > s && !py_space(win[e - 1]);
                        uint32_t p1, p2, p3, p4, p5;
                        {
                            const uint32_t base = s & ~15u;
                            uint32_t tm[HEAD_SPAN / 32];
#pragma unroll
                            for (int j = 0; j < HEAD_SPAN / 32; ++j) {
                                const uint4 v0 = *reinterpret_cast<const uint4 *>(win + base + 32 * j);
                                const uint4 v1 = *reinterpret_cast<const uint4 *>(win + base + 32 * j + 16);
                                tm[j] = mask16(v0, IsTab()) | (mask16(v1, IsTab()) << 16);
                            }
                            tm[0] &= 0xFFFFFFFFu << (s - base);
                            const uint32_t rel_e = e - base;
#pragma unroll
                            for (int j = 0; j < HEAD_SPAN / 32; ++j)
                                if (rel_e < 32u * (j + 1)) tm[j] = rel_e <= 32u * j ? 0u : (tm[j] & ((1u << (rel_e - 32u * j)) - 1u));
                            // the first 5 tab positions (relative to base), one byte each, newest in the low byte
                            uint32_t r0 = 0, r1 = 0, nt = 0;
#pragma unroll
                            for (int j = 0; j < HEAD_SPAN / 32; ++j) {
                                uint32_t m = tm[j];
                                while (m && nt < 5) {
                                    const uint32_t pos = 32u * j + uint32_t(__ffs(m) - 1);
                                    m &= m - 1;
                                    r1 = __funnelshift_l(r0, r1, 8);
                                    r0 = (r0 << 8) | pos;
                                    ++nt;
                                }
                            }
                            plain &= nt == 5;
                            p1 = base + (r1 & 255u), p2 = base + (r0 >> 24), p3 = base + ((r0 >> 16) & 255u);
                            p4 = base + ((r0 >> 8) & 255u), p5 = base + (r0 & 255u);
                        }
                        // column 6, the path [ps, pe): walk to its tab; count token starts (a non-delimiter byte
                        // right after a delimiter) and look for ',' on the way
                        ps = p5 + 1;
                        pe = e;
                        uint32_t ntok = 0, comma = 0;
                        if (plain) {
                            uint32_t carry = 0;   // delimiter flag of the byte in front of the word, at bit 7
                            uint32_t keep = 0xFFFFFFFFu << (8u * (ps & 3u));
                            for (uint32_t w0 = ps & ~3u; w0 < e; w0 += 4) {
                                const uint32_t w = lds32(win, w0);
                                const uint32_t tb = eq_bytes(w, 0x09090909u) & keep;
                                const uint32_t d = delim_bytes(w);
                                if (tb) {
                                    const uint32_t j = uint32_t(__ffs(tb) - 1) >> 3;          // byte of the tab in this word
                                    pe = w0 + j;
                                    keep &= j ? (0xFFFFFFFFu >> (8u * (4u - j))) : 0u;
                                }
                                ntok += __popc(((d << 8) | carry) & ~d & SVJG_H8 & keep);
                                comma |= eq_bytes(w, 0x2C2C2C2Cu) & keep;
                                if (tb) break;
                                carry = d >> 24;
                                keep = 0xFFFFFFFFu;
                            }
                            plain = pe < e && pe > ps;
                        }
                        // columns 7-12: tab bitmap of the TAIL_SPAN bytes around the end of the path
                        if (plain) {
                            const uint32_t p6 = pe;
                            const uint32_t base = p6 & ~15u;
                            uint32_t tm[TAIL_SPAN / 32];
#pragma unroll
                            for (int j = 0; j < TAIL_SPAN / 32; ++j) {
                                const uint4 v0 = *reinterpret_cast<const uint4 *>(win + base + 32 * j);
                                const uint4 v1 = *reinterpret_cast<const uint4 *>(win + base + 32 * j + 16);
                                tm[j] = mask16(v0, IsTab()) | (mask16(v1, IsTab()) << 16);
                            }
                            tm[0] &= 0xFFFFFFFEu << (p6 - base);                               // tabs after p6
                            const uint32_t rel_e = e - base;
#pragma unroll
                            for (int j = 0; j < TAIL_SPAN / 32; ++j)
                                if (rel_e < 32u * (j + 1)) tm[j] = rel_e <= 32u * j ? 0u : (tm[j] & ((1u << (rel_e - 32u * j)) - 1u));
                            uint32_t r0 = 0, r1 = 0, nt = 0;
#pragma unroll
                            for (int j = 0; j < TAIL_SPAN / 32; ++j) {
                                uint32_t m = tm[j];
                                while (m && nt < 6) {
                                    const uint32_t pos = 32u * j + uint32_t(__ffs(m) - 1);
                                    m &= m - 1;
                                    r1 = __funnelshift_l(r0, r1, 8);
                                    r0 = (r0 << 8) | pos;
                                    ++nt;
                                }
                            }
                            if (nt == 5 && rel_e <= TAIL_SPAN) {                               // no tab after column 12
                                r1 = __funnelshift_l(r0, r1, 8);
                                r0 = (r0 << 8) | rel_e;
                                ++nt;
                            }
                            plain = nt == 6;
                            const uint32_t p7 = base + ((r1 >> 8) & 255u), p8 = base + (r1 & 255u), p9 = base + (r0 >> 24),
                                           p10 = base + ((r0 >> 16) & 255u), p11 = base + ((r0 >> 8) & 255u),
                                           p12 = base + (r0 & 255u);
                            if (plain) {
                                // no empty integer column, Alen not zero, digits only in columns 2-4 and 7-12
                                const bool w_ok = (p2 - p1 > 1u) & (p3 - p2 > 1u) & (p4 - p3 > 1u) & (p7 - p6 > 1u) &
                                                  (p8 - p7 > 1u) & (p9 - p8 > 1u) & (p10 - p9 > 1u) & (p11 - p10 > 1u) &
                                                  (p12 - p11 > 1u);
                                plain = w_ok && win[p10 + 1] != '0';
                                if (plain)
                                    plain = count_nondigits(win, p1 + 1, p4) == 2u && count_nondigits(win, p6 + 1, p12) == 5u;
                            }
                        }
                        if (!plain) {
                            slow_line(a, win, s, e, off, len, loc);
                        } else if (!is_delim(win[ps])) {
                            // bare name or GFA-style a+,b+ (extract_nodes :369-373): one piece without ',' is one node
                            if (comma) slow_line(a, win, s, e, off, len, loc);
                        } else if (ntok >= 2) {                                                // :133
                            loc.n_multi++;
                            if ((a.flags & FLAG_FORCE_GENERAL) || ntok > TCAP) general_line(a, win, ps, pe, e, off, len, loc);
                            else want = ntok;
                        }
                    }
                }
                __syncwarp();
                // slots on the token-parallel route for the lanes that want them (in lane order, no holes)
                const uint32_t wb = __ballot_sync(0xFFFFFFFFu, want != 0);
                if (!wb) continue;
                const uint32_t incl = warp_incl_scan(want, lane);
                const uint32_t t0 = n_tok + incl - want;
                const uint32_t li = n_ml + __popc(wb & lt_mask);
                const bool fits = want && !route_full && t0 + want <= TCAP && li < MCAP;
                const uint32_t nofit = __ballot_sync(0xFFFFFFFFu, want && !fits);
                if (nofit) {                       // everything from the first lane without room takes the exact route
                    const int first = __ffs(nofit) - 1;
                    n_tok = __shfl_sync(0xFFFFFFFFu, t0, first);
                    n_ml = __shfl_sync(0xFFFFFFFFu, li, first);
                    route_full = true;
                } else {
                    n_tok += __shfl_sync(0xFFFFFFFFu, incl, 31);
                    n_ml += __popc(wb);
                }
                if (want && !fits) {
                    general_line(a, win, ps, pe, e, off, len, loc);
                } else if (want) {
                    MLine L;
                    L.s = uint16_t(s);
                    L.e = uint16_t(e);
                    L.p6 = uint16_t(pe);
                    L.tok0 = uint16_t(t0);
                    L.ntok = uint16_t(want);
                    L.flags = has_nl ? LF_HAS_NL : 0;
                    L.ps = uint16_t(ps);
                    L.pad = 0;
                    ml[li] = L;
                    // token records: maximal runs of non-delimiter bytes, from the start / end flags
                    const uint32_t a0 = ps & ~3u, a1 = (pe - 1u) & ~3u;
                    const uint32_t lom = 0xFFFFFFFFu << (8u * (ps & 3u)), him = 0xFFFFFFFFu >> (8u * (3u - ((pe - 1u) & 3u)));
                    uint32_t t = t0, cur = 0xFFFFFFFFu, carry = 0;
                    for (uint32_t w0 = a0; w0 <= a1; w0 += 4) {
                        const uint32_t d = delim_bytes(lds32(win, w0));
                        const uint32_t pd = (d << 8) | carry;                       // "previous byte is a delimiter"
                        uint32_t st = pd & ~d & SVJG_H8, en = d & ~pd & SVJG_H8;
                        if (w0 == a0) st &= lom, en &= lom;
                        if (w0 == a1) st &= him, en &= him;
                        carry = d >> 24;
                        uint32_t ev = st | en;
                        while (ev) {
                            const uint32_t bit = uint32_t(__ffs(ev) - 1);
                            ev &= ev - 1;
                            const uint32_t pos = w0 + (bit >> 3);
                            if ((st >> bit) & 1u) {
                                cur = pos;
                                t_b[t] = uint16_t(pos);
                                t_line[t] = uint16_t(li);
                            } else if (cur != 0xFFFFFFFFu) {
                                t_l[t++] = uint16_t(pos - cur);
                                cur = 0xFFFFFFFFu;
                            }
                        }
                    }
                    if (cur != 0xFFFFFFFFu) t_l[t++] = uint16_t(pe - cur);
                }
            }
            __syncwarp();
            if (stop_after == 2) continue;

            // ---- phase C: one lane per path node
            for (uint32_t t = lane; t < n_tok; t += 32) {
                const uint32_t b = t_b[t], l = t_l[t];
                // name hash, 4 bytes a step, and the colons on the way
                const uint32_t al = b & ~3u, sh = (b & 3u) * 8u;
                const uint32_t *wp = reinterpret_cast<const uint32_t *>(win + al);
                uint32_t curw = wp[0];
                TokHash h = tok_init();
                uint32_t ncolon = 0, cpos = 0;
                for (uint32_t i = 0; i < l; i += 4) {
                    uint32_t nxt = wp[(i >> 2) + 1];
                    uint32_t w = __funnelshift_r(curw, nxt, sh);
                    curw = nxt;
                    uint32_t rem = l - i;
                    if (rem < 4) w &= (1u << (8 * rem)) - 1u;
                    tok_step(h, w);
                    uint32_t f = eq_bytes(w, 0x3A3A3A3Au);
                    if (f) {
                        ncolon += __popc(f);
                        cpos = i + ((31 - __clz(f)) >> 3);
                    }
                }
                const uint64_t hv = tok_value(h, l);
                // chrom:start-end  or  chrom:pos.<anything>
                uint32_t fl = win[b - 1] == '>' ? TF_PLUS : 0;
                uint32_t sval = 0;
                int64_t nlen = 0;
                if (ncolon == 1) {
                    uint32_t q = b + cpos + 1, end = b + l;
                    uint32_t v0 = 0, nd0 = 0;
                    for (; q < end; ++q) {
                        uint32_t d = uint32_t(win[q]) - '0';
                        if (d > 9) break;
                        v0 = v0 * 10 + d;
                        ++nd0;
                    }
                    if (nd0 >= 1 && nd0 <= 9 && q < end) {
                        uint32_t c = win[q];
                        if (c == '.') {
                            Rec<SmemSrc> rec(a, SmemSrc{win}, 0, 0, loc);
                            Rec<SmemSrc>::Tok tk{b, l};
                            if (rec.alt_lookup(hv, tk, nlen) && nlen > 0 && nlen <= 0x7FFFFFFF) fl |= TF_ALT | TF_PLAIN;
                        } else if (c == '-') {
                            uint32_t v1 = 0, nd1 = 0;
                            for (++q; q < end; ++q) {
                                uint32_t d = uint32_t(win[q]) - '0';
                                if (d > 9) break;
                                v1 = v1 * 10 + d;
                                ++nd1;
                            }
                            nlen = int64_t(v1) - int64_t(v0) + 1;
                            if (q == end && nd1 >= 1 && nd1 <= 9 && nlen > 0) fl |= TF_PLAIN;
                        }
                    }
                    sval = v0;
                }
                t_hash[t] = hv;
                t_len[t] = int32_t(nlen);
                t_sval[t] = sval;
                t_flags[t] = uint8_t(fl);
            }
            __syncwarp();
            if (stop_after == 3) continue;

            // ---- phase D: per multi-node line, coordinates + running sums + overlap verdicts;
            //      per node, "could this name occur inside an earlier one" (first-occurrence rules :206, :269-271)
            for (uint32_t k = lane; k < n_ml; k += 32) {
                const MLine L = ml[k];
                // Tlen, Ts, Te: digit-only columns 7-9 (validated in phase B)
                int64_t v[3];
                uint32_t q = uint32_t(L.p6) + 1, too_long = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    int64_t x = 0;
                    uint32_t nd = 0;
                    for (;; ++q) {
                        uint32_t d = uint32_t(win[q]) - '0';
                        if (d > 9) break;
                        nd += (x != 0 || d != 0);
                        x = x * 10 + int64_t(d);
                    }
                    too_long |= nd > 18;
                    v[j] = x;
                    ++q;
                }
                if (too_long) {
                    // the reference has bigints; this implementation stops at 18 digits and says so
                    report(a, SVJG_BAD_RANGE, uint32_t(tile_start) + (uint32_t(L.s) - HEAD));
                    ml[k].flags = uint16_t(L.flags | LF_SKIP);
                    continue;
                }
                const int64_t ts = v[1], tail = v[0] - v[2] - 1;
                uint32_t bad = 0;
                int64_t total = 0;
                for (uint32_t t = L.tok0; t < uint32_t(L.tok0) + L.ntok; ++t) {
                    bad |= !(t_flags[t] & TF_PLAIN);
                    total += t_len[t];
                }
                int64_t pre = t_len[L.tok0];
                for (uint32_t t = uint32_t(L.tok0) + 1; t < uint32_t(L.tok0) + L.ntok; ++t) {
                    const bool ok = (pre - ts >= a.d_over) && (total - pre - tail >= a.d_over);   // :269-273
                    if (ok) t_flags[t] |= TF_OK;
                    pre += t_len[t];
                }
                if (bad) ml[k].flags = uint16_t(L.flags | LF_GENERAL);
            }
            __syncwarp();
            for (uint32_t t = lane; t < n_tok; t += 32) {
                const uint32_t li = t_line[t];
                const uint32_t t0 = ml[li].tok0;
                const uint32_t sv = t_sval[t], kind = t_flags[t] & TF_ALT;
                bool clash = false;
                for (uint32_t j = t0; j < t; ++j) clash |= (t_sval[j] == sv) && ((t_flags[j] & TF_ALT) == kind);
                if (clash) ml[li].flags |= LF_GENERAL;   // benign race: every writer sets the same bit
            }
            __syncwarp();
            if (stop_after == 4) continue;

            // ---- phase E: one lane per link (node t with its predecessor), both keys
            for (uint32_t t = lane; t < n_tok; t += 32) {
                const uint32_t li = t_line[t];
                const MLine L = ml[li];
                if (t == L.tok0 || (L.flags & (LF_GENERAL | LF_SKIP))) continue;
                const uint32_t fb = t_flags[t];
                const bool ok = fb & TF_OK;
                if (!ok && !all_links) continue;
                const uint32_t off = uint32_t(tile_start) + (uint32_t(L.s) - HEAD);
                const uint32_t len = uint32_t(L.e) - L.s + ((L.flags & LF_HAS_NL) ? 1u : 0u);
                Rec<SmemSrc> rec(a, SmemSrc{win}, off, len, loc);
                rec.stage_sv = h_sv;
                rec.stage_ml = h_ml;
                rec.stage_n = s_nhits;
                rec.stage_line = li;
                Rec<SmemSrc>::Tok A{t_b[t - 1], t_l[t - 1]}, B{t_b[t], t_l[t]};
                rec.link(A, t_hash[t - 1], t_flags[t - 1] & TF_PLUS, B, t_hash[t], fb & TF_PLUS, true, ok);
                if (rec.err) report(a, rec.err, off);
            }
            for (uint32_t k = lane; k < n_ml; k += 32) {
                const MLine L = ml[k];
                if (!(L.flags & LF_GENERAL) || (L.flags & LF_SKIP)) continue;
                const uint32_t off = uint32_t(tile_start) + (uint32_t(L.s) - HEAD);
                const uint32_t len = uint32_t(L.e) - L.s + ((L.flags & LF_HAS_NL) ? 1u : 0u);
                general_line(a, win, L.ps, L.p6, L.e, off, len, loc);
            }
        }
        __syncwarp();
    }
    // hits of the last tile
    {
        const uint32_t n_staged = min(*s_nhits, uint32_t(HCAP));
        if (n_staged) {
            unsigned long long hbase = 0;
            if (lane == 0) hbase = atomicAdd(a.stats + 0, (unsigned long long)n_staged);
            hbase = __shfl_sync(0xFFFFFFFFu, hbase, 0);
            flush_hits(a, h_sv, h_ml, ml, n_staged, prev_tile_start, hbase, lane);
        }
    }

    // ---- per-warp statistics
    uint64_t v1 = loc.n_rec, v2 = loc.n_multi, v3 = loc.n_checks, v6 = loc.n_generic;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        v1 += __shfl_xor_sync(0xFFFFFFFFu, v1, d);
        v2 += __shfl_xor_sync(0xFFFFFFFFu, v2, d);
        v3 += __shfl_xor_sync(0xFFFFFFFFu, v3, d);
        v6 += __shfl_xor_sync(0xFFFFFFFFu, v6, d);
    }
    if (lane == 0) {
        if (v1) atomicAdd(a.stats + 1, (unsigned long long)v1);
        if (v2) atomicAdd(a.stats + 2, (unsigned long long)v2);
        if (v3) atomicAdd(a.stats + 3, (unsigned long long)v3);
        if (v6) atomicAdd(a.stats + 6, (unsigned long long)v6);
    }
}

__global__ void reset_kernel(uint32_t *counts, uint64_t n, unsigned long long *stats) {
    uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
    uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
    for (uint64_t k = i; k < n; k += stride) counts[k] = 0;
    if (i < 8) stats[i] = (i == 5) ? ~0ull : 0ull;
}

int g_grid_cap = 0;   // blocks resident at once (SM count x occupancy), per process

}  // namespace

extern "C" int svjg_filter_reset(uint32_t *d_counts, uint32_t num_sv, svjg_filter_stats *d_stats, void *stream) {
    if (!d_counts || !d_stats) return set_error(SVJG_E_ARG, "svjg_filter_reset: NULL argument");
    uint64_t n = uint64_t(num_sv) * 2;
    int blocks = int(std::min<uint64_t>(1184, (n + 255) / 256));
    if (blocks < 1) blocks = 1;
    reset_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_counts, n, reinterpret_cast<unsigned long long *>(d_stats));
    SVJG_CUDA(cudaGetLastError());
    return SVJG_OK;
}

extern "C" int svjg_filter_device(const svjg_tables *t, const uint8_t *d_gaf, uint64_t n_bytes, uint64_t base_offset,
                                  int64_t d_over, uint32_t *d_counts, uint32_t *d_hit_sv2, uint32_t *d_hit_off,
                                  uint32_t *d_hit_len, uint64_t hit_cap, svjg_filter_stats *d_stats, void *stream) {
    if (!t || t->device < 0) return set_error(SVJG_E_ARG, "svjg_filter_device: tables are not on a device");
    if (!d_counts || !d_stats || (n_bytes && !d_gaf)) return set_error(SVJG_E_ARG, "svjg_filter_device: NULL argument");
    if (hit_cap && (!d_hit_sv2 || !d_hit_off || !d_hit_len)) return set_error(SVJG_E_ARG, "svjg_filter_device: NULL hit buffer");
    if (reinterpret_cast<uintptr_t>(d_gaf) & 15) return set_error(SVJG_E_ARG, "svjg_filter_device: d_gaf must be 16-byte aligned");
    if (n_bytes >= 0xFFFF0000ull) return set_error(SVJG_E_ARG, "svjg_filter_device: shard must be smaller than 4 GiB");
    if (n_bytes == 0) return SVJG_OK;
    static bool configured = false;
    if (!configured) {
        SVJG_CUDA(cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        int dev = 0, sms = 0, occ = 0;
        SVJG_CUDA(cudaGetDevice(&dev));
        SVJG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SVJG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_kernel, THREADS, SMEM_BYTES));
        if (occ < 1) return set_error(SVJG_E_CUDA, "filter kernel does not fit on an SM");
        g_grid_cap = sms * occ;
        configured = true;
    }
    FilterArgs a;
    a.gaf = d_gaf;
    a.n = n_bytes;
    a.base = base_offset;
    a.d_over = d_over;
    a.tb = t->dev;
    a.counts = d_counts;
    a.hit_sv2 = d_hit_sv2;
    a.hit_off = d_hit_off;
    a.hit_len = d_hit_len;
    a.hit_cap = hit_cap;
    a.stats = reinterpret_cast<unsigned long long *>(d_stats);
    a.n_tiles = uint32_t((n_bytes + TILE - 1) / TILE);
    a.flags = t->filter_flags;
    static const char *stop_env = getenv("SVJG_STOP_AFTER");   // profiling hook: run the phases up to A/B/C/D only
    if (stop_env && stop_env[0] >= 'A' && stop_env[0] <= 'D') a.flags |= uint32_t(stop_env[0] - 'A' + 1) << 8;
    int grid = int(std::min<uint32_t>((a.n_tiles + WARPS - 1) / WARPS, uint32_t(g_grid_cap)));
    filter_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
    SVJG_CUDA(cudaGetLastError());
    return SVJG_OK;
}
