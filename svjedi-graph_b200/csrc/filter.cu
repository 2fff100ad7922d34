// Kernels 1-3 of the hot path, fused: GAF ingest (newline scan + field split),
// alignment -> allele resolution through the device link hash, and the per-SV
// counter / hit-tuple reduction.  One pass over the GAF bytes.
//
// Reference semantics restated (filter-alignments.py): per-line loop :123-166,
// read_gaf_line :184-198, extract_nodes :351-373, get_aln_links :200-219,
// reverse_link :221-225, check_bkpt_overlap :258-273, get_node_len :343-349.
//
// Layout: a persistent grid walks 32 KiB tiles of the byte buffer.  A tile plus
// 8 KiB of look-ahead (and the 16 bytes in front of it) is staged into shared
// memory by one TMA bulk copy (cp.async.bulk + mbarrier).  Per tile:
//   A  the CTA finds every newline with 16-byte SWAR loads and an ordered block
//      scan (line starts in shared memory);
//   B  one thread per line splits the 12 columns, validates the integers and
//      counts the path tokens — uniform work, warps stay converged; lines with
//      >= 2 path nodes are compacted into a shared-memory queue;
//   C  one thread per queued line resolves its links.  The fast path streams the
//      path twice (lengths/total, then prefix sums + hash probes) and needs no
//      per-token storage; it hands the line to the general routine whenever a
//      token could be a substring of an earlier one (the first-occurrence rule
//      of :206 and list.index() of :269-271 only differ from "own position"
//      in that case), or any name is not of the plain chrom:start-end form.
// A line belongs to the tile its first byte is in; a line running past the
// window is handled by the general routine reading global memory.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>

#include "svjg_internal.h"

namespace cg = cooperative_groups;
using namespace svjg;

namespace {

constexpr int TILE = 32768;
constexpr int LOOKAHEAD = 8192;
constexpr int HEAD = 16;
constexpr int WIN = HEAD + TILE + LOOKAHEAD;   // 40976 = 16 * 2561
constexpr int THREADS = 256;
constexpr int NWARPS = THREADS / 32;
constexpr int CHUNKS = WIN / 16;
constexpr int ROUNDS = (CHUNKS + THREADS - 1) / THREADS;   // 11
constexpr int NL_CAP = CHUNKS;                 // more newlines than this => some line < 16 bytes
constexpr int QCAP = 1024;                     // multi-node lines queued per tile (overflow: handled in place)
constexpr int OFF_NL = WIN;
constexpr int OFF_WTOT = OFF_NL + ((NL_CAP * 2 + 15) & ~15);
constexpr int OFF_Q = OFF_WTOT + ((ROUNDS * NWARPS * 4 + 15) & ~15);
constexpr int SMEM_BYTES = OFF_Q + QCAP * 8;
static_assert(WIN <= 65536, "line starts are stored as 16-bit window offsets");

constexpr uint32_t FLAG_EXACT_CHECKS = SVJG_FLAG_EXACT_CHECKS;   // probe links whose overlap test fails too
constexpr uint32_t FLAG_FORCE_GENERAL = SVJG_FLAG_FORCE_GENERAL; // test hook: every multi-node line through the general routine
constexpr uint32_t COMMA_PATH = 0xFFFFFFFFu;

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

    // ------------------------------------------------------------------ phase B
    // read_gaf_line (:184-198) on line.rstrip() (:126): 12 tab-separated columns,
    // 9 of them int().  Returns the number of path tokens for a '<'/'>' path,
    // COMMA_PATH for any other non-empty path.  Sets err where the reference raises.
    __device__ uint32_t parse_fields(P s, P e) {
        loc.n_rec++;
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
                // numeric columns: plain digit strings are checked on the fly
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
                    if (!plain || nd == 0 || nd > 18) r = parse_int(src, pos, f, v);
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

    // phase C re-reads Tlen, Ts, Te (columns 7-9) instead of carrying 24 bytes per queued line
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

    // --------------------------------------------------------- shared pieces
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
        for (uint32_t i = 0; i < t.l; ++i) tok_step(h, src[t.b + i]);
        return tok_value(h);
    }

    __device__ bool names_match(uint32_t off, const Tok &t) const {
        const uint8_t *q = a.tb.blob + off;
        for (uint32_t i = 0; i < t.l; ++i)
            if (__ldg(q + i) != src[t.b + i]) return false;
        return true;
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
                (lo.w >> 16) == tr_.l && names_match(lo.z, tl_) && names_match(lo.z + tl_.l, tr_)) {
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

    __device__ void emit(uint32_t sv2) {
        // warp-aggregated: one counter atomic per distinct SV allele among the
        // lanes that are here together, one cursor atomic for all of them
        cg::coalesced_group active = cg::coalesced_threads();
        cg::coalesced_group same = cg::labeled_partition(active, sv2);
        if (same.thread_rank() == 0) atomicAdd(a.counts + sv2, same.size());
        unsigned long long base = 0;
        if (active.thread_rank() == 0) base = atomicAdd(a.stats + 0, (unsigned long long)active.size());
        base = active.shfl(base, 0) + active.thread_rank();
        if (base < a.hit_cap) {
            a.hit_sv2[base] = sv2;
            a.hit_off[base] = line_off;
            a.hit_len[base] = line_len;
        }
    }

    // forward and reverse key of one link (:141-148) and their entries (:150-166).
    // ok_known: the overlap verdict is already there (fast path); otherwise it is
    // computed once, on the first key that has entries, by the general overlap().
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

    // ------------------------------------------------------- phase C, fast path
    // One token of a '<'/'>' path, streamed: hash, node length and the two
    // "could be a substring of an earlier token" signals.
    struct Scan {
        Tok t;
        uint64_t hash, sig;
        int64_t len;
        uint32_t start_val;
        int plus;          // 1 when the delimiter in front is '>'
        bool alt, plain;   // plain: exactly one ':', then digits-digits (<= 9 each) or digits.<anything>
    };

    __device__ bool scan_tok(P &cur, Scan &o) const {
        while (cur < pe && is_delim(src[cur])) ++cur;
        if (cur >= pe) return false;
        o.plus = src[cur - 1] == '>';
        o.t.b = cur;
        TokHash h = tok_init();
        uint64_t sig = 0;
        uint32_t v0 = 0, v1 = 0, nd0 = 0, nd1 = 0, dash = 0, colons = 0, prev = 0;
        bool dot = false, junk = false;
        for (; cur < pe; ++cur) {
            uint32_t c = src[cur];
            if (is_delim(c)) break;
            tok_step(h, c);
            if (cur != o.t.b) sig |= 1ull << ((c ^ (prev << 3) ^ (prev >> 2)) & 63);   // adjacent byte pair
            prev = c;
            uint32_t d = c - '0';
            if (d <= 9) {
                if (dot) {
                } else if (dash == 0) {
                    v0 = v0 * 10 + d;
                    ++nd0;
                } else {
                    v1 = v1 * 10 + d;
                    ++nd1;
                }
            } else if (c == ':') {
                ++colons;
                v0 = v1 = nd0 = nd1 = dash = 0;
                dot = junk = false;
            } else if (c == '-' && !dot) {
                ++dash;
            } else if (c == '.' && dash == 0) {
                dot = true;
            } else if (!dot) {
                junk = true;
            }
        }
        o.t.l = uint32_t(cur - o.t.b);
        o.hash = tok_value(h);
        o.sig = sig;
        o.alt = dot;
        o.start_val = v0;
        o.plain = colons == 1 && !junk && nd0 >= 1 && nd0 <= 9 && (dot || (dash == 1 && nd1 >= 1 && nd1 <= 9));
        o.len = int64_t(v1) - int64_t(v0) + 1;
        return true;
    }

    // false -> the general routine must take the line
    __device__ bool fast() {
        // pass 1: total path length; bail out on anything that is not plain
        int64_t total = 0;
        uint64_t seen = 0;
        uint32_t lo[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, hi[2] = {0, 0};
        uint32_t n = 0;
        P cur = ps;
        Scan t;
        while (scan_tok(cur, t)) {
            if (!t.plain) return false;
            if (t.alt && !alt_lookup(t.hash, t.t, t.len)) return false;
            if (t.len <= 0) return false;
            // t can only sit inside an earlier token x if every adjacent byte pair of t
            // occurs in x, and (single ':' in both) x has the same start digits and kind
            int k = t.alt;
            if (n && (t.sig & ~seen) == 0 && t.start_val >= lo[k] && t.start_val <= hi[k]) return false;
            seen |= t.sig;
            lo[k] = min(lo[k], t.start_val);
            hi[k] = max(hi[k], t.start_val);
            total += t.len;
            ++n;
        }
        // pass 2: every token is its own first occurrence, so strand = own delimiter
        // (:206), index = own position (:269-271) and the overlap sums are a prefix
        // sum and its complement (:269-273)
        const int64_t tail = tlen - te - 1;
        const bool all = a.flags & FLAG_EXACT_CHECKS;
        cur = ps;
        Scan A, B;
        scan_tok(cur, A);
        if (A.alt) alt_lookup(A.hash, A.t, A.len);
        int64_t pre = A.len;
        for (uint32_t i = 1; i < n; ++i) {
            scan_tok(cur, B);
            if (B.alt) alt_lookup(B.hash, B.t, B.len);
            bool ok = (pre - ts >= a.d_over) && (total - pre - tail >= a.d_over);
            if (ok || all) {
                link(A.t, A.hash, A.plus, B.t, B.hash, B.plus, true, ok);
                if (err) return true;
            }
            pre += B.len;
            A = B;
        }
        return true;
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

    // links of a line already split by parse_fields() / reparse_coords()
    __device__ void resolve() {
        if (angle && !(a.flags & FLAG_FORCE_GENERAL) && fast()) return;
        if (!err) general();
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

// 0x80 in every byte of w that equals '\n' (exact, no borrow artefacts)
__device__ __forceinline__ uint32_t nl_bytes(uint32_t w) {
    uint32_t x = w ^ 0x0A0A0A0Au;
    uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | x | 0x7F7F7F7Fu);
}
// bits 7,15,23,31 -> bits 0..3
__device__ __forceinline__ uint32_t pack4(uint32_t m) { return (((m >> 7) * 0x00204081u) >> 21) & 0xFu; }

__global__ void __launch_bounds__(THREADS, 2) filter_kernel(const FilterArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *win = smem;
    uint16_t *nl = reinterpret_cast<uint16_t *>(smem + OFF_NL);
    uint32_t *wtot = reinterpret_cast<uint32_t *>(smem + OFF_WTOT);
    ushort4 *queue = reinterpret_cast<ushort4 *>(smem + OFF_Q);   // (start, path start, path end, line end)
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t s_total, s_qn;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) mbar_init(&mbar, 1);
    __syncthreads();
    uint32_t phase = 0;
    Local loc;

    for (uint32_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        const uint64_t tile_start = uint64_t(tile) * TILE;
        const uint64_t g0 = tile_start ? tile_start - HEAD : 0;
        const uint32_t dst0 = tile_start ? 0 : HEAD;
        uint64_t g1 = tile_start + TILE + LOOKAHEAD;
        if (g1 > a.n) g1 = a.n;
        const bool at_eof = (g1 == a.n);
        const uint32_t nbytes = uint32_t(g1 - g0);
        const uint32_t bulk = nbytes & ~15u;
        const uint32_t valid_end = dst0 + nbytes;

        if (tid == 0) {
            s_qn = 0;
            if (bulk) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&mbar, bulk);
                bulk_g2s(win + dst0, a.gaf + g0, bulk, &mbar);
            }
        }
        for (uint32_t i = bulk + tid; i < nbytes; i += THREADS) win[dst0 + i] = __ldg(a.gaf + g0 + i);
        if (dst0 && tid < HEAD) win[tid] = tid == HEAD - 1 ? '\n' : 0;   // "newline" in front of byte 0 of the file
        for (uint32_t i = valid_end + tid; i < WIN; i += THREADS) win[i] = 0;
        if (bulk) {
            mbar_wait(&mbar, phase);
            phase ^= 1;
        }
        __syncthreads();

        // ---- phase A: newline scan, ordered compaction of line starts
        uint32_t masks[ROUNDS];
        uint32_t excl[ROUNDS];
#pragma unroll
        for (int r = 0; r < ROUNDS; ++r) {
            int c = r * THREADS + tid;
            uint32_t m = 0;
            if (c < CHUNKS) {
                uint4 v = *reinterpret_cast<const uint4 *>(win + c * 16);
                m = pack4(nl_bytes(v.x)) | (pack4(nl_bytes(v.y)) << 4) | (pack4(nl_bytes(v.z)) << 8) |
                    (pack4(nl_bytes(v.w)) << 12);
            }
            masks[r] = m;
            uint32_t cnt = __popc(m), inc = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += o;
            }
            excl[r] = inc - cnt;
            if (lane == 31) wtot[r * NWARPS + warp] = inc;
        }
        __syncthreads();
        if (warp == 0) {
            constexpr int PER = (ROUNDS * NWARPS + 31) / 32;
            uint32_t v[PER], sum = 0;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                int idx = lane * PER + k;
                v[k] = idx < ROUNDS * NWARPS ? wtot[idx] : 0;
                sum += v[k];
            }
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += o;
            }
            uint32_t run = inc - sum;
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                int idx = lane * PER + k;
                if (idx < ROUNDS * NWARPS) wtot[idx] = run;
                run += v[k];
            }
            if (lane == 31) s_total = inc;
        }
        __syncthreads();
        const uint32_t m_total = s_total;
        if (m_total > NL_CAP) {
            if (tid == 0) report(a, SVJG_BAD_SHORTLINE, tile_start);
        } else {
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r) {
                uint32_t m = masks[r], o = wtot[r * NWARPS + warp] + excl[r];
                uint32_t basepos = uint32_t(r * THREADS + tid) * 16;
                while (m) {
                    int b = __ffs(m) - 1;
                    m &= m - 1;
                    nl[o++] = uint16_t(basepos + b);
                }
            }
        }
        __syncthreads();

        if (m_total <= NL_CAP) {
            // ---- phase B: one thread per line, columns + integers + token count
            for (uint32_t k0 = 0; k0 < m_total; k0 += THREADS) {
                uint32_t k = k0 + tid;
                bool queued = false;
                uint32_t start = 0, end = 0, rps = 0, rpe = 0;
                if (k < m_total) {
                    start = uint32_t(nl[k]) + 1;
                    bool owned = start >= HEAD && start < HEAD + TILE && start < valid_end;
                    if (owned) {
                        uint32_t off = uint32_t(tile_start) + (start - HEAD);
                        bool in_win = (k + 1 < m_total) || at_eof;
                        if (in_win) {
                            end = (k + 1 < m_total) ? nl[k + 1] : valid_end;
                            uint32_t len = end - start + ((k + 1 < m_total) ? 1u : 0u);
                            Rec<SmemSrc> rec(a, SmemSrc{win}, off, len, loc);
                            uint32_t ntok = rec.parse_fields(start, end);
                            if (rec.err) {
                                report(a, rec.err, off);
                            } else if (ntok == COMMA_PATH) {
                                rec.general();
                                if (rec.err) report(a, rec.err, off);
                            } else if (ntok >= 2) {
                                loc.n_multi++;
                                queued = true;
                                rps = rec.ps;
                                rpe = rec.pe;
                            }
                        } else {
                            // the line runs past the staged window: general routine on global memory
                            uint64_t e = tile_start + TILE + LOOKAHEAD;
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
                    }
                }
                // compaction of the multi-node lines (one shared atomic per warp)
                uint32_t bal = __ballot_sync(0xFFFFFFFFu, queued);
                if (bal) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&s_qn, __popc(bal));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (queued) {
                        uint32_t q = base + __popc(bal & ((1u << lane) - 1));
                        if (q < QCAP) {
                            queue[q] = make_ushort4(uint16_t(start), uint16_t(rps), uint16_t(rpe), uint16_t(end));
                        } else {
                            // queue full: resolve in place
                            uint32_t off = uint32_t(tile_start) + (start - HEAD);
                            uint32_t len = end - start + ((end < valid_end && win[end] == '\n') ? 1u : 0u);
                            Rec<SmemSrc> rec(a, SmemSrc{win}, off, len, loc);
                            rec.ps = rps;
                            rec.pe = rpe;
                            rec.reparse_coords(end);
                            rec.resolve();
                            if (rec.err) report(a, rec.err, off);
                        }
                    }
                }
            }
            __syncthreads();
            // ---- phase C: one thread per multi-node line
            const uint32_t qn = min(s_qn, uint32_t(QCAP));
            for (uint32_t q = tid; q < qn; q += THREADS) {
                ushort4 e4 = queue[q];
                uint32_t start = e4.x, end = e4.w;
                uint32_t off = uint32_t(tile_start) + (start - HEAD);
                uint32_t len = end - start + ((end < valid_end && win[end] == '\n') ? 1u : 0u);
                Rec<SmemSrc> rec(a, SmemSrc{win}, off, len, loc);
                rec.ps = e4.y;
                rec.pe = e4.z;
                rec.reparse_coords(end);
                rec.resolve();
                if (rec.err) report(a, rec.err, off);
            }
        }
        __syncthreads();
    }

    // ---- per-CTA statistics
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
    int grid = int(std::min<uint32_t>(a.n_tiles, uint32_t(g_grid_cap)));
    filter_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(a);
    SVJG_CUDA(cudaGetLastError());
    return SVJG_OK;
}
