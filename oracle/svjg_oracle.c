/* svjg_oracle.c — CPU restatement in C of the reference's informative-alignment filter.
 * TEST INFRASTRUCTURE ONLY: the checker for full-size inputs, where the line-by-line Python oracle
 * (oracle/svjg_oracle.py) would take minutes.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu-baseline legs may load it; nothing under svjedi-graph_b200/ does.
 *
 * Parity status: PINNED.  tests/test_c_oracle.py holds it against the committed outputs of the
 * unmodified reference (tests/golden: c1, s2-s4, the 2 500 damaged lines of fuzz_lines.json.gz, the
 * 300 damaged link tables of fuzz_edges.json) and against the Python oracle on generated workloads.
 *
 * It follows filter-alignments.py string by string (paths relative to /root/reference):
 *   :123-126, :184-198  line.rstrip().split("\t"), the int() columns, the id:f: float / Alen rule
 *   :351-373            extract_nodes: split at '<' '>' , or the comma form with the last char dropped
 *   :200-219            strand of a node = the character before the FIRST occurrence of its name
 *   :141-148, :221-225  forward key "nL@sL@nR@sR", reverse key, probed in that order
 *   :258-273, :328-349  overlap sums over nodes.index() (first index), node length from the name or
 *                       from the GFA's alt-node table
 *   :153-166            the test is made once per (key, entry); an entry is appended when it passes
 * Python semantics that matter are restated (str.rstrip, int() with sign / blanks / underscores,
 * float() spellings).  What is NOT restated is reported instead of guessed: a line with a byte >= 0x80
 * (Unicode blanks and digits) or an integer of more than 18 digits gives ORACLE_UNSUPPORTED.
 * Lines end at '\n' only: a caller with carriage returns translates them first (text mode), as for
 * the CUDA path. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_RAISES 1       /* the reference raises on this line (exit status 1)  */
#define ORACLE_UNSUPPORTED 2  /* a spelling this restatement does not cover         */
#define ORACLE_HITS_FULL 3    /* hit_cap too small                                  */
#define ORACLE_NOMEM 4

#define POISON 0xFFFFFFFFu        /* an svs_edges entry the reference raises on once its overlap test has passed (:160, :166) */
#define POISON_ALWAYS 0xFFFFFFFEu /* an entry that cannot be unpacked into (sv id, allele): raises as soon as it is reached (:153) */

typedef struct {
    const char *p;
    size_t n;
} str;

typedef struct {
    /* link key -> entry range */
    const char *keys;
    const uint64_t *key_off;
    uint32_t n_keys;
    const uint32_t *ent_begin;
    const uint32_t *ent;
    uint32_t *key_slots, key_mask;
    /* alt node name -> sequence length */
    const char *alts;
    const uint64_t *alt_off;
    const int64_t *alt_len;
    uint32_t n_alt;
    uint32_t *alt_slots, alt_mask;
} tables;

static uint64_t fnv(const char *p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
    return h ^ (h >> 29);
}

static uint32_t pow2(uint32_t n) {
    uint32_t c = 16;
    while (c < n * 2u + 2u) c <<= 1;
    return c;
}

static int build_index(const char *blob, const uint64_t *off, uint32_t n, uint32_t **slots, uint32_t *mask) {
    uint32_t cap = pow2(n);
    uint32_t *s = (uint32_t *)malloc(sizeof(uint32_t) * cap);
    if (!s) return ORACLE_NOMEM;
    memset(s, 0xFF, sizeof(uint32_t) * cap);
    for (uint32_t i = 0; i < n; ++i) {            /* a later duplicate never occurs: keys come from a dict */
        uint32_t j = (uint32_t)fnv(blob + off[i], (size_t)(off[i + 1] - off[i])) & (cap - 1);
        while (s[j] != 0xFFFFFFFFu) j = (j + 1) & (cap - 1);
        s[j] = i;
    }
    *slots = s;
    *mask = cap - 1;
    return ORACLE_OK;
}

static uint32_t lookup(const char *blob, const uint64_t *off, const uint32_t *slots, uint32_t mask, const char *p, size_t n) {
    uint32_t j = (uint32_t)fnv(p, n) & mask;
    for (;; j = (j + 1) & mask) {
        uint32_t i = slots[j];
        if (i == 0xFFFFFFFFu) return i;
        if (off[i + 1] - off[i] == n && memcmp(blob + off[i], p, n) == 0) return i;
    }
}

/* str.isspace() for ASCII: what rstrip() and int() strip */
static int py_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

/* int(s): 0 ok, ORACLE_RAISES ValueError, ORACLE_UNSUPPORTED beyond 18 digits */
static int py_int(str s, int64_t *out) {
    while (s.n && py_space((unsigned char)s.p[0])) s.p++, s.n--;
    while (s.n && py_space((unsigned char)s.p[s.n - 1])) s.n--;
    int neg = 0;
    if (s.n && (s.p[0] == '+' || s.p[0] == '-')) {
        neg = s.p[0] == '-';
        s.p++, s.n--;
    }
    if (!s.n) return ORACLE_RAISES;
    int64_t v = 0;
    int digits = 0, prev_digit = 0;
    for (size_t i = 0; i < s.n; ++i) {
        char c = s.p[i];
        if (c >= '0' && c <= '9') {
            if (v || c != '0')
                if (++digits > 18) return ORACLE_UNSUPPORTED;
            v = v * 10 + (c - '0');
            prev_digit = 1;
        } else if (c == '_' && prev_digit) {
            prev_digit = 0;
        } else {
            return ORACLE_RAISES;
        }
    }
    if (!prev_digit) return ORACLE_RAISES;
    *out = neg ? -v : v;
    return ORACLE_OK;
}

static int lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }
static int ieq(str s, const char *w) {
    size_t n = strlen(w);
    if (s.n != n) return 0;
    for (size_t i = 0; i < n; ++i)
        if (lower((unsigned char)s.p[i]) != w[i]) return 0;
    return 1;
}
/* digits with single underscores between them, at least one digit; returns chars consumed or 0 */
static size_t digit_run(const char *p, size_t n) {
    size_t i = 0;
    int prev_digit = 0;
    while (i < n) {
        if (p[i] >= '0' && p[i] <= '9') prev_digit = 1;
        else if (p[i] == '_' && prev_digit && i + 1 < n && p[i + 1] >= '0' && p[i + 1] <= '9') prev_digit = 0;
        else break;
        ++i;
    }
    return i;
}
/* float(s) accepts it?  (value not needed: filter-alignments.py:193-194 only stores it) */
static int py_float_ok(str s) {
    while (s.n && py_space((unsigned char)s.p[0])) s.p++, s.n--;
    while (s.n && py_space((unsigned char)s.p[s.n - 1])) s.n--;
    if (s.n && (s.p[0] == '+' || s.p[0] == '-')) s.p++, s.n--;
    if (!s.n) return 0;
    if (ieq(s, "inf") || ieq(s, "infinity") || ieq(s, "nan")) return 1;
    size_t a = digit_run(s.p, s.n), i = a, b = 0;
    if (i < s.n && s.p[i] == '.') {
        ++i;
        b = digit_run(s.p + i, s.n - i);
        i += b;
    }
    if (!a && !b) return 0;
    if (i < s.n && (s.p[i] == 'e' || s.p[i] == 'E')) {
        ++i;
        if (i < s.n && (s.p[i] == '+' || s.p[i] == '-')) ++i;
        size_t e = digit_run(s.p + i, s.n - i);
        if (!e) return 0;
        i += e;
    }
    return i == s.n;
}

static const char *find(str hay, str needle) {     /* str.find: first occurrence; needle.n > 0 */
    if (needle.n > hay.n) return NULL;
    const char *end = hay.p + (hay.n - needle.n);
    for (const char *q = hay.p; q <= end; ++q) {
        q = (const char *)memchr(q, needle.p[0], (size_t)(end - q) + 1);
        if (!q) return NULL;
        if (memcmp(q, needle.p, needle.n) == 0) return q;
    }
    return NULL;
}

/* filter-alignments.py:343-349 */
static int node_len(const tables *t, str name, __int128 *out) {
    size_t c = name.n;
    while (c > 0 && name.p[c - 1] != ':') --c;      /* name.split(":")[-1] */
    str last = {name.p + c, name.n - c};
    if (memchr(last.p, '.', last.n)) {
        uint32_t i = t->n_alt ? lookup(t->alts, t->alt_off, t->alt_slots, t->alt_mask, name.p, name.n) : 0xFFFFFFFFu;
        if (i == 0xFFFFFFFFu) return ORACLE_RAISES;                 /* KeyError */
        *out = t->alt_len[i];
        return ORACLE_OK;
    }
    /* last.split("-"): pieces 0 and 1 */
    const char *d0 = (const char *)memchr(last.p, '-', last.n);
    if (!d0) return ORACLE_RAISES;                                   /* IndexError */
    str p0 = {last.p, (size_t)(d0 - last.p)};
    const char *rest = d0 + 1;
    size_t rest_n = last.n - p0.n - 1;
    const char *d1 = (const char *)memchr(rest, '-', rest_n);
    str p1 = {rest, d1 ? (size_t)(d1 - rest) : rest_n};
    int64_t s, e;
    /* get_node_end is evaluated before get_node_start in `end - start + 1`; either failure raises */
    int r1 = py_int(p1, &e), r0 = py_int(p0, &s);
    if (r1 == ORACLE_RAISES || r0 == ORACLE_RAISES) return ORACLE_RAISES;
    if (r1 || r0) return ORACLE_UNSUPPORTED;
    *out = (__int128)e - s + 1;
    return ORACLE_OK;
}

typedef struct {
    str *tok;
    char *strand;
    uint32_t *first;
    size_t cap;
    char *key;
    size_t key_cap;
} scratch;

static int grow(scratch *sc, size_t n) {
    if (n <= sc->cap) return ORACLE_OK;
    size_t cap = sc->cap ? sc->cap : 64;
    while (cap < n) cap *= 2;
    str *a = (str *)realloc(sc->tok, cap * sizeof(str));
    if (a) sc->tok = a;
    char *b = (char *)realloc(sc->strand, cap);
    if (b) sc->strand = b;
    uint32_t *c = (uint32_t *)realloc(sc->first, cap * sizeof(uint32_t));
    if (c) sc->first = c;
    if (!a || !b || !c) return ORACLE_NOMEM;
    sc->cap = cap;
    return ORACLE_OK;
}

typedef struct {
    uint32_t *counts;
    uint32_t *hit_sv2;
    uint64_t *hit_off;
    uint32_t *hit_len;
    uint64_t hit_cap, n_hits, n_records, n_multi;
} sink;

/* one line (without its '\n'); line_off / line_len describe the raw line for the hit tuples */
static int do_line(const tables *t, scratch *sc, sink *out, str line, uint64_t line_off, uint32_t line_len, int64_t d_over) {
    for (size_t i = 0; i < line.n; ++i)
        if ((unsigned char)line.p[i] >= 0x80) return ORACLE_UNSUPPORTED;
    while (line.n && py_space((unsigned char)line.p[line.n - 1])) line.n--;       /* line.rstrip() */
    /* split("\t"): the first 12 columns */
    str col[12];
    int nc = 0;
    size_t a = 0;
    while (nc < 12) {
        const char *tab = (const char *)memchr(line.p + a, '\t', line.n - a);
        if (!tab) {
            col[nc].p = line.p + a, col[nc].n = line.n - a;
            ++nc;
            break;
        }
        col[nc].p = line.p + a, col[nc].n = (size_t)(tab - (line.p + a));
        ++nc;
        a = (size_t)(tab - line.p) + 1;
    }
    if (nc < 12) return ORACLE_RAISES;
    static const int INTS[9] = {1, 2, 3, 6, 7, 8, 9, 10, 11};
    int64_t v[12] = {0};
    int unsupported = 0;
    for (int k = 0; k < 9; ++k) {
        int r = py_int(col[INTS[k]], &v[INTS[k]]);
        if (r == ORACLE_RAISES) return r;
        if (r) unsupported = 1;
    }
    if (unsupported) return ORACLE_UNSUPPORTED;
    /* "id:f:" in line: float(line.split("id:f:")[-1].split("\t")[0]); else Am / Alen */
    {
        const char *last = NULL;
        str hay = line, tag = {"id:f:", 5};
        for (;;) {
            const char *q = find(hay, tag);
            if (!q) break;
            last = q;
            hay.n -= (size_t)(q + 5 - hay.p);
            hay.p = q + 5;
        }
        if (last) {
            const char *b = last + 5, *e = (const char *)memchr(b, '\t', (size_t)(line.p + line.n - b));
            str f = {b, e ? (size_t)(e - b) : (size_t)(line.p + line.n - b)};
            if (!py_float_ok(f)) return ORACLE_RAISES;
        } else if (v[10] == 0) {
            return ORACLE_RAISES;                                     /* ZeroDivisionError */
        }
    }
    out->n_records++;
    const str path = col[5];
    const int64_t tlen = v[6], ts = v[7], te = v[8];
    if (!path.n) return ORACLE_RAISES;                               /* path[0] of an empty string */
    /* extract_nodes (:351-373) */
    size_t n = 0;
    if (path.p[0] == '<' || path.p[0] == '>') {
        size_t b = 0;
        for (size_t i = 0; i <= path.n; ++i) {
            if (i == path.n || path.p[i] == '<' || path.p[i] == '>') {
                if (i > b) {
                    if (grow(sc, n + 1)) return ORACLE_NOMEM;
                    sc->tok[n].p = path.p + b, sc->tok[n].n = i - b;
                    ++n;
                }
                b = i + 1;
            }
        }
    } else {
        size_t b = 0;
        for (size_t i = 0; i <= path.n; ++i) {
            if (i == path.n || path.p[i] == ',') {
                if (i > b) {                                          /* non-empty piece, last char dropped */
                    if (grow(sc, n + 1)) return ORACLE_NOMEM;
                    sc->tok[n].p = path.p + b, sc->tok[n].n = i - b - 1;
                    ++n;
                }
                b = i + 1;
            }
        }
    }
    if (n < 2) return ORACLE_OK;
    out->n_multi++;
    /* strands (:206) and first indices */
    for (size_t i = 0; i < n; ++i) {
        if (!sc->tok[i].n) return ORACLE_RAISES;
        const char *q = find(path, sc->tok[i]);
        if (!q || q == path.p) return ORACLE_RAISES;
        sc->strand[i] = q[-1] == '>' ? '+' : '-';
        sc->first[i] = (uint32_t)i;
        for (size_t j = 0; j < i; ++j)
            if (sc->tok[j].n == sc->tok[i].n && memcmp(sc->tok[j].p, sc->tok[i].p, sc->tok[i].n) == 0) {
                sc->first[i] = (uint32_t)j;
                break;
            }
    }
    const __int128 tail_clip = (__int128)tlen - te - 1;
    for (size_t i = 1; i < n; ++i) {
        const str na = sc->tok[i - 1], nb = sc->tok[i];
        const char sa = sc->strand[i - 1], sb = sc->strand[i];
        const size_t kn = na.n + nb.n + 6;
        if (kn > sc->key_cap) {
            char *k = (char *)realloc(sc->key, kn * 2);
            if (!k) return ORACLE_NOMEM;
            sc->key = k, sc->key_cap = kn * 2;
        }
        int verdict = -1;                                            /* overlap test not made yet */
        for (int dir = 0; dir < 2; ++dir) {
            char *k = sc->key;
            const str l = dir ? nb : na, r = dir ? na : nb;
            const char sl = dir ? (sb == '+' ? '-' : '+') : sa, sr = dir ? (sa == '+' ? '-' : '+') : sb;
            memcpy(k, l.p, l.n), k += l.n;
            *k++ = '@', *k++ = sl, *k++ = '@';
            memcpy(k, r.p, r.n), k += r.n;
            *k++ = '@', *k++ = sr;
            const uint32_t ki = t->n_keys ? lookup(t->keys, t->key_off, t->key_slots, t->key_mask, sc->key, (size_t)(k - sc->key)) : 0xFFFFFFFFu;
            if (ki == 0xFFFFFFFFu) continue;
            for (uint32_t e = t->ent_begin[ki]; e < t->ent_begin[ki + 1]; ++e) {
                if (t->ent[e] == POISON_ALWAYS) return ORACLE_RAISES;
                if (verdict < 0) {                                   /* the same sums for every entry of this link */
                    __int128 left = 0, right = 0, len;
                    int rc;
                    for (size_t j = 0; j <= sc->first[i - 1]; ++j) {
                        if ((rc = node_len(t, sc->tok[j], &len))) return rc;
                        left += len;
                    }
                    for (size_t j = sc->first[i]; j < n; ++j) {
                        if ((rc = node_len(t, sc->tok[j], &len))) return rc;
                        right += len;
                    }
                    verdict = (left - ts >= d_over) && (right - tail_clip >= d_over);
                }
                if (!verdict) continue;
                const uint32_t sv2 = t->ent[e];
                if (sv2 == POISON) return ORACLE_RAISES;
                out->counts[sv2]++;
                if (out->hit_sv2) {
                    if (out->n_hits >= out->hit_cap) return ORACLE_HITS_FULL;
                    out->hit_sv2[out->n_hits] = sv2;
                    out->hit_off[out->n_hits] = line_off;
                    out->hit_len[out->n_hits] = line_len;
                }
                out->n_hits++;
            }
        }
    }
    return ORACLE_OK;
}

/* Filters gaf[0:n) (lines end at '\n').  Tables as flat arrays (oracle/c_oracle.py builds them):
 *   keys / key_off[n_keys+1]      link key strings back to back
 *   ent_begin[n_keys+1], ent[]    entries of key i: 2*sv_index + allele, or POISON
 *   alts / alt_off[n_alt+1], alt_len[n_alt]
 * counts[2*num_sv] is accumulated (not cleared).  hit_* may be NULL (counts only).
 * stats[0..3] = hits, records (lines that parse), records with >= 2 nodes, offset of the failing line.
 * Returns ORACLE_OK or the code of the FIRST line that does not pass (processing stops there, like the
 * reference's traceback). */
int svjg_oracle_filter(const char *gaf, uint64_t n, uint64_t base_off,
                       const char *keys, const uint64_t *key_off, uint32_t n_keys, const uint32_t *ent_begin, const uint32_t *ent,
                       const char *alts, const uint64_t *alt_off, const int64_t *alt_len, uint32_t n_alt,
                       int64_t d_over, uint32_t *counts, uint32_t *hit_sv2, uint64_t *hit_off, uint32_t *hit_len,
                       uint64_t hit_cap, uint64_t *stats) {
    tables t = {keys, key_off, n_keys, ent_begin, ent, NULL, 0, alts, alt_off, alt_len, n_alt, NULL, 0};
    scratch sc = {0};
    sink out = {counts, hit_sv2, hit_off, hit_len, hit_cap, 0, 0, 0};
    int rc = ORACLE_OK;
    if (n_keys && (rc = build_index(keys, key_off, n_keys, &t.key_slots, &t.key_mask))) goto done;
    if (n_alt && (rc = build_index(alts, alt_off, n_alt, &t.alt_slots, &t.alt_mask))) goto done;
    stats[3] = 0;
    for (uint64_t p = 0; p < n;) {
        const char *nl = (const char *)memchr(gaf + p, '\n', (size_t)(n - p));
        const uint64_t next = nl ? (uint64_t)(nl - gaf) + 1 : n;
        str line = {gaf + p, (size_t)((nl ? (uint64_t)(nl - gaf) : n) - p)};
        rc = do_line(&t, &sc, &out, line, base_off + p, (uint32_t)(next - p), d_over);
        if (rc) {
            stats[3] = base_off + p;
            break;
        }
        p = next;
    }
done:
    stats[0] = out.n_hits, stats[1] = out.n_records, stats[2] = out.n_multi;
    free(t.key_slots), free(t.alt_slots), free(sc.tok), free(sc.strand), free(sc.first), free(sc.key);
    return rc;
}
