// Host side of the graph tables: parses <prefix>_svs_edges.json and the GFA and
// builds the device image (link-key hash -> SV entries, alt-node hash -> length).
// Reference behaviour restated: filter-alignments.py:95-98 (json.load into
// d_link_sv) and :103-113 (alt_node_len).  Written from scratch; no reference
// code is reused.
#include "svjg_internal.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>

namespace svjg {

// SVJG_TIMING=1: wall clock of the host stages on stderr (profiling hook, no effect on results)
struct StageTimer {
    bool on;
    std::chrono::steady_clock::time_point t0;
    StageTimer() : on(getenv("SVJG_TIMING") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "svjg timing: %-28s %8.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
        t0 = t1;
    }
};


// ---------------------------------------------------------------------------
// a small strict JSON reader (only what svs_edges.json needs, but any valid
// JSON value is skipped correctly)
// ---------------------------------------------------------------------------
struct JsonIn {
    const char *p, *end;
    std::string err;
    bool fail(const char *m) {
        if (err.empty()) err = m;
        return false;
    }
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    }
    bool lit(const char *s) {
        size_t n = strlen(s);
        if (size_t(end - p) >= n && memcmp(p, s, n) == 0) {
            p += n;
            return true;
        }
        return false;
    }
    static void put_utf8(std::string &out, uint32_t cp) {
        if (cp < 0x80) out.push_back(char(cp));
        else if (cp < 0x800) {
            out.push_back(char(0xC0 | (cp >> 6)));
            out.push_back(char(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            out.push_back(char(0xE0 | (cp >> 12)));
            out.push_back(char(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(char(0x80 | (cp & 0x3F)));
        } else {
            out.push_back(char(0xF0 | (cp >> 18)));
            out.push_back(char(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back(char(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(char(0x80 | (cp & 0x3F)));
        }
    }
    bool hex4(uint32_t &v) {
        if (end - p < 4) return fail("truncated \\u escape");
        v = 0;
        for (int i = 0; i < 4; ++i) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
            else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
            else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
            else return fail("bad \\u escape");
        }
        return true;
    }
    bool string(std::string &out) {
        out.clear();
        if (p >= end || *p != '"') return fail("expected string");
        // nearly every string of these files has no escape: one memchr to the closing quote, one copy
        if (const char *q = static_cast<const char *>(memchr(p + 1, '"', size_t(end - p - 1)))) {
            const size_t n = size_t(q - p - 1);
            if (!memchr(p + 1, '\\', n) && !has_control(p + 1, n)) {
                out.assign(p + 1, n);
                p = q + 1;
                return true;
            }
        }
        ++p;
        while (p < end) {
            unsigned char c = (unsigned char)*p++;
            if (c == '"') return true;
            if (c < 0x20) return fail("control character in a string");     // json.load is strict about these
            if (c != '\\') {
                out.push_back(char(c));
                continue;
            }
            if (p >= end) break;
            char e = *p++;
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t cp;
                    if (!hex4(cp)) return false;
                    if (cp >= 0xD800 && cp < 0xDC00 && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        const char *save = p;
                        p += 2;
                        uint32_t lo;
                        if (!hex4(lo)) return false;
                        if (lo >= 0xDC00 && lo < 0xE000) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        else p = save;
                    }
                    put_utf8(out, cp);
                    break;
                }
                default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    // any byte below 0x20 (json.load rejects them inside strings), eight bytes a step
    static bool has_control(const char *b, size_t n) {
        size_t i = 0;
        uint64_t hit = 0;
        for (; i + 8 <= n; i += 8) {
            uint64_t w;
            memcpy(&w, b + i, 8);
            // bytes < 0x20: the top three bits are clear
            hit |= ~(w | (w << 1) | (w << 2)) & 0x8080808080808080ull;
        }
        for (; i < n; ++i) hit |= uint64_t((unsigned char)b[i] < 0x20);
        return hit != 0;
    }
    // a string whose content is not needed (the GAF lines of informative_aln.json: hundreds of MB):
    // jump from quote to quote; a quote is the end unless an odd number of backslashes precedes it.
    // Escapes are still checked as string() would, so that a malformed file fails the same way.
    bool skip_string() {
        if (p >= end || *p != '"') return fail("expected string");
        const char *q = p + 1;
        for (;;) {
            const char *e = static_cast<const char *>(memchr(q, '"', size_t(end - q)));
            if (!e) {
                p = end;
                return fail("unterminated string");
            }
            size_t nb = 0;
            while (e - nb > p + 1 && e[-1 - long(nb)] == '\\') ++nb;
            if (nb % 2 == 0) {
                if (has_control(p + 1, size_t(e - p - 1))) return fail("control character in a string");
                // validate the escapes of [p+1, e)
                for (const char *b = static_cast<const char *>(memchr(p + 1, '\\', size_t(e - p - 1))); b;) {
                    const char esc = b[1];
                    size_t adv = 2;
                    if (esc == 'u') {
                        if (e - b < 6) return fail("truncated \\u escape");
                        for (int i = 2; i < 6; ++i)
                            if (!isxdigit((unsigned char)b[i])) return fail("bad \\u escape");
                        adv = 6;
                    } else if (!strchr("\"\\/bfnrt", esc) || esc == 0) {
                        return fail("bad escape");
                    }
                    const char *from = b + adv;
                    b = from < e ? static_cast<const char *>(memchr(from, '\\', size_t(e - from))) : nullptr;
                }
                p = e + 1;
                return true;
            }
            q = e + 1;
        }
    }
    // scalar kinds for the allele slot
    enum Kind { K_INT, K_FLOAT, K_TRUE, K_FALSE, K_NULL, K_STRING, K_ARRAY, K_OBJECT };
    bool skip_value(Kind &kind, long long *ival = nullptr, std::string *sval = nullptr) {
        ws();
        if (p >= end) return fail("unexpected end");
        char c = *p;
        if (c == '"') {
            kind = K_STRING;
            return sval ? string(*sval) : skip_string();
        }
        if (c == '{') {
            kind = K_OBJECT;
            ++p;
            ws();
            if (p < end && *p == '}') {
                ++p;
                return true;
            }
            for (;;) {
                ws();
                std::string k;
                if (!string(k)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                Kind kk;
                if (!skip_value(kk)) return false;
                ws();
                if (p < end && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < end && *p == '}') {
                    ++p;
                    return true;
                }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            kind = K_ARRAY;
            ++p;
            ws();
            if (p < end && *p == ']') {
                ++p;
                return true;
            }
            for (;;) {
                Kind kk;
                if (!skip_value(kk)) return false;
                ws();
                if (p < end && *p == ',') {
                    ++p;
                    continue;
                }
                if (p < end && *p == ']') {
                    ++p;
                    return true;
                }
                return fail("expected ',' or ']'");
            }
        }
        if (lit("true")) {
            kind = K_TRUE;
            return true;
        }
        if (lit("false")) {
            kind = K_FALSE;
            return true;
        }
        if (lit("null")) {
            kind = K_NULL;
            return true;
        }
        if (lit("NaN") || lit("Infinity") || lit("-Infinity")) {
            kind = K_FLOAT;
            return true;
        }
        // number, by json's own pattern  -?(0|[1-9]\d*)(\.\d+)?([eE][-+]?\d+)?  : a part that is not complete
        // is left unread and trips the delimiter check of the caller, as in the reference
        const char *s = p;
        bool is_float = false;
        if (p < end && *p == '-') ++p;
        if (p >= end || *p < '0' || *p > '9') return fail("bad value");
        if (*p == '0') ++p;
        else
            while (p < end && *p >= '0' && *p <= '9') ++p;
        if (end - p >= 2 && *p == '.' && p[1] >= '0' && p[1] <= '9') {
            is_float = true;
            ++p;
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        if (p < end && (*p == 'e' || *p == 'E')) {
            const char *q = p + 1;
            if (q < end && (*q == '+' || *q == '-')) ++q;
            if (q < end && *q >= '0' && *q <= '9') {
                is_float = true;
                while (q < end && *q >= '0' && *q <= '9') ++q;
                p = q;
            }
        }
        kind = is_float ? K_FLOAT : K_INT;
        if (!is_float && ival) {
            // clamp: anything outside [-2, 1] is an out-of-range list index anyway
            long long v = 0;
            bool neg = (*s == '-');
            const char *d = s + (neg ? 1 : 0);
            int nd = 0;
            for (; d < p; ++d, ++nd) v = nd < 18 ? v * 10 + (*d - '0') : 1000;
            *ival = neg ? -v : v;
        }
        return true;
    }
};

struct RawKey {
    std::string key;
    bool poison_key = false;
    std::vector<std::pair<std::string, int>> ents;  // allele -1 = poison entry
};

// one member of the top-level object: "nL@sL@nR@sR": [["sv id", allele], ...]   (in.p at the opening quote of the key)
static bool parse_member(JsonIn &in, RawKey &rk) {
    if (!in.string(rk.key)) return false;
    in.ws();
    if (in.p >= in.end || *in.p != ':') {
        in.fail("expected ':'");
        return false;
    }
    ++in.p;
    in.ws();
    if (in.p < in.end && *in.p == '[') {
        ++in.p;
        in.ws();
        if (in.p < in.end && *in.p == ']') {
            ++in.p;
        } else {
            for (;;) {
                in.ws();
                // one entry: [ "sv id", allele ]
                if (in.p < in.end && *in.p == '[') {
                    ++in.p;
                    int n_el = 0;
                    std::string sv;
                    bool sv_is_str = false;
                    int allele = -1;
                    in.ws();
                    if (in.p < in.end && *in.p == ']') {
                        ++in.p;
                    } else {
                        for (;;) {
                            JsonIn::Kind kind;
                            long long iv = 0;
                            std::string sval;
                            if (!in.skip_value(kind, &iv, &sval)) break;
                            if (n_el == 0 && kind == JsonIn::K_STRING) {
                                sv_is_str = true;
                                sv.swap(sval);
                            }
                            if (n_el == 1) {
                                // list index semantics of Python: 0,1,-1,-2 and bools are valid
                                if (kind == JsonIn::K_INT) {
                                    if (iv == 0 || iv == -2) allele = 0;
                                    else if (iv == 1 || iv == -1) allele = 1;
                                } else if (kind == JsonIn::K_TRUE) allele = 1;
                                else if (kind == JsonIn::K_FALSE) allele = 0;
                            }
                            ++n_el;
                            in.ws();
                            if (in.p < in.end && *in.p == ',') {
                                ++in.p;
                                continue;
                            }
                            if (in.p < in.end && *in.p == ']') {
                                ++in.p;
                                break;
                            }
                            in.fail("expected ',' or ']'");
                            break;
                        }
                    }
                    if (!in.err.empty()) break;
                    if (n_el != 2) rk.poison_key = true;  // tuple unpack raises as soon as the key is found
                    else if (!sv_is_str || sv.find(':') == std::string::npos) rk.ents.push_back({std::string(), -1});
                    else rk.ents.push_back({sv, allele});
                } else {
                    JsonIn::Kind kind;
                    if (!in.skip_value(kind)) break;
                    // a 2-character string would unpack; everything else raises
                    rk.poison_key = true;
                }
                in.ws();
                if (in.p < in.end && *in.p == ',') {
                    ++in.p;
                    continue;
                }
                if (in.p < in.end && *in.p == ']') {
                    ++in.p;
                    break;
                }
                in.fail("expected ',' or ']'");
                break;
            }
        }
    } else {
        JsonIn::Kind kind;
        if (!in.skip_value(kind)) return false;
        rk.poison_key = true;
    }
    return in.err.empty();
}

// json.dumps(sort_keys=True) wrote the keys in ascending order: no key can repeat while that holds, and the
// dictionary of seen keys is only built once it stops holding.  Duplicate keys: the last one wins (dict).
struct KeySink {
    std::vector<RawKey> &keys;
    std::unordered_map<std::string, size_t> seen;
    bool ascending = true;
    void add(RawKey &&rk) {
        if (ascending && (keys.empty() || keys.back().key < rk.key)) {
            keys.push_back(std::move(rk));
            return;
        }
        if (ascending) {
            ascending = false;
            for (size_t i = 0; i < keys.size(); ++i) seen.emplace(keys[i].key, i);
        }
        auto it = seen.find(rk.key);
        if (it == seen.end()) {
            seen.emplace(rk.key, keys.size());
            keys.push_back(std::move(rk));
        } else {
            keys[it->second] = std::move(rk);
        }
    }
};

// The file as json.dumps(indent=4) writes it (construct-graph.py:554) has every top-level member start with
// "\n    \"" -- a raw line feed cannot occur inside a JSON string, deeper levels are indented further -- so a large
// file is cut there into pieces that are parsed side by side.  A piece must consist of whole members, each
// followed by a comma (the last piece by the closing brace and nothing else): if every piece does, the cuts were
// member boundaries and the pieces together are the sequential parse.  Anything else (another layout, a damaged
// file): false, and the sequential parser decides and reports.
static bool parse_edges_pieces(const char *text, size_t len, std::vector<RawKey> &keys) {
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t n_pieces = std::min<size_t>(std::min<unsigned>(hw ? hw : 1u, 8u), len >> 24);      // >= 16 MiB each
    if (n_pieces < 2) return false;
    JsonIn head{text, text + len, {}};
    head.ws();
    if (head.p >= head.end || *head.p != '{') return false;
    ++head.p;
    head.ws();
    if (head.p >= head.end || *head.p != '"') return false;
    std::vector<const char *> cut{head.p};
    static const char mark[] = "\n    \"";
    for (size_t k = 1; k < n_pieces; ++k) {
        const char *from = std::max(cut.back(), text + len / n_pieces * k);
        const void *hit = memmem(from, size_t(text + len - from), mark, 6);
        if (!hit) break;
        const char *q = static_cast<const char *>(hit) + 5;
        if (q > cut.back()) cut.push_back(q);
    }
    cut.push_back(text + len);
    const size_t n = cut.size() - 1;
    if (n < 2) return false;
    std::vector<std::vector<RawKey>> part(n);
    std::vector<char> ok(n, 0);
    std::vector<std::thread> pool;
    for (size_t k = 0; k < n; ++k)
        pool.emplace_back([&, k] {
            JsonIn in{cut[k], cut[k + 1], {}};
            const bool last = k + 1 == n;
            for (;;) {
                RawKey rk;
                if (!parse_member(in, rk)) return;
                part[k].push_back(std::move(rk));
                in.ws();
                if (in.p < in.end && *in.p == ',') {
                    ++in.p;
                    in.ws();
                    if (in.p == in.end) {            // the piece ends behind a comma: fine unless it is the last one
                        ok[k] = !last;
                        return;
                    }
                    continue;
                }
                if (last && in.p < in.end && *in.p == '}') {
                    ++in.p;
                    in.ws();
                    ok[k] = in.p == in.end;
                }
                return;
            }
        });
    for (auto &t : pool) t.join();
    for (size_t k = 0; k < n; ++k)
        if (!ok[k]) return false;
    KeySink sink{keys, {}, true};
    for (auto &v : part)
        for (auto &rk : v) sink.add(std::move(rk));
    return true;
}

static bool parse_edges(const char *text, size_t len, std::vector<RawKey> &keys, std::string &err) {
    if (parse_edges_pieces(text, len, keys)) return true;
    keys.clear();
    JsonIn in{text, text + len, {}};
    // json.load() accepts a UTF-8 BOM-less document; skip leading whitespace
    in.ws();
    if (in.p >= in.end || *in.p != '{') {
        err = "svs_edges: top level is not a JSON object";
        return false;
    }
    ++in.p;
    KeySink sink{keys, {}, true};
    in.ws();
    if (in.p < in.end && *in.p == '}') {
        ++in.p;
    } else {
        for (;;) {
            in.ws();
            RawKey rk;
            if (!parse_member(in, rk)) break;
            sink.add(std::move(rk));
            in.ws();
            if (in.p < in.end && *in.p == ',') {
                ++in.p;
                continue;
            }
            if (in.p < in.end && *in.p == '}') {
                ++in.p;
                break;
            }
            in.fail("expected ',' or '}'");
            break;
        }
    }
    if (!in.err.empty()) {
        err = "svs_edges: " + in.err;
        return false;
    }
    in.ws();
    if (in.p != in.end) {
        err = "svs_edges: trailing data after the JSON object";
        return false;
    }
    return true;
}

// ---------------------------------------------------------------------------
// GFA: alt-node lengths   (filter-alignments.py:103-113)
// ---------------------------------------------------------------------------
static inline bool py_space(unsigned char c) {
    return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f);
}

// str.rstrip() also removes the non-ASCII white space of Unicode; as UTF-8, ending at `e`: its length or 0
static inline int py_space_utf8_before(const char *b, const char *e) {
    const unsigned char *u = (const unsigned char *)e;
    if (e - b >= 2 && u[-2] == 0xC2 && (u[-1] == 0x85 || u[-1] == 0xA0)) return 2;
    if (e - b >= 3) {
        const unsigned a = u[-3], m = u[-2], z = u[-1];
        if (a == 0xE1 && m == 0x9A && z == 0x80) return 3;                                          // U+1680
        if (a == 0xE2 && m == 0x80 && ((z >= 0x80 && z <= 0x8A) || z == 0xA8 || z == 0xA9 || z == 0xAF)) return 3;
        if (a == 0xE2 && m == 0x81 && z == 0x9F) return 3;                                          // U+205F
        if (a == 0xE3 && m == 0x80 && z == 0x80) return 3;                                          // U+3000
    }
    return 0;
}

static bool scan_gfa(const char *text, size_t len, std::vector<std::pair<std::string, int64_t>> &alts,
                     std::string &err) {
    std::unordered_map<std::string, size_t> seen;
    const char *p = text, *end = text + len;
    // the reference reads the GFA in text mode: "\n", "\r\n" and a lone "\r" all end a line
    const char *cr = (const char *)memchr(text, '\r', len);
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', size_t(end - p));
        const char *le = nl ? nl : end;            // line without its line end
        const char *next = nl ? nl + 1 : end;
        if (cr && cr < p) cr = (const char *)memchr(p, '\r', size_t(end - p));
        if (cr && cr < le) {
            le = cr;
            next = (cr + 1 < end && cr[1] == '\n') ? cr + 2 : cr + 1;
        }
        if (*p == 'S') {
            // cols of the raw line (with its newline, as the reference splits before stripping)
            const char *t1 = (const char *)memchr(p, '\t', size_t(le - p));
            if (!t1) {
                err = "GFA: S line without a name column";
                return false;
            }
            const char *name_b = t1 + 1;
            const char *t2 = (const char *)memchr(name_b, '\t', size_t(le - name_b));
            const char *name_e = t2 ? t2 : le;     // two columns: the name would keep the "\n", and [2] raises below
            // last ':' piece
            const char *q = name_e;
            while (q > name_b && q[-1] != ':') --q;
            bool alt = memchr(q, '.', size_t(name_e - q)) != nullptr;
            if (alt) {
                // line.rstrip().split("\t")[2]
                const char *re = le;
                for (;;) {
                    if (re > p && py_space((unsigned char)re[-1])) {
                        --re;
                        continue;
                    }
                    const int k = py_space_utf8_before(p, re);
                    if (!k) break;
                    re -= k;
                }
                if (!t2 || t2 >= re) {   // fewer than 3 columns once stripped -> IndexError in the reference
                    err = "GFA: alt-node S line without a sequence column";
                    return false;
                }
                const char *seq_b = t2 + 1;
                const char *seq_e = (const char *)memchr(seq_b, '\t', size_t(re - seq_b));
                if (!seq_e) seq_e = re;
                int64_t n = 0;   // len() of a str counts code points
                for (const char *c = seq_b; c < seq_e; ++c) n += ((unsigned char)*c & 0xC0) != 0x80;
                std::string name(name_b, size_t(name_e - name_b));
                auto it = seen.find(name);
                if (it == seen.end()) {
                    seen.emplace(name, alts.size());
                    alts.push_back({std::move(name), n});
                } else {
                    alts[it->second].second = n;
                }
            }
        }
        p = next;
    }
    return true;
}

// ---------------------------------------------------------------------------
// build
// ---------------------------------------------------------------------------
static uint64_t hash_bytes(const char *s, size_t n) {
    TokHash h = tok_init();
    for (size_t i = 0; i < n; i += 4) {
        uint32_t w = 0;
        for (size_t k = 0; k < 4 && i + k < n; ++k) w |= uint32_t((unsigned char)s[i + k]) << (8 * k);
        tok_step(h, w);
    }
    return tok_value(h, uint32_t(n));
}

// exact key of a plain node name (see PNodeSlot); false for any other name.  Read from the end, as the
// scan kernel reads it: digits, '-' or '.', digits, ':', chrom.
static bool plain_key(const std::string &name, PNodeSlot &k) {
    const size_t n = name.size();
    auto digit = [&](size_t i) { return name[i] >= '0' && name[i] <= '9'; };
    size_t q = n;
    while (q > 0 && digit(q - 1)) --q;
    const size_t n2 = n - q;                                   // digits of the second number
    if (n2 < 1 || n2 > 9 || q == 0 || (name[q - 1] != '-' && name[q - 1] != '.')) return false;
    const size_t sep = q - 1;
    size_t p = sep;
    while (p > 0 && digit(p - 1)) --p;
    const size_t n1 = sep - p;                                 // digits of the first number
    if (n1 < 1 || n1 > 9 || p == 0 || name[p - 1] != ':') return false;
    const size_t c = p - 1;                                    // the colon; chrom = name[0, c)
    if (c > 15) return false;
    if ((n1 > 1 && name[p] == '0') || (n2 > 1 && name[q] == '0')) return false;
    uint8_t chrom[16] = {0};
    for (size_t i = 0; i < c; ++i) chrom[i] = uint8_t(name[i]);
    chrom[15] = uint8_t(c);
    uint32_t a = 0, b = 0;
    for (size_t i = p; i < sep; ++i) a = a * 10 + uint32_t(name[i] - '0');
    for (size_t i = q; i < n; ++i) b = b * 10 + uint32_t(name[i] - '0');
    memcpy(&k.c0, chrom, 8);
    memcpy(&k.c1, chrom + 8, 8);
    k.a = a;
    k.b = b | (name[sep] == '.' ? PN_ALT : 0u);
    return true;
}

static uint32_t pow2_at_least(uint64_t n) {
    uint64_t c = 2;
    while (c < n) c <<= 1;
    return uint32_t(c);
}

// strings -> dense ids in the order they first appear: open addressing over the ids, the hashes kept for the
// compares and for growing (std::unordered_map spends most of a million-SV build in its nodes)
struct FlatIndex {
    std::vector<uint32_t> slot;                 // id + 1, 0 = free
    std::vector<std::string_view> keys;         // views the caller keeps alive
    std::vector<uint64_t> hashes;
    explicit FlatIndex(size_t expect) { slot.assign(size_t(pow2_at_least(expect * 2 + 16)), 0); }
    static uint64_t hash(std::string_view v) {
        uint64_t h = 0xcbf29ce484222325ull;
        size_t i = 0;
        for (; i + 8 <= v.size(); i += 8) {
            uint64_t w;
            memcpy(&w, v.data() + i, 8);
            h = (h ^ w) * 0x9E3779B97F4A7C15ull;
            h ^= h >> 29;
        }
        uint64_t w = 0;
        memcpy(&w, v.data() + i, v.size() - i);
        h = (h ^ w ^ (uint64_t(v.size()) << 56)) * 0x9E3779B97F4A7C15ull;
        return h ^ (h >> 32);
    }
    void grow() {
        std::vector<uint32_t> bigger(slot.size() * 2, 0);
        const size_t mask = bigger.size() - 1;
        for (size_t id = 0; id < keys.size(); ++id) {
            size_t i = size_t(hashes[id]) & mask;
            while (bigger[i]) i = (i + 1) & mask;
            bigger[i] = uint32_t(id) + 1;
        }
        slot.swap(bigger);
    }
    uint32_t intern(std::string_view v) {
        const uint64_t h = hash(v);
        size_t mask = slot.size() - 1, i = size_t(h) & mask;
        while (slot[i]) {
            const uint32_t id = slot[i] - 1;
            if (hashes[id] == h && keys[id] == v) return id;
            i = (i + 1) & mask;
        }
        if ((keys.size() + 1) * 2 > slot.size()) {
            grow();
            mask = slot.size() - 1;
            i = size_t(h) & mask;
            while (slot[i]) i = (i + 1) & mask;
        }
        slot[i] = uint32_t(keys.size()) + 1;
        keys.push_back(v);
        hashes.push_back(h);
        return uint32_t(keys.size()) - 1;
    }
};

static bool build(svjg_tables *t, std::vector<RawKey> &keys, std::vector<std::pair<std::string, int64_t>> &alts,
                  std::string &err) {
    StageTimer tm;
    t->n_keys = uint32_t(keys.size());
    // entries on which the reference raises make skipping any probe unsafe
    for (auto &k : keys) {
        if (k.poison_key) t->filter_flags |= SVJG_FLAG_EXACT_CHECKS;
        for (auto &e : k.ents)
            if (e.second < 0) t->filter_flags |= SVJG_FLAG_EXACT_CHECKS;
    }

    // Two halves that do not depend on each other run side by side: (1) the sv ids and the entry lists,
    // (2) the readings of the link keys and the node names.
    std::vector<uint32_t> ent_begin(keys.size() + 1, 0);
    std::string err1;
    std::thread sv_side([&] {
        StageTimer tm1;
        // distinct sv ids, byte-sorted == the order json.dumps(sort_keys=True) prints them
        size_t n_ents = 0;
        for (auto &k : keys) n_ents += k.ents.size();
        FlatIndex ids(std::min<size_t>(n_ents, 1u << 20));   // views into keys[].ents, which outlive the build
        std::vector<uint32_t> ent_id;                          // per entry: its sv id in order of first appearance
        ent_id.reserve(n_ents);
        for (auto &k : keys)
            for (auto &e : k.ents) ent_id.push_back(e.second >= 0 ? ids.intern(std::string_view(e.first)) : 0u);
        if (ids.keys.size() >= 0x7FFFFFFFull) {
            err1 = "too many distinct sv ids";
            return;
        }
        std::vector<uint32_t> order(ids.keys.size()), rank(ids.keys.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = uint32_t(i);
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return ids.keys[a] < ids.keys[b]; });
        t->sv_ids.reserve(order.size());
        for (size_t i = 0; i < order.size(); ++i) {
            rank[order[i]] = uint32_t(i);
            t->sv_ids.emplace_back(ids.keys[order[i]]);
        }
        t->entries.reserve(n_ents + 1);
        size_t at = 0;
        for (size_t ki = 0; ki < keys.size(); ++ki) {
            ent_begin[ki] = uint32_t(t->entries.size());
            for (auto &e : keys[ki].ents) {
                t->entries.push_back(e.second < 0 ? ENTRY_POISON : rank[ent_id[at]] * 2 + uint32_t(e.second));
                ++at;
            }
        }
        ent_begin[keys.size()] = uint32_t(t->entries.size());
        tm1.lap("    (sv ids + entries)");
    });

    struct Parse {
        uint32_t key, split;   // nL = key[0:split), sL = key[split+1], nR = key[split+3 : len-2)
        uint32_t idl, idr;     // node ids of the two names
    };
    std::vector<Parse> parses;
    for (size_t ki = 0; ki < keys.size(); ++ki) {
        const std::string &k = keys[ki].key;
        size_t n = k.size();
        if (n < 7 || k[n - 2] != '@' || (k[n - 1] != '+' && k[n - 1] != '-')) continue;
        // every "@+@" / "@-@" with a non-empty name on both sides is a possible reading of the key
        for (size_t s = 1; s + 3 < n - 2; ++s) {
            if (k[s] == '@' && k[s + 2] == '@' && (k[s + 1] == '+' || k[s + 1] == '-')) parses.push_back({uint32_t(ki), uint32_t(s), 0, 0});
        }
    }

    // node names: every name a link key can be read with, plus the GFA's alt nodes
    // the views point into keys[].key and alts[].first, neither of which changes during the build
    FlatIndex node_id(keys.size() + alts.size());
    std::vector<std::pair<std::string, int64_t>> nodes;      // name, alt sequence length or -1
    auto intern = [&](std::string_view name) -> uint32_t {
        const uint32_t id = node_id.intern(name);
        if (id == nodes.size()) nodes.push_back({std::string(name), -1});
        return id;
    };
    for (auto &a : alts) nodes[intern(std::string_view(a.first))].second = a.second;

    std::vector<uint8_t> roles;                              // per node: bit s = left node with strand s, bit 2+s = right
    std::string err2;
    for (auto &ps : parses) {
        const std::string &k = keys[ps.key].key;
        size_t n = k.size();
        size_t len_l = ps.split, off_r = ps.split + 3, len_r = n - 2 - off_r;
        if (len_l > 0xFFFF || len_r > 0xFFFF) {
            err2 = "node name longer than 65535 bytes in svs_edges key";
            break;
        }
        uint32_t sl = k[ps.split + 1] == '+', sr = k[n - 1] == '+';
        const std::string_view kv(k);
        ps.idl = intern(kv.substr(0, len_l));
        ps.idr = intern(kv.substr(off_r, len_r));
        if (roles.size() < nodes.size()) roles.resize(nodes.size(), 0);
        roles[ps.idl] |= uint8_t(1u << sl);
        roles[ps.idr] |= uint8_t(4u << sr);
        if (nodes.size() >= 0x7FFFFFFFull) {
            err2 = "too many distinct node names";
            break;
        }
    }
    tm.lap("    (node names)");
    sv_side.join();
    tm.lap("  sv ids, entries | node names");
    if (!err1.empty() || !err2.empty()) {
        err = !err1.empty() ? err1 : err2;
        return false;
    }
    if (t->entries.empty()) t->entries.push_back(ENTRY_POISON);  // never empty on the device

    // the name-hash node table and the plain-node table need the nodes only: beside the link table
    t->n_alt = uint32_t(alts.size());
    t->n_nodes = uint32_t(nodes.size());
    std::string err3, err4;
    std::thread name_side([&] {
        uint32_t ncap = pow2_at_least(nodes.size() * 2 + 2);
        t->nodes.assign(ncap, NodeSlot{});
        for (size_t id = 0; id < nodes.size(); ++id) {
            const std::string &name = nodes[id].first;
            NodeSlot s{};
            s.hash = node_hash(hash_bytes(name.data(), name.size()));
            s.name_off = uint32_t(t->blob.size());
            s.name_len = uint32_t(name.size());
            s.seq_len = nodes[id].second;
            s.id1 = uint32_t(id) + 1;
            // every name starts on a 4-byte boundary and is zero padded: the kernel compares words
            t->blob.insert(t->blob.end(), name.begin(), name.end());
            t->blob.resize((t->blob.size() + 3) & ~size_t(3), 0);
            if (t->blob.size() >= 0xFFFF0000ull) {
                err3 = "node name blob exceeds 4 GiB";
                return;
            }
            uint32_t i = uint32_t(s.hash) & (ncap - 1);
            while (t->nodes[i].id1) i = (i + 1) & (ncap - 1);
            t->nodes[i] = s;
        }
        // pad the blob so 4-byte reads at the tail stay in bounds
        t->blob.insert(t->blob.end(), 16, 0);
    });
    // plain names once more under their exact key (same ids)
    std::thread plain_side([&] {
        std::vector<PNodeSlot> plain;
        for (size_t id = 0; id < nodes.size(); ++id) {
            PNodeSlot s{};
            if (!plain_key(nodes[id].first, s)) continue;
            if (id + 1 > PN_ID_MASK) {
                err4 = "too many distinct node names";
                return;
            }
            s.id1 = (uint32_t(id) + 1) | (uint32_t(id < roles.size() ? roles[id] : 0) << 28);
            const int64_t sl = nodes[id].second;
            s.alt_len = (sl > 0 && sl <= 0x7FFFFFFF) ? uint32_t(sl) : PN_NO_LEN;
            plain.push_back(s);
        }
        uint32_t pcap = pow2_at_least(plain.size() * 2 + 2);
        t->pnodes.assign(pcap, PNodeSlot{});
        for (auto &s : plain) {
            uint32_t i = pnode_hash(s.c0, s.c1, s.a, s.b) & (pcap - 1);
            while (t->pnodes[i].id1) i = (i + 1) & (pcap - 1);
            t->pnodes[i] = s;
        }
    });

    std::string err5;
    uint32_t cap = pow2_at_least(parses.size() * 2 + 2);
    t->links.assign(cap, LinkSlot{});
    for (auto &ps : parses) {
        const std::string &k = keys[ps.key].key;
        uint32_t sl = k[ps.split + 1] == '+', sr = k[k.size() - 1] == '+';
        uint32_t cnt = ent_begin[ps.key + 1] - ent_begin[ps.key];
        if (cnt >= (1u << 28)) {
            err5 = "too many entries under one link key";
            break;
        }
        LinkSlot s{};
        s.key = link_key(ps.idl, sl, ps.idr, sr);
        s.val = cnt == 1 ? t->entries[ent_begin[ps.key]] : ent_begin[ps.key];
        s.meta = (cnt << 4) | (keys[ps.key].poison_key ? 8u : 0u) | 1u;
        uint32_t i = link_hash(s.key) & (cap - 1);
        while (t->links[i].meta & 1u) i = (i + 1) & (cap - 1);
        t->links[i] = s;
    }
    t->n_link_slots = uint32_t(parses.size());
    name_side.join();
    plain_side.join();
    tm.lap("  link | name-hash | plain-node tables");
    for (const std::string *e : {&err5, &err3, &err4})
        if (!e->empty()) {
            err = *e;
            return false;
        }
    return true;
}

static bool read_file(const char *path, std::string &out) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    // a regular file: one read into a string of its size; anything else (a pipe): append as it comes
    if (fseek(f, 0, SEEK_END) == 0) {
        const long n = ftell(f);
        if (n > 0 && fseek(f, 0, SEEK_SET) == 0) {
            out.resize(size_t(n));
            const size_t got = fread(&out[0], 1, size_t(n), f);
            out.resize(got);
        } else {
            rewind(f);
        }
    }
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    bool ok = !ferror(f);
    fclose(f);
    return ok;
}

}  // namespace svjg

// ---------------------------------------------------------------------------
// informative_aln.json -> per-key list lengths   (predict-genotype.py:67-68, :219-226)
// ---------------------------------------------------------------------------
struct svjg_aln_counts {
    std::vector<std::string> keys;        // byte-sorted, distinct (the last duplicate wins, like a dict)
    std::vector<uint32_t> counts;         // [n][2]; UINT32_MAX where the reference would raise on lookup
};

namespace svjg {

// number of items `for x in value` yields in Python; false if not iterable (TypeError)
static bool iter_len(JsonIn &in, uint32_t &n) {
    in.ws();
    n = 0;
    if (in.p < in.end && *in.p == '[') {
        ++in.p;
        in.ws();
        if (in.p < in.end && *in.p == ']') {
            ++in.p;
            return true;
        }
        for (;;) {
            JsonIn::Kind k;
            if (!in.skip_value(k)) return true;   // error is recorded in `in`
            ++n;
            in.ws();
            if (in.p < in.end && *in.p == ',') {
                ++in.p;
                continue;
            }
            if (in.p < in.end && *in.p == ']') {
                ++in.p;
                return true;
            }
            in.fail("expected ',' or ']'");
            return true;
        }
    }
    JsonIn::Kind k;
    std::string sval;
    const char *before = in.p;
    if (!in.skip_value(k, nullptr, &sval)) return true;
    if (k == JsonIn::K_STRING) {          // iterating a str yields its code points
        for (unsigned char c : sval) n += (c & 0xC0) != 0x80;
        return true;
    }
    if (k == JsonIn::K_OBJECT) {          // iterating a dict yields its keys; count them by re-scanning
        JsonIn sub{before, in.p, {}};
        ++sub.p;
        sub.ws();
        if (sub.p < sub.end && *sub.p == '}') return true;
        std::unordered_set<std::string> seen;                  // a repeated key is one key of the dict
        for (;;) {
            std::string key;
            sub.ws();
            if (!sub.string(key)) return true;
            sub.ws();
            ++sub.p;                      // ':'
            JsonIn::Kind kk;
            if (!sub.skip_value(kk)) return true;
            if (seen.insert(key).second) ++n;
            sub.ws();
            if (sub.p < sub.end && *sub.p == ',') {
                ++sub.p;
                continue;
            }
            return true;
        }
    }
    return false;
}

static bool parse_aln_counts(const char *text, size_t len, svjg_aln_counts *out, std::string &err) {
    JsonIn in{text, text + len, {}};
    in.ws();
    if (in.p >= in.end || *in.p != '{') {
        err = "informative_aln: top level is not a JSON object";
        return false;
    }
    ++in.p;
    std::unordered_map<std::string, std::pair<uint32_t, uint32_t>> d;
    in.ws();
    if (in.p < in.end && *in.p == '}') {
        ++in.p;
    } else {
        for (;;) {
            in.ws();
            std::string key;
            if (!in.string(key)) break;
            in.ws();
            if (in.p >= in.end || *in.p != ':') {
                in.fail("expected ':'");
                break;
            }
            ++in.p;
            in.ws();
            std::pair<uint32_t, uint32_t> c{UINT32_MAX, UINT32_MAX};
            if (in.p < in.end && *in.p == '[') {
                // value[0] and value[1] are iterated (:220-221); anything after them is ignored
                ++in.p;
                uint32_t idx = 0, n0 = 0, n1 = 0;
                bool ok0 = false, ok1 = false;
                in.ws();
                if (in.p < in.end && *in.p == ']') {
                    ++in.p;
                } else {
                    for (;;) {
                        if (idx == 0) ok0 = iter_len(in, n0);
                        else if (idx == 1) ok1 = iter_len(in, n1);
                        else {
                            JsonIn::Kind k;
                            if (!in.skip_value(k)) break;
                        }
                        if (!in.err.empty()) break;
                        ++idx;
                        in.ws();
                        if (in.p < in.end && *in.p == ',') {
                            ++in.p;
                            continue;
                        }
                        if (in.p < in.end && *in.p == ']') {
                            ++in.p;
                            break;
                        }
                        in.fail("expected ',' or ']'");
                        break;
                    }
                }
                if (idx >= 2 && ok0 && ok1) c = {n0, n1};
            } else {
                JsonIn::Kind k;
                std::string sval;
                if (!in.skip_value(k, nullptr, &sval)) break;
                if (k == JsonIn::K_STRING) {            // "ab"[0] and "ab"[1] are 1-character strings
                    uint32_t cp = 0;
                    for (unsigned char ch : sval) cp += (ch & 0xC0) != 0x80;
                    if (cp >= 2) c = {1u, 1u};
                }
            }
            if (!in.err.empty()) break;
            d[key] = c;
            in.ws();
            if (in.p < in.end && *in.p == ',') {
                ++in.p;
                continue;
            }
            if (in.p < in.end && *in.p == '}') {
                ++in.p;
                break;
            }
            in.fail("expected ',' or '}'");
            break;
        }
    }
    if (!in.err.empty()) {
        err = "informative_aln: " + in.err;
        return false;
    }
    in.ws();
    if (in.p != in.end) {
        err = "informative_aln: trailing data after the JSON object";
        return false;
    }
    out->keys.reserve(d.size());
    for (auto &kv : d) out->keys.push_back(kv.first);
    std::sort(out->keys.begin(), out->keys.end());
    out->counts.reserve(out->keys.size() * 2);
    for (auto &k : out->keys) {
        auto &c = d[k];
        out->counts.push_back(c.first);
        out->counts.push_back(c.second);
    }
    return true;
}

}  // namespace svjg

using namespace svjg;

extern "C" int svjg_aln_counts_from_memory(const char *json, size_t len, svjg_aln_counts **out) {
    if (!json || !out) return set_error(SVJG_E_ARG, "svjg_aln_counts_from_memory: NULL argument");
    svjg_aln_counts *c = new svjg_aln_counts();
    std::string err;
    if (!parse_aln_counts(json, len, c, err)) {
        delete c;
        return set_error(SVJG_E_JSON, err);
    }
    *out = c;
    return SVJG_OK;
}
extern "C" int svjg_aln_counts_load(const char *json_path, svjg_aln_counts **out) {
    if (!json_path || !out) return set_error(SVJG_E_ARG, "svjg_aln_counts_load: NULL argument");
    std::string text;
    if (!read_file(json_path, text)) return set_error(SVJG_E_IO, std::string("cannot read ") + json_path);
    return svjg_aln_counts_from_memory(text.data(), text.size(), out);
}
extern "C" void svjg_aln_counts_free(svjg_aln_counts *c) { delete c; }
extern "C" uint32_t svjg_aln_counts_num(const svjg_aln_counts *c) { return c ? uint32_t(c->keys.size()) : 0; }
extern "C" const char *svjg_aln_counts_key(const svjg_aln_counts *c, uint32_t i, uint32_t *len) {
    if (!c || i >= c->keys.size()) return nullptr;
    if (len) *len = uint32_t(c->keys[i].size());
    return c->keys[i].data();
}
extern "C" const uint32_t *svjg_aln_counts_data(const svjg_aln_counts *c) { return c ? c->counts.data() : nullptr; }
extern "C" uint32_t svjg_aln_counts_find(const svjg_aln_counts *c, const char *key, uint32_t len) {
    if (!c || !key) return UINT32_MAX;
    std::string k(key, len);
    auto it = std::lower_bound(c->keys.begin(), c->keys.end(), k);
    if (it == c->keys.end() || *it != k) return UINT32_MAX;
    return uint32_t(it - c->keys.begin());
}

extern "C" uint64_t svjg_tables_image_hash(const svjg_tables *t) {
    if (!t) return 0;
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void *p, size_t n) {
        const unsigned char *b = static_cast<const unsigned char *>(p);
        for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
        h = (h ^ n) * 1099511628211ull;
    };
    mix(t->links.data(), t->links.size() * sizeof(LinkSlot));
    mix(t->nodes.data(), t->nodes.size() * sizeof(NodeSlot));
    mix(t->pnodes.data(), t->pnodes.size() * sizeof(PNodeSlot));
    mix(t->blob.data(), t->blob.size());
    mix(t->entries.data(), t->entries.size() * sizeof(uint32_t));
    for (const std::string &s : t->sv_ids) mix(s.data(), s.size());
    const uint32_t tail[5] = {t->n_keys, t->n_link_slots, t->n_alt, t->n_nodes, t->filter_flags};
    mix(tail, sizeof tail);
    return h;
}

extern "C" int64_t svjg_tables_alt_node_len(const svjg_tables *t, const char *name, uint32_t len) {
    if (!t || (!name && len) || t->nodes.empty()) return -1;
    // the probe the kernels make: token hash of the name, then the name bytes
    const uint64_t h = node_hash(hash_bytes(name, len));
    const uint32_t mask = uint32_t(t->nodes.size()) - 1;
    int64_t found = -1;
    for (uint32_t i = uint32_t(h) & mask; t->nodes[i].id1; i = (i + 1) & mask) {
        const NodeSlot &s = t->nodes[i];
        if (s.hash == h && s.name_len == len && (len == 0 || memcmp(t->blob.data() + s.name_off, name, len) == 0)) {
            found = s.seq_len;
            break;
        }
    }
    // a plain name is resolved through its exact key on the fast path: both tables must agree
    PNodeSlot k{};
    if (plain_key(std::string(name ? name : "", len), k)) {
        const uint32_t pmask = uint32_t(t->pnodes.size()) - 1;
        int64_t pfound = -1;
        for (uint32_t i = pnode_hash(k.c0, k.c1, k.a, k.b) & pmask; t->pnodes[i].id1; i = (i + 1) & pmask) {
            const PNodeSlot &s = t->pnodes[i];
            if (s.c0 == k.c0 && s.c1 == k.c1 && s.a == k.a && s.b == k.b) {
                pfound = s.alt_len == PN_NO_LEN ? -1 : int64_t(s.alt_len);
                break;
            }
        }
        const int64_t want = (found > 0 && found <= 0x7FFFFFFF) ? found : -1;   // what the fast path may hold
        if (pfound != want) return -2;
    }
    return found;
}

extern "C" int svjg_tables_from_memory(const char *edges_json, size_t edges_len, const char *gfa, size_t gfa_len,
                                       svjg_tables **out) {
    if (!edges_json || !out || (!gfa && gfa_len)) return set_error(SVJG_E_ARG, "svjg_tables_from_memory: NULL argument");
    std::vector<RawKey> keys;
    std::vector<std::pair<std::string, int64_t>> alts;
    std::string err;
    StageTimer tm;
    // the GFA is scanned on a thread of its own while the JSON is parsed
    std::string gfa_err;
    bool gfa_ok = true;
    std::thread gfa_side([&] { gfa_ok = scan_gfa(gfa, gfa_len, alts, gfa_err); });
    const bool edges_ok = parse_edges(edges_json, edges_len, keys, err);
    gfa_side.join();
    if (!edges_ok) return set_error(SVJG_E_JSON, err);
    if (!gfa_ok) return set_error(SVJG_E_INPUT, gfa_err);
    tm.lap("svs_edges.json parse | GFA scan");
    svjg_tables *t = new svjg_tables();
    if (!build(t, keys, alts, err)) {
        delete t;
        return set_error(SVJG_E_ARG, err);
    }
    tm.lap("table build");
    *out = t;
    return SVJG_OK;
}

extern "C" int svjg_tables_load(const char *svs_edges_json_path, const char *gfa_path, svjg_tables **out) {
    if (!svs_edges_json_path || !gfa_path || !out) return set_error(SVJG_E_ARG, "svjg_tables_load: NULL argument");
    std::string edges, gfa;
    bool gfa_read = false;
    std::thread gfa_side([&] { gfa_read = read_file(gfa_path, gfa); });
    const bool edges_read = read_file(svs_edges_json_path, edges);
    gfa_side.join();
    if (!edges_read) return set_error(SVJG_E_IO, std::string("cannot read ") + svs_edges_json_path);
    if (!gfa_read) return set_error(SVJG_E_IO, std::string("cannot read ") + gfa_path);
    return svjg_tables_from_memory(edges.data(), edges.size(), gfa.data(), gfa.size(), out);
}

// a second handle with the same host image and no device image: one handle per GPU of a node
extern "C" int svjg_tables_clone(const svjg_tables *t, svjg_tables **out) {
    if (!t || !out) return set_error(SVJG_E_ARG, "svjg_tables_clone: NULL argument");
    svjg_tables *c = new svjg_tables();
    c->sv_ids = t->sv_ids;
    c->links = t->links;
    c->nodes = t->nodes;
    c->pnodes = t->pnodes;
    c->blob = t->blob;
    c->entries = t->entries;
    c->n_keys = t->n_keys;
    c->n_link_slots = t->n_link_slots;
    c->n_alt = t->n_alt;
    c->n_nodes = t->n_nodes;
    c->filter_flags = t->filter_flags;
    *out = c;
    return SVJG_OK;
}

extern "C" int svjg_tables_set_flags(svjg_tables *t, uint32_t flags) {
    if (!t) return set_error(SVJG_E_ARG, "svjg_tables_set_flags: NULL tables");
    t->filter_flags |= flags & (SVJG_FLAG_EXACT_CHECKS | SVJG_FLAG_FORCE_GENERAL);
    return SVJG_OK;
}

extern "C" uint32_t svjg_tables_num_sv(const svjg_tables *t) { return t ? uint32_t(t->sv_ids.size()) : 0; }
extern "C" uint32_t svjg_tables_num_links(const svjg_tables *t) { return t ? t->n_keys : 0; }
extern "C" uint32_t svjg_tables_num_alt_nodes(const svjg_tables *t) { return t ? t->n_alt : 0; }
extern "C" uint64_t svjg_tables_device_bytes(const svjg_tables *t) {
    if (!t) return 0;
    return t->links.size() * sizeof(LinkSlot) + t->nodes.size() * sizeof(NodeSlot) + t->blob.size() +
           t->entries.size() * sizeof(uint32_t) + t->pnodes.size() * sizeof(PNodeSlot);
}
extern "C" const char *svjg_tables_sv_id(const svjg_tables *t, uint32_t i, uint32_t *len) {
    if (!t || i >= t->sv_ids.size()) return nullptr;
    if (len) *len = uint32_t(t->sv_ids[i].size());
    return t->sv_ids[i].data();
}
extern "C" uint32_t svjg_tables_find_sv(const svjg_tables *t, const char *sv_id, uint32_t len) {
    if (!t || !sv_id) return UINT32_MAX;
    std::string key(sv_id, len);
    auto it = std::lower_bound(t->sv_ids.begin(), t->sv_ids.end(), key);
    if (it == t->sv_ids.end() || *it != key) return UINT32_MAX;
    return uint32_t(it - t->sv_ids.begin());
}
