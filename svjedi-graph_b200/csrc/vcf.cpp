// VCF side of the genotype stage on the host: the SV key of every record (predict-genotype.py:118-211),
// the gate inputs of :216 and the output text of :248-271, without a Python-level loop over the records
// (C5: 1 M records).  svjg/genotype.py::parse_vcf / format_vcf are the line-by-line statement of the same
// rules; tests/test_vcf_native.py holds the two against each other on every fixture.
//
// Only spellings whose Python meaning is reproduced exactly are accepted: a file with a non-ASCII
// byte, or a POS/END of more than 18 digits, is answered with SVJG_E_UNSUPPORTED and the caller uses
// the Python statement instead.
#include <cstdio>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <thread>
#include <vector>

#include "../../include/svjg.h"
#include "svjg_internal.h"

using svjg::set_error;
using sv = std::string_view;

namespace {

const char FORMAT_LINES[] =
    "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
    "##FORMAT=<ID=DP,Number=1,Type=Float,Description=\"Total number of informative read alignments across all alleles "
    "(after normalization for unbalanced SVs)\">\n"
    "##FORMAT=<ID=AD,Number=2,Type=Float,Description=\"Number of informative read alignments supporting each allele "
    "(after normalization by breakpoint number for unbalanced SVs)\">\n"
    "##FORMAT=<ID=PL,Number=3,Type=Integer,Description=\"Phred-scaled likelihood for each genotype\">\n"
    "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE\n";

struct Rec {
    uint64_t head_off;   // into text
    uint32_t head_len;
    uint32_t key_len;    // UINT32_MAX: no key
    uint64_t key_off;    // into keys
};
struct Hdr {
    uint32_t before;     // printed in front of this record index
    uint32_t len;        // UINT32_MAX: the FORMAT block + column header
    uint64_t off;
};

inline bool starts_with(sv s, sv p) { return s.size() >= p.size() && memcmp(s.data(), p.data(), p.size()) == 0; }

// s.split(sep)[1] (npos-safe): the text between the first and the second occurrence of sep; false if sep is absent
bool second_piece(sv s, sv sep, sv &out) {
    size_t a = s.find(sep);
    if (a == sv::npos) return false;
    a += sep.size();
    size_t b = s.find(sep, a);
    out = s.substr(a, b == sv::npos ? sv::npos : b - a);
    return true;
}
inline sv upto_semicolon(sv s) {
    size_t p = s.find(';');
    return p == sv::npos ? s : s.substr(0, p);
}
inline sv first_field(sv info) { return upto_semicolon(info); }
inline sv last_field(sv info) {
    size_t p = info.rfind(';');
    return p == sv::npos ? info : info.substr(p + 1);
}

// get_info of predict-genotype.py:77-87; false where the reference raises IndexError
bool info_value(sv info, sv tag /* "END=" */, sv sep_tag /* ";END=" */, sv &out) {
    sv piece;
    if (starts_with(first_field(info), tag)) {
        second_piece(info, tag, piece);
        out = upto_semicolon(piece);
        return true;
    }
    if (!second_piece(info, sep_tag, piece)) return false;
    out = starts_with(last_field(info), tag) ? piece : upto_semicolon(piece);
    return true;
}

inline bool py_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13) || (c >= 0x1c && c <= 0x1f); }

// int(str) for ASCII text.  0 ok, 1 ValueError, 2 more than 18 digits
int py_int(sv s, int64_t &out) {
    while (!s.empty() && py_space((unsigned char)s.front())) s.remove_prefix(1);
    while (!s.empty() && py_space((unsigned char)s.back())) s.remove_suffix(1);
    bool neg = false;
    if (!s.empty() && (s.front() == '+' || s.front() == '-')) {
        neg = s.front() == '-';
        s.remove_prefix(1);
    }
    if (s.empty()) return 1;
    int64_t v = 0;
    int digits = 0;
    bool prev_digit = false;
    for (char c : s) {
        if (c >= '0' && c <= '9') {
            if (v || c != '0') {
                if (++digits > 18) return 2;
            }
            v = v * 10 + (c - '0');
            prev_digit = true;
        } else if (c == '_' && prev_digit) {
            prev_digit = false;            // a single underscore between two digits
        } else {
            return 1;
        }
    }
    if (!prev_digit) return 1;             // trailing underscore
    out = neg ? -v : v;
    return 0;
}

}  // namespace

struct svjg_vcf {
    std::string text;                // the file as text mode reads it ("\r\n" and "\r" are "\n")
    std::string keys;
    std::vector<Rec> recs;
    std::vector<uint8_t> type;
    std::vector<Hdr> hdrs;
};

#define SVJG_VCF_FAIL(msg)                                                                             \
    do {                                                                                               \
        delete v;                                                                                      \
        return set_error(SVJG_E_INPUT, std::string("VCF line ") + std::to_string(line_no) + ": " + msg); \
    } while (0)

extern "C" int svjg_vcf_parse(const char *data, size_t len, int translate_cr, svjg_vcf **out) {
    if ((!data && len) || !out) return set_error(SVJG_E_ARG, "svjg_vcf_parse: NULL argument");
    {
        uint64_t acc = 0;
        size_t i = 0;
        for (; i + 8 <= len; i += 8) {
            uint64_t w;
            memcpy(&w, data + i, 8);
            acc |= w;
        }
        for (; i < len; ++i) acc |= uint64_t((unsigned char)data[i]);
        if (acc & 0x8080808080808080ull)
            return set_error(SVJG_E_UNSUPPORTED, "VCF with non-ASCII bytes: left to the Python statement of the rules");
    }
    svjg_vcf *v = new svjg_vcf();
    if (translate_cr && len && memchr(data, '\r', len)) {
        v->text.reserve(len);
        for (size_t i = 0; i < len; ++i) {
            if (data[i] == '\r') {
                v->text.push_back('\n');
                if (i + 1 < len && data[i + 1] == '\n') ++i;
            } else {
                v->text.push_back(data[i]);
            }
        }
    } else {
        v->text.assign(data ? data : "", len);
    }
    const std::string &t = v->text;
    std::unordered_map<std::string, uint32_t> ins_seen;   // running count per POS string, any chromosome (:151-155)
    size_t p = 0, line_no = 0;
    std::string key;
    while (p < t.size()) {
        ++line_no;
        size_t nl = t.find('\n', p);
        const size_t next = nl == std::string::npos ? t.size() : nl + 1;
        const sv line(t.data() + p, next - p);                                  // with its "\n"
        const sv body(t.data() + p, (nl == std::string::npos ? t.size() : nl) - p);   // line.rstrip("\n")
        p = next;
        if (starts_with(line, "##FORMAT")) continue;
        if (starts_with(line, "##")) {
            v->hdrs.push_back({uint32_t(v->recs.size()), uint32_t(line.size()), uint64_t(line.data() - t.data())});
            continue;
        }
        if (starts_with(line, "#C")) {
            v->hdrs.push_back({uint32_t(v->recs.size()), UINT32_MAX, 0});
            continue;
        }
        // columns 0..7 and where the 8th ends
        sv col[8];
        size_t a = 0;
        int nc = 0;
        size_t end8 = sv::npos;      // offset of the tab after column 7, npos if the line has exactly 8 columns
        while (nc < 8) {
            size_t tab = body.find('\t', a);
            if (tab == sv::npos) {
                col[nc++] = body.substr(a);
                a = sv::npos;
                break;
            }
            col[nc++] = body.substr(a, tab - a);
            a = tab + 1;
            if (nc == 8) end8 = tab;
        }
        if (nc < 8) SVJG_VCF_FAIL("fewer than 8 columns");
        const sv chrom = col[0], pos = col[1], alt = col[4], info = col[7];
        sv svtype;
        if (info.find("SVTYPE") != sv::npos) {
            sv piece;
            if (!second_piece(info, "SVTYPE=", piece)) SVJG_VCF_FAIL("SVTYPE without a value");
            svtype = starts_with(last_field(info), "SVTYPE=") ? piece : upto_semicolon(piece);
        }
        const bool is_del = svtype == "DEL", is_inv = svtype == "INV", is_ins = svtype == "INS", is_bnd = svtype == "BND";
        bool has_key = false;
        bool is_short = false;          // |length| < 50
        key.clear();
        if (!is_bnd && !is_ins) {
            sv end;
            if (!info_value(info, "END=", ";END=", end)) SVJG_VCF_FAIL("INFO has no END=");
            if (is_del || is_inv) {
                int64_t e = 0, s = 0;
                const int r1 = py_int(end, e), r2 = py_int(pos, s);
                if (r1 == 1) SVJG_VCF_FAIL("non-integer POS/END");      // int(end) is evaluated first
                if (r1 == 2) {
                    delete v;
                    return set_error(SVJG_E_UNSUPPORTED, "POS/END of more than 18 digits: left to the Python statement of the rules");
                }
                if (r2 == 1) SVJG_VCF_FAIL("non-integer POS/END");
                if (r2 == 2) {
                    delete v;
                    return set_error(SVJG_E_UNSUPPORTED, "POS/END of more than 18 digits: left to the Python statement of the rules");
                }
                const int64_t d = e - s;
                is_short = d > -50 && d < 50;
                key.append(chrom).append(":").append(svtype).append("-").append(pos).append("-").append(end);
                has_key = true;
            }
        } else if (is_ins) {
            const uint32_t k = ++ins_seen[std::string(pos)];
            key.append(chrom).append(":INS-").append(pos).append("-").append(std::to_string(k));
            has_key = true;
            is_short = alt.size() < 50;                                   // len(ALT), ASCII
        } else {
            key = "wrong_format";
            has_key = true;
            for (char br : {'[', ']'}) {
                if (alt.find(br) == sv::npos) continue;
                sv piece[2];
                int np = 0;
                size_t b = 0;
                while (b <= alt.size() && np < 2) {                       // the first two non-empty pieces of alt.split(br)
                    size_t e = alt.find(br, b);
                    if (e == sv::npos) e = alt.size();
                    if (e > b) piece[np++] = alt.substr(b, e - b);
                    b = e + 1;
                }
                if (np < 2) SVJG_VCF_FAIL("malformed BND ALT");
                key.clear();
                key.append(chrom).append(":BND-");
                if (piece[1].find(':') != sv::npos)
                    key.append(pos).append(1, br).append(piece[1]).append(1, br);
                else
                    key.append(1, br).append(piece[0]).append(1, br).append(pos);
                break;
            }
        }
        uint8_t code = is_del ? 0 : is_ins ? 1 : is_inv ? 2 : is_bnd ? 3 : 255;
        if (code != 255 && is_short) code |= 0x80;
        Rec r{};
        r.head_off = uint64_t(body.data() - t.data());
        r.head_len = uint32_t(end8 == sv::npos ? body.size() : end8);
        if (has_key) {
            r.key_off = v->keys.size();
            r.key_len = uint32_t(key.size());
            v->keys.append(key);
        } else {
            r.key_len = UINT32_MAX;
        }
        if (v->recs.size() >= 0xFFFFFFF0u) SVJG_VCF_FAIL("too many records");
        v->recs.push_back(r);
        v->type.push_back(code);
    }
    *out = v;
    return SVJG_OK;
}

extern "C" void svjg_vcf_free(svjg_vcf *v) { delete v; }
extern "C" uint32_t svjg_vcf_num_records(const svjg_vcf *v) { return v ? uint32_t(v->recs.size()) : 0; }
extern "C" const uint8_t *svjg_vcf_svtype(const svjg_vcf *v) { return v ? v->type.data() : nullptr; }
extern "C" const char *svjg_vcf_key(const svjg_vcf *v, uint32_t i, uint32_t *len) {
    if (!v || i >= v->recs.size() || v->recs[i].key_len == UINT32_MAX) return nullptr;
    if (len) *len = v->recs[i].key_len;
    return v->keys.data() + v->recs[i].key_off;
}

extern "C" int svjg_vcf_index_tables(const svjg_vcf *v, const svjg_tables *t, uint32_t *sv_index) {
    if (!v || !t || (!sv_index && !v->recs.empty())) return set_error(SVJG_E_ARG, "svjg_vcf_index_tables: NULL argument");
    for (size_t i = 0; i < v->recs.size(); ++i) {
        const Rec &r = v->recs[i];
        sv_index[i] = r.key_len == UINT32_MAX ? UINT32_MAX : svjg_tables_find_sv(t, v->keys.data() + r.key_off, r.key_len);
    }
    return SVJG_OK;
}

extern "C" int svjg_vcf_index_counts(const svjg_vcf *v, const svjg_aln_counts *c, uint32_t *sv_index, uint8_t *svtype) {
    if (!v || !c || ((!sv_index || !svtype) && !v->recs.empty())) return set_error(SVJG_E_ARG, "svjg_vcf_index_counts: NULL argument");
    const uint32_t *counts = svjg_aln_counts_data(c);
    for (size_t i = 0; i < v->recs.size(); ++i) {
        const Rec &r = v->recs[i];
        uint8_t ty = v->type[i];
        sv_index[i] = UINT32_MAX;
        svtype[i] = ty;
        if (r.key_len == UINT32_MAX) continue;
        const uint32_t j = svjg_aln_counts_find(c, v->keys.data() + r.key_off, r.key_len);
        if (j == UINT32_MAX) continue;
        const bool gated = ty != 255 && (ty & 0x3F) <= 3 && !(ty & 0x80);
        if (gated && counts[size_t(j) * 2] == UINT32_MAX)
            return set_error(SVJG_E_INPUT, "informative_aln entry of '" + std::string(v->keys.data() + r.key_off, r.key_len) +
                                               "' is not a pair of lists (the reference raises here)");
        sv_index[i] = j;
        if (ty != 255) svtype[i] = ty | 0x40;
    }
    return SVJG_OK;
}

namespace {
inline char *put_u64(char *o, uint64_t x) {
    char b[24];
    int n = 0;
    do {
        b[n++] = char('0' + x % 10);
        x /= 10;
    } while (x);
    while (n) *o++ = b[--n];
    return o;
}
inline char *put_i64(char *o, int64_t x) {
    if (x < 0) {
        *o++ = '-';
        return put_u64(o, uint64_t(0) - uint64_t(x));
    }
    return put_u64(o, uint64_t(x));
}
// str() of a count held in half units: int when the allele was not normalised, else one decimal
inline char *put_half(char *o, uint64_t twice, bool halved) {
    o = put_u64(o, twice >> 1);
    if (halved) {
        *o++ = '.';
        *o++ = (twice & 1) ? '5' : '0';
    }
    return o;
}
inline char *put_mem(char *o, const char *p, size_t n) {
    memcpy(o, p, n);
    return o + n;
}
}  // namespace

extern "C" int svjg_vcf_format(const svjg_vcf *v, const uint8_t *gt, const uint8_t *flags, const uint32_t *ad2,
                               const int64_t *pl, char **out, uint64_t *out_len, uint64_t *n_genotyped) {
    if (!v || !out || !out_len) return set_error(SVJG_E_ARG, "svjg_vcf_format: NULL argument");
    const size_t n = v->recs.size();
    if (n && (!gt || !flags || !ad2 || !pl)) return set_error(SVJG_E_ARG, "svjg_vcf_format: NULL argument");
    for (size_t i = 0; i < n; ++i)
        if ((flags[i] & SVJG_GT_GENOTYPED) && gt[i] > 3) return set_error(SVJG_E_ARG, "svjg_vcf_format: gt code out of range");
    static const char *GT_TEXT[4] = {"0/0", "0/1", "1/1", "./."};
    // upper bound: every head and header line is a piece of the text; the sample column is at most
    // 13 + 3 + 1 + 22 + 1 + 45 + 1 + 62 + 1 bytes; untouched pages of the block cost nothing
    size_t cap = v->text.size() + n * 160 + 16;
    for (const Hdr &h : v->hdrs)
        if (h.len == UINT32_MAX) cap += sizeof FORMAT_LINES;
    char *buf = (char *)malloc(cap);
    if (!buf) return set_error(SVJG_E_NOMEM, "svjg_vcf_format: out of memory");
    const char *text = v->text.data();
    auto put_hdr = [&](char *o, const Hdr &h) {
        return h.len == UINT32_MAX ? put_mem(o, FORMAT_LINES, sizeof FORMAT_LINES - 1) : put_mem(o, text + h.off, h.len);
    };
    // the sample column of record i (at most 150 bytes)
    auto put_sample = [&](char *o, size_t i) {
        o = put_mem(o, "\tGT:DP:AD:PL\t", 13);
        const uint8_t f = flags[i];
        if (f & SVJG_GT_GENOTYPED) {
            const uint64_t t1 = ad2[2 * i], t2 = ad2[2 * i + 1];
            const bool h0 = f & SVJG_GT_HALVED_0, h1 = f & SVJG_GT_HALVED_1;
            o = put_mem(o, GT_TEXT[gt[i]], 3);
            *o++ = ':';
            o = put_half(o, t1 + t2, h0 || h1);
            *o++ = ':';
            o = put_half(o, t1, h0);
            *o++ = ',';
            o = put_half(o, t2, h1);
            *o++ = ':';
            o = put_i64(o, pl[3 * i]);
            *o++ = ',';
            o = put_i64(o, pl[3 * i + 1]);
            *o++ = ',';
            o = put_i64(o, pl[3 * i + 2]);
        } else {
            o = put_mem(o, "./.:0:0,0:.,.,.", 15);
        }
        *o++ = '\n';
        return o;
    };
    // where every record starts: its head is copied as it is, the sample column is rendered once to learn
    // its length (cheap next to the copies); then the records are filled in by several threads
    std::vector<uint64_t> at(n + 1, 0);
    uint64_t genotyped = 0;
    {
        char tmp[192];
        size_t hi = 0;
        uint64_t pos = 0;
        for (size_t i = 0; i < n; ++i) {
            while (hi < v->hdrs.size() && v->hdrs[hi].before <= i) {
                const Hdr &h = v->hdrs[hi++];
                pos += h.len == UINT32_MAX ? sizeof FORMAT_LINES - 1 : h.len;
            }
            at[i] = pos;
            genotyped += (flags[i] & SVJG_GT_GENOTYPED) != 0;
            pos += v->recs[i].head_len + uint64_t(put_sample(tmp, i) - tmp);
        }
        at[n] = pos;
    }
    auto fill = [&](size_t lo, size_t hi_rec) {
        size_t hi = 0;
        while (hi < v->hdrs.size() && v->hdrs[hi].before < lo) ++hi;          // header lines in front of record lo: not ours
        // (those with before == lo sit right in front of record lo: ours, they end where it starts)
        for (size_t i = lo; i < hi_rec; ++i) {
            uint64_t back = 0;
            size_t h2 = hi;
            while (h2 < v->hdrs.size() && v->hdrs[h2].before <= i) {
                back += v->hdrs[h2].len == UINT32_MAX ? sizeof FORMAT_LINES - 1 : v->hdrs[h2].len;
                ++h2;
            }
            char *o = buf + at[i] - back;
            while (hi < h2) o = put_hdr(o, v->hdrs[hi++]);
            const Rec &r = v->recs[i];
            o = put_mem(o, text + r.head_off, r.head_len);
            put_sample(o, i);
        }
    };
    const uint64_t body = at[n];
    unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (body < (4u << 20) || n < 64) n_thr = 1;
    if (n_thr == 1) {
        fill(0, n);
    } else {
        // ranges of about equal bytes
        std::vector<size_t> cutv(n_thr + 1, 0);
        for (unsigned p = 1; p < n_thr; ++p)
            cutv[p] = size_t(std::lower_bound(at.begin(), at.begin() + n, body * p / n_thr) - at.begin());
        cutv[n_thr] = n;
        for (unsigned p = 1; p <= n_thr; ++p) cutv[p] = std::max(cutv[p], cutv[p - 1]);
        std::vector<std::thread> pool;
        for (unsigned p = 1; p < n_thr; ++p) pool.emplace_back(fill, cutv[p], cutv[p + 1]);
        fill(cutv[0], cutv[1]);
        for (auto &th : pool) th.join();
    }
    // header lines behind the last record
    char *o = buf + at[n];
    {
        size_t hi = 0;
        while (hi < v->hdrs.size() && (n == 0 ? false : v->hdrs[hi].before <= n - 1)) ++hi;
        while (hi < v->hdrs.size()) o = put_hdr(o, v->hdrs[hi++]);
    }
    *out = buf;
    *out_len = uint64_t(o - buf);
    if (n_genotyped) *n_genotyped = genotyped;
    return SVJG_OK;
}

extern "C" void svjg_buffer_free(char *p) { free(p); }
