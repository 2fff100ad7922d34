// informative_aln.json rendered on the device (filter-alignments.py:160-175):
//   json.dumps({sv id: [[ref lines], [alt lines]]}, sort_keys=True, indent=4)
// from the hit tuples the filter left in device memory and the GAF bytes resident there.
//
//   1  the counters ARE the list lengths: an exclusive scan gives every (sv, allele) list its segment
//      of the ordered hit array; the hits are scattered into their segments (any order) ...
//   2  ... and ranked inside them by line offset = the order the reference appends in (:166)
//   3  one warp per hit: the text it stores (the line up to the first "cg:Z:", :166) and the bytes its
//      JSON string takes (json.dumps escapes: \" \\ \n \r \t \b \f, \u00XX for other control bytes)
//   4  scans: element sizes -> element positions, key sizes -> key positions, total size
//   5  one warp per hit writes its element, one thread per key the framing around its two lists
// The text crosses PCIe once, as text.  Anything the renderer does not restate -- a non-ASCII byte in a
// stored line (Python decodes UTF-8), a list beyond 64 Ki entries (the ranking is quadratic in the list)
// -- makes the call decline (SVJG_E_UNSUPPORTED) and the caller use svjg_emit_informative_json on the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "svjg_internal.h"

using namespace svjg;

namespace {

constexpr int T = 256;                          // block size of the flat kernels
constexpr uint32_t MAX_LIST = 65536;            // longest list ranked here

// ---- exclusive scan of n u64 values (in place), two levels; *total = the sum
constexpr int SCAN_ITEMS = 8;                   // per thread
constexpr int SCAN_BLOCK = T * SCAN_ITEMS;      // per block

__global__ void __launch_bounds__(T) scan_partial(const uint64_t *in, uint64_t n, uint64_t *block_sum) {
    __shared__ uint64_t warp_sum[T / 32];
    const uint64_t base = uint64_t(blockIdx.x) * SCAN_BLOCK + uint64_t(threadIdx.x) * SCAN_ITEMS;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < T / 32; ++w) t += warp_sum[w];
        block_sum[blockIdx.x] = t;
    }
}
// one block: exclusive scan of the block sums in place; the grand total behind them
__global__ void __launch_bounds__(1024) scan_sums(uint64_t *block_sum, uint32_t n_blocks) {
    __shared__ uint64_t carry, warp_tot[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < n_blocks; b0 += 1024) {
        const uint32_t i = b0 + threadIdx.x;
        const uint64_t v = i < n_blocks ? block_sum[i] : 0;
        uint64_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (int(threadIdx.x & 31) >= d) incl += o;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        uint64_t before = carry;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += warp_tot[w];
        if (i < n_blocks) block_sum[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sum[n_blocks] = carry;
}
__global__ void __launch_bounds__(T) scan_final(uint64_t *data, uint64_t n, const uint64_t *block_sum) {
    __shared__ uint64_t warp_tot[T / 32];
    const uint64_t base = uint64_t(blockIdx.x) * SCAN_BLOCK + uint64_t(threadIdx.x) * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? data[base + k] : 0;
        s += v[k];
    }
    uint64_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (int(threadIdx.x & 31) >= d) incl += o;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint64_t before = block_sum[blockIdx.x];
    for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) before += warp_tot[w];
    uint64_t run = before + incl - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

// data[0..n) -> exclusive prefix sums in place, data[n] = total (data has n + 1 entries); tmp: n / SCAN_BLOCK + 2 entries
void exclusive_scan(uint64_t *data, uint64_t n, uint64_t *tmp, cudaStream_t st) {
    const uint32_t nb = uint32_t((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    if (nb == 0) {
        cudaMemsetAsync(data, 0, 8, st);
        return;
    }
    scan_partial<<<nb, T, 0, st>>>(data, n, tmp);
    scan_sums<<<1, 1024, 0, st>>>(tmp, nb);
    scan_final<<<nb, T, 0, st>>>(data, n, tmp);
    cudaMemcpyAsync(data + n, tmp + nb, 8, cudaMemcpyDeviceToDevice, st);
}

// ---- 1: list lengths -> u64, for the scan that gives the segments
__global__ void widen_counts(const uint32_t *counts, uint64_t n2, uint64_t *seg, uint32_t *flags) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n2) {
        const uint32_t c = counts[i];
        seg[i] = c;
        if (c > MAX_LIST) atomicOr(flags, 2u);
    }
}
// every hit into its list's segment, in the order the atomics fall
__global__ void scatter_hits(const uint32_t *sv2, const uint64_t *off, const uint32_t *len, uint64_t n, const uint64_t *seg,
                             uint32_t *cursor, uint64_t *s_off, uint32_t *s_len) {
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = sv2[i];
    const uint64_t p = seg[k] + atomicAdd(cursor + k, 1u);
    s_off[p] = off[i];
    s_len[p] = len[i];
}
// 2: a hit's place in its list = the number of hits of the list in front of it in the file (equal
// offsets -- a line appended twice -- are the same text: any order among them)
__global__ void rank_hits(const uint64_t *seg, uint32_t n2, const uint64_t *s_off, const uint32_t *s_len, uint64_t *r_off,
                          uint32_t *r_len, uint32_t *r_key) {
    // one warp per list, lists of any length up to MAX_LIST
    const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (k >= n2) return;
    const uint64_t b = seg[k], e = seg[k + 1];
    for (uint64_t i = b + lane; i < e; i += 32) {
        const uint64_t mine = s_off[i];
        uint64_t r = 0;
        for (uint64_t j = b; j < e; ++j) {
            const uint64_t o = s_off[j];
            r += (o < mine) || (o == mine && j < i);
        }
        r_off[b + r] = mine;
        r_len[b + r] = s_len[i];
        r_key[b + r] = k;
    }
}

// the bytes of the file at offset `off`: in the range that holds it (this device's memory or a peer's)
__device__ __forceinline__ const uint8_t *line_ptr(const LineSrc &s, uint64_t off) {
    uint32_t k = 0;
    for (uint32_t j = 1; j < s.n; ++j) k += off >= s.start[j];
    return s.base[k] + (off - s.start[k]);
}

// bytes a byte takes inside a JSON string (json.dumps, ensure_ascii); 0 = not restated here (non-ASCII)
__device__ __forceinline__ uint32_t json_width(uint32_t c) {
    if (c >= 0x80u) return 0u;
    if (c >= 0x20u) return (c == '"' || c == '\\') ? 2u : 1u;
    return (c == '\n' || c == '\r' || c == '\t' || c == '\b' || c == '\f') ? 2u : 6u;
}

// 3: one warp per hit (in list order): where "cg:Z:" cuts the line, and the size of its element:
// 14 bytes of separator and indent + two quotes + the escaped text
__global__ void __launch_bounds__(T) size_hits(const LineSrc gaf, const uint64_t *r_off, uint32_t *r_len, uint64_t n, uint64_t *esz,
                                               uint32_t *flags) {
    const uint64_t i = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint8_t *p = line_ptr(gaf, r_off[i]);
    uint32_t len = r_len[i];
    // the first "cg:Z:" (line.split("cg:Z:")[0], :166)
    uint32_t cut = len;
    for (uint32_t b0 = 0; b0 + 5 <= len && cut == len; b0 += 32) {
        const uint32_t q = b0 + lane;
        const bool hit = q + 5 <= len && p[q] == 'c' && p[q + 1] == 'g' && p[q + 2] == ':' && p[q + 3] == 'Z' && p[q + 4] == ':';
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
        if (bal) cut = b0 + uint32_t(__ffs(bal) - 1);
    }
    len = cut;
    uint32_t w = 0;
    bool odd = false;
    for (uint32_t q = lane; q < len; q += 32) {
        const uint32_t x = json_width(p[q]);
        odd |= x == 0u;
        w += x;
    }
    w = __reduce_add_sync(0xFFFFFFFFu, w);
    if (__any_sync(0xFFFFFFFFu, odd) && lane == 0) atomicOr(flags, 1u);
    if (lane == 0) {
        r_len[i] = len;
        esz[i] = 16ull + w;
    }
}

struct KeyText {              // the JSON strings of the sv ids (quotes and escapes included), uploaded once per table set
    const uint8_t *blob;
    const uint32_t *off;      // [num_sv + 1]
};

// 4: bytes of one key with its two lists; 0 when nothing was appended to it (:163)
__global__ void size_keys(const uint64_t *seg, const uint64_t *epos, KeyText kt, uint32_t num_sv, uint64_t *ksz) {
    const uint32_t sv = blockIdx.x * blockDim.x + threadIdx.x;
    if (sv >= num_sv) return;
    const uint64_t s0 = seg[2 * sv], s1 = seg[2 * sv + 1], s2 = seg[2 * sv + 2];
    uint64_t z = 0;
    if (s2 > s0) {
        const uint64_t l0 = s1 > s0 ? epos[s1] - epos[s0] + 10 : 2, l1 = s2 > s1 ? epos[s2] - epos[s1] + 10 : 2;
        z = 6 + (kt.off[sv + 1] - kt.off[sv]) + 12 + l0 + 10 + l1 + 6;
    }
    ksz[sv] = z;
}

__device__ __forceinline__ void put(uint8_t *dst, const char *s, uint32_t n) {
    for (uint32_t k = 0; k < n; ++k) dst[k] = uint8_t(s[k]);
}

// 5a: the framing of every key that has something: ',\n    "key": [\n        ' ... '\n    ]'
// Keys [sv_lo, sv_hi) go to out[position - base]; the thread behind the last key of all closes the file.
__global__ void render_keys(const uint64_t *seg, const uint64_t *epos, const uint64_t *kpos, KeyText kt, uint32_t num_sv, uint32_t sv_lo,
                            uint32_t sv_hi, uint64_t base, uint8_t *out) {
    const uint32_t sv = sv_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (sv > sv_hi) return;
    if (sv == sv_hi) {
        if (sv_hi != num_sv) return;
        const uint64_t total = kpos[num_sv];                       // the end of the file
        if (total) put(out + (total - base), "\n}", 2);
        else put(out, "{}", 2);
        return;
    }
    const uint64_t s0 = seg[2 * sv], s1 = seg[2 * sv + 1], s2 = seg[2 * sv + 2];
    if (s2 == s0) return;
    uint8_t *p = out + (kpos[sv] - base);
    put(p, kpos[sv] ? ",\n    " : "{\n    ", 6);
    p += 6;
    const uint32_t kb = kt.off[sv], kn = kt.off[sv + 1] - kb;
    for (uint32_t k = 0; k < kn; ++k) p[k] = kt.blob[kb + k];
    p += kn;
    put(p, ": [\n        ", 12);
    p += 12;
    if (s1 > s0) {
        p += epos[s1] - epos[s0];
        put(p, "\n        ]", 10);
        p += 10;
    } else {
        put(p, "[]", 2);
        p += 2;
    }
    put(p, ",\n        ", 10);
    p += 10;
    if (s2 > s1) {
        p += epos[s2] - epos[s1];
        put(p, "\n        ]", 10);
        p += 10;
    } else {
        put(p, "[]", 2);
        p += 2;
    }
    put(p, "\n    ]", 6);
}

// 5b: one warp per hit: '[\n' or ',\n' + 12 blanks, then the line as a JSON string.  Hits [h_lo, h_hi) of the
// ordered list (= the hits of a range of keys) go to out[position - base].
__global__ void __launch_bounds__(T) render_hits(const LineSrc gaf, const uint64_t *r_off, const uint32_t *r_len, const uint32_t *r_key,
                                                 uint64_t h_lo, uint64_t h_hi, const uint64_t *seg, const uint64_t *epos, const uint64_t *kpos,
                                                 KeyText kt, uint64_t base, uint8_t *out) {
    const uint64_t i = h_lo + ((uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (i >= h_hi) return;
    const uint32_t k = r_key[i], sv = k >> 1;
    const uint64_t s0 = seg[2 * sv], s1 = seg[2 * sv + 1], lb = seg[k];
    uint64_t at = kpos[sv] + 6 + (kt.off[sv + 1] - kt.off[sv]) + 12;                          // the first list
    if (k & 1u) at += (s1 > s0 ? epos[s1] - epos[s0] + 10 : 2) + 10;                          // ... the second
    uint8_t *dst = out + (at + (epos[i] - epos[lb]) - base);
    if (lane < 14) dst[lane] = lane == 0 ? (i == lb ? '[' : ',') : (lane == 1 ? '\n' : ' ');
    if (lane == 14) dst[14] = '"';
    dst += 15;
    const uint8_t *p = line_ptr(gaf, r_off[i]);
    const uint32_t len = r_len[i];
    auto hex = [](uint32_t h) { return uint8_t(h < 10u ? '0' + h : 'a' + h - 10u); };
    for (uint32_t b0 = 0; b0 < len; b0 += 32) {
        const uint32_t q = b0 + lane;
        const uint32_t c = q < len ? p[q] : 0u, w = q < len ? json_width(c) : 0u;
        uint32_t incl = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (int(lane) >= d) incl += o;
        }
        uint8_t *d8 = dst + (incl - w);
        if (w == 1u) {
            d8[0] = uint8_t(c);
        } else if (w == 2u) {
            d8[0] = '\\';
            d8[1] = c == '\n' ? 'n' : c == '\r' ? 'r' : c == '\t' ? 't' : c == '\b' ? 'b' : c == '\f' ? 'f' : uint8_t(c);
        } else if (w == 6u) {
            d8[0] = '\\', d8[1] = 'u', d8[2] = '0', d8[3] = '0', d8[4] = hex(c >> 4), d8[5] = hex(c & 15u);
        }
        dst += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) dst[0] = '"';
}

}  // namespace

namespace svjg {

// device image of the key strings (built on first use)
struct JsonKeys {
    uint8_t *d_blob = nullptr;
    uint32_t *d_off = nullptr;
};

static void json_string_ascii(const std::string &s, std::string &o, bool &ok) {
    // json.dumps of a str with ensure_ascii: the same rules as the line text, plus \uXXXX for non-ASCII code points
    static const char hex[] = "0123456789abcdef";
    o.push_back('"');
    for (size_t i = 0; i < s.size();) {
        const uint8_t c = uint8_t(s[i]);
        if (c < 0x80) {
            switch (c) {
                case '"': o += "\\\""; break;
                case '\\': o += "\\\\"; break;
                case '\n': o += "\\n"; break;
                case '\r': o += "\\r"; break;
                case '\t': o += "\\t"; break;
                case '\b': o += "\\b"; break;
                case '\f': o += "\\f"; break;
                default:
                    if (c < 0x20) {
                        o += "\\u00";
                        o.push_back(hex[c >> 4]);
                        o.push_back(hex[c & 15]);
                    } else {
                        o.push_back(char(c));
                    }
            }
            ++i;
            continue;
        }
        uint32_t cp;
        int extra;
        if ((c & 0xE0) == 0xC0) cp = c & 0x1F, extra = 1;
        else if ((c & 0xF0) == 0xE0) cp = c & 0x0F, extra = 2;
        else if ((c & 0xF8) == 0xF0) cp = c & 0x07, extra = 3;
        else { ok = false; return; }
        if (i + size_t(extra) >= s.size()) { ok = false; return; }
        for (int k = 1; k <= extra; ++k) {
            if ((uint8_t(s[i + k]) & 0xC0) != 0x80) { ok = false; return; }
            cp = (cp << 6) | (uint8_t(s[i + k]) & 0x3F);
        }
        auto u4 = [&](uint32_t v) {
            o += "\\u";
            o.push_back(hex[(v >> 12) & 15]);
            o.push_back(hex[(v >> 8) & 15]);
            o.push_back(hex[(v >> 4) & 15]);
            o.push_back(hex[v & 15]);
        };
        if (cp >= 0x10000) {
            cp -= 0x10000;
            u4(0xD800 + (cp >> 10));
            u4(0xDC00 + (cp & 0x3FF));
        } else {
            u4(cp);
        }
        i += size_t(extra) + 1;
    }
    o.push_back('"');
}

void free_json_keys(svjg_tables *t) {
    if (!t->json_keys) return;
    cudaFree(t->json_keys->d_blob);
    cudaFree(t->json_keys->d_off);
    delete t->json_keys;
    t->json_keys = nullptr;
}

// ---- the plan of a rendering: everything up to the sizes (steps 1-4); the text itself is rendered from it whole
// (json_render_device) or key range by key range (json_render_range)
struct JsonPlan {
    uint8_t *ws = nullptr;            // one stream-ordered allocation holding the arrays below
    uint64_t *seg = nullptr, *epos = nullptr, *kpos = nullptr;
    uint64_t *r_off = nullptr;
    uint32_t *r_len = nullptr, *r_key = nullptr;
    LineSrc src{};
    KeyText kt{};
    uint32_t num_sv = 0;
    uint64_t n_hits = 0;
    uint64_t total = 0;               // bytes of the text without the closing "\n}" (or "{}" of an empty dictionary)
};

void json_plan_free(JsonPlan *plan, cudaStream_t st) {
    if (!plan) return;
    if (plan->ws) cudaFreeAsync(plan->ws, st);
    delete plan;
}

// Steps 1-4 on `st` from hits in DEVICE memory; synchronises `st` once to learn the size and whether the renderer
// declines.  d_counts must be the counters the same filter pass(es) produced: they are the list lengths.
int json_plan(svjg_tables *t, const LineSrc &src, const uint32_t *d_hit_sv2, const uint64_t *d_hit_off, const uint32_t *d_hit_len,
              uint64_t n_hits, const uint32_t *d_counts, JsonPlan **out_plan, cudaStream_t st) {
    const uint32_t num_sv = uint32_t(t->sv_ids.size());
    const uint64_t n2 = uint64_t(num_sv) * 2;
    if (!t->json_keys) {
        std::string blob;
        std::vector<uint32_t> off(num_sv + 1, 0);
        bool ok = true;
        for (uint32_t i = 0; i < num_sv && ok; ++i) {
            json_string_ascii(t->sv_ids[i], blob, ok);
            off[i + 1] = uint32_t(blob.size());
            if (blob.size() >= 0xFFFF0000ull) ok = false;
        }
        if (!ok) return set_error(SVJG_E_UNSUPPORTED, "sv ids the device renderer does not restate");
        JsonKeys *k = new JsonKeys();
        t->json_keys = k;
        SVJG_CUDA(cudaMalloc(&k->d_blob, blob.size() + 16));
        SVJG_CUDA(cudaMalloc(&k->d_off, off.size() * 4));
        SVJG_CUDA(cudaMemcpy(k->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
        SVJG_CUDA(cudaMemcpy(k->d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice));
    }
    const KeyText kt{t->json_keys->d_blob, t->json_keys->d_off};

    // one allocation: segments [n2 + 1], element positions [n + 1], key positions [num_sv + 1], ranked hits -- kept
    // for the rendering -- then cursors [n2], scattered hits, scan block sums, flags
    auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
    const uint64_t n = n_hits;
    const size_t o_seg = 0, o_epos = o_seg + up((n2 + 1) * 8), o_kpos = o_epos + up((n + 1) * 8),
                 o_roff = o_kpos + up((size_t(num_sv) + 1) * 8), o_rlen = o_roff + up(n * 8), o_rkey = o_rlen + up(n * 4),
                 o_cur = o_rkey + up(n * 4), o_soff = o_cur + up(n2 * 4), o_slen = o_soff + up(n * 8), o_tmp = o_slen + up(n * 4),
                 o_flags = o_tmp + up((std::max<uint64_t>(n, n2) / SCAN_BLOCK + 4) * 8), total = o_flags + 256;
    uint8_t *ws = nullptr;
    if (cudaError_t me = cudaMallocAsync(reinterpret_cast<void **>(&ws), total, st)) {
        if (me != cudaErrorMemoryAllocation) return cuda_fail(int(me), "json_plan: scratch");
        cudaGetLastError();                                       // not sticky: the host emitter takes over
        return set_error(SVJG_E_UNSUPPORTED, "no device memory for the renderer's scratch: the host emitter writes the text");
    }
    JsonPlan *plan = new JsonPlan();
    plan->ws = ws;
    plan->seg = reinterpret_cast<uint64_t *>(ws + o_seg);
    plan->epos = reinterpret_cast<uint64_t *>(ws + o_epos);
    plan->kpos = reinterpret_cast<uint64_t *>(ws + o_kpos);
    plan->r_off = reinterpret_cast<uint64_t *>(ws + o_roff);
    plan->r_len = reinterpret_cast<uint32_t *>(ws + o_rlen);
    plan->r_key = reinterpret_cast<uint32_t *>(ws + o_rkey);
    plan->src = src;
    plan->kt = kt;
    plan->num_sv = num_sv;
    plan->n_hits = n;
    uint64_t *seg = plan->seg, *epos = plan->epos, *kpos = plan->kpos, *s_off = reinterpret_cast<uint64_t *>(ws + o_soff),
             *tmp = reinterpret_cast<uint64_t *>(ws + o_tmp);
    uint32_t *cursor = reinterpret_cast<uint32_t *>(ws + o_cur), *s_len = reinterpret_cast<uint32_t *>(ws + o_slen),
             *flags = reinterpret_cast<uint32_t *>(ws + o_flags);
    auto fail = [&](int rc) {
        json_plan_free(plan, st);
        return rc;
    };
    auto blocks = [](uint64_t items) { return unsigned(std::max<uint64_t>(1, (items + T - 1) / T)); };
    cudaMemsetAsync(cursor, 0, n2 * 4, st);
    cudaMemsetAsync(flags, 0, 4, st);
    widen_counts<<<blocks(n2), T, 0, st>>>(d_counts, n2, seg, flags);
    exclusive_scan(seg, n2, tmp, st);
    if (n) {
        scatter_hits<<<blocks(n), T, 0, st>>>(d_hit_sv2, d_hit_off, d_hit_len, n, seg, cursor, s_off, s_len);
        rank_hits<<<blocks(n2 * 32), T, 0, st>>>(seg, uint32_t(n2), s_off, s_len, plan->r_off, plan->r_len, plan->r_key);
        size_hits<<<blocks(n * 32), T, 0, st>>>(src, plan->r_off, plan->r_len, n, epos, flags);
    }
    exclusive_scan(epos, n, tmp, st);
    size_keys<<<blocks(num_sv), T, 0, st>>>(seg, epos, kt, num_sv, kpos);
    exclusive_scan(kpos, num_sv, tmp, st);
    uint64_t h_total = 0, h_seg_total = 0;
    uint32_t h_flags = 0;
    cudaMemcpyAsync(&h_total, kpos + num_sv, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&h_seg_total, seg + n2, 8, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&h_flags, flags, 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(cuda_fail(int(e), "json_plan"));
    if (h_seg_total != n) return fail(set_error(SVJG_E_ARG, "counters and hit list do not belong together"));
    if (h_flags & 1u) return fail(set_error(SVJG_E_UNSUPPORTED, "a stored line holds non-ASCII bytes: the host emitter decodes UTF-8"));
    if (h_flags & 2u) return fail(set_error(SVJG_E_UNSUPPORTED, "a list beyond 64 Ki entries: the host emitter sorts it"));
    plan->total = h_total;
    *out_plan = plan;
    return SVJG_OK;
}

uint64_t json_plan_bytes(const JsonPlan *plan) { return plan->total + 2; }
uint32_t json_plan_keys(const JsonPlan *plan) { return plan->num_sv; }

// positions of the keys, on the host: pos[sv] .. pos[sv + 1] are the bytes of key sv in the text (none if it has no
// hits), pos[num_sv] = the text without its closing two bytes; first_hit[sv] .. first_hit[sv + 1] its hits in the
// ordered list.  Both arrays have num_sv + 1 entries.  Synchronises `st`.
int json_plan_key_positions(const JsonPlan *plan, uint64_t *pos, uint64_t *first_hit, cudaStream_t st) {
    const size_t n1 = size_t(plan->num_sv) + 1;
    SVJG_CUDA(cudaMemcpyAsync(pos, plan->kpos, n1 * 8, cudaMemcpyDeviceToHost, st));
    SVJG_CUDA(cudaMemcpy2DAsync(first_hit, 8, plan->seg, 16, 8, n1, cudaMemcpyDeviceToHost, st));     // seg[2 * sv]
    SVJG_CUDA(cudaStreamSynchronize(st));
    return SVJG_OK;
}

// Step 5 for the keys [sv_lo, sv_hi) = the hits [h_lo, h_hi) of the ordered list: their bytes [base, pos[sv_hi]) go to
// d_out[0 ..), base = pos[sv_lo]; the range that ends with the last key also gets the two closing bytes.
// Asynchronous on `st`.
int json_render_range(const JsonPlan *plan, uint32_t sv_lo, uint32_t sv_hi, uint64_t base, uint64_t h_lo, uint64_t h_hi, uint8_t *d_out,
                      cudaStream_t st) {
    auto blocks = [](uint64_t items) { return unsigned(std::max<uint64_t>(1, (items + T - 1) / T)); };
    render_keys<<<blocks(uint64_t(sv_hi - sv_lo) + 1), T, 0, st>>>(plan->seg, plan->epos, plan->kpos, plan->kt, plan->num_sv, sv_lo, sv_hi,
                                                                    base, d_out);
    if (h_hi > h_lo)
        render_hits<<<blocks((h_hi - h_lo) * 32), T, 0, st>>>(plan->src, plan->r_off, plan->r_len, plan->r_key, h_lo, h_hi, plan->seg,
                                                              plan->epos, plan->kpos, plan->kt, base, d_out);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return cuda_fail(int(le), "json_render_range: launch");
    return SVJG_OK;
}

// Renders the whole text into *d_out (cudaMallocAsync on `st`, the caller frees it with cudaFreeAsync) from hits in
// DEVICE memory; *out_len on the host after the call.
int json_render_device(svjg_tables *t, const LineSrc &src, const uint32_t *d_hit_sv2, const uint64_t *d_hit_off,
                       const uint32_t *d_hit_len, uint64_t n_hits, const uint32_t *d_counts, uint8_t **d_out, uint64_t *out_len,
                       cudaStream_t st) {
    JsonPlan *plan = nullptr;
    if (int rc = json_plan(t, src, d_hit_sv2, d_hit_off, d_hit_len, n_hits, d_counts, &plan, st)) return rc;
    const uint64_t out_bytes = json_plan_bytes(plan);
    uint8_t *out = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&out), out_bytes, st);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        json_plan_free(plan, st);
        return set_error(SVJG_E_UNSUPPORTED, "no device memory for the text: the host emitter writes it");
    }
    if (e != cudaSuccess) {
        json_plan_free(plan, st);
        return cuda_fail(int(e), "json_render_device: output buffer");
    }
    const int rc = json_render_range(plan, 0, plan->num_sv, 0, 0, plan->n_hits, out, st);
    json_plan_free(plan, st);
    if (rc) {
        cudaFreeAsync(out, st);
        return rc;
    }
    *d_out = out;
    *out_len = out_bytes;
    return SVJG_OK;
}

}  // namespace svjg
