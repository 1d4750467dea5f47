// zstd_core.h -- a Zstandard frame decoder (RFC 8878) written for one sequential thread per page: Parquet's ZSTD codec,
// the default of the Polars writer the reference's data preparation uses (apps/convert.py:5-19).
//
// Pure functions of bytes, compiled under nvcc (device: parquet.cu k_pq_inflate) and under g++
// (tests/native/pq_core_check.cpp, checked against Arrow's zstd encoder).  Written from the format specification;
// no zstd source was consulted.  Not supported (reported, never guessed): dictionaries, skippable frames.
//
// Memory: the caller provides a ZstdWork (decoding tables, ~11 KB) and a literals buffer of min(128 KB, frame
// content size) bytes per concurrently decoded frame.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZS_HD __host__ __device__ __forceinline__
#define ZS_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ZS_HD static inline
#define ZS_HD_NOINLINE static
#endif

namespace qkzstd {

enum { ZS_OK = 0, ZS_E_MAGIC = 1, ZS_E_HEADER = 2, ZS_E_DICT = 3, ZS_E_BLOCK = 4, ZS_E_LITERALS = 5, ZS_E_HUFFMAN = 6,
       ZS_E_FSE = 7, ZS_E_SEQUENCES = 8, ZS_E_OVERFLOW = 9, ZS_E_SIZE = 10 };

constexpr int ZS_BLOCK_MAX = 128 * 1024;
constexpr int HUF_LOG_MAX = 11;
constexpr int LL_LOG_MAX = 9, ML_LOG_MAX = 9, OF_LOG_MAX = 8;

struct FseEntry {
    uint16_t base;      // new_state = base + read(nbits)
    uint8_t symbol;
    uint8_t nbits;
};
struct FseTable {
    int log;            // accuracy log; -1 = not set
    FseEntry e[1 << LL_LOG_MAX];
};
struct FseTableSmall {  // offsets (log <= 8) and Huffman weights (log <= 6)
    int log;
    FseEntry e[1 << OF_LOG_MAX];
};
struct ZstdWork {
    uint16_t huf[1 << HUF_LOG_MAX];     // (nbits << 8) | symbol
    int huf_log;                         // 0 = no table yet
    FseTable ll, ml;
    FseTableSmall of;
    FseTableSmall wt;                    // scratch: Huffman weights' FSE table
    uint8_t weights[256];
    int16_t norm[64];                    // scratch: normalised counts of the table being built (max 53 symbols)
    uint16_t next[64];                   // scratch: per-symbol state counters
    uint64_t rep[3];
};

ZS_HD int highbit(uint32_t v) {         // index of the highest set bit, v > 0
    int n = 0;
    while (v >>= 1) n++;
    return n;
}

// ---------------------------------------------------------------- forward bit reader (FSE table descriptions)
struct FwdBits {
    const uint8_t* p;
    int64_t len;        // bytes
    int64_t bit;        // next bit to read
    bool ok;
};
ZS_HD uint32_t fwd_read(FwdBits& b, int n) {
    uint32_t v = 0;
    for (int i = 0; i < n; i++) {
        const int64_t at = b.bit + i;
        if ((at >> 3) >= b.len) { b.ok = false; return 0; }
        v |= (uint32_t)((b.p[at >> 3] >> (at & 7)) & 1) << i;
    }
    b.bit += n;
    return v;
}

// ---------------------------------------------------------------- backward bit reader (Huffman and FSE streams)
// The stream is a little-endian integer; its highest set bit is a marker, the bits below it are consumed from the top.
struct BackBits {
    const uint8_t* p;
    int64_t pos;        // bits not consumed yet (position of the next bit to read is pos-1 downwards); may go negative
    bool ok;
};
ZS_HD void back_init(BackBits& b, const uint8_t* p, int64_t len) {
    b.p = p;
    b.ok = len > 0 && p[len - 1] != 0;
    b.pos = b.ok ? (len - 1) * 8 + highbit(p[len - 1]) : 0;
}
// the n (<= 32) bits below pos, WITHOUT consuming; bits below the start of the stream read as zero.  Touches only the
// bytes that hold those bits (at most five).
ZS_HD uint32_t back_peek(const BackBits& b, int n) {
    if (n <= 0 || b.pos <= 0) return 0;
    const int64_t lo = b.pos - n, lo_c = lo < 0 ? 0 : lo;
    const int64_t first = lo_c >> 3, last = (b.pos - 1) >> 3;
    uint64_t acc = 0;
    for (int64_t k = 0; first + k <= last; k++) acc |= (uint64_t)b.p[first + k] << (8 * k);
    acc >>= (lo_c & 7);
    const int avail = (int)(b.pos - lo_c);                 // 1..32
    acc &= (avail >= 64) ? ~0ULL : ((1ULL << avail) - 1);
    return lo < 0 ? (uint32_t)(acc << (-lo)) : (uint32_t)acc;
}
ZS_HD uint32_t back_read(BackBits& b, int n) {
    const uint32_t v = back_peek(b, n);
    b.pos -= n;
    return v;
}

// ---------------------------------------------------------------- FSE
// Reads a table description (normalised counts) from src[0..len) into norm[]; returns bytes consumed, <0 on error.
ZS_HD int64_t fse_read_norm(const uint8_t* src, int64_t len, int max_symbols, int max_log, int16_t* norm, int* n_symbols, int* log_out) {
    FwdBits b{src, len, 0, true};
    const int log = 5 + (int)fwd_read(b, 4);
    if (!b.ok || log > max_log) return -1;
    int remaining = 1 << log, sym = 0;
    while (remaining > 0 && sym < max_symbols) {
        const int bits = highbit((uint32_t)remaining + 1) + 1;
        uint32_t val = fwd_read(b, bits);
        if (!b.ok) return -1;
        const uint32_t lower = (1u << (bits - 1)) - 1;
        const uint32_t threshold = (1u << bits) - 1 - (uint32_t)(remaining + 1);
        if ((val & lower) < threshold) {
            b.bit -= 1;
            val &= lower;
        } else if (val > lower) {
            val -= threshold;
        }
        const int proba = (int)val - 1;
        remaining -= proba < 0 ? -proba : proba;
        norm[sym++] = (int16_t)proba;
        if (proba == 0) {
            uint32_t rep = fwd_read(b, 2);
            while (true) {
                for (uint32_t i = 0; i < rep && sym < max_symbols; i++) norm[sym++] = 0;
                if (rep == 3) rep = fwd_read(b, 2); else break;
            }
            if (!b.ok) return -1;
        }
    }
    if (remaining != 0) return -1;
    *n_symbols = sym;
    *log_out = log;
    return (b.bit + 7) >> 3;
}

// Builds the decoding table of size 1 << log from normalised counts (norm[s] == -1: "less than one").
ZS_HD bool fse_build(FseEntry* e, int log, const int16_t* norm, int n_symbols, uint16_t* next) {
    const int size = 1 << log;
    int high = size;
    for (int s = 0; s < n_symbols; s++) {
        if (norm[s] == -1) { e[--high].symbol = (uint8_t)s; next[s] = 1; }
    }
    const int step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    int pos = 0;
    for (int s = 0; s < n_symbols; s++) {
        if (norm[s] <= 0) continue;
        next[s] = (uint16_t)norm[s];
        for (int i = 0; i < norm[s]; i++) {
            e[pos].symbol = (uint8_t)s;
            do { pos = (pos + step) & mask; } while (pos >= high);
        }
    }
    if (pos != 0) return false;
    for (int i = 0; i < size; i++) {
        const int s = e[i].symbol;
        const uint32_t ns = next[s]++;
        const int nb = log - highbit(ns);
        e[i].nbits = (uint8_t)nb;
        e[i].base = (uint16_t)((ns << nb) - size);
    }
    return true;
}
ZS_HD void fse_rle(FseEntry* e, int* log, int symbol) {
    *log = 0;
    e[0].symbol = (uint8_t)symbol; e[0].nbits = 0; e[0].base = 0;
}

// ---------------------------------------------------------------- Huffman
// weights[0..n) given (the last symbol's weight is implied); fills work.huf / huf_log.
ZS_HD bool huf_build(ZstdWork& w, int n) {
    uint32_t total = 0;
    for (int i = 0; i < n; i++) {
        if (w.weights[i] > HUF_LOG_MAX) return false;
        if (w.weights[i]) total += 1u << (w.weights[i] - 1);
    }
    if (total == 0 || n >= 256) return false;
    const int log = highbit(total) + 1;
    if (log > HUF_LOG_MAX) return false;
    const uint32_t left = (1u << log) - total;
    if (left == 0 || (left & (left - 1))) return false;
    w.weights[n] = (uint8_t)(highbit(left) + 1);
    n += 1;
    // table positions by increasing weight, symbols in natural order inside a weight
    uint32_t start[HUF_LOG_MAX + 2];
    uint32_t count[HUF_LOG_MAX + 2];
    for (int i = 0; i <= HUF_LOG_MAX + 1; i++) count[i] = 0;
    for (int i = 0; i < n; i++) count[w.weights[i]]++;
    uint32_t at = 0;
    for (int wt = 1; wt <= log; wt++) { start[wt] = at; at += count[wt] << (wt - 1); }
    if (at != (1u << log)) return false;
    for (int s = 0; s < n; s++) {
        const int wt = w.weights[s];
        if (!wt) continue;
        const uint32_t span = 1u << (wt - 1);
        const uint16_t ent = (uint16_t)(((log + 1 - wt) << 8) | s);
        for (uint32_t i = 0; i < span; i++) w.huf[start[wt] + i] = ent;
        start[wt] += span;
    }
    w.huf_log = log;
    return true;
}

// Huffman tree description at src[0..len): returns bytes consumed or <0.
ZS_HD int64_t huf_read_tree(ZstdWork& w, const uint8_t* src, int64_t len) {
    if (len < 1) return -1;
    const int hb = src[0];
    int n = 0;
    int64_t used;
    if (hb >= 128) {                                   // direct: 4-bit weights
        n = hb - 127;
        used = 1 + (n + 1) / 2;
        if (used > len) return -1;
        for (int i = 0; i < n; i++) w.weights[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    } else {                                           // FSE-compressed weights, two interleaved states
        used = 1 + hb;
        if (hb == 0 || used > len) return -1;
        int nsym = 0, log = 0;
        const int64_t hdr = fse_read_norm(src + 1, hb, 13, 6, w.norm, &nsym, &log);      // weights 0..12
        if (hdr < 0 || hdr >= hb) return -1;
        if (!fse_build(w.wt.e, log, w.norm, nsym, w.next)) return -1;
        BackBits b;
        back_init(b, src + 1 + hdr, hb - hdr);
        if (!b.ok) return -1;
        uint32_t s1 = back_read(b, log), s2 = back_read(b, log);
        while (true) {
            if (n >= 254) return -1;
            w.weights[n++] = w.wt.e[s1].symbol;
            s1 = w.wt.e[s1].base + back_read(b, w.wt.e[s1].nbits);
            if (b.pos < 0) { w.weights[n++] = w.wt.e[s2].symbol; break; }
            if (n >= 254) return -1;
            w.weights[n++] = w.wt.e[s2].symbol;
            s2 = w.wt.e[s2].base + back_read(b, w.wt.e[s2].nbits);
            if (b.pos < 0) { w.weights[n++] = w.wt.e[s1].symbol; break; }
        }
    }
    if (!huf_build(w, n)) return -1;
    return used;
}

ZS_HD bool huf_decode_stream(const ZstdWork& w, const uint8_t* src, int64_t len, uint8_t* out, int64_t n) {
    BackBits b;
    back_init(b, src, len);
    if (!b.ok) return false;
    for (int64_t i = 0; i < n; i++) {
        const uint16_t ent = w.huf[back_peek(b, w.huf_log)];
        out[i] = (uint8_t)ent;
        b.pos -= ent >> 8;
    }
    return b.pos == 0;
}

// ---------------------------------------------------------------- sequence code tables
ZS_HD void ll_code(int c, uint32_t* base, int* bits) {
    if (c < 16) { *base = (uint32_t)c; *bits = 0; return; }
    const uint32_t B[20] = {16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536};
    const uint8_t N[20] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    *base = B[c - 16]; *bits = N[c - 16];
}
ZS_HD void ml_code(int c, uint32_t* base, int* bits) {
    if (c < 32) { *base = (uint32_t)c + 3; *bits = 0; return; }
    const uint32_t B[21] = {35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387, 32771, 65539};
    const uint8_t N[21] = {1, 1, 1, 1, 2, 2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
    *base = B[c - 32]; *bits = N[c - 32];
}

// sets table `which` (0 LL, 1 OF, 2 ML) according to `mode`; returns bytes consumed from src or <0
ZS_HD int64_t seq_table(ZstdWork& w, int which, int mode, const uint8_t* src, int64_t len) {
    FseEntry* e = which == 0 ? w.ll.e : which == 1 ? w.of.e : w.ml.e;
    int* log = which == 0 ? &w.ll.log : which == 1 ? &w.of.log : &w.ml.log;
    const int max_sym = which == 0 ? 36 : which == 1 ? 32 : 53;
    const int max_log = which == 0 ? LL_LOG_MAX : which == 1 ? OF_LOG_MAX : ML_LOG_MAX;
    if (mode == 0) {                                   // predefined distributions
        const int8_t LL[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
        const int8_t OF[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};
        const int8_t ML[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                               1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
        const int n = which == 0 ? 36 : which == 1 ? 29 : 53;
        for (int i = 0; i < n; i++) w.norm[i] = which == 0 ? LL[i] : which == 1 ? OF[i] : ML[i];
        *log = which == 1 ? 5 : 6;
        return fse_build(e, *log, w.norm, n, w.next) ? 0 : -1;
    }
    if (mode == 1) {                                   // RLE: one symbol
        if (len < 1 || src[0] >= max_sym) return -1;
        fse_rle(e, log, src[0]);
        return 1;
    }
    if (mode == 2) {                                   // FSE table description
        int nsym = 0, l = 0;
        const int64_t used = fse_read_norm(src, len, max_sym, max_log, w.norm, &nsym, &l);
        if (used < 0) return -1;
        if (!fse_build(e, l, w.norm, nsym, w.next)) return -1;
        *log = l;
        return used;
    }
    return *log >= 0 ? 0 : -1;                         // repeat: the previous table must exist
}

ZS_HD void copy_match(uint8_t* dst, int64_t op, int64_t offset, int64_t len) {
    for (int64_t i = 0; i < len; i++) dst[op + i] = dst[op - offset + i];
}

// ---------------------------------------------------------------- one compressed block
ZS_HD_NOINLINE int zstd_block(ZstdWork& w, const uint8_t* src, int64_t len, uint8_t* dst, int64_t dst_cap, int64_t& op, uint8_t* lit,
                              int64_t lit_cap) {
    if (len < 1) return ZS_E_LITERALS;
    // ---- literals section
    const int ltype = src[0] & 3, sf = (src[0] >> 2) & 3;
    int64_t regen, comp = 0, hdr;
    int streams = 1;
    if (ltype < 2) {
        if (sf == 0 || sf == 2) { hdr = 1; regen = src[0] >> 3; }
        else if (sf == 1) { hdr = 2; if (len < 2) return ZS_E_LITERALS; regen = (src[0] >> 4) | ((int64_t)src[1] << 4); }
        else { hdr = 3; if (len < 3) return ZS_E_LITERALS; regen = (src[0] >> 4) | ((int64_t)src[1] << 4) | ((int64_t)src[2] << 12); }
    } else {
        if (len < 3) return ZS_E_LITERALS;
        const uint64_t v = (uint64_t)src[0] | ((uint64_t)src[1] << 8) | ((uint64_t)src[2] << 16) |
                           (len > 3 ? (uint64_t)src[3] << 24 : 0) | (len > 4 ? (uint64_t)src[4] << 32 : 0);
        if (sf < 2) { hdr = 3; regen = (v >> 4) & 0x3ff; comp = (v >> 14) & 0x3ff; streams = sf == 0 ? 1 : 4; }
        else if (sf == 2) { hdr = 4; regen = (v >> 4) & 0x3fff; comp = (v >> 18) & 0x3fff; streams = 4; }
        else { hdr = 5; regen = (v >> 4) & 0x3ffff; comp = (v >> 22) & 0x3ffff; streams = 4; }
        if (hdr > len) return ZS_E_LITERALS;
    }
    if (regen > lit_cap || regen > ZS_BLOCK_MAX) return ZS_E_LITERALS;
    int64_t ip = hdr;
    if (ltype == 0) {
        if (ip + regen > len) return ZS_E_LITERALS;
        for (int64_t i = 0; i < regen; i++) lit[i] = src[ip + i];
        ip += regen;
    } else if (ltype == 1) {
        if (ip + 1 > len) return ZS_E_LITERALS;
        for (int64_t i = 0; i < regen; i++) lit[i] = src[ip];
        ip += 1;
    } else {
        if (ip + comp > len) return ZS_E_LITERALS;
        const uint8_t* hs = src + ip;
        int64_t hl = comp;
        if (ltype == 2) {
            const int64_t used = huf_read_tree(w, hs, hl);
            if (used < 0) return ZS_E_HUFFMAN;
            hs += used; hl -= used;
        } else if (w.huf_log == 0) {
            return ZS_E_HUFFMAN;                          // treeless without a previous table
        }
        if (streams == 1) {
            if (!huf_decode_stream(w, hs, hl, lit, regen)) return ZS_E_HUFFMAN;
        } else {
            if (hl < 6) return ZS_E_HUFFMAN;
            const int64_t s1 = hs[0] | (hs[1] << 8), s2 = hs[2] | (hs[3] << 8), s3 = hs[4] | (hs[5] << 8);
            const int64_t s4 = hl - 6 - s1 - s2 - s3;
            if (s4 < 0) return ZS_E_HUFFMAN;
            const int64_t q = (regen + 3) / 4;
            if (3 * q > regen) return ZS_E_HUFFMAN;
            const uint8_t* p = hs + 6;
            if (!huf_decode_stream(w, p, s1, lit, q)) return ZS_E_HUFFMAN;
            if (!huf_decode_stream(w, p + s1, s2, lit + q, q)) return ZS_E_HUFFMAN;
            if (!huf_decode_stream(w, p + s1 + s2, s3, lit + 2 * q, q)) return ZS_E_HUFFMAN;
            if (!huf_decode_stream(w, p + s1 + s2 + s3, s4, lit + 3 * q, regen - 3 * q)) return ZS_E_HUFFMAN;
        }
        ip += comp;
    }
    // ---- sequences section
    if (ip >= len) return ZS_E_SEQUENCES;
    int64_t nseq = src[ip++];
    if (nseq >= 128) {
        if (nseq < 255) { if (ip >= len) return ZS_E_SEQUENCES; nseq = ((nseq - 128) << 8) + src[ip++]; }
        else { if (ip + 2 > len) return ZS_E_SEQUENCES; nseq = src[ip] + ((int64_t)src[ip + 1] << 8) + 0x7f00; ip += 2; }
    }
    int64_t lp = 0;                                        // literals consumed
    if (nseq > 0) {
        if (ip >= len) return ZS_E_SEQUENCES;
        const int modes = src[ip++];
        if (modes & 3) return ZS_E_SEQUENCES;
        for (int t = 0; t < 3; t++) {
            const int64_t used = seq_table(w, t, (modes >> (6 - 2 * t)) & 3, src + ip, len - ip);
            if (used < 0) return ZS_E_FSE;
            ip += used;
        }
        BackBits b;
        back_init(b, src + ip, len - ip);
        if (!b.ok) return ZS_E_SEQUENCES;
        uint32_t sl = back_read(b, w.ll.log), so = back_read(b, w.of.log), sm = back_read(b, w.ml.log);
        for (int64_t i = 0; i < nseq; i++) {
            const int oc = w.of.e[so].symbol, mc = w.ml.e[sm].symbol, lc = w.ll.e[sl].symbol;
            if (oc > 31 || mc > 52 || lc > 35) return ZS_E_SEQUENCES;
            uint64_t ov = ((uint64_t)1 << oc);
            // up to 31 extra offset bits: read in two steps (back_read serves at most 32 bits at a time)
            ov += oc > 16 ? (((uint64_t)back_read(b, oc - 16) << 16) | back_read(b, 16)) : back_read(b, oc);
            uint32_t mb, lb; int mn, ln;
            ml_code(mc, &mb, &mn);
            ll_code(lc, &lb, &ln);
            const int64_t mlen = (int64_t)mb + back_read(b, mn);
            const int64_t llen = (int64_t)lb + back_read(b, ln);
            if (i + 1 < nseq) {
                sl = w.ll.e[sl].base + back_read(b, w.ll.e[sl].nbits);
                sm = w.ml.e[sm].base + back_read(b, w.ml.e[sm].nbits);
                so = w.of.e[so].base + back_read(b, w.of.e[so].nbits);
            }
            if (b.pos < 0) return ZS_E_SEQUENCES;
            // repeat-offset rules
            uint64_t offset;
            if (ov > 3) {
                offset = ov - 3;
                w.rep[2] = w.rep[1]; w.rep[1] = w.rep[0]; w.rep[0] = offset;
            } else {
                const int idx = (int)ov + (llen == 0 ? 1 : 0);       // 1..4
                if (idx == 1) {
                    offset = w.rep[0];
                } else {
                    offset = idx == 4 ? w.rep[0] - 1 : w.rep[idx - 1];
                    if (idx > 2) w.rep[2] = w.rep[1];
                    w.rep[1] = w.rep[0];
                    w.rep[0] = offset;
                }
            }
            if (lp + llen > regen || op + llen + mlen > dst_cap) return ZS_E_OVERFLOW;
            for (int64_t k = 0; k < llen; k++) dst[op + k] = lit[lp + k];
            op += llen; lp += llen;
            if (offset == 0 || (int64_t)offset > op) return ZS_E_SEQUENCES;
            copy_match(dst, op, (int64_t)offset, mlen);
            op += mlen;
        }
        if (b.pos != 0) return ZS_E_SEQUENCES;
    }
    if (op + (regen - lp) > dst_cap) return ZS_E_OVERFLOW;
    for (int64_t k = 0; k < regen - lp; k++) dst[op + k] = lit[lp + k];
    op += regen - lp;
    return ZS_OK;
}

// ---------------------------------------------------------------- frames
// Decompresses the frame(s) in src[0..len) into dst[0..dst_len) exactly.  lit: >= min(128 KB, dst_len) bytes.
ZS_HD_NOINLINE int zstd_decompress(ZstdWork& w, const uint8_t* src, int64_t len, uint8_t* dst, int64_t dst_len, uint8_t* lit, int64_t lit_cap) {
    int64_t ip = 0, op = 0;
    while (ip < len) {
        if (ip + 5 > len) return ZS_E_HEADER;
        if (!(src[ip] == 0x28 && src[ip + 1] == 0xB5 && src[ip + 2] == 0x2F && src[ip + 3] == 0xFD)) return ZS_E_MAGIC;
        const int fhd = src[ip + 4];
        ip += 5;
        const int fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
        if (fhd & 0x08) return ZS_E_HEADER;
        if (!single) ip += 1;                              // window descriptor: the whole output is addressable here
        if (did) return ZS_E_DICT;
        const int fcs_bytes = fcs_flag == 0 ? single : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
        if (ip + fcs_bytes > len) return ZS_E_HEADER;
        ip += fcs_bytes;                                   // the caller already knows the size (page header)
        w.huf_log = 0; w.ll.log = w.ml.log = w.of.log = -1;
        w.rep[0] = 1; w.rep[1] = 4; w.rep[2] = 8;
        const int64_t frame_start = op;
        while (true) {
            if (ip + 3 > len) return ZS_E_BLOCK;
            const uint32_t bh = src[ip] | (src[ip + 1] << 8) | ((uint32_t)src[ip + 2] << 16);
            ip += 3;
            const int last = bh & 1, type = (bh >> 1) & 3;
            const int64_t bsize = bh >> 3;
            if (type == 0) {
                if (ip + bsize > len || op + bsize > dst_len) return ZS_E_OVERFLOW;
                for (int64_t i = 0; i < bsize; i++) dst[op + i] = src[ip + i];
                ip += bsize; op += bsize;
            } else if (type == 1) {
                if (ip + 1 > len || op + bsize > dst_len) return ZS_E_OVERFLOW;
                for (int64_t i = 0; i < bsize; i++) dst[op + i] = src[ip];
                ip += 1; op += bsize;
            } else if (type == 2) {
                if (ip + bsize > len || bsize > ZS_BLOCK_MAX) return ZS_E_BLOCK;
                // matches may reach back into earlier blocks of the frame but not before it
                int64_t bop = op - frame_start;
                const int rc = zstd_block(w, src + ip, bsize, dst + frame_start, dst_len - frame_start, bop, lit, lit_cap);
                if (rc) return rc;
                op = frame_start + bop;
                ip += bsize;
            } else {
                return ZS_E_BLOCK;
            }
            if (last) break;
        }
        if (checksum) { if (ip + 4 > len) return ZS_E_HEADER; ip += 4; }
    }
    return op == dst_len ? ZS_OK : ZS_E_SIZE;
}

}  // namespace qkzstd
