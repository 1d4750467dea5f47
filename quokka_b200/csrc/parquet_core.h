// parquet_core.h -- the format arithmetic of the Parquet column-chunk decoder, shared by the host-side page walker
// and the device decode kernel (parquet.cu).  Everything here is a pure function of bytes, written so that it
// compiles both under nvcc (host + device) and under plain g++ (tests/native/pq_core_check.cpp drives the very same
// functions on the CPU against pyarrow-written files).
//
// Stands in for the Arrow C++ Parquet reader the reference scans with (pyquokka/dataset/unordered_readers.py:51,98-99).
// Format facts follow the public parquet-format specification (Encodings.md: PLAIN, RLE/bit-packed hybrid,
// RLE_DICTIONARY; parquet.thrift: PageHeader) -- no Arrow source was consulted.
#pragma once
#include <stdint.h>
#include "../../include/qk.h"

#if defined(__CUDACC__)
#define QK_HD __host__ __device__ __forceinline__
#define QK_HD_MEMBER __host__ __device__ __forceinline__
#else
#define QK_HD static inline
#define QK_HD_MEMBER inline
#endif

namespace qkpq {

// ---------------------------------------------------------------- unaligned little-endian loads
// `base` must be 8-byte aligned and the buffer padded by >= 8 bytes past the last byte ever addressed.
QK_HD uint64_t load_u64_at(const uint8_t* base, int64_t off) {
    const uint64_t* w = (const uint64_t*)(base + (off & ~(int64_t)7));
    const int sh = (int)(off & 7) * 8;
    const uint64_t lo = w[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (w[1] << (64 - sh));
}
QK_HD uint32_t load_u32_at(const uint8_t* base, int64_t off) {
    const uint32_t* w = (const uint32_t*)(base + (off & ~(int64_t)3));
    const int sh = (int)(off & 3) * 8;
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return (lo >> sh) | (w[1] << (32 - sh));
}

// ---------------------------------------------------------------- run table
// Every data page is described by runs (include/qk.h: qk_pq_run): a PLAIN page is one run, a dictionary-coded
// page one run per RLE / bit-packed group.  dense_start is strictly increasing, the table ends with a sentinel
// whose dense_start is the total value count.
QK_HD int64_t find_run(const qk_pq_run* runs, int64_t n_runs, int64_t t) {
    int64_t lo = 0, hi = n_runs;                   // last run with dense_start <= t
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (runs[mid].dense_start <= t) lo = mid; else hi = mid;
    }
    return lo;
}

// k-th value of a bit-packed group sequence starting at byte `off` (LSB-first packing, width 0..32)
QK_HD uint32_t unpack_at(const uint8_t* base, int64_t off, int bw, int64_t k) {
    if (bw == 0) return 0;
    const int64_t bit = k * bw;
    const uint64_t w = load_u64_at(base, off + (bit >> 3));      // 7 + 32 bits always fit in one 64-bit window
    return (uint32_t)((w >> (bit & 7)) & ((bw == 32) ? 0xffffffffULL : ((1ULL << bw) - 1)));
}

// the dictionary index / raw element the run yields for its k-th value
QK_HD uint32_t run_index(const uint8_t* base, const qk_pq_run& r, int64_t k) {
    return r.kind == QK_PQ_RUN_RLE ? (uint32_t)r.payload : unpack_at(base, r.payload, r.bit_width, k);
}

// ---------------------------------------------------------------- varints and the hybrid run headers
struct Cursor {
    const uint8_t* p;
    int64_t pos, end;
    bool ok;
};
QK_HD uint64_t read_uvarint(Cursor& c) {
    uint64_t v = 0;
    int shift = 0;
    while (true) {
        if (c.pos >= c.end || shift > 63) { c.ok = false; return 0; }
        const uint8_t b = c.p[c.pos++];
        v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) return v;
        shift += 7;
    }
}
QK_HD int64_t read_zigzag(Cursor& c) {
    const uint64_t u = read_uvarint(c);
    return (int64_t)(u >> 1) ^ -(int64_t)(u & 1);
}

// One step of the RLE / bit-packed hybrid: parses the next run header at the cursor and advances past the run's
// payload.  Returns false at a malformed stream.  count = values the run encodes (before clipping to the page).
struct HybridRun {
    int kind;          // QK_PQ_RUN_RLE / QK_PQ_RUN_PACKED
    int64_t count;
    int64_t payload;   // RLE: the repeated value; PACKED: byte position of the packed groups
};
QK_HD bool next_hybrid_run(Cursor& c, int bw, HybridRun& out) {
    const uint64_t h = read_uvarint(c);
    if (!c.ok) return false;
    if (h & 1) {
        const int64_t groups = (int64_t)(h >> 1);
        out.kind = QK_PQ_RUN_PACKED;
        out.count = groups * 8;
        out.payload = c.pos;
        if (groups <= 0) { c.ok = false; return false; }
        const int64_t bytes = groups * bw;
        // a writer may truncate the padding of a page's final group: stop at the stream end (the padded values are
        // clipped away by the page's value count)
        c.pos = (c.pos + bytes > c.end) ? c.end : c.pos + bytes;
    } else {
        const int nbytes = (bw + 7) / 8;
        out.kind = QK_PQ_RUN_RLE;
        out.count = (int64_t)(h >> 1);
        if (out.count <= 0 || c.pos + nbytes > c.end) { c.ok = false; return false; }
        uint64_t v = 0;
        for (int i = 0; i < nbytes; i++) v |= (uint64_t)c.p[c.pos + i] << (8 * i);
        out.payload = (int64_t)v;
        c.pos += nbytes;
    }
    return true;
}

// ---------------------------------------------------------------- value t of a run table
// EB = bytes per output element (1, 4, 8).  `status` bit 0 reports a dictionary index outside the dictionary.
template <int EB> struct ElemOf;
template <> struct ElemOf<1> { typedef uint8_t type; };
template <> struct ElemOf<4> { typedef uint32_t type; };
template <> struct ElemOf<8> { typedef uint64_t type; };

template <int EB>
QK_HD typename ElemOf<EB>::type decode_value(const uint8_t* bytes, const qk_pq_run& r, int64_t t, const void* dictionary,
                                             int64_t dict_len, int* bad) {
    typedef typename ElemOf<EB>::type T;
    const int64_t k = t - r.dense_start;
    switch (r.kind) {
        case QK_PQ_RUN_PLAIN:
            if (EB == 8) return (T)load_u64_at(bytes, r.payload + k * 8);
            if (EB == 4) return (T)load_u32_at(bytes, r.payload + k * 4);
            return (T)bytes[r.payload + k];
        case QK_PQ_RUN_BOOL:
            return (T)((bytes[r.payload + (k >> 3)] >> (k & 7)) & 1);
        default: {
            int64_t e = (int64_t)r.dict_base + (int64_t)run_index(bytes, r, k);
            if (e >= dict_len || e < 0) { *bad = 1; e = 0; }
            return dict_len > 0 ? ((const T*)dictionary)[e] : (T)0;
        }
    }
}

// ---------------------------------------------------------------- the value section of one data page -> runs
// Shared by the host walker (uncompressed chunks: qk_parquet_walk_chunk) and the device walker (pages inflated on
// the device: k_pq_page_runs).  `elem`: 0 BOOLEAN, 4 / 8 fixed width, -1 BYTE_ARRAY (dictionary-coded only).
// Sink: bool operator()(int kind, int64_t dense_start, int64_t payload, int bit_width, int32_t dict_base) -- false = full.
enum { PQ_OK = 0, PQ_E_FULL = 1, PQ_E_SHORT = 2, PQ_E_BAD_RUN = 3, PQ_E_ENCODING = 4, PQ_E_PLAIN_BYTE_ARRAY = 5, PQ_E_WIDTH = 6 };
enum { PQ_ENC_PLAIN = 0, PQ_ENC_PLAIN_DICT = 2, PQ_ENC_RLE = 3, PQ_ENC_BIT_PACKED = 4, PQ_ENC_RLE_DICT = 8 };

template <class Sink>
QK_HD int walk_values(const uint8_t* bytes, int64_t v0, int64_t page_end, int64_t nv, int encoding, int elem,
                      int32_t dict_base, int64_t dense, Sink& sink) {
    if (nv <= 0) return PQ_OK;
    if (encoding == PQ_ENC_PLAIN) {
        if (elem < 0) return PQ_E_PLAIN_BYTE_ARRAY;
        const int64_t need = elem == 0 ? (nv + 7) / 8 : nv * elem;
        if (v0 + need > page_end) return PQ_E_SHORT;
        return sink(elem == 0 ? QK_PQ_RUN_BOOL : QK_PQ_RUN_PLAIN, dense, v0, 0, 0) ? PQ_OK : PQ_E_FULL;
    }
    const bool bool_rle = encoding == PQ_ENC_RLE && elem == 0;
    if (!(encoding == PQ_ENC_RLE_DICT || encoding == PQ_ENC_PLAIN_DICT || bool_rle)) return PQ_E_ENCODING;
    // dictionary indices: [bit width byte][hybrid runs]; RLE-coded BOOLEAN values (the V2 default):
    // [4-byte length][hybrid runs of width 1], decoded through the two-entry identity dictionary {0, 1}
    if (v0 + (bool_rle ? 4 : 1) > page_end) return PQ_E_SHORT;
    const int bw = bool_rle ? 1 : bytes[v0];
    if (bw > 32) return PQ_E_WIDTH;
    const int32_t base = bool_rle ? 0 : dict_base;
    Cursor rc{bytes, v0 + (bool_rle ? 4 : 1), page_end, true};
    int64_t left = nv;
    while (left > 0) {
        HybridRun r;
        if (!next_hybrid_run(rc, bw, r)) return PQ_E_BAD_RUN;
        const int64_t cnt = r.count < left ? r.count : left;
        if (r.kind == QK_PQ_RUN_PACKED && r.payload + (cnt * bw + 7) / 8 > page_end) return PQ_E_BAD_RUN;
        if (!sink(r.kind, dense, r.payload, bw, base)) return PQ_E_FULL;
        dense += cnt;
        left -= cnt;
    }
    return PQ_OK;
}

QK_HD int level_bits(int max_level) {
    int b = 0;
    while ((1 << b) <= max_level) b++;
    return b;
}
// byte-wise (bounds-exact) bit unpack for the level check
QK_HD uint32_t unpack_bytes(const uint8_t* p, int64_t off, int bw, int64_t k) {
    uint32_t v = 0;
    for (int i = 0; i < bw; i++) {
        const int64_t bit = k * bw + i;
        v |= (uint32_t)((p[off + (bit >> 3)] >> (bit & 7)) & 1) << i;
    }
    return v;
}
// 1 = the `n` definition levels encoded in [pos, end) all equal max_def (no null), 0 = some null, -1 = malformed
QK_HD int levels_all_defined(const uint8_t* p, int64_t pos, int64_t end, int64_t n, int max_def) {
    const int bw = level_bits(max_def);
    Cursor c{p, pos, end, true};
    int64_t left = n;
    while (left > 0) {
        HybridRun r;
        if (!next_hybrid_run(c, bw, r)) return -1;
        const int64_t cnt = r.count < left ? r.count : left;
        if (r.kind == QK_PQ_RUN_RLE) {
            if (r.payload != max_def) return 0;
        } else {
            if (r.payload + (cnt * bw + 7) / 8 > end) return -1;
            for (int64_t k = 0; k < cnt; k++)
                if ((int)unpack_bytes(p, r.payload, bw, k) != max_def) return 0;
        }
        left -= cnt;
    }
    return 1;
}

// ---------------------------------------------------------------- pages inflated on the device
// Where the value section of an inflated data page starts: V1 pages carry their definition levels in front
// ([4-byte length][hybrid runs]); V2 pages were inflated without their (uncompressed) level bytes.
// status bits: 1 malformed, 2 the page holds nulls.
QK_HD int64_t page_values_start(const uint8_t* img, const qk_pq_page& p, int* status) {
    int64_t v0 = p.dst_offset;
    const int64_t end = p.dst_offset + p.dst_bytes;
    if (p.kind == QK_PQ_PAGE_DATA_V1 && p.max_def > 0) {
        if (v0 + 4 > end) { *status |= 1; return end; }
        const int64_t len = (int64_t)img[v0] | ((int64_t)img[v0 + 1] << 8) | ((int64_t)img[v0 + 2] << 16) | ((int64_t)img[v0 + 3] << 24);
        if (v0 + 4 + len > end) { *status |= 1; return end; }
        const int ok = levels_all_defined(img, v0 + 4, v0 + 4 + len, p.num_values, p.max_def);
        if (ok < 0) *status |= 1; else if (ok == 0) *status |= 2;
        v0 += 4 + len;
    }
    return v0;
}

struct CountSink {
    int64_t n;
    QK_HD_MEMBER bool operator()(int, int64_t, int64_t, int, int32_t) { n++; return true; }
};
struct FillSink {
    qk_pq_run* out;
    int64_t n, cap;
    QK_HD_MEMBER bool operator()(int kind, int64_t dense, int64_t payload, int bw, int32_t base) {
        if (n >= cap) return false;
        qk_pq_run& r = out[n++];
        r.dense_start = dense; r.payload = payload; r.dict_base = base; r.kind = (uint8_t)kind; r.bit_width = (uint8_t)bw; r.reserved = 0;
        return true;
    }
};

// runs of one inflated data page: counted (runs == nullptr) or written to runs[0 .. cap).  Returns the count.
QK_HD int64_t page_runs(const uint8_t* img, const qk_pq_page& p, int elem, qk_pq_run* runs, int64_t cap, int* status) {
    if (p.kind == QK_PQ_PAGE_DICT) return 0;
    const int64_t v0 = page_values_start(img, p, status);
    const int64_t end = p.dst_offset + p.dst_bytes;
    int rc;
    int64_t n;
    if (runs) {
        FillSink s{runs, 0, cap};
        rc = walk_values(img, v0, end, p.num_values, p.encoding, elem, p.dict_base, p.dense_start, s);
        n = s.n;
    } else {
        CountSink s{0};
        rc = walk_values(img, v0, end, p.num_values, p.encoding, elem, p.dict_base, p.dense_start, s);
        n = s.n;
    }
    if (rc != PQ_OK) *status |= (rc == PQ_E_ENCODING || rc == PQ_E_PLAIN_BYTE_ARRAY) ? 4 : 1;
    return n;
}

// ---------------------------------------------------------------- Snappy (raw format, as Parquet's SNAPPY codec)
// A stream is [uvarint uncompressed length] then elements; an element is a literal (bytes follow) or a copy of
// `len` bytes from `offset` bytes back in the output (overlap allowed: offset < len repeats a pattern).
struct SnappyElem {
    int is_copy;
    int64_t len;
    int64_t arg;       // literal: position of its bytes in src; copy: offset back into the output
};
QK_HD bool snappy_next(const uint8_t* src, int64_t& ip, int64_t end, SnappyElem& e) {
    if (ip >= end) return false;
    const uint8_t tag = src[ip++];
    const int t = tag & 3;
    if (t == 0) {
        int64_t len = (tag >> 2) + 1;
        if (len > 60) {
            const int nb = (int)len - 60;
            if (ip + nb > end) return false;
            len = 0;
            for (int i = 0; i < nb; i++) len |= (int64_t)src[ip + i] << (8 * i);
            len += 1;
            ip += nb;
        }
        if (ip + len > end) return false;
        e.is_copy = 0; e.len = len; e.arg = ip;
        ip += len;
        return true;
    }
    e.is_copy = 1;
    if (t == 1) {
        if (ip + 1 > end) return false;
        e.len = 4 + ((tag >> 2) & 7);
        e.arg = ((int64_t)(tag >> 5) << 8) | src[ip];
        ip += 1;
    } else if (t == 2) {
        if (ip + 2 > end) return false;
        e.len = (tag >> 2) + 1;
        e.arg = (int64_t)src[ip] | ((int64_t)src[ip + 1] << 8);
        ip += 2;
    } else {
        if (ip + 4 > end) return false;
        e.len = (tag >> 2) + 1;
        e.arg = (int64_t)src[ip] | ((int64_t)src[ip + 1] << 8) | ((int64_t)src[ip + 2] << 16) | ((int64_t)src[ip + 3] << 24);
        ip += 4;
    }
    return true;
}
// lane `lane` of `nlanes` writes its share of element e at output position op.  Hazard-free across lanes: a copy reads
// only bytes in front of op (dst[op - offset + i % offset]), which earlier elements wrote.
QK_HD void snappy_apply(uint8_t* dst, int64_t op, const uint8_t* src, const SnappyElem& e, int lane, int nlanes) {
    if (!e.is_copy) {
        for (int64_t i = lane; i < e.len; i += nlanes) dst[op + i] = src[e.arg + i];
    } else if (e.arg >= e.len) {
        for (int64_t i = lane; i < e.len; i += nlanes) dst[op + i] = dst[op - e.arg + i];
    } else {
        for (int64_t i = lane; i < e.len; i += nlanes) dst[op + i] = dst[op - e.arg + i % e.arg];
    }
}

}  // namespace qkpq
