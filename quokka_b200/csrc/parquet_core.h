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
#else
#define QK_HD static inline
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

}  // namespace qkpq
