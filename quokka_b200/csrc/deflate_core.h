// deflate_core.h -- a DEFLATE (RFC 1951) decoder with gzip (RFC 1952) / zlib (RFC 1950) framing, written for one sequential
// thread per page: Parquet's GZIP codec (the default of Athena / Glue / older Hive writers).
//
// Pure functions of bytes, compiled under nvcc (device: parquet.cu k_pq_inflate) and under g++
// (tests/native/pq_core_check.cpp, checked against Arrow's gzip encoder and zlib).  Written from the RFCs; canonical Huffman
// codes are decoded length by length from two small count / symbol arrays -- the classic table-free scheme (the one zlib's
// contrib/puff documents) -- so the per-thread state is ~1.5 KB.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DF_HD __host__ __device__ __forceinline__
#define DF_HD_NOINLINE __host__ __device__ __noinline__
#else
#define DF_HD static inline
#define DF_HD_NOINLINE static
#endif

namespace qkdeflate {

enum { DF_OK = 0, DF_E_HEADER = 1, DF_E_BLOCK = 2, DF_E_STORED = 3, DF_E_CODES = 4, DF_E_SYMBOL = 5, DF_E_DISTANCE = 6, DF_E_OVERFLOW = 7,
       DF_E_SIZE = 8, DF_E_TRUNCATED = 9 };

struct Huffman {
    uint16_t count[16];      // codes of each length
    uint16_t symbol[288];    // symbols ordered by code
};
struct InflateWork {
    Huffman lit, dist;
    uint16_t lengths[352];      // 19 code-length lengths, then (from 32) up to 286 + 30 literal / distance lengths
};

struct Bits {                // LSB-first forward reader
    const uint8_t* p;
    int64_t len, pos;        // bytes, next byte
    uint32_t buf;
    int cnt;
    bool ok;
};
DF_HD uint32_t bits(Bits& b, int n) {           // n <= 16
    while (b.cnt < n) {
        if (b.pos >= b.len) { b.ok = false; return 0; }
        b.buf |= (uint32_t)b.p[b.pos++] << b.cnt;
        b.cnt += 8;
    }
    const uint32_t v = b.buf & ((1u << n) - 1);
    b.buf >>= n;
    b.cnt -= n;
    return v;
}

// canonical code from code lengths; returns false for an over-subscribed set (incomplete sets are allowed where the
// format allows them: a single distance code)
DF_HD bool construct(Huffman& h, const uint16_t* length, int n, bool* complete) {
    for (int i = 0; i < 16; i++) h.count[i] = 0;
    for (int s = 0; s < n; s++) h.count[length[s]]++;
    int left = 1;
    for (int len = 1; len < 16; len++) {
        left <<= 1;
        left -= h.count[len];
        if (left < 0) return false;
    }
    uint16_t offs[16];
    offs[1] = 0;
    for (int len = 1; len < 15; len++) offs[len + 1] = (uint16_t)(offs[len] + h.count[len]);
    for (int s = 0; s < n; s++)
        if (length[s]) h.symbol[offs[length[s]]++] = (uint16_t)s;
    *complete = left == 0;
    return true;
}
DF_HD int decode(Bits& b, const Huffman& h) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len < 16; len++) {
        code |= (int)bits(b, 1);
        if (!b.ok) return -1;
        const int count = h.count[len];
        if (code - count < first) return h.symbol[index + (code - first)];
        index += count;
        first += count;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

DF_HD int inflate_codes(Bits& b, const Huffman& lit, const Huffman& dist, uint8_t* dst, int64_t cap, int64_t& op) {
    const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097,
                                6145, 8193, 12289, 16385, 24577};
    const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    while (true) {
        int sym = decode(b, lit);
        if (sym < 0) return b.ok ? DF_E_SYMBOL : DF_E_TRUNCATED;
        if (sym < 256) {
            if (op >= cap) return DF_E_OVERFLOW;
            dst[op++] = (uint8_t)sym;
        } else if (sym == 256) {
            return DF_OK;
        } else {
            sym -= 257;
            if (sym >= 29) return DF_E_SYMBOL;
            const int64_t len = LBASE[sym] + (int64_t)bits(b, LEXT[sym]);
            const int ds = decode(b, dist);
            if (ds < 0 || ds >= 30) return b.ok ? DF_E_DISTANCE : DF_E_TRUNCATED;
            const int64_t d = DBASE[ds] + (int64_t)bits(b, DEXT[ds]);
            if (!b.ok) return DF_E_TRUNCATED;
            if (d > op) return DF_E_DISTANCE;
            if (op + len > cap) return DF_E_OVERFLOW;
            for (int64_t i = 0; i < len; i++) dst[op + i] = dst[op - d + i];
            op += len;
        }
    }
}

// one DEFLATE stream at b -> dst[op...]
DF_HD_NOINLINE int inflate_stream(InflateWork& w, Bits& b, uint8_t* dst, int64_t cap, int64_t& op) {
    while (true) {
        const int last = (int)bits(b, 1), type = (int)bits(b, 2);
        if (!b.ok) return DF_E_TRUNCATED;
        if (type == 0) {
            b.buf = 0; b.cnt = 0;                                   // to the byte boundary
            if (b.pos + 4 > b.len) return DF_E_TRUNCATED;
            const uint32_t n = b.p[b.pos] | (b.p[b.pos + 1] << 8), nn = b.p[b.pos + 2] | (b.p[b.pos + 3] << 8);
            b.pos += 4;
            if ((n ^ nn) != 0xffff) return DF_E_STORED;
            if (b.pos + n > b.len) return DF_E_TRUNCATED;
            if (op + n > cap) return DF_E_OVERFLOW;
            for (uint32_t i = 0; i < n; i++) dst[op + i] = b.p[b.pos + i];
            op += n; b.pos += n;
        } else if (type == 1) {
            bool complete;
            for (int s = 0; s < 144; s++) w.lengths[s] = 8;
            for (int s = 144; s < 256; s++) w.lengths[s] = 9;
            for (int s = 256; s < 280; s++) w.lengths[s] = 7;
            for (int s = 280; s < 288; s++) w.lengths[s] = 8;
            construct(w.lit, w.lengths, 288, &complete);
            for (int s = 0; s < 30; s++) w.lengths[s] = 5;
            construct(w.dist, w.lengths, 30, &complete);
            const int rc = inflate_codes(b, w.lit, w.dist, dst, cap, op);
            if (rc) return rc;
        } else if (type == 2) {
            const int nlen = (int)bits(b, 5) + 257, ndist = (int)bits(b, 5) + 1, ncode = (int)bits(b, 4) + 4;
            if (!b.ok) return DF_E_TRUNCATED;
            if (nlen > 286 || ndist > 30) return DF_E_CODES;
            const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            for (int i = 0; i < 19; i++) w.lengths[i] = 0;
            for (int i = 0; i < ncode; i++) w.lengths[ORDER[i]] = (uint16_t)bits(b, 3);
            if (!b.ok) return DF_E_TRUNCATED;
            bool complete;
            if (!construct(w.lit, w.lengths, 19, &complete) || !complete) return DF_E_CODES;   // the code-length code, held in w.lit
            int idx = 0;
            while (idx < nlen + ndist) {
                int sym = decode(b, w.lit);
                if (sym < 0) return b.ok ? DF_E_CODES : DF_E_TRUNCATED;
                if (sym < 16) {
                    w.lengths[32 + idx++] = (uint16_t)sym;          // lengths are staged past the 19 code-length entries
                } else {
                    int len = 0, rep;
                    if (sym == 16) {
                        if (idx == 0) return DF_E_CODES;
                        len = w.lengths[32 + idx - 1];
                        rep = 3 + (int)bits(b, 2);
                    } else if (sym == 17) {
                        rep = 3 + (int)bits(b, 3);
                    } else {
                        rep = 11 + (int)bits(b, 7);
                    }
                    if (!b.ok) return DF_E_TRUNCATED;
                    if (idx + rep > nlen + ndist) return DF_E_CODES;
                    while (rep--) w.lengths[32 + idx++] = (uint16_t)len;
                }
            }
            if (w.lengths[32 + 256] == 0) return DF_E_CODES;        // no end-of-block code
            if (!construct(w.lit, w.lengths + 32, nlen, &complete) || (!complete && nlen - w.lit.count[0] != 1)) return DF_E_CODES;
            if (!construct(w.dist, w.lengths + 32 + nlen, ndist, &complete) || (!complete && ndist - w.dist.count[0] != 1)) return DF_E_CODES;
            const int rc = inflate_codes(b, w.lit, w.dist, dst, cap, op);
            if (rc) return rc;
        } else {
            return DF_E_BLOCK;
        }
        if (last) return DF_OK;
    }
}

// gzip member(s) or a zlib stream in src[0..len) -> exactly dst_len bytes.  Checksums are not verified (the page's
// own CRC, when a writer sets it, lives in the page header).
DF_HD_NOINLINE int gzip_decompress(InflateWork& w, const uint8_t* src, int64_t len, uint8_t* dst, int64_t dst_len) {
    int64_t ip = 0, op = 0;
    while (ip < len) {
        int trailer;
        if (len - ip >= 10 && src[ip] == 0x1f && src[ip + 1] == 0x8b) {
            if (src[ip + 2] != 8) return DF_E_HEADER;
            const int flg = src[ip + 3];
            ip += 10;
            if (flg & 4) { if (ip + 2 > len) return DF_E_HEADER; ip += 2 + (src[ip] | (src[ip + 1] << 8)); }
            if (flg & 8) { while (ip < len && src[ip]) ip++; ip++; }
            if (flg & 16) { while (ip < len && src[ip]) ip++; ip++; }
            if (flg & 2) ip += 2;
            trailer = 8;
        } else if (len - ip >= 2 && (src[ip] & 0x0f) == 8 && ((src[ip] << 8) | src[ip + 1]) % 31 == 0) {
            if (src[ip + 1] & 0x20) return DF_E_HEADER;            // preset dictionary
            ip += 2;
            trailer = 4;
        } else {
            return DF_E_HEADER;
        }
        if (ip >= len) return DF_E_TRUNCATED;
        Bits b{src + ip, len - ip, 0, 0, 0, true};
        const int rc = inflate_stream(w, b, dst, dst_len, op);
        if (rc) return rc;
        ip += b.pos + trailer;                                     // whole bytes consumed; bits left in b.buf are padding
        if (ip > len) return DF_E_TRUNCATED;
    }
    return op == dst_len ? DF_OK : DF_E_SIZE;
}

}  // namespace qkdeflate
