// TEST-ONLY host build of the decoder core: drives quokka_b200/csrc/parquet_core.h's decode_value -- the same
// function the device kernel k_pq_decode calls per value -- in a plain loop, so that the format arithmetic is
// checked against pyarrow in the CPU container.  Built by tests/test_parquet_decode.py with g++; never shipped.
#include "../../quokka_b200/csrc/parquet_core.h"

extern "C" int pq_check_decode(const uint8_t* bytes, const qk_pq_run* runs, int64_t n_runs, int64_t n_values,
                               const void* dictionary, int64_t dict_len, int elem_bytes, void* out) {
    int bad = 0;
    const int64_t tile = 2048;                       // the kernel's CTA tile: same windowed search
    for (int64_t t0 = 0; t0 < n_values; t0 += tile) {
        const int64_t t1 = t0 + tile < n_values ? t0 + tile : n_values;
        const int64_t r0 = qkpq::find_run(runs, n_runs, t0), r1 = qkpq::find_run(runs, n_runs, t1 - 1);
        for (int64_t t = t0; t < t1; t++) {
            const qk_pq_run r = runs[r0 + qkpq::find_run(runs + r0, r1 - r0 + 1, t)];
            if (elem_bytes == 8) ((uint64_t*)out)[t] = qkpq::decode_value<8>(bytes, r, t, dictionary, dict_len, &bad);
            else if (elem_bytes == 4) ((uint32_t*)out)[t] = qkpq::decode_value<4>(bytes, r, t, dictionary, dict_len, &bad);
            else ((uint8_t*)out)[t] = qkpq::decode_value<1>(bytes, r, t, dictionary, dict_len, &bad);
        }
    }
    return bad;
}

// The paged (compressed-chunk) pipeline: the same per-page functions the kernels k_pq_inflate / k_pq_page_runs call.
// The warp of k_pq_inflate is emulated lane by lane per element (snappy_apply is hazard-free across lanes); warps take
// pages round-robin and ZSTD pages use their warp's workspace slot, as in the kernel.
#include "../../quokka_b200/csrc/zstd_core.h"
#include "../../quokka_b200/csrc/deflate_core.h"
static const size_t ZSTD_WORK = (sizeof(qkzstd::ZstdWork) + 15) / 16 * 16;
extern "C" size_t pq_check_slot_bytes() { return ZSTD_WORK + qkzstd::ZS_BLOCK_MAX; }

extern "C" void pq_check_inflate(const uint8_t* bytes, qk_pq_page* pages, int64_t n_pages, uint8_t* scratch, uint8_t* work,
                                 int64_t n_slots) {
    using namespace qkpq;
    const int64_t n_warps = work ? n_slots : n_pages;
    for (int64_t warp = 0; warp < n_warps; warp++)
    for (int64_t pi = warp; pi < n_pages; pi += n_warps) {
        qk_pq_page& p = pages[pi];
        uint8_t* dst = scratch + p.dst_offset;
        const uint8_t* src = bytes + p.src_offset;
        if (p.compressed == QK_PQ_CODEC_NONE) {
            const int64_t n = p.src_bytes < p.dst_bytes ? p.src_bytes : p.dst_bytes;
            for (int64_t i = 0; i < n; i++) dst[i] = src[i];
            if (p.src_bytes != p.dst_bytes) p.status |= 8;
            continue;
        }
        if (p.compressed == QK_PQ_CODEC_GZIP) {
            if (warp >= n_slots || !work) { p.status |= 16; continue; }
            qkdeflate::InflateWork& w = *(qkdeflate::InflateWork*)(work + warp * (ZSTD_WORK + qkzstd::ZS_BLOCK_MAX));
            if (qkdeflate::gzip_decompress(w, src, p.src_bytes, dst, p.dst_bytes)) p.status |= 8;
            continue;
        }
        if (p.compressed == QK_PQ_CODEC_ZSTD) {
            if (warp >= n_slots || !work) { p.status |= 16; continue; }
            uint8_t* slot = work + warp * (ZSTD_WORK + qkzstd::ZS_BLOCK_MAX);
            if (qkzstd::zstd_decompress(*(qkzstd::ZstdWork*)slot, src, p.src_bytes, dst, p.dst_bytes, slot + ZSTD_WORK, qkzstd::ZS_BLOCK_MAX))
                p.status |= 8;
            continue;
        }
        Cursor c{src, 0, p.src_bytes, true};
        const uint64_t ulen = read_uvarint(c);
        int64_t ip = c.pos, op = 0;
        bool bad = !c.ok || (int64_t)ulen != p.dst_bytes;
        while (!bad && ip < p.src_bytes) {
            SnappyElem e;
            if (!snappy_next(src, ip, p.src_bytes, e) || op + e.len > p.dst_bytes || (e.is_copy && (e.arg <= 0 || e.arg > op))) { bad = true; break; }
            for (int lane = 31; lane >= 0; lane--) snappy_apply(dst, op, src, e, lane, 32);      // any lane order must do
            op += e.len;
        }
        if (bad || op != p.dst_bytes) p.status |= 8;
    }
}

extern "C" void pq_check_page_runs(const uint8_t* img, qk_pq_page* pages, int64_t n_pages, int elem, const int64_t* run_offsets,
                                   qk_pq_run* runs, int64_t runs_cap) {
    using namespace qkpq;
    for (int64_t pi = 0; pi < n_pages; pi++) {
        const qk_pq_page p = pages[pi];
        int status = 0;
        if (run_offsets) {
            const int64_t at = run_offsets[pi];
            const int64_t cap = at + p.n_runs <= runs_cap ? p.n_runs : 0;
            const int64_t n = page_runs(img, p, elem, cap ? runs + at : nullptr, cap, &status);
            if (n != p.n_runs || (p.n_runs && !cap)) status |= 1;
        } else {
            const int64_t n = page_runs(img, p, elem, nullptr, 0, &status);
            pages[pi].n_runs = n > 0x7fffffffLL ? 0x7fffffff : (int32_t)n;
        }
        if (status) pages[pi].status |= status;
    }
}

// Zstandard alone: the frame decoder of quokka_b200/csrc/zstd_core.h on one stream.
#include <stdlib.h>
extern "C" int pq_check_zstd(const uint8_t* src, int64_t len, uint8_t* dst, int64_t dst_len) {
    qkzstd::ZstdWork* w = (qkzstd::ZstdWork*)malloc(sizeof(qkzstd::ZstdWork));
    const int64_t lit_cap = dst_len < qkzstd::ZS_BLOCK_MAX ? dst_len : qkzstd::ZS_BLOCK_MAX;
    uint8_t* lit = (uint8_t*)malloc(lit_cap > 0 ? lit_cap : 1);
    const int rc = qkzstd::zstd_decompress(*w, src, len, dst, dst_len, lit, lit_cap);
    free(lit);
    free(w);
    return rc;
}
extern "C" int pq_zstd_work_bytes() { return (int)sizeof(qkzstd::ZstdWork); }

// DEFLATE / gzip alone: quokka_b200/csrc/deflate_core.h on one stream.
extern "C" int pq_check_gzip(const uint8_t* src, int64_t len, uint8_t* dst, int64_t dst_len) {
    qkdeflate::InflateWork* w = (qkdeflate::InflateWork*)malloc(sizeof(qkdeflate::InflateWork));
    const int rc = qkdeflate::gzip_decompress(*w, src, len, dst, dst_len);
    free(w);
    return rc;
}
