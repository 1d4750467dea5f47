// parquet.cu -- Parquet column chunks decoded in HBM (include/qk.h: qk_parquet_walk_chunk, qk_parquet_decode).
//
// Stands in for the Arrow C++ reader the reference scans with (pyquokka/dataset/unordered_readers.py:51,98-99).
// Split of labour: the HOST walks the page headers (a few hundred bytes per page, Thrift compact protocol) and the
// run headers of dictionary-coded pages (one byte or two per <= 504 values) and writes a run table; the DEVICE turns
// the raw page bytes, copied as they lie in the file, into Arrow-layout columns -- one thread per value, output
// fully coalesced, every value byte read once.  Algorithmic bytes per value: encoded bytes in + elem_bytes out.
#include "common.cuh"
#include "parquet_core.h"

namespace {
using namespace qkpq;

// ------------------------------------------------------------------------------------ Thrift compact protocol
enum { T_STOP = 0, T_TRUE = 1, T_FALSE = 2, T_BYTE = 3, T_I16 = 4, T_I32 = 5, T_I64 = 6, T_DOUBLE = 7, T_BINARY = 8,
       T_LIST = 9, T_SET = 10, T_MAP = 11, T_STRUCT = 12 };

void skip_value(Cursor& c, int type, int depth);

void skip_struct(Cursor& c, int depth) {
    if (depth > 16) { c.ok = false; return; }
    while (c.ok) {
        if (c.pos >= c.end) { c.ok = false; return; }
        const uint8_t b = c.p[c.pos++];
        if (b == T_STOP) return;
        if ((b >> 4) == 0) read_zigzag(c);              // long-form field id
        skip_value(c, b & 0x0f, depth + 1);
    }
}
void skip_value(Cursor& c, int type, int depth) {
    if (depth > 16) { c.ok = false; return; }
    switch (type) {
        case T_TRUE: case T_FALSE: return;                    // value lives in the field header
        case T_BYTE: c.pos += 1; break;
        case T_I16: case T_I32: case T_I64: read_uvarint(c); break;
        case T_DOUBLE: c.pos += 8; break;
        case T_BINARY: { const uint64_t n = read_uvarint(c); c.pos += (int64_t)n; } break;
        case T_LIST: case T_SET: {
            if (c.pos >= c.end) { c.ok = false; return; }
            const uint8_t h = c.p[c.pos++];
            uint64_t n = h >> 4;
            if (n == 15) n = read_uvarint(c);
            const int et = h & 0x0f;
            for (uint64_t i = 0; i < n && c.ok; i++) {
                if (et == T_TRUE || et == T_FALSE) c.pos += 1;   // list elements of type bool take a byte each
                else skip_value(c, et, depth + 1);
            }
        } break;
        case T_MAP: {
            const uint64_t n = read_uvarint(c);
            if (n) {
                if (c.pos >= c.end) { c.ok = false; return; }
                const uint8_t kv = c.p[c.pos++];
                for (uint64_t i = 0; i < n && c.ok; i++) { skip_value(c, kv >> 4, depth + 1); skip_value(c, kv & 0x0f, depth + 1); }
            }
        } break;
        case T_STRUCT: skip_struct(c, depth + 1); break;
        default: c.ok = false;
    }
    if (c.pos > c.end) c.ok = false;
}

// Iterates the fields of the struct at the cursor: returns the field id and wire type, 0 at STOP.
int next_field(Cursor& c, int& last_id, int& type) {
    if (c.pos >= c.end) { c.ok = false; return 0; }
    const uint8_t b = c.p[c.pos++];
    if (b == T_STOP) return 0;
    type = b & 0x0f;
    const int delta = b >> 4;
    last_id = delta ? last_id + delta : (int)read_zigzag(c);
    return c.ok ? last_id : 0;
}

// parquet.thrift PageHeader and the three page-type headers, reduced to what the decoder needs
enum { PAGE_DATA = 0, PAGE_INDEX = 1, PAGE_DICT = 2, PAGE_DATA_V2 = 3 };
enum { ENC_PLAIN = 0, ENC_PLAIN_DICT = 2, ENC_RLE = 3, ENC_BIT_PACKED = 4, ENC_RLE_DICT = 8 };
struct PageHeader {
    int type = -1;
    int64_t uncompressed = -1, compressed = -1;
    int64_t num_values = -1;
    int encoding = -1, def_encoding = -1, rep_encoding = -1;
    int64_t num_nulls = 0, def_bytes = 0, rep_bytes = 0;          // V2
    bool is_compressed = true;                                     // V2 (default true)
};

void parse_inner(Cursor& c, PageHeader& h, int which) {
    int id = 0, type = 0;
    while (c.ok && next_field(c, id, type)) {
        const bool is_int = type == T_I32 || type == T_I16 || type == T_I64;
        if (which == PAGE_DATA && is_int && id >= 1 && id <= 4) {
            const int64_t v = read_zigzag(c);
            if (id == 1) h.num_values = v; else if (id == 2) h.encoding = (int)v;
            else if (id == 3) h.def_encoding = (int)v; else h.rep_encoding = (int)v;
        } else if (which == PAGE_DICT && is_int && id >= 1 && id <= 2) {
            const int64_t v = read_zigzag(c);
            if (id == 1) h.num_values = v; else h.encoding = (int)v;
        } else if (which == PAGE_DATA_V2 && is_int && id >= 1 && id <= 6) {
            const int64_t v = read_zigzag(c);
            if (id == 1) h.num_values = v; else if (id == 2) h.num_nulls = v; else if (id == 4) h.encoding = (int)v;
            else if (id == 5) h.def_bytes = v; else if (id == 6) h.rep_bytes = v;
        } else if (which == PAGE_DATA_V2 && id == 7 && (type == T_TRUE || type == T_FALSE)) {
            h.is_compressed = type == T_TRUE;
        } else {
            skip_value(c, type, 0);
        }
    }
}

bool parse_page_header(Cursor& c, PageHeader& h) {
    int id = 0, type = 0;
    while (c.ok && next_field(c, id, type)) {
        if (id == 1 && type == T_I32) h.type = (int)read_zigzag(c);
        else if (id == 2 && type == T_I32) h.uncompressed = read_zigzag(c);
        else if (id == 3 && type == T_I32) h.compressed = read_zigzag(c);
        else if (id == 5 && type == T_STRUCT) parse_inner(c, h, PAGE_DATA);
        else if (id == 7 && type == T_STRUCT) parse_inner(c, h, PAGE_DICT);
        else if (id == 8 && type == T_STRUCT) parse_inner(c, h, PAGE_DATA_V2);
        else skip_value(c, type, 0);
    }
    return c.ok && h.type >= 0 && h.compressed >= 0;
}

const char* encoding_name(int e) {
    switch (e) {
        case 0: return "PLAIN"; case 2: return "PLAIN_DICTIONARY"; case 3: return "RLE"; case 4: return "BIT_PACKED";
        case 5: return "DELTA_BINARY_PACKED"; case 6: return "DELTA_LENGTH_BYTE_ARRAY"; case 7: return "DELTA_BYTE_ARRAY";
        case 8: return "RLE_DICTIONARY"; case 9: return "BYTE_STREAM_SPLIT"; default: return "unknown";
    }
}

int level_bits(int max_level) {
    int b = 0;
    while ((1 << b) <= max_level) b++;
    return b;
}

// byte-wise (bounds-exact) bit unpack for the host-side level check
uint32_t unpack_bytes(const uint8_t* p, int64_t off, int bw, int64_t k) {
    uint32_t v = 0;
    for (int i = 0; i < bw; i++) {
        const int64_t bit = k * bw + i;
        v |= (uint32_t)((p[off + (bit >> 3)] >> (bit & 7)) & 1) << i;
    }
    return v;
}

// true when the `n` definition levels encoded in [pos, end) all equal max_def (i.e. the page holds no null)
bool levels_all_defined(const uint8_t* p, int64_t pos, int64_t end, int64_t n, int max_def, bool& malformed) {
    const int bw = level_bits(max_def);
    Cursor c{p, pos, end, true};
    int64_t left = n;
    while (left > 0) {
        HybridRun r;
        if (!next_hybrid_run(c, bw, r)) { malformed = true; return false; }
        const int64_t cnt = r.count < left ? r.count : left;
        if (r.kind == QK_PQ_RUN_RLE) {
            if (r.payload != max_def) return false;
        } else {
            if (r.payload + (cnt * bw + 7) / 8 > end) { malformed = true; return false; }
            for (int64_t k = 0; k < cnt; k++)
                if ((int)unpack_bytes(p, r.payload, bw, k) != max_def) return false;
        }
        left -= cnt;
    }
    return true;
}

// ------------------------------------------------------------------------------------ device decode
constexpr int PQ_THREADS = 256;
constexpr int PQ_ITEMS = 8;     // values per thread: a CTA covers 2048 consecutive values

template <int EB>
__global__ void __launch_bounds__(PQ_THREADS) k_pq_decode(const uint8_t* __restrict__ bytes, const qk_pq_run* __restrict__ runs,
                                                          int64_t n_runs, int64_t n_values, const void* __restrict__ dictionary,
                                                          int64_t dict_len, typename ElemOf<EB>::type* __restrict__ out,
                                                          int32_t* __restrict__ status) {
    __shared__ int64_t s_run[2];
    const int64_t tile = (int64_t)blockIdx.x * (PQ_THREADS * PQ_ITEMS);
    const int64_t tile_end = min(tile + (int64_t)PQ_THREADS * PQ_ITEMS, n_values);
    // the runs that overlap this tile: two searches of the whole table, then every thread searches only that window
    if (threadIdx.x == 0) s_run[0] = find_run(runs, n_runs, tile);
    if (threadIdx.x == 32) s_run[1] = find_run(runs, n_runs, tile_end - 1);
    __syncthreads();
    const int64_t r0 = s_run[0], nr = s_run[1] - s_run[0] + 1;
    int bad = 0;
#pragma unroll
    for (int j = 0; j < PQ_ITEMS; j++) {
        const int64_t t = tile + (int64_t)j * PQ_THREADS + threadIdx.x;
        if (t < tile_end) {
            const qk_pq_run r = runs[r0 + find_run(runs + r0, nr, t)];
            out[t] = decode_value<EB>(bytes, r, t, dictionary, dict_len, &bad);
        }
    }
    if (bad && status) atomicOr(status, 1);
}

}  // namespace

// ------------------------------------------------------------------------------------ C-ABI
extern "C" {

int qk_parquet_walk_chunk(const uint8_t* bytes, int64_t chunk_offset, int64_t chunk_bytes, int64_t num_values,
                          int32_t physical_type, int32_t max_def_level, int32_t compression, int32_t dict_base,
                          qk_pq_run* runs, int64_t runs_cap, int64_t* n_runs, int64_t* dense, qk_pq_chunk_info* info) {
    const char* who = "qk_parquet_walk_chunk";
    if (!bytes || !n_runs || !dense || !info || (!runs && runs_cap > 0)) QK_FAIL(QK_ERR_INVALID, "%s: null argument", who);
    if (chunk_offset < 0 || chunk_bytes < 0 || num_values < 0 || *n_runs < 0 || *dense < 0) QK_FAIL(QK_ERR_INVALID, "%s: negative size", who);
    if (max_def_level < 0 || max_def_level > 1) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: nested columns (max definition level %d) are not supported", who, max_def_level);
    int elem = 0;
    switch (physical_type) {
        case QK_PQ_BOOLEAN: elem = 0; break;
        case QK_PQ_INT32: case QK_PQ_FLOAT: elem = 4; break;
        case QK_PQ_INT64: case QK_PQ_DOUBLE: elem = 8; break;
        case QK_PQ_BYTE_ARRAY: elem = -1; break;                      // only through a dictionary
        default: QK_FAIL(QK_ERR_UNSUPPORTED, "%s: physical type %d (INT96 / FIXED_LEN_BYTE_ARRAY) is not supported", who, physical_type);
    }
    int64_t nr = *n_runs, d = *dense;
    qk_pq_chunk_info ci;
    ci.dict_offset = -1; ci.dict_bytes = 0; ci.n_values = 0; ci.dict_num_values = 0; ci.n_data_pages = 0;
    int64_t pos = chunk_offset;
    const int64_t end = chunk_offset + chunk_bytes;
    auto push = [&](int kind, int64_t payload, int bw, int32_t base) -> bool {
        if (nr >= runs_cap) return false;
        qk_pq_run& r = runs[nr++];
        r.dense_start = d; r.payload = payload; r.dict_base = base; r.kind = (uint8_t)kind; r.bit_width = (uint8_t)bw; r.reserved = 0;
        return true;
    };
    while (ci.n_values < num_values) {
        if (pos >= end) QK_FAIL(QK_ERR_INVALID, "%s: chunk ends after %lld of %lld values", who, (long long)ci.n_values, (long long)num_values);
        Cursor c{bytes, pos, end, true};
        PageHeader h;
        if (!parse_page_header(c, h)) QK_FAIL(QK_ERR_INVALID, "%s: malformed page header at byte %lld", who, (long long)pos);
        const int64_t data = c.pos, page_end = data + h.compressed;
        if (page_end > end) QK_FAIL(QK_ERR_INVALID, "%s: page at byte %lld runs past the chunk", who, (long long)pos);
        pos = page_end;
        if (h.type == PAGE_INDEX) continue;
        if (h.type == PAGE_DICT) {
            if (compression != 0) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: compressed pages (codec %d) are not supported yet", who, compression);
            if (h.encoding != ENC_PLAIN && h.encoding != ENC_PLAIN_DICT)
                QK_FAIL(QK_ERR_UNSUPPORTED, "%s: dictionary page encoding %s is not supported", who, encoding_name(h.encoding));
            if (elem > 0 && h.num_values * elem > h.compressed) QK_FAIL(QK_ERR_INVALID, "%s: dictionary page shorter than its %lld values", who, (long long)h.num_values);
            if (elem == 0) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: dictionary-coded BOOLEAN column", who);
            ci.dict_offset = data; ci.dict_bytes = h.compressed; ci.dict_num_values = (int32_t)h.num_values;
            continue;
        }
        if (h.type != PAGE_DATA && h.type != PAGE_DATA_V2) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: page type %d", who, h.type);
        if (h.num_values < 0 || ci.n_values + h.num_values > num_values) QK_FAIL(QK_ERR_INVALID, "%s: page value counts exceed the chunk's %lld values", who, (long long)num_values);
        int64_t v0 = data;
        if (h.type == PAGE_DATA) {
            if (compression != 0) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: compressed pages (codec %d) are not supported yet", who, compression);
            if (max_def_level > 0) {
                if (h.def_encoding != ENC_RLE) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: definition levels encoded as %s", who, encoding_name(h.def_encoding));
                if (v0 + 4 > page_end) QK_FAIL(QK_ERR_INVALID, "%s: truncated definition levels", who);
                const int64_t len = (int64_t)bytes[v0] | ((int64_t)bytes[v0 + 1] << 8) | ((int64_t)bytes[v0 + 2] << 16) | ((int64_t)bytes[v0 + 3] << 24);
                if (v0 + 4 + len > page_end) QK_FAIL(QK_ERR_INVALID, "%s: truncated definition levels", who);
                bool malformed = false;
                if (!levels_all_defined(bytes, v0 + 4, v0 + 4 + len, h.num_values, max_def_level, malformed)) {
                    if (malformed) QK_FAIL(QK_ERR_INVALID, "%s: malformed definition levels", who);
                    QK_FAIL(QK_ERR_UNSUPPORTED, "%s: the column holds nulls (validity is outside the hot path)", who);
                }
                v0 += 4 + len;
            }
        } else {
            if (h.num_nulls != 0) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: the column holds nulls (validity is outside the hot path)", who);
            if (h.rep_bytes != 0) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: repeated (nested) columns are not supported", who);
            if (compression != 0 && h.is_compressed) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: compressed pages (codec %d) are not supported yet", who, compression);
            v0 += h.rep_bytes + h.def_bytes;
            if (v0 > page_end) QK_FAIL(QK_ERR_INVALID, "%s: level bytes exceed the page", who);
        }
        const int64_t nv = h.num_values;
        if (h.encoding == ENC_PLAIN) {
            if (elem < 0) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: PLAIN BYTE_ARRAY values (strings are supported as dictionary codes only)", who);
            const int64_t need = elem == 0 ? (nv + 7) / 8 : nv * elem;
            if (v0 + need > page_end) QK_FAIL(QK_ERR_INVALID, "%s: PLAIN page shorter than its %lld values", who, (long long)nv);
            if (nv > 0 && !push(elem == 0 ? QK_PQ_RUN_BOOL : QK_PQ_RUN_PLAIN, v0, 0, 0)) QK_FAIL(QK_ERR_CAPACITY, "%s: run table full (%lld)", who, (long long)runs_cap);
            d += nv;
        } else if (h.encoding == ENC_RLE_DICT || h.encoding == ENC_PLAIN_DICT || (h.encoding == ENC_RLE && elem == 0)) {
            // dictionary indices: [bit width byte][hybrid runs]; RLE-coded BOOLEAN values (the V2 default):
            // [4-byte length][hybrid runs of width 1] -- decoded through the two-entry identity dictionary {0, 1}
            const bool bool_rle = h.encoding == ENC_RLE;
            if (!bool_rle && ci.dict_offset < 0) QK_FAIL(QK_ERR_INVALID, "%s: dictionary-coded page without a dictionary page", who);
            if (nv > 0) {
                if (v0 + (bool_rle ? 4 : 1) > page_end) QK_FAIL(QK_ERR_INVALID, "%s: empty run-encoded page", who);
                const int bw = bool_rle ? 1 : bytes[v0];
                if (bw > 32) QK_FAIL(QK_ERR_INVALID, "%s: index bit width %d", who, bw);
                const int32_t base = bool_rle ? 0 : dict_base;
                Cursor rc{bytes, v0 + (bool_rle ? 4 : 1), page_end, true};
                int64_t left = nv;
                while (left > 0) {
                    HybridRun r;
                    if (!next_hybrid_run(rc, bw, r)) QK_FAIL(QK_ERR_INVALID, "%s: malformed RLE / bit-packed run in the page at byte %lld", who, (long long)data);
                    const int64_t cnt = r.count < left ? r.count : left;
                    if (r.kind == QK_PQ_RUN_PACKED && r.payload + (cnt * bw + 7) / 8 > page_end)
                        QK_FAIL(QK_ERR_INVALID, "%s: bit-packed run past the page end", who);
                    if (!push(r.kind, r.payload, bw, base)) QK_FAIL(QK_ERR_CAPACITY, "%s: run table full (%lld)", who, (long long)runs_cap);
                    d += cnt;
                    left -= cnt;
                }
            }
        } else {
            QK_FAIL(QK_ERR_UNSUPPORTED, "%s: value encoding %s is not supported", who, encoding_name(h.encoding));
        }
        ci.n_values += nv;
        ci.n_data_pages++;
    }
    *n_runs = nr;
    *dense = d;
    *info = ci;
    return QK_OK;
}

int qk_parquet_decode(const uint8_t* bytes, int64_t n_bytes, const qk_pq_run* runs, int64_t n_runs, int64_t n_values,
                      const void* dictionary, int64_t dict_len, int32_t elem_bytes, void* out, int32_t* status, void* stream) {
    const char* who = "qk_parquet_decode";
    if (n_values < 0 || n_runs < 0 || n_bytes < 0 || dict_len < 0) QK_FAIL(QK_ERR_INVALID, "%s: negative size", who);
    if (n_values == 0) return QK_OK;
    if (!bytes || !runs || !out || n_runs == 0) QK_FAIL(QK_ERR_INVALID, "%s: null argument", who);
    if (((uintptr_t)bytes & 7) || ((uintptr_t)runs & 7)) QK_FAIL(QK_ERR_INVALID, "%s: bytes / runs must be 8-byte aligned", who);
    if (dict_len > 0 && !dictionary) QK_FAIL(QK_ERR_INVALID, "%s: null dictionary", who);
    if ((uintptr_t)out % (elem_bytes > 0 ? elem_bytes : 1)) QK_FAIL(QK_ERR_INVALID, "%s: misaligned output", who);
    const int64_t per_cta = (int64_t)PQ_THREADS * PQ_ITEMS;
    const int64_t grid = (n_values + per_cta - 1) / per_cta;
    if (grid > 0x7fffffffLL) QK_FAIL(QK_ERR_INVALID, "%s: too many values", who);
    cudaStream_t s = (cudaStream_t)stream;
    switch (elem_bytes) {
        case 1: k_pq_decode<1><<<(unsigned)grid, PQ_THREADS, 0, s>>>(bytes, runs, n_runs, n_values, dictionary, dict_len, (uint8_t*)out, status); break;
        case 4: k_pq_decode<4><<<(unsigned)grid, PQ_THREADS, 0, s>>>(bytes, runs, n_runs, n_values, dictionary, dict_len, (uint32_t*)out, status); break;
        case 8: k_pq_decode<8><<<(unsigned)grid, PQ_THREADS, 0, s>>>(bytes, runs, n_runs, n_values, dictionary, dict_len, (uint64_t*)out, status); break;
        default: QK_FAIL(QK_ERR_INVALID, "%s: elem_bytes must be 1, 4 or 8", who);
    }
    QK_LAUNCH_CHECK(who);
    return QK_OK;
}

}  // extern "C"
