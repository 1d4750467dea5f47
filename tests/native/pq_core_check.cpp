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
