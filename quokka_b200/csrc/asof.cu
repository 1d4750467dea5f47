// asof.cu -- K7 backward as-of join by key (SortedAsofExecutor, ts_executors.py:324-383).
//
// Both inputs are time-sorted.  The right side is segmented by its `by` code with the STABLE partition
// of partition.cu (time order survives inside a segment, so no sort is needed), then every left row does
// an upper-bound binary search inside its key's segment: the LAST right row with r_time <= l_time, the
// tie rule of Polars / pandas.  HBM-bound: right side read + written once (time 8 B + index 4 B per row),
// left side 12 B/row read + 4 B/row written, plus ~log2(segment) cached probes per left row; neighbouring
// left rows of one key search neighbouring positions, so the probes hit L2.
#include "common.cuh"

namespace qk {
namespace {

__global__ void __launch_bounds__(256) k_asof_prepare(const long long* r_time, const int32_t* dest, int64_t n,
                                                      long long* sorted_time, int32_t* sorted_idx) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t d = dest[i];
        sorted_time[d] = r_time[i];
        sorted_idx[d] = (int32_t)i;
    }
}

__global__ void __launch_bounds__(256) k_asof_search(const long long* l_time, const int32_t* l_by, int64_t n_left, int n_by,
                                                     const long long* sorted_time, const int32_t* sorted_idx,
                                                     const int64_t* seg, int32_t* out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_left; i += (int64_t)gridDim.x * blockDim.x) {
        const int b = l_by[i];
        int32_t r = -1;
        if (b >= 0 && b < n_by) {
            const long long t = l_time[i];
            int64_t lo = seg[b], hi = seg[b + 1];
            const int64_t first = lo;
            while (lo < hi) {                       // upper bound: first position with time > t
                const int64_t mid = (lo + hi) >> 1;
                if (sorted_time[mid] <= t) lo = mid + 1; else hi = mid;
            }
            if (lo > first) r = sorted_idx[lo - 1];
        }
        out[i] = r;
    }
}

}  // namespace
}  // namespace qk

using namespace qk;

static size_t asof_fixed_bytes(int64_t n_right, int32_t n_by) {
    return align_up((size_t)n_right * 4, 256)            // dest
         + align_up((size_t)(n_by + 1) * 8, 256)         // segment offsets
         + align_up((size_t)n_right * 8, 256)            // sorted_time
         + align_up((size_t)n_right * 4, 256);           // sorted_idx
}

extern "C" size_t qk_asof_workspace_bytes(int64_t n_right, int32_t n_by) {
    if (n_right < 0 || n_by <= 0) return 0;
    return asof_fixed_bytes(n_right, n_by) + qk_partition_workspace_bytes(n_right, n_by);
}

extern "C" int qk_asof_backward(const qk_column* l_time, const qk_column* l_by, const qk_column* r_time,
                                const qk_column* r_by, int32_t n_by, int32_t* out_ridx, void* workspace,
                                size_t ws_bytes, void* stream) {
    const char* who = "qk_asof_backward";
    if (int rc = check_col(l_time, who)) return rc;
    if (int rc = check_col(l_by, who)) return rc;
    if (int rc = check_col(r_time, who)) return rc;
    if (int rc = check_col(r_by, who)) return rc;
    if (l_time->dtype != QK_I64 || r_time->dtype != QK_I64) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: time columns must be int64", who);
    if (l_by->dtype != QK_I32 || r_by->dtype != QK_I32) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: by columns must be dense int32 codes", who);
    if (l_time->length != l_by->length || r_time->length != r_by->length) QK_FAIL(QK_ERR_INVALID, "%s: time / by length mismatch", who);
    if (n_by <= 0) QK_FAIL(QK_ERR_INVALID, "%s: n_by must be positive", who);
    const int64_t nl = l_time->length, nr = r_time->length;
    if (nl == 0) return QK_OK;
    if (!out_ridx) QK_FAIL(QK_ERR_INVALID, "%s: null output", who);
    if (!workspace || ws_bytes < qk_asof_workspace_bytes(nr, n_by)) QK_FAIL(QK_ERR_CAPACITY, "%s: workspace too small", who);
    char* w = (char*)workspace;
    int32_t* dest = (int32_t*)w; w += align_up((size_t)nr * 4, 256);
    int64_t* seg = (int64_t*)w; w += align_up((size_t)(n_by + 1) * 8, 256);
    long long* sorted_time = (long long*)w; w += align_up((size_t)nr * 8, 256);
    int32_t* sorted_idx = (int32_t*)w; w += align_up((size_t)nr * 4, 256);
    if (int rc = qk_partition_plan(r_by, n_by, QK_PART_CODE, dest, seg, w, ws_bytes - asof_fixed_bytes(nr, n_by), stream)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int cap = sm_count() * 16;
    if (nr > 0) {
        int64_t nb = (nr + 255) / 256; if (nb > cap) nb = cap;
        k_asof_prepare<<<(unsigned)nb, 256, 0, st>>>((const long long*)r_time->data, dest, nr, sorted_time, sorted_idx);
        QK_LAUNCH_CHECK("k_asof_prepare");
    }
    int64_t nb = (nl + 255) / 256; if (nb > cap) nb = cap;
    k_asof_search<<<(unsigned)nb, 256, 0, st>>>((const long long*)l_time->data, (const int32_t*)l_by->data, nl, n_by,
                                                sorted_time, sorted_idx, seg, out_ridx);
    QK_LAUNCH_CHECK("k_asof_search");
    return QK_OK;
}
