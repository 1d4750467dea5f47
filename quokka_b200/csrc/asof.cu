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


// ================================================================================================================
// Sorted-merge as-of (the default when the per-key table fits shared memory).  Both inputs are time-sorted, so the
// join is one sweep over the merged timeline carrying last[key] = row of the newest right row of every key:
// a right row updates its entry, a left row reads it.  The timeline is cut into P chunks of equal merged length
// (merge-path diagonals), one warp per chunk with its table in shared memory:
//   1  k_asof_bounds   P + 1 diagonals -> (right, left) split points; ties: right rows first (r_time <= l_time matches)
//   2  k_asof_local    every warp sweeps its RIGHT rows' keys only -> last right row per key inside the chunk
//   3  k_asof_carry    one thread per key folds the chunk tables front to back -> the table valid at each chunk's start
//   4  k_asof_sweep    every warp re-sweeps its chunk, windows of 32 right + 32 left rows: cross ranks by binary search over
//                      the other side's lanes (shuffles), every emitted left row takes the newest same-key right row among
//                      the window's visible ones (ballot) or else the table, emitted right rows then update the table.
// No sort, no scatter: reads 12 B per right row + 4 B again in step 2, 12 B per left row, writes 4 B per left row.
constexpr long long T_INF = 0x7fffffffffffffffLL;

__global__ void __launch_bounds__(128) k_asof_bounds(const long long* r_time, long long nr, const long long* l_time, long long nl,
                                                     int P, long long* rb, long long* lb) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > P) return;
    const long long total = nr + nl;
    const long long d = c == P ? total : (long long)((__int128)total * c / P);
    long long lo = d > nl ? d - nl : 0, hi = d < nr ? d : nr;
    while (lo < hi) {                                   // merge path: right rows first on ties
        const long long mid = (lo + hi) >> 1;
        if (r_time[mid] <= l_time[d - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    rb[c] = lo; lb[c] = d - lo;
}

__global__ void __launch_bounds__(32) k_asof_local(const int32_t* r_by, const long long* rb, int n_by, int32_t* tables) {
    extern __shared__ int32_t tab[];
    const int lane = threadIdx.x;
    for (int s = lane; s < n_by; s += 32) tab[s] = -1;
    __syncwarp();
    const long long qa = rb[blockIdx.x], qb = rb[blockIdx.x + 1];
    for (long long q0 = qa; q0 < qb; q0 += 512) {       // 16 independent loads per lane in flight
        int sym[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) { const long long q = q0 + u * 32 + lane; sym[u] = q < qb ? r_by[q] : -1; }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const long long q = q0 + u * 32 + lane;
            const bool ok = sym[u] >= 0 && sym[u] < n_by;
            const unsigned peers = __match_any_sync(0xffffffffu, ok ? sym[u] : -1 - lane);
            if (ok && (peers >> lane) == 1u) tab[sym[u]] = (int32_t)q;      // newest lane of its key in this step
            __syncwarp();
        }
    }
    __syncwarp();
    int32_t* out = tables + (size_t)blockIdx.x * n_by;
    for (int s = lane; s < n_by; s += 32) out[s] = tab[s];
}

// tables[c][s]: in = newest right row of key s inside chunk c (-1 none); out = newest right row of key s BEFORE chunk c,
// as the caller will see it (local rows + r_base, older rows = carry_in's value)
__global__ void __launch_bounds__(256) k_asof_carry(int32_t* tables, int P, int n_by, const int32_t* carry_in, int32_t r_base, int32_t* carry_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_by) return;
    int32_t run = carry_in ? carry_in[s] : -1;
    for (int c = 0; c < P; ++c) {
        const int32_t v = tables[(size_t)c * n_by + s];
        tables[(size_t)c * n_by + s] = run;
        if (v >= 0) run = v + r_base;
    }
    if (carry_out) carry_out[s] = run;
}

__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    const int lo = __shfl_sync(0xffffffffu, (int)(unsigned)(unsigned long long)v, src);
    const int hi = __shfl_sync(0xffffffffu, (int)((unsigned long long)v >> 32), src);
    return (long long)(((unsigned long long)(unsigned)hi << 32) | (unsigned)lo);
}

// (A variant that staged the four input streams through per-warp cp.async rings in shared memory and ranked by binary search
// over those rings was measured slower -- 18.1 ms vs 12.2 ms for 240 M rows: 500 instructions per step and one warp fewer per SM --
// profiles/r02_asof_sweep_ring_variant.txt; the sweep below reads its windows through L1 with a software prefetch.)
__global__ void __launch_bounds__(32) k_asof_sweep(const long long* r_time, const int32_t* r_by, const long long* l_time, const int32_t* l_by,
                                                   const long long* rb, const long long* lb, int n_by, const int32_t* tables, int32_t r_base,
                                                   int32_t* out) {
    extern __shared__ int32_t tab[];
    const int lane = threadIdx.x;
    const int32_t* before = tables + (size_t)blockIdx.x * n_by;
    for (int s = lane; s < n_by; s += 32) tab[s] = before[s];
    __syncwarp();
    long long qi = rb[blockIdx.x], ti = lb[blockIdx.x];
    const long long qb = rb[blockIdx.x + 1], tb = lb[blockIdx.x + 1];
    while (qi < qb || ti < tb) {
        // the window: the next 32 rows of either side (+inf past the chunk's end, so they sort last)
        const bool qv = qi + lane < qb, tv = ti + lane < tb;
        const long long Q = qv ? r_time[qi + lane] : T_INF, T = tv ? l_time[ti + lane] : T_INF;
        int qs = qv ? r_by[qi + lane] : -1, ts = tv ? l_by[ti + lane] : -1;
        if (qi + lane + 512 < qb) { asm volatile("prefetch.global.L1 [%0];" ::"l"(r_time + qi + lane + 512)); asm volatile("prefetch.global.L1 [%0];" ::"l"(r_by + qi + lane + 512)); }
        if (ti + lane + 512 < tb) { asm volatile("prefetch.global.L1 [%0];" ::"l"(l_time + ti + lane + 512)); asm volatile("prefetch.global.L1 [%0];" ::"l"(l_by + ti + lane + 512)); }
        if (qs < 0 || qs >= n_by) qs = -1;
        // cross ranks: right row k precedes every left row with T >= Q[k]; left row k follows every right row with Q <= T[k]
        int nlt = 0, nle = 0;                       // # window left rows with T < Q (mine);  # window right rows with Q <= T (mine)
        {
            int lo = 0, hi = 32;
#pragma unroll
            for (int it = 0; it < 6; ++it) {
                const int mid = (lo + hi) >> 1;
                const long long v = shfl_ll(T, mid & 31);
                if (lo < hi) { if (v < Q) lo = mid + 1; else hi = mid; }
            }
            nlt = lo;
            lo = 0; hi = 32;
#pragma unroll
            for (int it = 0; it < 6; ++it) {
                const int mid = (lo + hi) >> 1;
                const long long v = shfl_ll(Q, mid & 31);
                if (lo < hi) { if (v <= T) lo = mid + 1; else hi = mid; }
            }
            nle = lo;
        }
        const bool q_emit = qv && lane + nlt < 32, t_emit = tv && lane + nle < 32;
        const unsigned qmask = __ballot_sync(0xffffffffu, q_emit), tmask = __ballot_sync(0xffffffffu, t_emit);
        // left rows of this step, one at a time: newest visible same-key right row of the window, else the table
        unsigned todo = tmask;
        int answer = -1;
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int sym = __shfl_sync(0xffffffffu, ts, k);
            const int vis = __shfl_sync(0xffffffffu, nle, k);                       // right lanes [0, vis) precede left row k
            const unsigned same = __ballot_sync(0xffffffffu, qs == sym && sym >= 0) & (vis >= 32 ? 0xffffffffu : ((1u << vis) - 1u));
            if (lane == k) answer = same ? (int)(qi + (31 - __clz(same))) + r_base : ((sym >= 0 && sym < n_by) ? tab[sym] : -1);
        }
        if (t_emit) out[ti + lane] = answer;
        __syncwarp();
        // emitted right rows update the table (the newest lane of each key wins)
        {
            const unsigned peers = __match_any_sync(0xffffffffu, (q_emit && qs >= 0) ? qs : -1 - lane);
            if (q_emit && qs >= 0 && ((peers & qmask) >> lane) == 1u) tab[qs] = (int32_t)(qi + lane) + r_base;
        }
        __syncwarp();
        qi += __popc(qmask); ti += __popc(tmask);
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

// ---- sorted-merge as-of ------------------------------------------------------------------------------------------
static int asof_merge_chunks(int32_t n_by, size_t* smem_out) {
    const size_t smem = align_up((size_t)n_by * 4, 128);
    if (smem > 160 * 1024) return 0;                               // table too large for shared memory: partition path
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 16) per_sm = 16;
    if (per_sm < 1) per_sm = 1;
    *smem_out = smem;
    return sm_count() * per_sm;
}

extern "C" size_t qk_asof_merge_workspace_bytes(int64_t n_left, int64_t n_right, int32_t n_by) {
    size_t smem;
    const int P = n_by > 0 ? asof_merge_chunks(n_by, &smem) : 0;
    if (P == 0 || n_left < 0 || n_right < 0) return 0;
    return align_up((size_t)(P + 1) * 16, 256) + align_up((size_t)P * n_by * 4, 256);
}

extern "C" int qk_asof_merge(const qk_column* l_time, const qk_column* l_by, const qk_column* r_time, const qk_column* r_by,
                             int32_t n_by, const int32_t* carry_in, int32_t r_base, int32_t* carry_out, int32_t* out_ridx,
                             void* workspace, size_t ws_bytes, void* stream) {
    const char* who = "qk_asof_merge";
    if (int rc = check_col(l_time, who)) return rc;
    if (int rc = check_col(l_by, who)) return rc;
    if (int rc = check_col(r_time, who)) return rc;
    if (int rc = check_col(r_by, who)) return rc;
    if (l_time->dtype != QK_I64 || r_time->dtype != QK_I64) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: time columns must be int64", who);
    if (l_by->dtype != QK_I32 || r_by->dtype != QK_I32) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: by columns must be dense int32 codes", who);
    if (l_time->length != l_by->length || r_time->length != r_by->length) QK_FAIL(QK_ERR_INVALID, "%s: time / by length mismatch", who);
    if (n_by <= 0) QK_FAIL(QK_ERR_INVALID, "%s: n_by must be positive", who);
    const int64_t nl = l_time->length, nr = r_time->length;
    if (r_base < 0 || (int64_t)r_base + nr > 0x7fffffffLL) QK_FAIL(QK_ERR_INVALID, "%s: right row numbers exceed int32", who);
    size_t smem = 0;
    const int P0 = asof_merge_chunks(n_by, &smem);
    int P = P0;
    if (P == 0) QK_FAIL(QK_ERR_UNSUPPORTED, "%s: %d keys do not fit a shared-memory table; use qk_asof_backward", who, n_by);
    if (nl > 0 && !out_ridx) QK_FAIL(QK_ERR_INVALID, "%s: null output", who);
    if (!workspace || ws_bytes < qk_asof_merge_workspace_bytes(nl, nr, n_by)) QK_FAIL(QK_ERR_CAPACITY, "%s: workspace too small", who);
    const int64_t total = nl + nr;
    if (total < (int64_t)P * 64) P = (int)(total / 64 > 0 ? total / 64 : 1);     // tiny inputs: fewer, fuller chunks
    cudaStream_t st = (cudaStream_t)stream;
    long long* rb = (long long*)workspace;
    long long* lb = rb + (P + 1);
    int32_t* tables = (int32_t*)((char*)workspace + align_up((size_t)(P0 + 1) * 16, 256));
    k_asof_bounds<<<(P + 1 + 127) / 128, 128, 0, st>>>((const long long*)r_time->data, nr, (const long long*)l_time->data, nl, P, rb, lb);
    QK_LAUNCH_CHECK("k_asof_bounds");
    QK_CUDA(cudaFuncSetAttribute(k_asof_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QK_CUDA(cudaFuncSetAttribute(k_asof_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_asof_local<<<P, 32, smem, st>>>((const int32_t*)r_by->data, rb, n_by, tables);
    QK_LAUNCH_CHECK("k_asof_local");
    k_asof_carry<<<(n_by + 255) / 256, 256, 0, st>>>(tables, P, n_by, carry_in, r_base, carry_out);
    QK_LAUNCH_CHECK("k_asof_carry");
    k_asof_sweep<<<P, 32, smem, st>>>((const long long*)r_time->data, (const int32_t*)r_by->data, (const long long*)l_time->data,
                                      (const int32_t*)l_by->data, rb, lb, n_by, tables, r_base, out_ridx);
    QK_LAUNCH_CHECK("k_asof_sweep");
    return QK_OK;
}
