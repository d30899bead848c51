// Shared helpers for the ctgcn_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/ctgcn_b200.h"

namespace ctgcn {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CTGCN_CUDA_OK(expr)                                                                   \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ctgcn::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,             \
                             cudaGetErrorString(_e));                                         \
            return CTGCN_ECUDA;                                                               \
        }                                                                                     \
    } while (0)

#define CTGCN_LAUNCH_OK(name)                                                                 \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            ctgcn::set_error("launch of %s failed at %s:%d: %s", name, __FILE__, __LINE__,    \
                             cudaGetErrorString(_e));                                         \
            return CTGCN_ECUDA;                                                               \
        }                                                                                     \
        ctgcn::count_launch();                                                                \
    } while (0)

#define CTGCN_REQUIRE(cond, ...)                                                              \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ctgcn::set_error(__VA_ARGS__);                                                    \
            return CTGCN_EINVAL;                                                              \
        }                                                                                     \
    } while (0)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Optional per-kernel-class device timing (ctgcn_prof_*): CUDA events recorded on the launching stream
// around each launch while enabled; read back (with a device synchronise) by ctgcn_prof_collect.
enum ProfClass { PROF_SPMM = 0, PROF_GRU = 1, PROF_LINEAR = 2, PROF_PACK = 3, PROF_SPMM_LINEAR = 4, PROF_NCLASS = 5 };
struct ProfScope {
    ProfScope(int cls, cudaStream_t st);
    ~ProfScope();
    int cls_;
    cudaStream_t st_;
    cudaEvent_t stop_ = nullptr;
};

}  // namespace ctgcn

struct ctgcn_plan {
    int64_t n_rows = 0, n_cols = 0;
    int k = 0;
    int64_t entries = 0;        // stored entries of the union CSR
    int64_t nnz_raw_sum = 0;    // Σ_i stored nnz of the K input matrices ("aggregated edges")
    int64_t nnz_coalesced = 0;  // Σ_i nnz after summing duplicates
    int64_t n_oneshot = 0;
    int64_t max_row_entries = 0;  // longest row of the union CSR (the fused CoreDiffusion kernel and the hub-row split look at it)
    int device = 0;
    int32_t* rowptr = nullptr;  // [n_rows + 1]
    int32_t* col = nullptr;     // [entries]
    float* val = nullptr;       // [entries]
    uint8_t* lvl = nullptr;     // [entries]  bits 0..6 level, bit 7 one-shot
    size_t bytes = 0;
    // Hub rows (power-law graphs, BASELINE.json configs[4]): the gather kernel gives one warp to one row, so a row with 10^5
    // entries would keep one warp busy long after the other rows are done.  Rows above hub_threshold entries (512 … : chosen at
    // plan build, plan.cu) are cut into segments of that many entries that run as rows of their own (partial sums per level, no
    // relu) and are then added up in segment order (deterministic) — see launch_cumspmm.  All NULL / 0 when the plan has no such row.
    static constexpr int HUB_THRESHOLD_MIN = 512, HUB_DMAX = 512;
    int hub_threshold = 0;
    int32_t* row_end = nullptr;   // [n_rows] rowptr[r + 1], but rowptr[r] for hub rows (emptied in the main pass)
    int32_t* seg_start = nullptr; // [n_seg] entry range of every segment
    int32_t* seg_end = nullptr;
    int32_t* hub_rows = nullptr;  // [n_hub] ascending
    int32_t* hub_seg_ptr = nullptr;  // [n_hub + 1] first segment of every hub row
    float* hub_scratch = nullptr; // [n_seg, k, HUB_DMAX] partial sums of one pass
    std::vector<int32_t> h_hub_rows, h_hub_seg_ptr;   // host copies (row-chunked launches pick their hub range)
};

namespace ctgcn {
// Optional per-row redirection of a [N, H] output: row r is written into the buffer of the node slice that owns it
// (balanced contiguous slices: the first `rem` slices hold base+1 rows).  With peer-mapped slice pointers this is the
// snapshot exchange of CTGCN.forward (models.py:248) fused into the producing kernel's epilogue: every GPU stores its
// snapshot's rows straight into the [rows, T, D] sequence buffer of the rank that runs the temporal GRU on them.
struct RowScatter {
    float* const* slices = nullptr;  // DEVICE array of n_slices base pointers; nullptr = disabled
    int n_slices = 0;
    long long base = 0, rem = 0;
    long long row_stride = 0;        // elements between consecutive rows of a slice buffer (T·D)
    long long col_offset = 0;        // element offset of this output inside a slice row (t·D)
    long long row_off = 0;           // added to the kernel's (chunk-local) row index: row-chunked CoreDiffusion launches
#ifdef __CUDACC__
    __device__ __forceinline__ float* row_ptr(long long row) const {
        row += row_off;
        const long long big = rem * (base + 1);
        long long g, local;
        if (row < big) {
            g = row / (base + 1);
            local = row - g * (base + 1);
        } else {
            g = rem + (row - big) / base;
            local = (row - big) - (g - rem) * base;
        }
        return slices[g] + local * row_stride + col_offset;
    }
#endif
};
}  // namespace ctgcn

// kernels' host launchers (defined in the respective .cu files)
namespace ctgcn {
// rows < 0: all rows from row0 on; u is the [rows, K, d] buffer of the selected row range
int launch_cumspmm(const ctgcn_plan* p, const float* x, int64_t ldx, int d, float* u, bool relu, cudaStream_t st,
                   int64_t row0 = 0, int64_t rows = -1);
int launch_cumspmm_bwd(const ctgcn_plan* pt, const float* g, int d, float* zo, float* zn, float* dx, int64_t lddx,
                       cudaStream_t st);
int launch_spmm_linear(const ctgcn_plan* p, const float* wt, const float* b, int64_t d_out, int act, float* y,
                       int64_t ldy, cudaStream_t st);
int launch_transpose(const float* src, int64_t rows, int64_t cols, float* dst, cudaStream_t st);
int launch_linear_simt(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* wt, const float* b,
                       int64_t d_out, int act, float* y, int64_t ldy, cudaStream_t st);
int launch_gru_simt(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h,
                    const float* wt_ih, const float* wt_hh, const float* b_ih, const float* b_hh, const float* ln_w,
                    const float* ln_b, float eps, int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc,
                    cudaStream_t st);
int launch_lstm_simt(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h,
                     const float* wt_ih, const float* wt_hh, const float* b_ih, const float* b_hh, const float* ln_w,
                     const float* ln_b, float eps, int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc,
                     cudaStream_t st);
}  // namespace ctgcn
