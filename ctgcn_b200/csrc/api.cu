// extern "C" surface of libctgcn_b200.so (declared in include/ctgcn_b200.h).
#include <stdarg.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace ctgcn {

static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_gru_impl{CTGCN_IMPL_AUTO};
static std::atomic<int> g_fusion{0};         // ctgcn_set_fusion: CoreDiffusion as one launch when the shapes allow (off until it beats two launches)
static constexpr size_t kDefaultChunkCap = (size_t)8 << 30;
static std::atomic<size_t> g_chunk_cap{kDefaultChunkCap};   // bound on the per-core-sums buffer of a CoreDiffusion call

// ---- per-kernel-class timing
static std::atomic<int> g_prof_on{0};
struct ProfRec {
    int cls;
    cudaEvent_t a, b;
};
static std::vector<ProfRec> g_prof_recs;
static std::mutex g_prof_mu;

ProfScope::ProfScope(int cls, cudaStream_t st) : cls_(cls), st_(st) {
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    cudaEvent_t a;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&stop_) != cudaSuccess) {
        stop_ = nullptr;
        return;
    }
    cudaEventRecord(a, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back({cls, a, stop_});
}
ProfScope::~ProfScope() {
    if (stop_) cudaEventRecord(stop_, st_);
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// tcgen05 path (gru_tc.cu); returns 1 when the shape is not supported by it
int launch_gru_tc(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* w_ih,
                  const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                  int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc, void* ws, size_t ws_bytes, cudaStream_t st);
// CTA-pair tcgen05 path (gru_tc2.cu), cg = 2, or the same kernel without pairing, cg = 1; returns 1 when the shape is not supported
int launch_gru_tc2(int cg, const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* w_ih,
                   const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                   int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc, void* ws, size_t ws_bytes, cudaStream_t st);
size_t gru_tc2_workspace_bytes(int d_in);
// one launch per step, h in global memory / L2 (gru_wide_tc.cu): H = 256, 384, 512 (and 128, as a cross-check of the other kernels)
int launch_gru_wide_tc(int cg, const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* w_ih,
                       const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                       int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc, void* ws, size_t ws_bytes, cudaStream_t st);
size_t gru_wide_tc_workspace_bytes(int d_in, int h);
bool gru_wide_tc_takes(int d_in, int h);
bool gru_tc2_takes(int d_in, int h);
// CoreDiffusion in ONE launch (gru_tc2.cu, fused build: the cumulative SpMM runs in gather warps of the GRU kernel)
size_t core_diffusion_fused_workspace_bytes(int64_t n, int k, int d_in);
int launch_core_diffusion_fused(const ctgcn_plan* plan, int64_t row0, int64_t rows, const float* x, int64_t ldx, int d_in, int h,
                                const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w,
                                const float* ln_b, float eps, float* y, int64_t yrs, const RowScatter* sc, void* ws, size_t ws_bytes,
                                cudaStream_t st);
// tcgen05 dense layer (linear_tc.cu); returns 1 when the shape is not supported by it
int launch_linear_tc(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* w, const float* b, int64_t d_out,
                     int act, float* y, int64_t ldy, void* ws, cudaStream_t st);
// tcgen05 dense layer for arbitrary widths (linear_gen_tc.cu: both operands streamed); returns 1 when the shape is not supported
int launch_linear_gen_tc(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* w, const float* b, int64_t d_out, int act,
                         float* y, int64_t ldy, void* ws, size_t ws_bytes, cudaStream_t st);
size_t linear_gen_tc_workspace_bytes(int64_t d_in, int64_t d_out);

}  // namespace ctgcn

using namespace ctgcn;

extern "C" int ctgcn_version(void) { return 100; }
extern "C" const char* ctgcn_last_error(void) { return g_err; }
extern "C" int64_t ctgcn_launch_count(void) { return g_launches.load(); }

extern "C" int ctgcn_device_check(void) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        set_error("no CUDA device available");
        return CTGCN_ENODEV;
    }
    if (prop.major != 10) {
        set_error("device %s is sm_%d%d; this library is built for sm_100a only", prop.name, prop.major, prop.minor);
        return CTGCN_ENODEV;
    }
    return CTGCN_OK;
}

extern "C" int ctgcn_prof_enable(int on) {
    g_prof_on.store(on ? 1 : 0);
    return CTGCN_OK;
}

extern "C" int ctgcn_prof_collect(double* ms, int64_t* counts, int reset) {
    CTGCN_REQUIRE(ms && counts, "prof_collect: NULL argument");
    CTGCN_CUDA_OK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int c = 0; c < PROF_NCLASS; ++c) {
        ms[c] = 0.0;
        counts[c] = 0;
    }
    for (auto& r : g_prof_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            ms[r.cls] += t;
            counts[r.cls] += 1;
        }
    }
    if (reset) {
        for (auto& r : g_prof_recs) {
            cudaEventDestroy(r.a);
            cudaEventDestroy(r.b);
        }
        g_prof_recs.clear();
    }
    return CTGCN_OK;
}

namespace ctgcn {
void set_gru_trace(long long* buf);
void set_gru2_trace(long long* buf);
void set_gru_wide_trace(long long* buf);
}
extern "C" int ctgcn_debug_gru_trace(int64_t* device_buf) {
    set_gru_trace(reinterpret_cast<long long*>(device_buf));
    set_gru2_trace(reinterpret_cast<long long*>(device_buf));
    set_gru_wide_trace(reinterpret_cast<long long*>(device_buf));
    return CTGCN_OK;
}

extern "C" int ctgcn_set_fusion(int on) {
    g_fusion.store(on ? 1 : 0);
    return CTGCN_OK;
}

extern "C" int ctgcn_set_gru_impl(int impl) {
    CTGCN_REQUIRE(impl >= CTGCN_IMPL_AUTO && impl <= CTGCN_IMPL_TC_WIDE, "set_gru_impl: unknown implementation %d", impl);
    g_gru_impl.store(impl);
    return CTGCN_OK;
}

extern "C" int ctgcn_cumspmm_fwd(const ctgcn_plan* plan, const float* x, int64_t ldx, int d, float* u, void* stream) {
    CTGCN_REQUIRE(plan && x && u, "cumspmm_fwd: NULL argument");
    CTGCN_REQUIRE(ldx >= d, "cumspmm_fwd: ldx < d");
    return launch_cumspmm(plan, x, ldx, d, u, true, (cudaStream_t)stream);
}

// plain / cumulative SpMM with the relu switchable (relu = 0: S_i itself; with a K = 1 plan a plain SpMM A·x)
extern "C" int ctgcn_cumspmm_fwd_ex(const ctgcn_plan* plan, const float* x, int64_t ldx, int d, int relu, float* u,
                                    void* stream) {
    CTGCN_REQUIRE(plan && x && u, "cumspmm_fwd_ex: NULL argument");
    CTGCN_REQUIRE(ldx >= d, "cumspmm_fwd_ex: ldx < d");
    return launch_cumspmm(plan, x, ldx, d, u, relu != 0, (cudaStream_t)stream);
}

// ---- backward of the cumulative SpMM w.r.t. x: workspace = Zo | Zn, each [n_cols(plan_t), K, d]
extern "C" size_t ctgcn_cumspmm_bwd_workspace_bytes(const ctgcn_plan* plan_t, int d) {
    if (!plan_t || d <= 0) return 0;
    return 2 * align_up((size_t)plan_t->n_cols * plan_t->k * d * sizeof(float), 256);
}

extern "C" int ctgcn_cumspmm_bwd(const ctgcn_plan* plan_t, const float* g, int d, float* dx, int64_t lddx, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    CTGCN_REQUIRE(plan_t && g && dx, "cumspmm_bwd: NULL argument");
    CTGCN_REQUIRE(lddx >= d, "cumspmm_bwd: lddx < d");
    const size_t need = ctgcn_cumspmm_bwd_workspace_bytes(plan_t, d);
    if (!workspace || workspace_bytes < need) {
        set_error("cumspmm_bwd: workspace of %zu bytes, need %zu", workspace_bytes, need);
        return CTGCN_ENOMEM;
    }
    float* zo = (float*)workspace;
    float* zn = (float*)((char*)workspace + need / 2);
    return launch_cumspmm_bwd(plan_t, g, d, zo, zn, dx, lddx, (cudaStream_t)stream);
}

// ---- GRU / LSTM: workspace = k-major copies of the two weight matrices (SIMT path) | tcgen05 packed weights (GRU only)
static int cell_gates(int cell) { return cell == CTGCN_CELL_LSTM ? 4 : 3; }
static size_t rnn_ws_simt(int cell, int d_in, int h) {
    return align_up((size_t)cell_gates(cell) * h * (d_in + h) * sizeof(float), 256);
}
static size_t gru_ws_tc(int d_in, int h) {   // packed bf16 hi|lo weights + biases (one-CTA kernel) or + bias-fold images (pair kernel)
    const size_t r1 = align_up((size_t)3 * h * (d_in + h) * 2 * sizeof(uint16_t), 256) + 4096;
    const size_t r2 = gru_tc2_takes(d_in, h) ? gru_tc2_workspace_bytes(d_in) : 0;
    // the wide kernel is the production path for H > 128 only; its (large, chunk-sized) workspace is not charged to 128-wide layers
    // unless it has been selected explicitly (CTGCN_IMPL_TC_WIDE, tests)
    const size_t r3 = (h > 128 || g_gru_impl.load() == CTGCN_IMPL_TC_WIDE) ? gru_wide_tc_workspace_bytes(d_in, h) : 0;
    const size_t r12 = r1 > r2 ? r1 : r2;
    return r12 > r3 ? r12 : r3;
}

extern "C" size_t ctgcn_rnn_workspace_bytes(int cell, int d_in, int h) {
    if (d_in <= 0 || h <= 0 || (cell != CTGCN_CELL_GRU && cell != CTGCN_CELL_LSTM)) return 0;
    return rnn_ws_simt(cell, d_in, h) + (cell == CTGCN_CELL_GRU ? gru_ws_tc(d_in, h) : 0);
}

extern "C" size_t ctgcn_gru_workspace_bytes(int d_in, int h) { return ctgcn_rnn_workspace_bytes(CTGCN_CELL_GRU, d_in, h); }

static int rnn_seq_impl(int cell, const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h,
                        const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w,
                        const float* ln_b, float eps, int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc,
                        void* workspace, size_t workspace_bytes, void* stream) {
    CTGCN_REQUIRE(cell == CTGCN_CELL_GRU || cell == CTGCN_CELL_LSTM, "rnn_seq_fwd: unknown cell %d", cell);
    CTGCN_REQUIRE(seq && w_ih && w_hh && ln_w && ln_b && (y || sc), "rnn_seq_fwd: NULL argument");
    CTGCN_REQUIRE((b_ih == nullptr) == (b_hh == nullptr), "rnn_seq_fwd: b_ih and b_hh must both be given or both NULL");
    CTGCN_REQUIRE(n >= 0 && steps >= 1 && d_in >= 1 && h >= 1, "rnn_seq_fwd: bad sizes n=%lld steps=%d d_in=%d h=%d",
                  (long long)n, steps, d_in, h);
    CTGCN_REQUIRE(mode == CTGCN_GRU_SUM_LN || mode == CTGCN_GRU_EACH_LN, "rnn_seq_fwd: unknown mode %d", mode);
    if (workspace_bytes < ctgcn_rnn_workspace_bytes(cell, d_in, h) || !workspace) {
        set_error("rnn_seq_fwd: workspace of %zu bytes, need %zu", workspace_bytes, ctgcn_rnn_workspace_bytes(cell, d_in, h));
        return CTGCN_ENOMEM;
    }
    if (n == 0) return CTGCN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int impl = g_gru_impl.load();
    if (cell == CTGCN_CELL_GRU && impl != CTGCN_IMPL_SIMT) {
        char* tc_ws = (char*)workspace + rnn_ws_simt(cell, d_in, h);
        int rc = 1;
        if (impl == CTGCN_IMPL_TC_WIDE || ((impl == CTGCN_IMPL_AUTO || impl == CTGCN_IMPL_TCGEN05 || impl == CTGCN_IMPL_TC_UNPAIRED) && h > 128))
            rc = launch_gru_wide_tc(impl == CTGCN_IMPL_TC_UNPAIRED ? 1 : 2, seq, srs, sss, n, steps, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc,
                                    tc_ws, gru_ws_tc(d_in, h), st);
        else if (impl != CTGCN_IMPL_TC_ONE_CTA_R1)
            rc = launch_gru_tc2(impl == CTGCN_IMPL_TC_UNPAIRED ? 1 : 2, seq, srs, sss, n, steps, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w,
                                ln_b, eps, mode, y, yrs, yss, sc, tc_ws, gru_ws_tc(d_in, h), st);
        if (rc == 1 && (impl == CTGCN_IMPL_TC_ONE_CTA_R1 || impl == CTGCN_IMPL_AUTO || impl == CTGCN_IMPL_TCGEN05))
            rc = launch_gru_tc(seq, srs, sss, n, steps, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc,
                               tc_ws, gru_ws_tc(d_in, h), st);
        if (rc <= 0) return rc;  // done or failed
        CTGCN_REQUIRE(impl != CTGCN_IMPL_TCGEN05, "gru_seq_fwd: no tensor-core kernel for d_in=%d h=%d", d_in, h);
    }
    const int g = cell_gates(cell);
    float* wt_ih = (float*)workspace;
    float* wt_hh = wt_ih + (size_t)g * h * d_in;
    int rc = launch_transpose(w_ih, (int64_t)g * h, d_in, wt_ih, st);
    if (rc) return rc;
    rc = launch_transpose(w_hh, (int64_t)g * h, h, wt_hh, st);
    if (rc) return rc;
    if (cell == CTGCN_CELL_LSTM)
        return launch_lstm_simt(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, st);
    return launch_gru_simt(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, st);
}

extern "C" int ctgcn_rnn_seq_fwd(int cell, const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h,
                                 const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                 const float* ln_w, const float* ln_b, float eps, int mode, float* y, int64_t yrs,
                                 int64_t yss, void* workspace, size_t workspace_bytes, void* stream) {
    return rnn_seq_impl(cell, seq, srs, sss, n, steps, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss,
                        nullptr, workspace, workspace_bytes, stream);
}

extern "C" int ctgcn_gru_seq_fwd(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h,
                                 const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                 const float* ln_w, const float* ln_b, float eps, int mode, float* y, int64_t yrs,
                                 int64_t yss, void* workspace, size_t workspace_bytes, void* stream) {
    return rnn_seq_impl(CTGCN_CELL_GRU, seq, srs, sss, n, steps, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y,
                        yrs, yss, nullptr, workspace, workspace_bytes, stream);
}

// ---- CoreDiffusion.forward: cumulative SpMM → U [n, K, d_in] (workspace) → GRU/LSTM over cores + Σ + LayerNorm
// Rows of one CoreDiffusion chunk: the per-core sums U [rows, K, d_in] of a chunk live in the workspace between the two
// kernels; when all rows would need more than the cap (cfg 5: 5 M × 20 × 256 × 4 B = 102 GB) the layer runs chunk by chunk —
// whole waves of the persistent sequence kernel (148 SMs × 128-row tiles) so that every chunk but the last fills the GPU.
static int64_t cd_chunk_rows(const ctgcn_plan* plan, int d_in) {
    const size_t per_row = (size_t)plan->k * d_in * sizeof(float);
    const size_t cap = g_chunk_cap.load();
    int64_t rows = plan->n_rows;
    if ((size_t)rows * per_row > cap) {
        rows = (int64_t)(cap / per_row);
        const int64_t wave = 148 * 128;
        rows = rows >= wave ? rows / wave * wave : (rows >= 128 ? rows / 128 * 128 : 128);
        if (rows > plan->n_rows) rows = plan->n_rows;
    }
    return rows;
}

// shapes the one-launch CoreDiffusion takes (the launcher has the final say: alignment, row lengths)
static bool cd_fused_eligible(const ctgcn_plan* plan, int cell, int d_in, int h) {
    return cell == CTGCN_CELL_GRU && h == 128 && d_in >= 32 && d_in <= 128 && d_in % 32 == 0 && plan->k <= 64 &&
           plan->n_rows == plan->n_cols;
}

extern "C" int ctgcn_set_workspace_cap(size_t bytes) {
    g_chunk_cap.store(bytes ? bytes : kDefaultChunkCap);
    return CTGCN_OK;
}

extern "C" size_t ctgcn_core_diffusion_rnn_workspace_bytes(const ctgcn_plan* plan, int cell, int d_in, int h) {
    if (!plan || d_in <= 0 || h <= 0) return 0;
    const size_t r = ctgcn_rnn_workspace_bytes(cell, d_in, h);
    if (!r) return 0;
    const size_t two_kernel = align_up((size_t)cd_chunk_rows(plan, d_in) * plan->k * d_in * sizeof(float), 256) + r;
    const size_t fused = cd_fused_eligible(plan, cell, d_in, h) && cd_chunk_rows(plan, d_in) == plan->n_rows
                             ? core_diffusion_fused_workspace_bytes(plan->n_rows, plan->k, d_in) : 0;
    return two_kernel > fused ? two_kernel : fused;
}

extern "C" size_t ctgcn_core_diffusion_workspace_bytes(const ctgcn_plan* plan, int d_in, int h) {
    return ctgcn_core_diffusion_rnn_workspace_bytes(plan, CTGCN_CELL_GRU, d_in, h);
}

static int core_diffusion_impl(const ctgcn_plan* plan, int cell, const float* x, int64_t ldx, int d_in, int h,
                               const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w,
                               const float* ln_b, float eps, float* y, int64_t ldy, const RowScatter* sc, void* workspace,
                               size_t workspace_bytes, void* stream) {
    CTGCN_REQUIRE(plan && x && (y || sc), "core_diffusion_fwd: NULL argument");
    CTGCN_REQUIRE(cell == CTGCN_CELL_GRU || cell == CTGCN_CELL_LSTM, "core_diffusion_fwd: unknown cell %d", cell);
    CTGCN_REQUIRE(plan->n_rows == plan->n_cols, "core_diffusion_fwd: adjacency plan must be square");
    CTGCN_REQUIRE(ldx >= d_in && (sc || ldy >= h), "core_diffusion_fwd: leading dimension too small");
    const size_t need = ctgcn_core_diffusion_rnn_workspace_bytes(plan, cell, d_in, h);
    if (!workspace || workspace_bytes < need) {
        set_error("core_diffusion_fwd: workspace of %zu bytes, need %zu", workspace_bytes, need);
        return CTGCN_ENOMEM;
    }
    {
        const int impl = g_gru_impl.load();
        if (g_fusion.load() && (impl == CTGCN_IMPL_AUTO || impl == CTGCN_IMPL_TCGEN05) && cd_fused_eligible(plan, cell, d_in, h) &&
            cd_chunk_rows(plan, d_in) == plan->n_rows) {
            int rc = launch_core_diffusion_fused(plan, 0, plan->n_rows, x, ldx, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, y,
                                                 ldy, sc, workspace, workspace_bytes, (cudaStream_t)stream);
            if (rc <= 0) return rc;      // done or failed; 1 = not for this path (alignment, hub rows) → two kernels
        }
    }
    float* u = (float*)workspace;
    const int64_t chunk = cd_chunk_rows(plan, d_in);
    const size_t u_bytes = align_up((size_t)chunk * plan->k * d_in * sizeof(float), 256);
    for (int64_t row0 = 0; row0 < plan->n_rows; row0 += chunk) {
        const int64_t rows = plan->n_rows - row0 < chunk ? plan->n_rows - row0 : chunk;
        int rc = launch_cumspmm(plan, x, ldx, d_in, u, true, (cudaStream_t)stream, row0, rows);
        if (rc) return rc;
        RowScatter sc_chunk;
        if (sc) {
            sc_chunk = *sc;
            sc_chunk.row_off = row0;
        }
        rc = rnn_seq_impl(cell, u, (int64_t)plan->k * d_in, d_in, rows, plan->k, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps,
                          CTGCN_GRU_SUM_LN, y ? y + row0 * ldy : nullptr, ldy, 0, sc ? &sc_chunk : nullptr,
                          (char*)workspace + u_bytes, workspace_bytes - u_bytes, stream);
        if (rc) return rc;
    }
    return CTGCN_OK;
}

static int make_scatter(const ctgcn_plan* plan, int h, float* const* slice_ptrs, int n_slices, int64_t slice_row_stride,
                        int64_t slice_col_offset, RowScatter* sc) {
    CTGCN_REQUIRE(plan && slice_ptrs && n_slices >= 1 && n_slices <= plan->n_rows, "core_diffusion_fwd_scatter: bad slice arguments");
    CTGCN_REQUIRE(slice_row_stride >= slice_col_offset + h && slice_col_offset >= 0, "core_diffusion_fwd_scatter: bad slice strides");
    sc->slices = slice_ptrs;
    sc->n_slices = n_slices;
    sc->base = plan->n_rows / n_slices;
    sc->rem = plan->n_rows % n_slices;
    sc->row_stride = slice_row_stride;
    sc->col_offset = slice_col_offset;
    return CTGCN_OK;
}

extern "C" int ctgcn_core_diffusion_fwd(const ctgcn_plan* plan, const float* x, int64_t ldx, int d_in, int h,
                                        const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                        const float* ln_w, const float* ln_b, float eps, float* y, int64_t ldy,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    return core_diffusion_impl(plan, CTGCN_CELL_GRU, x, ldx, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, y, ldy, nullptr,
                               workspace, workspace_bytes, stream);
}

extern "C" int ctgcn_core_diffusion_fwd_scatter(const ctgcn_plan* plan, const float* x, int64_t ldx, int d_in, int h,
                                                const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                                const float* ln_w, const float* ln_b, float eps, float* const* slice_ptrs,
                                                int n_slices, int64_t slice_row_stride, int64_t slice_col_offset,
                                                void* workspace, size_t workspace_bytes, void* stream) {
    RowScatter sc;
    int rc = make_scatter(plan, h, slice_ptrs, n_slices, slice_row_stride, slice_col_offset, &sc);
    if (rc) return rc;
    return core_diffusion_impl(plan, CTGCN_CELL_GRU, x, ldx, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, nullptr, 0, &sc,
                               workspace, workspace_bytes, stream);
}

extern "C" int ctgcn_core_diffusion_rnn_fwd(const ctgcn_plan* plan, int cell, const float* x, int64_t ldx, int d_in, int h,
                                            const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                            const float* ln_w, const float* ln_b, float eps, float* y, int64_t ldy,
                                            float* const* slice_ptrs, int n_slices, int64_t slice_row_stride,
                                            int64_t slice_col_offset, void* workspace, size_t workspace_bytes, void* stream) {
    if (y)
        return core_diffusion_impl(plan, cell, x, ldx, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, y, ldy, nullptr,
                                   workspace, workspace_bytes, stream);
    RowScatter sc;
    int rc = make_scatter(plan, h, slice_ptrs, n_slices, slice_row_stride, slice_col_offset, &sc);
    if (rc) return rc;
    return core_diffusion_impl(plan, cell, x, ldx, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, nullptr, 0, &sc, workspace,
                               workspace_bytes, stream);
}

// ---- MLP layers
extern "C" size_t ctgcn_linear_workspace_bytes(int64_t d_in, int64_t d_out) {
    if (d_in <= 0 || d_out <= 0) return 0;
    const size_t a = (size_t)d_in * d_out * sizeof(float), g = linear_gen_tc_workspace_bytes(d_in, d_out);
    return align_up(a > g ? a : g, 256);
}

extern "C" int ctgcn_linear_fwd(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* w, const float* b,
                                int64_t d_out, int act, float* y, int64_t ldy, void* workspace, size_t workspace_bytes,
                                void* stream) {
    CTGCN_REQUIRE(x && w && y, "linear_fwd: NULL argument");
    CTGCN_REQUIRE(n >= 0 && d_in >= 1 && d_out >= 1 && ldx >= d_in && ldy >= d_out, "linear_fwd: bad sizes");
    CTGCN_REQUIRE(act == CTGCN_ACT_NONE || act == CTGCN_ACT_SELU, "linear_fwd: unknown activation %d", act);
    if (!workspace || workspace_bytes < ctgcn_linear_workspace_bytes(d_in, d_out)) {
        set_error("linear_fwd: workspace of %zu bytes, need %zu", workspace_bytes, ctgcn_linear_workspace_bytes(d_in, d_out));
        return CTGCN_ENOMEM;
    }
    if (n == 0) return CTGCN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (g_gru_impl.load() != CTGCN_IMPL_SIMT) {   // the implementation selector covers every tensor-core kernel
        int rc = launch_linear_tc(x, ldx, n, d_in, w, b, d_out, act, y, ldy, workspace, st);
        if (rc == 1) rc = launch_linear_gen_tc(x, ldx, n, d_in, w, b, d_out, act, y, ldy, workspace, workspace_bytes, st);
        if (rc <= 0) return rc;                   // done or failed; 1 = shape not supported → SIMT kernel
    }
    float* wt = (float*)workspace;
    int rc = launch_transpose(w, d_out, d_in, wt, st);
    if (rc) return rc;
    return launch_linear_simt(x, ldx, n, d_in, wt, b, d_out, act, y, ldy, st);
}

extern "C" int ctgcn_spmm_linear_fwd(const ctgcn_plan* xp, const float* w, const float* b, int64_t d_out, int act, float* y,
                                     int64_t ldy, void* workspace, size_t workspace_bytes, void* stream) {
    CTGCN_REQUIRE(xp && w && y, "spmm_linear_fwd: NULL argument");
    CTGCN_REQUIRE(xp->k == 1, "spmm_linear_fwd: the feature plan must hold exactly one matrix (k=%d)", xp->k);
    CTGCN_REQUIRE(d_out >= 1 && ldy >= d_out, "spmm_linear_fwd: bad sizes");
    CTGCN_REQUIRE(act == CTGCN_ACT_NONE || act == CTGCN_ACT_SELU, "spmm_linear_fwd: unknown activation %d", act);
    const int64_t d_in = xp->n_cols;
    if (!workspace || workspace_bytes < ctgcn_linear_workspace_bytes(d_in, d_out)) {
        set_error("spmm_linear_fwd: workspace of %zu bytes, need %zu", workspace_bytes,
                  ctgcn_linear_workspace_bytes(d_in, d_out));
        return CTGCN_ENOMEM;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* wt = (float*)workspace;
    int rc = launch_transpose(w, d_out, d_in, wt, st);
    if (rc) return rc;
    return launch_spmm_linear(xp, wt, b, d_out, act, y, ldy, st);
}

// ---- host helper: exact core numbers (Batagelj–Zaversnik bucket peeling, O(n + m))
extern "C" int ctgcn_kcore_numbers(int64_t n, const int64_t* rowptr, const int32_t* col, int32_t* core) {
    CTGCN_REQUIRE(n >= 0 && rowptr && core && (col || rowptr[n] == 0), "kcore_numbers: NULL argument");
    if (n == 0) return CTGCN_OK;
    std::vector<int32_t> deg(n), pos(n), vert(n);
    int32_t md = 0;
    for (int64_t v = 0; v < n; ++v) {
        deg[v] = (int32_t)(rowptr[v + 1] - rowptr[v]);
        md = deg[v] > md ? deg[v] : md;
    }
    std::vector<int64_t> bin(md + 2, 0);
    for (int64_t v = 0; v < n; ++v) bin[deg[v]]++;
    int64_t start = 0;
    for (int32_t d = 0; d <= md; ++d) {
        int64_t c = bin[d];
        bin[d] = start;
        start += c;
    }
    for (int64_t v = 0; v < n; ++v) {
        pos[v] = (int32_t)bin[deg[v]];
        vert[pos[v]] = (int32_t)v;
        bin[deg[v]]++;
    }
    for (int32_t d = md; d >= 1; --d) bin[d] = bin[d - 1];
    bin[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t v = vert[i];
        core[v] = deg[v];
        for (int64_t e = rowptr[v]; e < rowptr[v + 1]; ++e) {
            const int32_t u = col[e];
            if (u < 0 || u >= n) {
                set_error("kcore_numbers: column index out of range");
                return CTGCN_EINVAL;
            }
            if (deg[u] > deg[v]) {
                const int32_t du = deg[u], pu = pos[u];
                const int32_t pw = (int32_t)bin[du], w = vert[pw];
                if (u != w) {
                    pos[u] = pw;
                    vert[pu] = w;
                    pos[w] = pu;
                    vert[pw] = u;
                }
                bin[du]++;
                deg[u]--;
            }
        }
    }
    return CTGCN_OK;
}
