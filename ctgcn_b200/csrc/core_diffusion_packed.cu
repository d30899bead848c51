// EXPERIMENTAL entry point (round-2 groundwork, never run): CoreDiffusion.forward (layers.py:38-63) for 128 → 128 GRU layers with
// the per-core sums handed from the SpMM to the GRU kernel PRE-SPLIT in the tensor-core operand layout and fetched there by bulk
// copies (spmm_packed.cu → gru_tc_packed_kernel; profiles/r02_gru_design.md step 3).  Separate from ctgcn_core_diffusion_fwd on
// purpose: nothing on the default path changes.  No row chunking: the whole [tiles, K, 64 KB] buffer lives in the workspace.
#include "common.cuh"

namespace ctgcn {
int launch_gru_tc_packed(const uint8_t* packed_u, int64_t n, int steps, const float* w_ih, const float* w_hh, const float* b_ih,
                         const float* b_hh, const float* ln_w, const float* ln_b, float eps, float* y, int64_t yrs, void* ws,
                         size_t ws_bytes, cudaStream_t st);
}
using namespace ctgcn;

extern "C" size_t ctgcn_cumspmm_packed_bytes(const ctgcn_plan* plan);
extern "C" int ctgcn_cumspmm_fwd_packed(const ctgcn_plan* plan, const float* x, int64_t ldx, int d, void* u, void* stream);

static size_t gru_ws_bytes() { return align_up((size_t)3 * 128 * 256 * 2 * sizeof(uint16_t), 256) + 4096; }

extern "C" size_t ctgcn_core_diffusion_packed_workspace_bytes(const ctgcn_plan* plan) {
    if (!plan) return 0;
    return align_up(ctgcn_cumspmm_packed_bytes(plan), 256) + gru_ws_bytes();
}

extern "C" int ctgcn_core_diffusion_fwd_packed(const ctgcn_plan* plan, const float* x, int64_t ldx, const float* w_ih,
                                               const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w,
                                               const float* ln_b, float eps, float* y, int64_t ldy, void* workspace,
                                               size_t workspace_bytes, void* stream) {
    CTGCN_REQUIRE(plan && x && w_ih && w_hh && ln_w && ln_b && y, "core_diffusion_fwd_packed: NULL argument");
    CTGCN_REQUIRE((b_ih == nullptr) == (b_hh == nullptr), "core_diffusion_fwd_packed: b_ih and b_hh must both be given or both NULL");
    CTGCN_REQUIRE(plan->n_rows == plan->n_cols && ldx >= 128 && ldy >= 128, "core_diffusion_fwd_packed: bad shapes");
    const size_t need = ctgcn_core_diffusion_packed_workspace_bytes(plan);
    if (!workspace || workspace_bytes < need) {
        set_error("core_diffusion_fwd_packed: workspace of %zu bytes, need %zu", workspace_bytes, need);
        return CTGCN_ENOMEM;
    }
    const size_t u_bytes = align_up(ctgcn_cumspmm_packed_bytes(plan), 256);
    int rc = ctgcn_cumspmm_fwd_packed(plan, x, ldx, 128, workspace, stream);
    if (rc) return rc;
    return launch_gru_tc_packed((const uint8_t*)workspace, plan->n_rows, plan->k, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, y, ldy,
                                (char*)workspace + u_bytes, workspace_bytes - u_bytes, (cudaStream_t)stream);
}
