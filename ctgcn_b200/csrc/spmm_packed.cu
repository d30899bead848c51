// EXPERIMENTAL (round-2 groundwork, never run): the cumulative SpMM storing U PRE-SPLIT in the tensor-core operand layout.
//
// profiles/r02_gru_design.md, step 3: two tiles in flight in the GRU kernel need U to stream through a bulk-copy ring, i.e. U has to
// sit in global memory exactly as the A operand wants it.  Same arithmetic as cumspmm_vec_kernel<1, 8, true> (spmm.cu, D = 128), same
// 4 bytes per element, but instead of fp32 rows [N, K, 128] it writes, per 128-row tile and level, one 64 KB image
//     [plane hi | lo][kb 0..15][row 0..127][8 bf16]          hi = bf16_rn(u), lo = bf16_rn(u − hi)   (what the GRU loaders build)
// A lane owns features 4l..4l+3 = half of a 16-byte unit: lane pairs exchange halves so that even lanes store a whole hi unit and
// odd lanes a whole lo unit — one 16-byte store per lane and level.  The question this variant answers on the GPU: does the
// scattered 16-byte store pattern (sectors completed by the neighbouring row's warp) cost the SpMM anything?
#include <cuda_bf16.h>

#include "common.cuh"

namespace ctgcn {
namespace {

constexpr int WARPS_PER_BLOCK = 8, TILE_M = 128, D = 128;
constexpr int PLANE = TILE_M * D * 2;          // 32 KB
constexpr int IMAGE = 2 * PLANE;               // 64 KB per (tile, level)

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    cumspmm_packed_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                          const uint8_t* __restrict__ lvl, const float* __restrict__ x, int64_t ldx, int k, int64_t n_rows,
                          uint8_t* __restrict__ u) {
    constexpr int UNROLL = 8;
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    float4 P = make_float4(0.f, 0.f, 0.f, 0.f), S = P;
    int cur = 0;
    // this lane's 16-byte unit inside a (tile, level) image: k-block lane/2, row row%128; odd lanes write the lo plane
    uint8_t* unit = u + (row >> 7) * (int64_t)k * IMAGE + ((lane & 1) ? PLANE : 0) + (lane >> 1) * (TILE_M * 16) + (row & 127) * 16;

    auto emit_until = [&](int lev) {
        while (cur < lev) {
            S.x += P.x;
            S.y += P.y;
            S.z += P.z;
            S.w += P.w;
            uint32_t h0, l0, h1, l1;
            split2(fmaxf(S.x, 0.f), fmaxf(S.y, 0.f), h0, l0);
            split2(fmaxf(S.z, 0.f), fmaxf(S.w, 0.f), h1, l1);
            const bool odd = lane & 1;
            const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? h0 : l0, 1);   // even lanes give their lo, odd lanes their hi
            const uint32_t r1 = __shfl_xor_sync(0xffffffffu, odd ? h1 : l1, 1);
            const uint4 v = odd ? make_uint4(r0, r1, l0, l1) : make_uint4(h0, h1, r0, r1);
            asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(unit + (int64_t)cur * IMAGE), "r"(v.x),
                         "r"(v.y), "r"(v.z), "r"(v.w)
                         : "memory");
            ++cur;
        }
    };

    for (int base = start; base < end; base += 32) {
        const int cnt = min(32, end - base);
        int c = 0, l = 0;
        float w = 0.f;
        if (lane < cnt) {
            c = __ldg(col + base + lane);
            w = __ldg(val + base + lane);
            l = __ldg(lvl + base + lane);
        }
        for (int j = 0; j < cnt; j += UNROLL) {
            float4 xv[UNROLL];
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) {
                const int cj = __shfl_sync(0xffffffffu, c, (j + q) & 31);
                xv[q] = (j + q < cnt) ? __ldg(reinterpret_cast<const float4*>(x + (int64_t)cj * ldx) + lane)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) {
                const float wj = __shfl_sync(0xffffffffu, w, (j + q) & 31);
                const int lj = __shfl_sync(0xffffffffu, l, (j + q) & 31);
                if (j + q < cnt) {  // warp-uniform
                    emit_until(lj & 127);
                    float4& acc = (lj & 128) ? S : P;
                    acc.x = fmaf(wj, xv[q].x, acc.x);
                    acc.y = fmaf(wj, xv[q].y, acc.y);
                    acc.z = fmaf(wj, xv[q].z, acc.z);
                    acc.w = fmaf(wj, xv[q].w, acc.w);
                }
            }
        }
    }
    emit_until(k);
}

}  // namespace
}  // namespace ctgcn

using namespace ctgcn;

extern "C" size_t ctgcn_cumspmm_packed_bytes(const ctgcn_plan* plan) {
    if (!plan) return 0;
    return (size_t)((plan->n_rows + TILE_M - 1) / TILE_M) * plan->k * IMAGE;
}

// U = relu(cumulative k-core sums) of a 128-wide x, pre-split (see the file header).  u: ctgcn_cumspmm_packed_bytes(plan) bytes,
// 16-byte aligned; rows ≥ n_rows of the last tile are left untouched.
extern "C" int ctgcn_cumspmm_fwd_packed(const ctgcn_plan* plan, const float* x, int64_t ldx, int d, void* u, void* stream) {
    CTGCN_REQUIRE(plan && x && u, "cumspmm_fwd_packed: NULL argument");
    CTGCN_REQUIRE(d == D && ldx >= d && (ldx % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(u) & 15) == 0,
                  "cumspmm_fwd_packed: needs d = 128 and 16-byte aligned rows");
    if (plan->n_rows == 0) return CTGCN_OK;
    ProfScope prof(PROF_SPMM, (cudaStream_t)stream);
    const unsigned blocks = (unsigned)((plan->n_rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    cumspmm_packed_kernel<<<blocks, WARPS_PER_BLOCK * 32, 0, (cudaStream_t)stream>>>(plan->rowptr, plan->col, plan->val, plan->lvl, x,
                                                                                    ldx, plan->k, plan->n_rows, (uint8_t*)u);
    CTGCN_LAUNCH_OK("cumspmm_packed_kernel");
    return CTGCN_OK;
}
