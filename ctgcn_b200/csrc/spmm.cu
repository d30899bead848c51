// Cumulative k-core SpMM over the level-tagged union CSR (HBM-bound gather kernel).
//
// Reference arithmetic (layers.py:41-48):  S_i = S_{i-1} + A_i·x,  U_i = relu(S_i),  i = 0..K-1,
// done there as K separate torch.sparse.mm calls over K nested COO matrices.  Here every distinct
// edge is gathered ONCE: a warp owns one output row and walks its entries in level order keeping
//   P = Σ_{nested entries with level ≤ i} w·x[col]       (what A_i·x contributes from nested entries)
//   S = running cumulative sum; one-shot entries (present in A_level only) are added to S directly.
// At the end of level i:  S += P;  U_i = relu(S).                     
//
// Memory behaviour: per entry 4 B col + 4 B val + 1 B level, read coalesced 32 entries at a time by the
// warp and broadcast with shuffles; per entry one feature row of x (4·D bytes) read with 128-bit loads,
// UNROLL rows in flight per lane; per (row, level) one coalesced 4·D-byte store.
#include "common.cuh"
#include <algorithm>

namespace ctgcn {
namespace {

constexpr int WARPS_PER_BLOCK = 8;

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void st_f4_stream(float* p, const float4& v) {
    // outputs are written once and consumed by the next kernel: do not pollute L1
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& x) {
    a.x = fmaf(w, x.x, a.x);
    a.y = fmaf(w, x.y, a.y);
    a.z = fmaf(w, x.z, a.z);
    a.w = fmaf(w, x.w, a.w);
}

// D % 4 == 0.  NV = float4 slots per lane: lane owns float4 indices lane + 32*v (< D/4).
// MINB = minimum resident blocks per SM the register allocation is held to: the 128-d build (NV = 1, 8 rows in flight per lane)
// compiled for 4 blocks (64 registers, 84 bytes of spills) measured 6 % FASTER than the unconstrained 80-register build at
// cfg 4 (2.96 vs 3.16 ms, bit-identical; profiles/r02_experiments.md) — more warps in flight hide the gather latency.
template <int NV, int UNROLL, bool RELU, int MINB>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, MINB)
    cumspmm_vec_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rowend, const int32_t* __restrict__ col,
                       const float* __restrict__ val, const uint8_t* __restrict__ lvl, const float* __restrict__ x,
                       int64_t ldx, int d, int k, int64_t n_rows, float* __restrict__ u) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int d4 = d >> 2;
    const int start = rowptr[row], end = rowend[row];   // rowend = rowptr + 1, or the plan's row_end / segment ends (hub rows)

    float4 P[NV], S[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) P[v] = S[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = 0;
    float* urow = u + row * (int64_t)k * d;

    auto emit_until = [&](int lev) {
        while (cur < lev) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                S[v].x += P[v].x;
                S[v].y += P[v].y;
                S[v].z += P[v].z;
                S[v].w += P[v].w;
                const int i4 = lane + 32 * v;
                if (i4 < d4) {
                    float4 o = S[v];
                    if (RELU) {
                        o.x = fmaxf(o.x, 0.f);
                        o.y = fmaxf(o.y, 0.f);
                        o.z = fmaxf(o.z, 0.f);
                        o.w = fmaxf(o.w, 0.f);
                    }
                    st_f4_stream(urow + (int64_t)cur * d + 4 * i4, o);
                }
            }
            ++cur;
        }
    };

    for (int base = start; base < end; base += 32) {
        const int cnt = min(32, end - base);
        int c = 0;
        float w = 0.f;
        int l = 0;
        if (lane < cnt) {
            c = __ldg(col + base + lane);
            w = __ldg(val + base + lane);
            l = __ldg(lvl + base + lane);
        }
        for (int j = 0; j < cnt; j += UNROLL) {
            float4 xv[UNROLL][NV];
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) {
                const int cj = __shfl_sync(0xffffffffu, c, (j + q) & 31);
                const float* xr = x + (int64_t)cj * ldx;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int i4 = lane + 32 * v;
                    xv[q][v] = (j + q < cnt && i4 < d4) ? ldg_f4(xr + 4 * i4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int q = 0; q < UNROLL; ++q) {
                const float wj = __shfl_sync(0xffffffffu, w, (j + q) & 31);
                const int lj = __shfl_sync(0xffffffffu, l, (j + q) & 31);
                if (j + q < cnt) {  // warp-uniform
                    emit_until(lj & 127);
                    if (lj & 128) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) fma4(S[v], wj, xv[q][v]);
                    } else {
#pragma unroll
                        for (int v = 0; v < NV; ++v) fma4(P[v], wj, xv[q][v]);
                    }
                }
            }
        }
    }
    emit_until(k);
}

// Any D ≤ 32·DPL: lane owns elements lane + 32*j.
template <int DPL, bool RELU>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    cumspmm_scalar_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                          const float* __restrict__ val, const uint8_t* __restrict__ lvl,
                          const float* __restrict__ x, int64_t ldx, int d, int k, int64_t n_rows,
                          float* __restrict__ u) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    float P[DPL], S[DPL];
#pragma unroll
    for (int j = 0; j < DPL; ++j) P[j] = S[j] = 0.f;
    int cur = 0;
    float* urow = u + row * (int64_t)k * d;
    auto emit_until = [&](int lev) {
        while (cur < lev) {
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                S[j] += P[j];
                const int e = lane + 32 * j;
                if (e < d) urow[(int64_t)cur * d + e] = RELU ? fmaxf(S[j], 0.f) : S[j];
            }
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
        for (int q = 0; q < cnt; ++q) {
            const int cj = __shfl_sync(0xffffffffu, c, q);
            const float wj = __shfl_sync(0xffffffffu, w, q);
            const int lj = __shfl_sync(0xffffffffu, l, q);
            const float* xr = x + (int64_t)cj * ldx;
            float xv[DPL];
#pragma unroll
            for (int j = 0; j < DPL; ++j) {
                const int e = lane + 32 * j;
                xv[j] = e < d ? __ldg(xr + e) : 0.f;
            }
            emit_until(lj & 127);
            if (lj & 128) {
#pragma unroll
                for (int j = 0; j < DPL; ++j) S[j] = fmaf(wj, xv[j], S[j]);
            } else {
#pragma unroll
                for (int j = 0; j < DPL; ++j) P[j] = fmaf(wj, xv[j], P[j]);
            }
        }
    }
    emit_until(k);
}

// y[row, :] = act( Σ_e val_e · wt[col_e, :] + b )   — sparse-input first MLP layer (layers.py:97,103 with a
// sparse COO x, e.g. the one-hot identity of helper.py:169-172: then this is a pure row gather of Wᵀ).
template <int DPL>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    spmm_linear_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                       const float* __restrict__ val, const float* __restrict__ wt, const float* __restrict__ bias,
                       int64_t d_out, int act, int64_t n_rows, float* __restrict__ y, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    for (int64_t f0 = 0; f0 < d_out; f0 += 32 * DPL) {
        float acc[DPL];
#pragma unroll
        for (int j = 0; j < DPL; ++j) acc[j] = 0.f;
        for (int base = start; base < end; base += 32) {
            const int cnt = min(32, end - base);
            int c = 0;
            float w = 0.f;
            if (lane < cnt) {
                c = __ldg(col + base + lane);
                w = __ldg(val + base + lane);
            }
            for (int q = 0; q < cnt; ++q) {
                const int cj = __shfl_sync(0xffffffffu, c, q);
                const float wj = __shfl_sync(0xffffffffu, w, q);
                const float* wr = wt + (int64_t)cj * d_out + f0;
#pragma unroll
                for (int j = 0; j < DPL; ++j) {
                    const int64_t e = lane + 32 * j;
                    if (f0 + e < d_out) acc[j] = fmaf(wj, __ldg(wr + e), acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
            const int64_t e = f0 + lane + 32 * j;
            if (e < d_out) {
                float v = acc[j] + (bias ? __ldg(bias + e) : 0.f);
                if (act == CTGCN_ACT_SELU) {
                    const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
                    v = scale * (v > 0.f ? v : alpha * expm1f(v));
                }
                y[row * ldy + e] = v;
            }
        }
    }
}

// ---- backward of the cumulative SpMM w.r.t. x.
// S_i = Σ_{j≤i} A_j x  ⇒  dx = Σ_j A_jᵀ Zo_j with Zo_j = Σ_{i≥j} g_i (g_i = dL/dS_i).  Over the union CSR of the TRANSPOSED
// list a nested entry of level ℓ (present in A_j for every j ≥ ℓ) contributes w·Σ_{j≥ℓ} Zo_j = w·Zn_ℓ, a one-shot entry w·Zo_ℓ:
//   dx[row] = Σ_e w_e · (one-shot ? Zo : Zn)[col_e, level_e, :]        — one gathered row per stored entry.
__global__ void suffix_sums_kernel(const float* __restrict__ g, int64_t n, int k, int d, float* __restrict__ zo,
                                   float* __restrict__ zn) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= n * d) return;
    const int64_t r = idx / d, f = idx % d;
    const int64_t base = r * (int64_t)k * d + f;
    float a = 0.f, b = 0.f;
    for (int i = k - 1; i >= 0; --i) {
        a += g[base + (int64_t)i * d];
        b += a;
        zo[base + (int64_t)i * d] = a;
        zn[base + (int64_t)i * d] = b;
    }
}

__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32)
    levelgather_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ val,
                       const uint8_t* __restrict__ lvl, const float* __restrict__ zo, const float* __restrict__ zn, int d, int k,
                       int64_t n_rows, float* __restrict__ dx, int64_t lddx) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const int start = rowptr[row], end = rowptr[row + 1];
    for (int f0 = 0; f0 < d; f0 += 128) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int base = start; base < end; base += 32) {
            const int cnt = min(32, end - base);
            int c = 0, l = 0;
            float w = 0.f;
            if (lane < cnt) {
                c = __ldg(col + base + lane);
                w = __ldg(val + base + lane);
                l = __ldg(lvl + base + lane);
            }
            for (int q = 0; q < cnt; ++q) {
                const int cj = __shfl_sync(0xffffffffu, c, q);
                const float wj = __shfl_sync(0xffffffffu, w, q);
                const int lj = __shfl_sync(0xffffffffu, l, q);
                const float* src = ((lj & 128) ? zo : zn) + ((int64_t)cj * k + (lj & 127)) * d + f0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = lane + 32 * j;
                    if (f0 + e < d) acc[j] = fmaf(wj, __ldg(src + e), acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = f0 + lane + 32 * j;
            if (e < d) dx[row * lddx + e] = acc[j];
        }
    }
}

__global__ void transpose_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, float* __restrict__ dst) {
    __shared__ float tile[32][33];
    const int64_t c0 = blockIdx.x * 32ll, r0 = blockIdx.y * 32ll;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * cols + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[c * rows + r] = tile[threadIdx.x][i];
    }
}

// Hub rows: u[hub row − row0, i, :] = relu?(Σ_segments partial[seg, i, :]), segments added in order.  Block = (hub row, level).
template <bool RELU>
__global__ void __launch_bounds__(128) hub_combine_kernel(const float* __restrict__ partial, const int32_t* __restrict__ hub_rows,
                                                          const int32_t* __restrict__ hub_seg_ptr, int h0, int s0, int64_t row0, int d, int k,
                                                          float* __restrict__ u) {
    const int hidx = h0 + blockIdx.x, i = blockIdx.y;
    const int sa = hub_seg_ptr[hidx] - s0, sb = hub_seg_ptr[hidx + 1] - s0;
    float* dst = u + ((int64_t)(hub_rows[hidx] - row0) * k + i) * d;
    for (int f = 4 * threadIdx.x; f < d; f += 4 * blockDim.x) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = sa; s < sb; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(partial + ((int64_t)s * k + i) * d + f);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }
        if (RELU) {
            acc.x = fmaxf(acc.x, 0.f);
            acc.y = fmaxf(acc.y, 0.f);
            acc.z = fmaxf(acc.z, 0.f);
            acc.w = fmaxf(acc.w, 0.f);
        }
        *reinterpret_cast<float4*>(dst + f) = acc;
    }
}

template <int NV, int UNROLL, bool RELU>
void launch_vec_one(const int32_t* rowptr, const int32_t* rowend, const ctgcn_plan* p, const float* x, int64_t ldx, int d, float* u,
                    cudaStream_t st, int64_t rows) {
    const unsigned blocks = (unsigned)((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    constexpr int MINB = NV == 1 ? 4 : 1;
    cumspmm_vec_kernel<NV, UNROLL, RELU, MINB><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(rowptr, rowend, p->col, p->val, p->lvl, x, ldx, d,
                                                                                       p->k, rows, u);
}

// Row range [row0, row0 + rows): the kernels index rows locally, so the range is selected by shifting the row-pointer array
// (entry offsets stay absolute) — u is the chunk's own [rows, K, d] buffer.
template <int NV, int UNROLL>
int launch_vec(const ctgcn_plan* p, const float* x, int64_t ldx, int d, float* u, bool relu, cudaStream_t st, int64_t row0,
               int64_t rows) {
    // hub rows of the range (plan.cu: build_hub_split): emptied in the main pass, then segments → partial sums → combine
    int h0 = 0, h1 = 0;
    const bool hubs = p->hub_scratch && d <= ctgcn_plan::HUB_DMAX;
    if (hubs) {
        h0 = (int)(std::lower_bound(p->h_hub_rows.begin(), p->h_hub_rows.end(), (int32_t)row0) - p->h_hub_rows.begin());
        h1 = (int)(std::lower_bound(p->h_hub_rows.begin(), p->h_hub_rows.end(), (int32_t)(row0 + rows)) - p->h_hub_rows.begin());
    }
    const int32_t* rowend = hubs ? p->row_end + row0 : p->rowptr + row0 + 1;
    if (relu) launch_vec_one<NV, UNROLL, true>(p->rowptr + row0, rowend, p, x, ldx, d, u, st, rows);
    else launch_vec_one<NV, UNROLL, false>(p->rowptr + row0, rowend, p, x, ldx, d, u, st, rows);
    CTGCN_LAUNCH_OK("cumspmm_vec_kernel");
    if (h1 > h0) {
        const int s0 = p->h_hub_seg_ptr[h0], s1 = p->h_hub_seg_ptr[h1];
        launch_vec_one<NV, UNROLL, false>(p->seg_start + s0, p->seg_end + s0, p, x, ldx, d, p->hub_scratch, st, s1 - s0);
        CTGCN_LAUNCH_OK("cumspmm_vec_kernel (hub segments)");
        const dim3 grid((unsigned)(h1 - h0), (unsigned)p->k);
        if (relu) hub_combine_kernel<true><<<grid, 128, 0, st>>>(p->hub_scratch, p->hub_rows, p->hub_seg_ptr, h0, s0, row0, d, p->k, u);
        else hub_combine_kernel<false><<<grid, 128, 0, st>>>(p->hub_scratch, p->hub_rows, p->hub_seg_ptr, h0, s0, row0, d, p->k, u);
        CTGCN_LAUNCH_OK("hub_combine_kernel");
    }
    return CTGCN_OK;
}

template <int DPL>
int launch_scalar(const ctgcn_plan* p, const float* x, int64_t ldx, int d, float* u, bool relu, cudaStream_t st, int64_t row0,
                  int64_t rows) {
    const unsigned blocks = (unsigned)((rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    if (relu)
        cumspmm_scalar_kernel<DPL, true><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(p->rowptr + row0, p->col, p->val, p->lvl, x, ldx,
                                                                                  d, p->k, rows, u);
    else
        cumspmm_scalar_kernel<DPL, false><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(p->rowptr + row0, p->col, p->val, p->lvl, x, ldx,
                                                                                   d, p->k, rows, u);
    CTGCN_LAUNCH_OK("cumspmm_scalar_kernel");
    return CTGCN_OK;
}

}  // namespace

int launch_cumspmm(const ctgcn_plan* p, const float* x, int64_t ldx, int d, float* u, bool relu, cudaStream_t st, int64_t row0,
                   int64_t rows) {
    CTGCN_REQUIRE(d >= 1 && d <= 1024, "cumspmm: feature width %d outside [1,1024]", d);
    if (rows < 0) rows = p->n_rows - row0;
    CTGCN_REQUIRE(row0 >= 0 && rows >= 0 && row0 + rows <= p->n_rows, "cumspmm: row range outside the plan");
    if (rows == 0) return CTGCN_OK;
    ProfScope prof(PROF_SPMM, st);
    const bool vec_ok = (d % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(u) & 15) == 0);
    if (vec_ok) {
        const int d4 = d / 4;
        if (d4 <= 32) return launch_vec<1, 8>(p, x, ldx, d, u, relu, st, row0, rows);
        if (d4 <= 64) return launch_vec<2, 4>(p, x, ldx, d, u, relu, st, row0, rows);
        if (d4 <= 128) return launch_vec<4, 2>(p, x, ldx, d, u, relu, st, row0, rows);
        return launch_vec<8, 1>(p, x, ldx, d, u, relu, st, row0, rows);
    }
    if (d <= 32) return launch_scalar<1>(p, x, ldx, d, u, relu, st, row0, rows);
    if (d <= 64) return launch_scalar<2>(p, x, ldx, d, u, relu, st, row0, rows);
    if (d <= 128) return launch_scalar<4>(p, x, ldx, d, u, relu, st, row0, rows);
    if (d <= 256) return launch_scalar<8>(p, x, ldx, d, u, relu, st, row0, rows);
    if (d <= 512) return launch_scalar<16>(p, x, ldx, d, u, relu, st, row0, rows);
    return launch_scalar<32>(p, x, ldx, d, u, relu, st, row0, rows);
}

int launch_cumspmm_bwd(const ctgcn_plan* pt, const float* g, int d, float* zo, float* zn, float* dx, int64_t lddx,
                       cudaStream_t st) {
    CTGCN_REQUIRE(d >= 1 && d <= 1024, "cumspmm_bwd: feature width %d outside [1,1024]", d);
    ProfScope prof(PROF_SPMM, st);
    const int64_t total = pt->n_cols * (int64_t)d;
    if (total > 0) {
        suffix_sums_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, pt->n_cols, pt->k, d, zo, zn);
        CTGCN_LAUNCH_OK("suffix_sums_kernel");
    }
    const unsigned blocks = (unsigned)((pt->n_rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    levelgather_kernel<<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(pt->rowptr, pt->col, pt->val, pt->lvl, zo, zn, d, pt->k,
                                                               pt->n_rows, dx, lddx);
    CTGCN_LAUNCH_OK("levelgather_kernel");
    return CTGCN_OK;
}

int launch_spmm_linear(const ctgcn_plan* p, const float* wt, const float* b, int64_t d_out, int act, float* y, int64_t ldy,
                       cudaStream_t st) {
    ProfScope prof(PROF_SPMM_LINEAR, st);
    const unsigned blocks = (unsigned)((p->n_rows + WARPS_PER_BLOCK - 1) / WARPS_PER_BLOCK);
    spmm_linear_kernel<4><<<blocks, WARPS_PER_BLOCK * 32, 0, st>>>(p->rowptr, p->col, p->val, wt, b, d_out, act, p->n_rows, y,
                                                                  ldy);
    CTGCN_LAUNCH_OK("spmm_linear_kernel");
    return CTGCN_OK;
}

int launch_transpose(const float* src, int64_t rows, int64_t cols, float* dst, cudaStream_t st) {
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32));
    CTGCN_REQUIRE(grid.y <= 65535, "transpose: too many rows (%lld)", (long long)rows);
    ProfScope prof(PROF_PACK, st);
    transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(src, rows, cols, dst);
    CTGCN_LAUNCH_OK("transpose_kernel");
    return CTGCN_OK;
}

}  // namespace ctgcn
