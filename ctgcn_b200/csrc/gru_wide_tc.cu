// tcgen05 GRU over a short sequence for WIDE hidden states (H = 128·u, u ≤ 4; d_in ≤ 1024, a multiple of 4): BASELINE.json
// configs[4] (256-d CoreDiffusion + temporal GRU, layers.py:59-62 / models.py:249-250), which the 128-wide kernels of gru_tc2.cu do
// not take — their design keeps U, h and both gate accumulator sets of a 128-row tile on the SM, and at H = 256 neither the
// operands (h alone is 128 KB as bf16 hi|lo planes) nor the accumulators (r, z, n_x, n_h × 256 = 1024 TMEM columns) fit.
//
// So the recurrence runs ONE STEP PER LAUNCH with h in global memory, over row chunks small enough that the two h buffers and the
// Σh buffer of a chunk stay in the 126 MB L2 (148·4 work units per launch; h never goes to DRAM), and a launch is a GEMM with a
// fused gate epilogue:
//   work unit = (128-row tile, 128 hidden features j ∈ [128u, 128u+128)); TMEM: r | z | n_x | n_h, 128 columns each (all 512)
//   for slice s of [x_i | h_{i-1}] (64 columns, fp32 rows → bf16 hi|lo planes by the loader warps):
//     for gate g ∈ {r, z, n}: D_g += A_s · W_{g,u,s}ᵀ    (4 K-steps × 3 split MMAs, M = 128, N = 128; the n gate of an h slice goes
//                                                         to its own columns: n = tanh(n_x + b_in + r ⊙ (n_h + b_hn)))
//   epilogue (8 warps, thread = row, 64 columns each): + biases, ex2/rcp gate math (tc_common.cuh, shared with gru_tc2.cu),
//     h_i = (1 − z) n + z h_{i-1} → global, Σh accumulated in place.
// Weights (pre-scaled by −log2e / 2·log2e, packed once per call) stream from the L2 through a 128 KB ring; every unit re-reads its
// (d_in + H)·384·4 B.  What binds is the SM's shared-memory data path (ncu: l1tex 78 %, L2 33 %, tensor pipe 41 % of a launch): an
// N = 128 MMA reads 8 KB of operands per 64 tensor cycles — hence CTA PAIRS (CG = 2, see WL below): each CTA holds half of every
// weight chunk.  Measured at 256 → 256 (profiles/r02_experiments.md): unit period 38.5 K cycles one-CTA → 36.3 K paired = 22.8 K MMA
// phase (18.4 K at the tensor peak) + 11-12 K epilogue (TMEM is full, so it does not overlap the next unit's MMAs) + hand-over;
// 265 → 274-285 TFLOP/s algorithmic (≈ 0.2 of the dense bf16 peak with 3 MMAs per product), 12× the fp32 kernel.
// LayerNorm (of Σh, or of every h_i in place for the temporal mode) is a separate row kernel per chunk.
//   warp 0       weight producer (cp.async.bulk + mbarrier tx)
//   warp 1       MMA issuer (leader CTA) / relay of "my half chunk has landed" (follower)
//   warps 4-11   loaders (16 rows each)
//   warps 12-19  gate epilogue
#include "common.cuh"
#include "tc_common.cuh"

namespace ctgcn {
namespace {
using namespace tc;

constexpr int TILE_M = 128, SLICE_K = 64, UNIT_N = 128, MAX_H = 512;
constexpr int PLANE = TILE_M * SLICE_K * 2;          // 16 KB: one bf16 plane of a 128 × 64 A slice
constexpr int BLOCK = 2 * PLANE;                     // hi | lo: an A slot, and a whole packed weight chunk (128 rows × 64 k)
// CG = CTAs per MMA.  CG = 2 (default): the two CTAs of a cluster work on the SAME feature block of two row tiles; every weight chunk
// is split between them (64 of its 128 rows each), the leader issues M = 256 `cta_group::2` MMAs, accumulators stay per CTA.  Per SM
// that halves the B-operand reads and the ring writes of the shared-memory data path, which is what bound the one-CTA build.
template <int CG>
struct WL {
    static constexpr int ROWS = UNIT_N / CG;         // weight rows of a chunk held by one CTA
    static constexpr int B_PLANE = ROWS * SLICE_K * 2;
    static constexpr int B_BLOCK = 2 * B_PLANE;      // this CTA's part of a chunk: 32 KB / 16 KB
    static constexpr int STAGES = 4 * CG;            // 128 KB of ring either way
    static constexpr int SM_A = 0;                   // 2 slots
    static constexpr int SM_B = SM_A + 2 * BLOCK;    // ring
    static constexpr int SM_BIAS = SM_B + STAGES * B_BLOCK;   // [4][MAX_H] floats: b_r, b_z, b_in, b_hn (pre-scaled)
    static constexpr int SM_BAR = SM_BIAS + 4 * MAX_H * 4;
    // B_PEER (leader only, CG = 2): "the follower's half of the chunk has landed", relayed by the follower's warp 1
    enum { A_READY = 0, A_FREE = 2, B_FULL = 4, B_EMPTY = 4 + STAGES, B_PEER = 4 + 2 * STAGES, ACC_FULL = 4 + 3 * STAGES,
           ACC_FREE = 5 + 3 * STAGES, NUM_BARS = 6 + 3 * STAGES };
    static constexpr int SM_TMEM_PTR = SM_BAR + NUM_BARS * 8;
    static constexpr int SMEM_BYTES = SM_TMEM_PTR + 16;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};
constexpr int NUM_LOADER_WARPS = 8, NUM_EPI_WARPS = 8, FIRST_LOADER_WARP = 4, FIRST_EPI_WARP = 12, THREADS = 640;
constexpr int REG_WG0 = 40, REG_LOAD = 64, REG_EPI = 152;      // setmaxnreg budgets out of 640 × 96
static_assert(128 * REG_WG0 + 256 * REG_LOAD + 256 * REG_EPI <= THREADS * 96, "register pool");
constexpr int SM_COUNT_SIZING = 148;                            // rows of a chunk: 148·UNITS_PER_CTA units (workspace sizing is device-independent)

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Packed weights: chunk (u, s, g) at ((u·(nx + nh) + s)·3 + g)·BLOCK, s < nx: W_ih columns [64s, 64s+64) (zero-padded), else W_hh;
// rows = the 128 features of unit u of gate g ∈ {r, z, n}.  Inside a chunk: rank 0's 128/CG rows (hi plane | lo plane), then rank 1's;
// element (row, k) of a plane at (k/8)·(ROWS·16) + row·16 + (k%8)·2.
// bias4 [4][MAX_H]: (b_ir + b_hr)·(−log2e), (b_iz + b_hz)·(−log2e), b_in·2log2e, b_hn·2log2e.
template <int CG>
__global__ void pack_wide_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                 const float* __restrict__ b_hh, int d_in, int h, int nx, int nh, uint8_t* __restrict__ packed,
                                 float* __restrict__ bias4) {
    using L = WL<CG>;
    constexpr float kLog2e = 1.4426950408889634f;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 4 * MAX_H) {
        const int g4 = t / MAX_H, f = t % MAX_H;
        float bv = 0.f;
        if (b_ih && f < h) {
            if (g4 == 0) bv = (b_ih[f] + b_hh[f]) * -kLog2e;
            else if (g4 == 1) bv = (b_ih[h + f] + b_hh[h + f]) * -kLog2e;
            else if (g4 == 2) bv = b_ih[2 * h + f] * (2.f * kLog2e);
            else bv = b_hh[2 * h + f] * (2.f * kLog2e);
        }
        bias4[t] = bv;
    }
    constexpr int UNITS = UNIT_N * (SLICE_K / 8);     // 16-byte units of one plane of a whole chunk
    const int nu = h / UNIT_N, ns = nx + nh;
    if (t >= nu * ns * 3 * UNITS) return;
    const int chunk = t / UNITS, unit = t % UNITS;
    const int g = chunk % 3, s = (chunk / 3) % ns, u = chunk / (3 * ns);
    const int kb = unit / UNIT_N, row = unit % UNIT_N;
    const bool is_x = s < nx;
    const int ktot = is_x ? d_in : h, k0 = (is_x ? s : s - nx) * SLICE_K + kb * 8;
    const float* src = (is_x ? w_ih : w_hh) + (int64_t)(g * h + u * UNIT_N + row) * ktot + k0;
    const float scale = g < 2 ? -kLog2e : 2.f * kLog2e;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = k0 + i < ktot ? src[i] * scale : 0.f;
    uint4 hi, lo;
    split8(v, hi, lo);
    const int rank = row / L::ROWS, rr = row % L::ROWS;
    uint8_t* dst = packed + (size_t)chunk * BLOCK + (size_t)rank * L::B_BLOCK + kb * (L::ROWS * 16) + rr * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + L::B_PLANE) = lo;
}

struct ParamsW {
    const float* x;          // step input rows [n, d_in]
    int64_t ldx;
    const float* hprev;      // [n, h] or NULL (first step: h = 0)
    int64_t ldh;
    float* hnew;             // [n, h] or NULL (last step of the Σ mode)
    int64_t ldn;
    float* sum;              // [n, h] running Σh, or NULL
    int64_t lds;
    int sum_add;             // 0: store, 1: add
    int64_t n;
    int d_in, h, nx, nh, ns_packed, nu;   // nh = 0 at the first step; ns_packed = slices per unit in the packed image
    const uint8_t* packed;
    const float* bias4;
    int num_units;
    long long* trace;        // optional [32 events][64 units] clock64 stamps of block 0 (ctgcn_debug_gru_trace), else NULL
};

#define WIDE_TRACE(e, t)                                                                       \
    do {                                                                                       \
        if (p.trace && blockIdx.x == 0 && (t) < 64) p.trace[(e) * 64 + (t)] = clock64();       \
    } while (0)
#define WIDE_TRACE_ADD(e, t, v)                                                                \
    do {                                                                                       \
        if (p.trace && blockIdx.x == 0 && (t) < 64) p.trace[(e) * 64 + (t)] = (v);             \
    } while (0)

template <int CG>
__global__ void __launch_bounds__(THREADS, 1) gru_wide_step_kernel(const ParamsW p) {
    using L = WL<CG>;
    constexpr int STAGES = L::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int i) { return sbase + L::SM_BAR + 8u * i; };
    const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
    const int cluster_id = (int)blockIdx.x / CG, nclusters = (int)gridDim.x / CG;
    // p.num_units counts (tile group of CG tiles, feature block) pairs; unit v of this cluster → feature block v % nu, tile (v / nu)·CG + rank
    const int my_units = (p.num_units - cluster_id + nclusters - 1) / nclusters;
    const int ns = p.nx + p.nh;
    if (threadIdx.x == 0) WIDE_TRACE(10, 0);

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(L::A_READY + b), CG * NUM_LOADER_WARPS);
            mbar_init(bar(L::A_FREE + b), 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(L::B_FULL + s), 1);
            mbar_init(bar(L::B_EMPTY + s), 1);
            mbar_init(bar(L::B_PEER + s), 1);
        }
        mbar_init(bar(L::ACC_FULL), 1);
        mbar_init(bar(L::ACC_FREE), CG * NUM_EPI_WARPS);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 4 * MAX_H; i += THREADS) reinterpret_cast<float*>(smem + L::SM_BIAS)[i] = p.bias4[i];
    if constexpr (CG == 2) {
        cluster_sync_all();                                     // barriers initialised in BOTH CTAs before anyone arrives remotely
        if (warp == 1) tmem_alloc2(sbase + L::SM_TMEM_PTR, 512);
        tc_fence_before();
        cluster_sync_all();
    } else {
        if (warp == 1) tmem_alloc(sbase + L::SM_TMEM_PTR, 512);
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::SM_TMEM_PTR);
    auto unit_of = [&](int t) { return cluster_id + t * nclusters; };

    if (warp < FIRST_LOADER_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_WG0));
        if (warp == 0) {
            // ================================================= weight producer (this CTA's part of every chunk)
            if (lane == 0) {
                uint32_t stage = 0, phase = 0;
                for (int t = 0; t < my_units; ++t) {
                    const int u = unit_of(t) % p.nu;
                    const uint8_t* src = p.packed + (size_t)u * p.ns_packed * 3 * BLOCK + (size_t)rank * L::B_BLOCK;
                    for (int c = 0; c < ns * 3; ++c) {
                        mbar_wait(bar(L::B_EMPTY + stage), phase ^ 1);
                        mbar_expect_tx(bar(L::B_FULL + stage), L::B_BLOCK);
                        bulk_g2s(sbase + L::SM_B + stage * L::B_BLOCK, src + (size_t)c * BLOCK, L::B_BLOCK, bar(L::B_FULL + stage));
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        } else if (warp == 1 && CG == 2 && rank != 0) {
            // ================================================= follower: relay "my half of the chunk has landed" to the leader
            uint32_t stage = 0, phase = 0;
            for (int t = 0; t < my_units; ++t)
                for (int c = 0; c < ns * 3; ++c) {
                    mbar_wait(bar(L::B_FULL + stage), phase);
                    if (lane == 0) mbar_arrive_remote(bar(L::B_PEER + stage), 0);
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
        } else if (warp == 1) {
            // ================================================= MMA issuer (leader of the pair)
            constexpr uint32_t idesc = umma_idesc_bf16(TILE_M * CG, UNIT_N);
            constexpr uint32_t A_STEP = (2 * TILE_M * 16) >> 4, A_PL = PLANE >> 4;
            constexpr uint32_t B_STEP = (2 * L::ROWS * 16) >> 4, B_PL = L::B_PLANE >> 4;
            uint32_t stage = 0, phase = 0, use = 0;              // use: running slice counter (slot = use & 1)
            for (int t = 0; t < my_units; ++t) {
                wait_pair<CG>(bar(L::ACC_FREE), (t & 1) ^ 1);
                tc_fence_after();
                WIDE_TRACE(0, t);
                long long wait_a = 0, wait_b = 0;
                for (int s = 0; s < ns; ++s, ++use) {
                    long long c0 = p.trace ? clock64() : 0;
                    wait_pair<CG>(bar(L::A_READY + (use & 1)), (use >> 1) & 1);
                    tc_fence_after();
                    if (p.trace) wait_a += clock64() - c0;
                    if (s == 0) WIDE_TRACE(1, t);
                    const uint32_t a0 = desc_lo(sbase + L::SM_A + (use & 1) * BLOCK, TILE_M * 16);
                    for (int g = 0; g < 3; ++g) {
                        c0 = p.trace ? clock64() : 0;
                        mbar_wait(bar(L::B_FULL + stage), phase);
                        if constexpr (CG == 2) mbar_wait_cluster(bar(L::B_PEER + stage), phase);
                        tc_fence_after();
                        if (p.trace) wait_b += clock64() - c0;
                        if (elect_one()) {
                            const uint32_t b0 = desc_lo(sbase + L::SM_B + stage * L::B_BLOCK, L::ROWS * 16);
                            const bool hn = g == 2 && s >= p.nx;                          // recurrent part of the n gate: own columns
                            const uint32_t d = tmem + (hn ? 3 : g) * UNIT_N;
                            const bool opens = hn ? s == p.nx : s == 0;
#pragma unroll
                            for (int ks = 0; ks < SLICE_K / 16; ++ks) {
                                const uint64_t ah = desc64(a0 + ks * A_STEP), al = desc64(a0 + A_PL + ks * A_STEP);
                                const uint64_t bh = desc64(b0 + ks * B_STEP), bl = desc64(b0 + B_PL + ks * B_STEP);
                                umma_cg<CG>(d, al, bh, idesc, (opens && ks == 0) ? 0u : 1u);  // small terms first
                                umma_cg<CG>(d, ah, bl, idesc, 1u);
                                umma_cg<CG>(d, ah, bh, idesc, 1u);
                            }
                            commit_cg<CG>(bar(L::B_EMPTY + stage));
                        }
                        __syncwarp();
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (elect_one()) commit_cg<CG>(bar(L::A_FREE + (use & 1)));
                    __syncwarp();
                }
                if (elect_one()) commit_cg<CG>(bar(L::ACC_FULL));
                __syncwarp();
                if (lane == 0) {
                    WIDE_TRACE(2, t);
                    WIDE_TRACE_ADD(8, t, wait_a);
                    WIDE_TRACE_ADD(9, t, wait_b);
                }
            }
        }
    } else if (warp < FIRST_EPI_WARP) {
        // ===================================================== loaders: 16 tile rows per warp, one 64-column slice at a time
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_LOAD));
        // A quarter-warp reads 128 contiguous bytes of ONE row (lane = 16-byte piece): 8 L1 data-pipe wavefronts per load instruction
        // instead of 32 for "8 rows per quarter" — this kernel is bound by the LSU data pipe (profiles/r02_experiments.md).  A lane then
        // holds half an operand unit and stores it with 64-bit stores (4-way bank conflicts: 8 wavefronts per store instruction,
        // 48 per KB of input in all against 72, without shuffles or extra latency).
        const int q4 = lane >> 3, p8 = lane & 7;
        const int row_base = 16 * (warp - FIRST_LOADER_WARP);
        uint32_t use = 0;
        for (int t = 0; t < my_units; ++t) {
            const int64_t tile_row0 = ((int64_t)(unit_of(t) / p.nu) * CG + rank) * TILE_M;
            for (int s = 0; s < ns; ++s, ++use) {
                const bool is_x = s < p.nx;
                const float* base = is_x ? p.x : p.hprev;
                const int64_t ld = is_x ? p.ldx : p.ldh;
                const int width = is_x ? p.d_in : p.h, col0 = (is_x ? s : s - p.nx) * SLICE_K;
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {                   // (row group of 4, 32-column half of the slice): 4 × 2
                    const int64_t srow = tile_row0 + row_base + 4 * (u & 3) + q4;
                    const int c0 = col0 + 32 * (u >> 2) + 4 * p8;
                    v[u] = srow < p.n && c0 + 4 <= width ? __ldg(reinterpret_cast<const float4*>(base + srow * ld + c0))
                                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (warp == FIRST_LOADER_WARP && lane == 0 && s == 0) WIDE_TRACE(6, t);
                mbar_wait(bar(L::A_FREE + (use & 1)), ((use >> 1) & 1) ^ 1);
                uint8_t* slot = smem + L::SM_A + (use & 1) * BLOCK;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int m = row_base + 4 * (u & 3) + q4, kb = 4 * (u >> 2) + (p8 >> 1);
                    uint2 hi, lo;
                    split2(v[u].x, v[u].y, hi.x, lo.x);
                    split2(v[u].z, v[u].w, hi.y, lo.y);
                    uint8_t* dst = slot + kb * (TILE_M * 16) + m * 16 + 8 * (p8 & 1);
                    *reinterpret_cast<uint2*>(dst) = hi;
                    *reinterpret_cast<uint2*>(dst + PLANE) = lo;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) arrive_leader<CG>(bar(L::A_READY + (use & 1)), rank);
                if (warp == FIRST_LOADER_WARP && lane == 0 && s == ns - 1) WIDE_TRACE(7, t);
            }
        }
    } else {
        // ===================================================== gate epilogue: thread = tile row, 64 of the unit's 128 features per warp
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_EPI));
        const int q = warp & 3, half = (warp - FIRST_EPI_WARP) >> 2;
        const int m = 32 * q + lane;
        const float* bias = reinterpret_cast<const float*>(smem + L::SM_BIAS);
        const uint32_t tmem_lane = tmem + ((uint32_t)(32 * q) << 16);
        for (int t = 0; t < my_units; ++t) {
            const int v = unit_of(t);
            const int u = v % p.nu;
            const int64_t row = ((int64_t)(v / p.nu) * CG + rank) * TILE_M + m;
            const bool ok = row < p.n;
            const int f0 = u * UNIT_N + 64 * half;              // first hidden feature of this warp's columns
            const float* hp = p.hprev ? p.hprev + row * p.ldh + f0 : nullptr;
            float* hn_out = p.hnew ? p.hnew + row * p.ldn + f0 : nullptr;
            float* sm = p.sum ? p.sum + row * p.lds + f0 : nullptr;
            // h_{i-1} of this thread's 64 features: fetched while the MMAs of the unit run (a load issued inside the gate loop waits a
            // full memory round trip under the loaders' traffic — 2-3 K cycles per 8 columns, measured as 30 K cycles per unit)
            float hold_all[64];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float4 hv = ok && hp ? __ldg(reinterpret_cast<const float4*>(hp + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                hold_all[4 * j] = hv.x;
                hold_all[4 * j + 1] = hv.y;
                hold_all[4 * j + 2] = hv.z;
                hold_all[4 * j + 3] = hv.w;
            }
            if (warp == FIRST_EPI_WARP && lane == 0) WIDE_TRACE(3, t);
            mbar_wait(bar(L::ACC_FULL), t & 1);
            tc_fence_after();
            if (warp == FIRST_EPI_WARP && lane == 0) WIDE_TRACE(4, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 64 * half + 8 * j;                // column inside the unit
                float ea[8], eb[8], gi[8], gh[8], hold[8], hn8[8];
                tmem_ld8(tmem_lane + c, ea);
                tmem_ld8(tmem_lane + UNIT_N + c, eb);
                tmem_ld8(tmem_lane + 2 * UNIT_N + c, gi);
                if (p.nh) tmem_ld8(tmem_lane + 3 * UNIT_N + c, gh);
                const int f = u * UNIT_N + c;
                const float4 br0 = *reinterpret_cast<const float4*>(bias + f), br1 = *reinterpret_cast<const float4*>(bias + f + 4);
                const float4 bz0 = *reinterpret_cast<const float4*>(bias + MAX_H + f), bz1 = *reinterpret_cast<const float4*>(bias + MAX_H + f + 4);
                const float4 bi0 = *reinterpret_cast<const float4*>(bias + 2 * MAX_H + f), bi1 = *reinterpret_cast<const float4*>(bias + 2 * MAX_H + f + 4);
                const float4 bh0 = *reinterpret_cast<const float4*>(bias + 3 * MAX_H + f), bh1 = *reinterpret_cast<const float4*>(bias + 3 * MAX_H + f + 4);
                const float br[8] = {br0.x, br0.y, br0.z, br0.w, br1.x, br1.y, br1.z, br1.w};
                const float bz[8] = {bz0.x, bz0.y, bz0.z, bz0.w, bz1.x, bz1.y, bz1.z, bz1.w};
                const float bi[8] = {bi0.x, bi0.y, bi0.z, bi0.w, bi1.x, bi1.y, bi1.z, bi1.w};
                const float bh[8] = {bh0.x, bh0.y, bh0.z, bh0.w, bh1.x, bh1.y, bh1.z, bh1.w};
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    ea[i] += br[i];
                    eb[i] += bz[i];
                    gi[i] += bi[i];
                    gh[i] = p.nh ? gh[i] + bh[i] : bh[i];
                    hold[i] = hold_all[8 * j + i];
                }
                gate_math<8>(ea, eb, gi, gh, hold, hn8);
                if (ok) {
                    if (hn_out) {
                        *reinterpret_cast<float4*>(hn_out + 8 * j) = make_float4(hn8[0], hn8[1], hn8[2], hn8[3]);
                        *reinterpret_cast<float4*>(hn_out + 8 * j + 4) = make_float4(hn8[4], hn8[5], hn8[6], hn8[7]);
                    }
                    if (sm) {                                   // Σh: one writer per element and launch → a reduction without a return value
                        if (p.sum_add) {
                            red_add4(sm + 8 * j, hn8[0], hn8[1], hn8[2], hn8[3]);
                            red_add4(sm + 8 * j + 4, hn8[4], hn8[5], hn8[6], hn8[7]);
                        } else {
                            *reinterpret_cast<float4*>(sm + 8 * j) = make_float4(hn8[0], hn8[1], hn8[2], hn8[3]);
                            *reinterpret_cast<float4*>(sm + 8 * j + 4) = make_float4(hn8[4], hn8[5], hn8[6], hn8[7]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (warp == FIRST_EPI_WARP && lane == 0) WIDE_TRACE(5, t);
            if (lane == 0) arrive_leader<CG>(bar(L::ACC_FREE), rank);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) WIDE_TRACE(11, 0);
    if constexpr (CG == 2) {
        cluster_sync_all();      // neither CTA may exit (or free TMEM) while the pair's MMAs can still touch its memory
        if (warp == 1) tmem_dealloc2(tmem, 512);
    } else {
        if (warp == 1) tmem_dealloc(tmem, 512);
    }
}

// LayerNorm of rows of width h (a multiple of 128, ≤ 512), one warp per row: src row r at src + (r / inner)·srs + (r % inner)·sss,
// dst likewise (in place allowed) or through the row scatter (Σ mode with a fused snapshot exchange).
__global__ void __launch_bounds__(256) ln_rows_wide_kernel(const float* src, int64_t srs, int64_t sss, int inner, int64_t rows,
                                                           int h, const float* __restrict__ ln_w, const float* __restrict__ ln_b, float eps,
                                                           float* dst, int64_t drs, int64_t dss, const RowScatter sc) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int64_t outer = r / inner, in = r % inner;
    const float* s = src + outer * srs + in * sss;
    float4 v[MAX_H / 128];
    const int nv = h / 128;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_H / 128; ++i)
        if (i < nv) {
            v[i] = *reinterpret_cast<const float4*>(s + 128 * i + 4 * lane);
            acc += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float mean = acc / (float)h;
    float var = 0.f;
#pragma unroll
    for (int i = 0; i < MAX_H / 128; ++i)
        if (i < nv) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            var += (a * a + b * b) + (c * c + d * d);
        }
#pragma unroll
    for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
    const float rstd = rsqrtf(var / (float)h + eps);
    float* d = sc.slices ? sc.row_ptr(r) : dst + outer * drs + in * dss;
#pragma unroll
    for (int i = 0; i < MAX_H / 128; ++i)
        if (i < nv) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(ln_w + 128 * i + 4 * lane));
            const float4 b = __ldg(reinterpret_cast<const float4*>(ln_b + 128 * i + 4 * lane));
            float4 o;
            o.x = (v[i].x - mean) * rstd * w.x + b.x;
            o.y = (v[i].y - mean) * rstd * w.y + b.y;
            o.z = (v[i].z - mean) * rstd * w.z + b.z;
            o.w = (v[i].w - mean) * rstd * w.w + b.w;
            float* q = d + 128 * i + 4 * lane;
            if ((reinterpret_cast<uintptr_t>(q) & 15) == 0) {
                *reinterpret_cast<float4*>(q) = o;
            } else {                                             // caller's [n, h] view with an odd row stride
                q[0] = o.x;
                q[1] = o.y;
                q[2] = o.z;
                q[3] = o.w;
            }
        }
}

// Work units per CTA and launch: measured 2 / 3 / 4 / 6 / 8 → 239 / 239 / 244 / 243.5 / 244 TFLOP/s (first version, 256 → 256): launch
// overhead is not what binds, so the smallest value on the plateau — the chunk's h / Σh buffers (38 MB each at H = 256) stay in the L2.
constexpr int UNITS_PER_CTA = 4;
int units_per_cta() { return UNITS_PER_CTA; }
long long* g_wide_trace = nullptr;
int64_t chunk_rows_of(int h, int upc) { return (int64_t)SM_COUNT_SIZING * upc * TILE_M / (h / UNIT_N); }
size_t packed_bytes(int d_in, int h) { return (size_t)(h / UNIT_N) * (ceil_div(d_in, SLICE_K) + h / SLICE_K) * 3 * BLOCK; }

bool wide_takes(int d_in, int h) {
    return h >= UNIT_N && h <= MAX_H && h % UNIT_N == 0 && d_in >= 8 && d_in <= 1024 && (d_in & 3) == 0;
}
size_t wide_workspace(int d_in, int h) {
    if (!wide_takes(d_in, h)) return 0;
    return align_up(packed_bytes(d_in, h), 256) + 4 * MAX_H * sizeof(float) + 3 * (size_t)chunk_rows_of(h, 4) * h * sizeof(float);
}

template <int CG>
int launch_wide(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* w_ih,
                       const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                       int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!wide_takes(d_in, h)) return 1;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(seq) || (srs & 3) || (sss & 3) || !al16(ln_w) || !al16(ln_b)) return 1;
    // temporal mode: the output slots are the h buffers of the recurrence (vector access); Σ mode: only the LayerNorm kernel writes y
    if (mode == CTGCN_GRU_EACH_LN && (!y || (sc && sc->slices) || !al16(y) || (yrs & 3) || (yss & 3))) return 1;
    const size_t need = wide_workspace(d_in, h);
    CTGCN_REQUIRE(ws && ws_bytes >= need, "gru_wide_tc: workspace of %zu bytes, need %zu", ws_bytes, need);

    const int nx = ceil_div(d_in, SLICE_K), nh = h / SLICE_K, nu = h / UNIT_N;
    uint8_t* packed = (uint8_t*)ws;
    float* bias4 = (float*)(packed + align_up(packed_bytes(d_in, h), 256));
    const int64_t chunk = chunk_rows_of(h, units_per_cta());
    float* hbuf[2] = {bias4 + 4 * MAX_H, bias4 + 4 * MAX_H + chunk * h};
    float* sumbuf = hbuf[1] + chunk * h;
    {
        ProfScope prof(PROF_PACK, st);
        int units = nu * (nx + nh) * 3 * UNIT_N * (SLICE_K / 8);
        if (units < 4 * MAX_H) units = 4 * MAX_H;
        pack_wide_kernel<CG><<<(units + 255) / 256, 256, 0, st>>>(w_ih, w_hh, b_ih, b_hh, d_in, h, nx, nh, packed, bias4);
        CTGCN_LAUNCH_OK("pack_wide_kernel");
    }
    int dev = 0, sm_count = 0;
    CTGCN_CUDA_OK(cudaGetDevice(&dev));
    CTGCN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    CTGCN_CUDA_OK(cudaFuncSetAttribute(gru_wide_step_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, WL<CG>::SMEM_BYTES));
    const bool each = mode == CTGCN_GRU_EACH_LN;
    ProfScope prof(PROF_GRU, st);
    for (int64_t row0 = 0; row0 < n; row0 += chunk) {
        const int64_t rows = n - row0 < chunk ? n - row0 : chunk;
        ParamsW p;
        p.n = rows;
        p.d_in = d_in;
        p.h = h;
        p.nx = nx;
        p.ns_packed = nx + nh;
        p.nu = nu;
        p.packed = packed;
        p.bias4 = bias4;
        const int tiles = (int)((rows + TILE_M - 1) / TILE_M);
        p.num_units = (tiles + CG - 1) / CG * nu;          // (group of CG tiles, feature block) pairs: one per cluster and turn
        p.trace = g_wide_trace;
        const int clusters = p.num_units < sm_count / CG ? p.num_units : sm_count / CG;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(clusters * CG));
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = WL<CG>::SMEM_BYTES;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        for (int i = 0; i < steps; ++i) {
            p.x = seq + row0 * srs + (int64_t)i * sss;
            p.ldx = srs;
            p.nh = i ? nh : 0;
            if (each) {
                float* yc = y + row0 * yrs;
                p.hprev = i ? yc + (int64_t)(i - 1) * yss : nullptr;
                p.ldh = yrs;
                p.hnew = yc + (int64_t)i * yss;
                p.ldn = yrs;
                p.sum = nullptr;
                p.lds = 0;
                p.sum_add = 0;
            } else {
                p.hprev = i ? hbuf[(i - 1) & 1] : nullptr;
                p.ldh = h;
                p.hnew = i + 1 < steps ? hbuf[i & 1] : nullptr;
                p.ldn = h;
                p.sum = sumbuf;
                p.lds = h;
                p.sum_add = i ? 1 : 0;
            }
            CTGCN_CUDA_OK(cudaLaunchKernelEx(&cfg, gru_wide_step_kernel<CG>, p));
            CTGCN_LAUNCH_OK("gru_wide_step_kernel");
        }
        const int64_t ln_rows = each ? rows * steps : rows;
        const int blocks = (int)((ln_rows + 7) / 8);
        if (each) {
            float* yc = y + row0 * yrs;
            ln_rows_wide_kernel<<<blocks, 256, 0, st>>>(yc, yrs, yss, steps, ln_rows, h, ln_w, ln_b, eps, yc, yrs, yss, RowScatter());
        } else {
            RowScatter s2;
            if (sc && sc->slices) {
                s2 = *sc;
                s2.row_off += row0;
            }
            ln_rows_wide_kernel<<<blocks, 256, 0, st>>>(sumbuf, h, 0, 1, ln_rows, h, ln_w, ln_b, eps, y ? y + row0 * yrs : nullptr, yrs, 0, s2);
        }
        CTGCN_LAUNCH_OK("ln_rows_wide_kernel");
    }
    return CTGCN_OK;
}

}  // namespace

void set_gru_wide_trace(long long* buf) { g_wide_trace = buf; }

bool gru_wide_tc_takes(int d_in, int h) { return wide_takes(d_in, h); }
// packed weights | bias4 | h ping-pong (2 × chunk × h) | Σh (chunk × h)
size_t gru_wide_tc_workspace_bytes(int d_in, int h) { return wide_workspace(d_in, h); }

// cg = CTAs per MMA (2: CTA pairs, the default; 1: the build without pairing, A/B and tests).
// returns 0 = done, <0 = error, 1 = shape / alignment not supported by this path
int launch_gru_wide_tc(int cg, const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* w_ih,
                       const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                       int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (cg == 2) return launch_wide<2>(seq, srs, sss, n, steps, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, ws, ws_bytes, st);
    return launch_wide<1>(seq, srs, sss, n, steps, d_in, h, w_ih, w_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, ws, ws_bytes, st);
}

}  // namespace ctgcn
