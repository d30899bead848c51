// tcgen05 (5th-gen tensor core) GRU over a short sequence + Σ/LayerNorm epilogue — the compute-bound half of
// CoreDiffusion.forward (layers.py:59-62) and the temporal GRU of CTGCN.forward (models.py:249-250).
//
// Numerics: the parity bar is 1e-4 relative in fp32 and single-pass bf16/tf32 tensor-core products miss it
// (2.7e-3 / 3.3e-4).  Every fp32 operand is therefore split a = hi + lo (two bf16 planes, 16
// significant bits) and each product is issued as three MMAs  hi·hi + lo·hi + hi·lo  accumulated in fp32 in
// TMEM (error ≈ 2^-17 per product, ~1e-5 on the layer output).
//
// One persistent CTA (512 threads) per SM owns 128-node tiles for the WHOLE sequence (h and Σh never leave the SM):
//   warp 0      weight producer: packed 24 KB chunks (192 gate rows × 32 k, bf16 hi|lo planes, exact shared-memory
//               images built once per call by pack_weights_kernel) stream from L2 through a 3-stage ring with
//               cp.async.bulk + mbarrier complete_tx
//   warp 1      MMA issuer (one elected lane): tcgen05.mma cta_group::1 kind::f16, M=128 N=192 K=16, operands described
//               by no-swizzle K-major shared-memory descriptors, accumulators in TMEM.  N = 192 keeps the operand
//               fetch (A 4 KB + B 6 KB per 96-cycle MMA) under the 128 B/cycle shared-memory bandwidth of the SM.
//   warps 4-7   input loaders: fp32 rows → bf16 hi/lo planes in UMMA core-matrix order, one step ahead of the MMAs
//   warps 8-15  gate math: tcgen05.ld the accumulators, ex2/rcp sigmoid & tanh, write h back as the next step's A operand,
//               keep Σh (or emit LayerNorm(h_s)), final LayerNorm.  setmaxnreg moves registers to them.
// A step is processed in two halves of 64 hidden features.  Each half owns one accumulator set of 256 TMEM columns
//   [ W_in·x | r | z | W_hn·h ]   (64 columns each)
// so the input part (A = U) is ONE N=192 MMA stream into columns [0,192) and the recurrent part (A = h) ONE N=192
// stream into columns [64,256); the gate math of one half overlaps the MMAs of the other.
//
// Shapes: H = 128, d_in ∈ {64, 128} (everything the 128-d configurations need); other shapes use gru_simt.cu.
#include "common.cuh"
#include "tc_common.cuh"

namespace ctgcn {
namespace {

constexpr int H = 128;
constexpr int TILE_M = 128;
constexpr int CHUNK_ROWS = 192, CHUNK_K = 32;
constexpr int CHUNK_PLANE = CHUNK_ROWS * CHUNK_K * 2;  // 12 KB: one bf16 plane of a weight chunk
constexpr int CHUNK_BYTES = 2 * CHUNK_PLANE;           // hi + lo = 24 KB
constexpr int STAGES = 3;
constexpr int NUM_WORKER_WARPS = 8, NUM_LOADER_WARPS = 4;
constexpr int FIRST_LOADER_WARP = 4, FIRST_WORKER_WARP = 8;   // warp 0 producer, warp 1 MMA, warps 2-3 idle
constexpr int THREADS = 32 * (FIRST_WORKER_WARP + NUM_WORKER_WARPS);
constexpr uint32_t TMEM_COLS = 512;
// accumulator set layout (columns inside a 256-column set)
constexpr uint32_t COL_IN = 0, COL_R = 64, COL_Z = 128, COL_HN = 192;

// ---- shared memory map (bytes)
constexpr int A_PLANE = TILE_M * H * 2;               // 32 KB: one bf16 plane of a 128×128 operand tile
constexpr int SM_U = 0;                               // U hi | U lo
constexpr int SM_H = SM_U + 2 * A_PLANE;              // h hi | h lo
constexpr int SM_W = SM_H + 2 * A_PLANE;              // weight ring
constexpr int SM_BIAS = SM_W + STAGES * CHUNK_BYTES;  // [4][H] fp32 (pre-scaled): b_r(+), b_z(+), b_in, b_hn
constexpr int SM_LN = SM_BIAS + 4 * H * 4;            // ln_w | ln_b
constexpr int SM_RED = SM_LN + 2 * H * 4;             // [2 buffers][2 column halves][128 rows] fp32
constexpr int SM_BAR = SM_RED + 2 * 2 * TILE_M * 4;
constexpr int NUM_BARS = 2 * STAGES + 7;
constexpr int SM_TMEM_PTR = SM_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = SM_TMEM_PTR + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum Bar { BAR_W_FULL = 0, BAR_W_EMPTY = STAGES, BAR_U_READY = 2 * STAGES, BAR_U_FREE, BAR_H_READY, BAR_ACC_FULL0,
           BAR_ACC_FULL1, BAR_ACC_FREE0, BAR_ACC_FREE1 };

using namespace tc;

// ------------------------------------------------------------------------------------------------ weight packing
// Packed order: part ∈ {X half0, X half1, H half0, H half1}, then k-chunk.  An X chunk holds the rows
// [n-gate | r | z] of W_ih for the half's 64 hidden features, an H chunk the rows [r | z | n-gate] of W_hh
// (192 rows × 32 k).  Image: bf16 hi plane (12 KB) then lo plane; element (row, k) at (k/8)·3072 + row·16 + (k%8)·2
// (no-swizzle K-major core matrices: LBO = 3072, SBO = 128).  With `prescale`, the r/z rows and biases are multiplied
// by −log2(e) and the n rows by 2·log2(e) so that sigmoid/tanh need a bare ex2 (gate math below).
__host__ __device__ constexpr int chunks_per_part(int k) { return k / CHUNK_K; }
constexpr int UNITS_PER_PLANE = CHUNK_ROWS * (CHUNK_K / 8);  // 16-byte units

__global__ void pack_weights_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                                    const float* __restrict__ b_ih, const float* __restrict__ b_hh, int d_in,
                                    uint8_t* __restrict__ packed, float* __restrict__ bias4, int prescale) {
    constexpr float kLog2e = 1.4426950408889634f;
    const int cx = chunks_per_part(d_in), chh = chunks_per_part(H);
    const int nchunks = 2 * cx + 2 * chh;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < 4 * H) {
        const int g = t / H, f = t % H;
        float v = 0.f;
        if (b_ih) {
            if (g == 0) v = b_ih[f] + b_hh[f];
            if (g == 1) v = b_ih[H + f] + b_hh[H + f];
            if (g == 2) v = b_ih[2 * H + f];
            if (g == 3) v = b_hh[2 * H + f];
        }
        bias4[t] = prescale ? v * (g < 2 ? -kLog2e : 2.f * kLog2e) : v;
    }
    if (t >= nchunks * UNITS_PER_PLANE) return;
    const int c = t / UNITS_PER_PLANE, unit = t % UNITS_PER_PLANE;
    const bool is_x = c < 2 * cx;
    const int cc = is_x ? c : c - 2 * cx;
    const int per = is_x ? cx : chh;
    const int ktot = is_x ? d_in : H;
    const int half = cc / per, kc = cc % per;
    const int kb = unit / CHUNK_ROWS, row = unit % CHUNK_ROWS;
    const int blk = row / 64, f = row % 64;
    const int gate = is_x ? (blk == 0 ? 2 : blk - 1) : blk;    // X: [n, r, z]   H: [r, z, n]
    const float* w = is_x ? w_ih : w_hh;
    const float* src = w + (int64_t)(gate * H + half * 64 + f) * ktot + kc * CHUNK_K + kb * 8;
    const float scale = prescale ? (gate < 2 ? -kLog2e : 2.f * kLog2e) : 1.f;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = src[i] * scale;
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* dst = packed + (size_t)c * CHUNK_BYTES + kb * (CHUNK_ROWS * 16) + row * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + CHUNK_PLANE) = lo;
}

// ------------------------------------------------------------------------------------------------ kernel pieces
struct Params {
    const float* seq;
    int64_t srs, sss, n;
    int steps, d_in;
    const uint8_t* packed;
    const float* bias4;
    const float* ln_w;
    const float* ln_b;
    float eps;
    float* y;
    int64_t yrs, yss;
    RowScatter sc;      // SUM_LN only: rows go to their node slice's buffer (fused snapshot exchange)
    int num_tiles;
    long long* trace;   // optional [24 events][64 steps] clock64 stamps of block 0 (ctgcn_debug_gru_trace), else NULL
};

// debug timeline: event e of global step gs of block 0
#define GRU_TRACE(e, gs)                                                                    \
    do {                                                                                    \
        if (p.trace && blockIdx.x == 0 && (gs) < 64u) p.trace[(e) * 64 + (gs)] = clock64(); \
    } while (0)

#ifdef GRU_EXP_NO_MUFU   // timing experiment: no special-function unit work at all (results are wrong)
__device__ __forceinline__ float ex2_approx(float v) { return v * 0.5f + 1.f; }
__device__ __forceinline__ float rcp_approx(float v) { return v * 0.25f + 0.5f; }
#else
__device__ __forceinline__ float ex2_approx(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float rcp_approx(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
#endif

// One 24 KB weight chunk (192 rows × 32 k): 2 K-steps × 3 split products (hi·hi, lo·hi, hi·lo) of N = 192 into 192
// consecutive accumulator columns starting at d_tmem.  a_lo32 / b_lo32: descriptor low words of the A hi plane at this
// chunk's first k and of the chunk's hi plane.
//   fresh       : the very first MMA overwrites all 192 columns (input part, first chunk)
//   split_first : recurrent part, first chunk — columns [0,128) (r|z) accumulate on the input part's result while
//                 columns [128,192) (W_hn·h) start fresh: that one MMA is issued as an N=128 and an N=64 instruction.
__device__ __forceinline__ void issue_chunk(uint32_t a_lo32, uint32_t b_lo32, uint32_t d_tmem, bool fresh, bool split_first) {
    constexpr uint32_t idesc192 = umma_idesc_bf16(TILE_M, 192), idesc128 = umma_idesc_bf16(TILE_M, 128),
                       idesc64 = umma_idesc_bf16(TILE_M, 64);
    constexpr uint32_t A_STEP = (2 * TILE_M * 16) >> 4, B_STEP = (2 * CHUNK_ROWS * 16) >> 4;
    constexpr uint32_t A_LO_PLANE = A_PLANE >> 4, B_LO_PLANE = CHUNK_PLANE >> 4;
#pragma unroll
    for (int ks = 0; ks < CHUNK_K / 16; ++ks) {
        const uint64_t ah = desc64(a_lo32 + ks * A_STEP), al = desc64(a_lo32 + A_LO_PLANE + ks * A_STEP);
        const uint64_t bh = desc64(b_lo32 + ks * B_STEP), bl = desc64(b_lo32 + B_LO_PLANE + ks * B_STEP);
        if (split_first && ks == 0) {
            umma_bf16(d_tmem, ah, bh, idesc128, 1u);
            umma_bf16(d_tmem + 128, ah, desc64(b_lo32 + ((128 * 16) >> 4)), idesc64, 0u);
        } else {
            umma_bf16(d_tmem, ah, bh, idesc192, (fresh && ks == 0) ? 0u : 1u);
        }
        umma_bf16(d_tmem, al, bh, idesc192, 1u);
        umma_bf16(d_tmem, ah, bl, idesc192, 1u);
    }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) gru_tc_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sbase + SM_BAR;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    const int cpx = chunks_per_part(p.d_in);   // weight chunks of one half of the input part
    constexpr int cph = chunks_per_part(H);    // … of the recurrent part
    const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(BAR_W_FULL + s), 1);
            mbar_init(bar(BAR_W_EMPTY + s), 1);
        }
        mbar_init(bar(BAR_U_READY), NUM_LOADER_WARPS);
        mbar_init(bar(BAR_U_FREE), 1);
        mbar_init(bar(BAR_H_READY), NUM_WORKER_WARPS);
        mbar_init(bar(BAR_ACC_FULL0), 1);
        mbar_init(bar(BAR_ACC_FULL1), 1);
        mbar_init(bar(BAR_ACC_FREE0), NUM_WORKER_WARPS);
        mbar_init(bar(BAR_ACC_FREE1), NUM_WORKER_WARPS);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 4 * H; i += THREADS) reinterpret_cast<float*>(smem + SM_BIAS)[i] = p.bias4[i];
    for (int i = threadIdx.x; i < H; i += THREADS) {
        reinterpret_cast<float*>(smem + SM_LN)[i] = p.ln_w[i];
        reinterpret_cast<float*>(smem + SM_LN)[H + i] = p.ln_b[i];
    }
    if (warp == 1) tmem_alloc(sbase + SM_TMEM_PTR, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_PTR);

    // 512 threads start with 128 registers each; the gate-math warps need more, the others far fewer: every role
    // branch starts with its warpgroup's setmaxnreg (warps 0-3: 56, loaders 4-7: 112, workers 8-15: 168)
    if (warp == 0) {
        // ===================================================== weight producer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const int nx = 2 * cpx;
            for (int t = 0; t < my_tiles; ++t) {
                for (int i = 0; i < p.steps; ++i) {
                    // consumption order of the MMA issuer: X half0, [H half0], X half1, [H half1]
                    // (packed order is X half0, X half1, H half0, H half1)
                    for (int seg = 0; seg < 4; ++seg) {
                        const bool rec = seg & 1;
                        if (rec && i == 0) continue;
                        const int half = seg >> 1;
                        const int first = rec ? nx + half * cph : half * cpx;
                        const int count = rec ? cph : cpx;
                        for (int c = first; c < first + count; ++c) {
                            mbar_wait(bar(BAR_W_EMPTY + stage), phase ^ 1);
                            mbar_expect_tx(bar(BAR_W_FULL + stage), CHUNK_BYTES);
                            bulk_g2s(sbase + SM_W + stage * CHUNK_BYTES, p.packed + (size_t)c * CHUNK_BYTES, CHUNK_BYTES,
                                     bar(BAR_W_FULL + stage));
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // All 32 lanes run the (warp-uniform) control flow and the barrier waits; one elected lane issues.
        {
            uint32_t stage = 0, phase = 0, gs = 0;
            const uint32_t u_desc = desc_lo(sbase + SM_U, TILE_M * 16), h_desc = desc_lo(sbase + SM_H, TILE_M * 16);
            // one part = one half (64 hidden features) of the input (A = U) or recurrent (A = h) contribution
            auto run_part = [&](uint32_t a_desc, int ktot, int half, bool recurrent) {
                const uint32_t d = tmem + half * 256 + (recurrent ? COL_R : COL_IN);
                for (int kc = 0; kc < ktot / CHUNK_K; ++kc) {
                    mbar_wait(bar(BAR_W_FULL + stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        issue_chunk(a_desc + kc * (CHUNK_K / 8) * ((TILE_M * 16) >> 4),
                                    desc_lo(sbase + SM_W + stage * CHUNK_BYTES, CHUNK_ROWS * 16), d, !recurrent && kc == 0,
                                    recurrent && kc == 0);
                        umma_commit(bar(BAR_W_EMPTY + stage));
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            };
            auto commit = [&](int b) {
                if (elect_one()) umma_commit(bar(b));
                __syncwarp();
            };
            for (int t = 0; t < my_tiles; ++t) {
                for (int i = 0; i < p.steps; ++i, ++gs) {
                    const uint32_t par = gs & 1;
                    if (lane == 0) GRU_TRACE(0, gs);
                    mbar_wait(bar(BAR_U_READY), par);
                    mbar_wait(bar(BAR_ACC_FREE0), par ^ 1);
                    tc_fence_after();
                    if (lane == 0) GRU_TRACE(1, gs);
                    run_part(u_desc, p.d_in, 0, false);
                    if (lane == 0) GRU_TRACE(2, gs);
                    if (i == 0) {
                        commit(BAR_ACC_FULL0);
                    } else {
                        // the recurrence h_{i-1} → gates → h_i is the critical chain: the first half's recurrent part
                        // goes ahead of the second half's input part (same chunk order in the producer)
                        mbar_wait(bar(BAR_H_READY), par ^ 1);
                        tc_fence_after();
                        if (lane == 0) GRU_TRACE(3, gs);
                        run_part(h_desc, H, 0, true);
                        commit(BAR_ACC_FULL0);
                        if (lane == 0) GRU_TRACE(4, gs);
                    }
                    mbar_wait(bar(BAR_ACC_FREE1), par ^ 1);
                    tc_fence_after();
                    if (lane == 0) GRU_TRACE(5, gs);
                    run_part(u_desc, p.d_in, 1, false);
                    commit(BAR_U_FREE);
                    if (lane == 0) GRU_TRACE(6, gs);
                    if (i > 0) run_part(h_desc, H, 1, true);
                    commit(BAR_ACC_FULL1);
                    if (lane == 0) GRU_TRACE(7, gs);
                }
            }
        }
    } else if (warp < FIRST_WORKER_WARP) {
        // ===================================================== input loaders (warps 4-7; warps 2-3 idle)
        if (warp < FIRST_LOADER_WARP) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        } else {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
            // A warp owns 32 tile rows.  Per load instruction its lanes cover 8 rows × 4 k-blocks (r = lane%8, c = lane/8):
            // 128 contiguous bytes per row (8 L1 wavefronts instead of 32 for a row-per-lane mapping) and the 16-byte
            // shared-memory stores of one 8-lane phase hit 8 consecutive rows of one k-block (conflict-free).
            const int r8 = lane & 7, c4 = lane >> 3;
            const int row_base = 32 * (warp - FIRST_LOADER_WARP);
            uint8_t* u_hi = smem + SM_U;
            const int nkg = p.d_in / 32;             // k-groups of 4 k-blocks
            const int nit = 4 * nkg;                 // (row-group, k-group) iterations per step: 16 for d_in = 128
            uint32_t gs = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const int64_t tile_row0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M;
                for (int i = 0; i < p.steps; ++i, ++gs) {
                    const float* base = p.seq + (int64_t)i * p.sss;
                    float4 v[16];
                    auto load_batch = [&](int it0) {   // 8 iterations = 16 LDG.128 in flight per lane
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int it = it0 + u, rg = it & 3, kg = it >> 2;
                            const int64_t srow = tile_row0 + row_base + 8 * rg + r8;
                            if (srow < p.n) {
                                const float* src = base + srow * p.srs + (4 * kg + c4) * 8;
                                v[2 * u] = __ldg(reinterpret_cast<const float4*>(src));
                                v[2 * u + 1] = __ldg(reinterpret_cast<const float4*>(src + 4));
                            } else {
                                v[2 * u] = v[2 * u + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                    };
                    auto store_batch = [&](int it0) {  // fp32 → bf16 hi/lo planes, 8 k-elements (16 B) per store
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int it = it0 + u, rg = it & 3, kg = it >> 2;
                            const int m = row_base + 8 * rg + r8, kb = 4 * kg + c4;
                            const float f8[8] = {v[2 * u].x, v[2 * u].y, v[2 * u].z, v[2 * u].w,
                                                 v[2 * u + 1].x, v[2 * u + 1].y, v[2 * u + 1].z, v[2 * u + 1].w};
                            uint4 hi, lo;
                            split8(f8, hi, lo);
#ifdef GRU_EXP_NO_U_STORE
                            if (f8[0] != 12345.678f) continue;   // timing experiment
#endif
                            *reinterpret_cast<uint4*>(u_hi + kb * (TILE_M * 16) + m * 16) = hi;
                            *reinterpret_cast<uint4*>(u_hi + A_PLANE + kb * (TILE_M * 16) + m * 16) = lo;
                        }
                    };
                    load_batch(0);   // global loads are issued BEFORE the buffer is free: their latency is off the loop
                    // the input part of the previous step's MMAs must have released the single U buffer
                    mbar_wait(bar(BAR_U_FREE), (gs & 1) ^ 1);
                    if (threadIdx.x == FIRST_LOADER_WARP * 32) GRU_TRACE(13, gs);
                    store_batch(0);
                    for (int it0 = 8; it0 < nit; it0 += 8) {
                        load_batch(it0);
                        store_batch(it0);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(BAR_U_READY));
                    if (threadIdx.x == FIRST_LOADER_WARP * 32) GRU_TRACE(14, gs);
                }
            }
        }
    } else {
        // ===================================================== workers (warps 8-15): gate math, h, Σh, LayerNorm
        asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
        const int ww = warp - FIRST_WORKER_WARP;
        const int q = warp & 3;          // TMEM lane quarter this warp may access
        const int ch = ww >> 2;          // which 32 of a half's 64 features this thread owns
        const int m = 32 * q + lane;     // row inside the tile
        const uint32_t tmem_lane = tmem + ((uint32_t)(32 * q) << 16);
        const float* bias = reinterpret_cast<const float*>(smem + SM_BIAS);
        const float* lnw = reinterpret_cast<const float*>(smem + SM_LN);
        float* red = reinterpret_cast<float*>(smem + SM_RED);
        uint8_t* h_hi = smem + SM_H;
        uint32_t gs = 0;

        // LayerNorm over the row: this thread holds 64 of its 128 values, the partner warp (other ch) the rest
        auto layer_norm_store = [&](const float (&v)[64], float* dst_row, bool valid) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) s += v[j];
            red[ch * TILE_M + m] = s;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float mean = (s + red[(ch ^ 1) * TILE_M + m]) * (1.f / H);
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float dlt = v[j] - mean;
                sq = fmaf(dlt, dlt, sq);
            }
            red[2 * TILE_M + ch * TILE_M + m] = sq;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float rstd = rsqrtf((sq + red[2 * TILE_M + (ch ^ 1) * TILE_M + m]) * (1.f / H) + p.eps);
            if (valid) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                    for (int j4 = 0; j4 < 32; j4 += 4) {
                        const int f = hf * 64 + ch * 32 + j4;
                        float4 o;
                        o.x = (v[hf * 32 + j4 + 0] - mean) * rstd * lnw[f + 0] + lnw[H + f + 0];
                        o.y = (v[hf * 32 + j4 + 1] - mean) * rstd * lnw[f + 1] + lnw[H + f + 1];
                        o.z = (v[hf * 32 + j4 + 2] - mean) * rstd * lnw[f + 2] + lnw[H + f + 2];
                        o.w = (v[hf * 32 + j4 + 3] - mean) * rstd * lnw[f + 3] + lnw[H + f + 3];
                        *reinterpret_cast<float4*>(dst_row + f) = o;
                    }
                }
            }
        };

        // SUM_LN result of a tile: normalised rows are staged in the (now idle) h buffer with a 16-byte XOR swizzle and
        // written out one whole 512-byte row per warp instruction — to y, or straight into the owning node slice's
        // (peer) buffer: NVLink wants full-line stores, not 32 scattered 16-byte pieces per instruction.
        auto layer_norm_store_rows = [&](const float (&v)[64], int64_t tile_row0) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) s += v[j];
            red[ch * TILE_M + m] = s;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float mean = (s + red[(ch ^ 1) * TILE_M + m]) * (1.f / H);
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
                const float dlt = v[j] - mean;
                sq = fmaf(dlt, dlt, sq);
            }
            red[2 * TILE_M + ch * TILE_M + m] = sq;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float rstd = rsqrtf((sq + red[2 * TILE_M + (ch ^ 1) * TILE_M + m]) * (1.f / H) + p.eps);
            float* stage = reinterpret_cast<float*>(smem + SM_H);     // [128 rows][32 chunks of 4 floats], chunk c of row r at c ^ (r & 31)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int j4 = 0; j4 < 32; j4 += 4) {
                    const int f = hf * 64 + ch * 32 + j4;
                    float4 o;
                    o.x = (v[hf * 32 + j4 + 0] - mean) * rstd * lnw[f + 0] + lnw[H + f + 0];
                    o.y = (v[hf * 32 + j4 + 1] - mean) * rstd * lnw[f + 1] + lnw[H + f + 1];
                    o.z = (v[hf * 32 + j4 + 2] - mean) * rstd * lnw[f + 2] + lnw[H + f + 2];
                    o.w = (v[hf * 32 + j4 + 3] - mean) * rstd * lnw[f + 3] + lnw[H + f + 3];
                    *reinterpret_cast<float4*>(stage + m * H + (((f >> 2) ^ (m & 31)) << 2)) = o;
                }
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            for (int rr = 0; rr < TILE_M / NUM_WORKER_WARPS; ++rr) {
                const int r = ww * (TILE_M / NUM_WORKER_WARPS) + rr;
                const int64_t grow = tile_row0 + r;
                if (grow >= p.n) break;   // warp-uniform
                const float4 o = *reinterpret_cast<const float4*>(stage + r * H + ((lane ^ (r & 31)) << 2));
                float* dst = p.sc.slices ? p.sc.row_ptr(grow) : p.y + grow * p.yrs;
                *reinterpret_cast<float4*>(dst + 4 * lane) = o;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the staging area is h again from the next tile's first step on
        };

        for (int t = 0; t < my_tiles; ++t) {
            const int64_t row = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + m;
            const bool valid = row < p.n;
            float acc_out[64];   // Σ_s h_s (SUM_LN) / h_s of the current step (EACH_LN) for this thread's 64 features
#pragma unroll
            for (int j = 0; j < 64; ++j) acc_out[j] = 0.f;

            for (int i = 0; i < p.steps; ++i, ++gs) {
                const uint32_t par = gs & 1;
                // gates, one half (64 hidden features) at a time; this thread owns 32 of them, 8 per pass
                float h0[32];   // first half of h_i, published only when no MMA reads h_{i-1} any more
                auto put_h8 = [&](const float (&f8)[8], int f) {
#ifdef GRU_EXP_NO_H_STORE
                    if (f8[0] != 12345.678f) return;   // timing experiment: never true in practice
#endif
                    uint4 hi, lo;
                    split8(f8, hi, lo);
                    const int kb = f >> 3;
                    *reinterpret_cast<uint4*>(h_hi + kb * (TILE_M * 16) + m * 16) = hi;
                    *reinterpret_cast<uint4*>(h_hi + A_PLANE + kb * (TILE_M * 16) + m * 16) = lo;
                };
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    if (threadIdx.x == FIRST_WORKER_WARP * 32) GRU_TRACE(8 + 2 * hf, gs);
                    mbar_wait(bar(BAR_ACC_FULL0 + hf), par);
                    tc_fence_after();
                    if (threadIdx.x == FIRST_WORKER_WARP * 32) GRU_TRACE(9 + 2 * hf, gs);
#pragma unroll
                    for (int sub = 0; sub < 4; ++sub) {
                        const int f0 = hf * 64 + ch * 32 + sub * 8;             // first of 8 features
                        const uint32_t col = hf * 256 + ch * 32 + sub * 8;      // + gate block
                        float gr[8], gz[8], gi[8], gh[8], hold[8];
                        tmem_ld8(tmem_lane + col + COL_R, gr);
                        tmem_ld8(tmem_lane + col + COL_Z, gz);
                        tmem_ld8(tmem_lane + col + COL_IN, gi);
                        if (i > 0) {
                            tmem_ld8(tmem_lane + col + COL_HN, gh);
                            const int kb = f0 >> 3;
                            const uint4 hi = *reinterpret_cast<const uint4*>(h_hi + kb * (TILE_M * 16) + m * 16);
                            const uint4 lo = *reinterpret_cast<const uint4*>(h_hi + A_PLANE + kb * (TILE_M * 16) + m * 16);
                            join8(hi, lo, hold);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) gh[j] = hold[j] = 0.f;
                        }
                        tmem_ld_wait();
                        // Gate math written stage by stage over the 8 features so that the 8 dependent chains
                        // (ex2 → rcp → ex2 → rcp) are interleaved.  Pre-activations are pre-scaled (pack_weights_kernel):
                        // sigmoid(a) = 1/(1 + 2^a'), tanh(s) = 1 − 2/(1 + 2^s'); one reciprocal serves r and z.
                        float hn8[8], ea[8], eb[8], zz[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            ea[j] = gr[j] + bias[f0 + j];
                            eb[j] = gz[j] + bias[H + f0 + j];
                            gi[j] += bias[2 * H + f0 + j];
                            gh[j] += bias[3 * H + f0 + j];
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            ea[j] = ex2_approx(ea[j]);
                            eb[j] = ex2_approx(eb[j]);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            ea[j] = 1.f + fminf(ea[j], 1e18f);     // clamped so that the product below stays finite
                            eb[j] = 1.f + fminf(eb[j], 1e18f);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) hn8[j] = rcp_approx(ea[j] * eb[j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            zz[j] = hn8[j] * ea[j];                                          // z
                            gi[j] = fmaf(hn8[j] * eb[j], gh[j], gi[j]);                      // W_in x + b_in + r ⊙ (W_hn h + b_hn)
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) gi[j] = ex2_approx(gi[j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) gi[j] = rcp_approx(1.f + gi[j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float nn = fmaf(-2.f, gi[j], 1.f);                         // tanh
                            hn8[j] = fmaf(zz[j], hold[j] - nn, nn);                          // (1 − z) n + z h
                        }
                        if (hf == 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) h0[sub * 8 + j] = hn8[j];
                        } else {
                            // every MMA that reads h_{i-1} has completed (acc1 full): h may be overwritten in place.
                            // The held-back first half is published piecewise here so that its ALU work hides
                            // under the MUFU latency of this pass.
                            const float f8[8] = {h0[sub * 8], h0[sub * 8 + 1], h0[sub * 8 + 2], h0[sub * 8 + 3],
                                                 h0[sub * 8 + 4], h0[sub * 8 + 5], h0[sub * 8 + 6], h0[sub * 8 + 7]};
                            put_h8(f8, ch * 32 + sub * 8);
                            put_h8(hn8, f0);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (MODE == CTGCN_GRU_SUM_LN) acc_out[hf * 32 + sub * 8 + j] += hn8[j];
                            else acc_out[hf * 32 + sub * 8 + j] = hn8[j];
                        }
                        if (threadIdx.x == FIRST_WORKER_WARP * 32) GRU_TRACE(16 + hf * 4 + sub, gs);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(BAR_ACC_FREE0 + hf));
                }
                // h_i is complete in shared memory: the next step's recurrent MMAs may read it
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BAR_H_READY));
                if (threadIdx.x == FIRST_WORKER_WARP * 32) GRU_TRACE(12, gs);
                if (MODE == CTGCN_GRU_EACH_LN) layer_norm_store(acc_out, p.y + row * p.yrs + (int64_t)i * p.yss, valid);
            }
            if (MODE == CTGCN_GRU_SUM_LN) layer_norm_store_rows(acc_out, row - m);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ self test
// One half-step of a GRU cell's pre-activations for d_in = 64 through exactly the packer, chunk images, bulk copies,
// descriptors, split-bf16 MMAs (incl. the split-first recurrent MMA) and TMEM loads of the GRU kernel:
//   out[128×256] = [ x·W_inᵀ | x·W_irᵀ + h·W_hrᵀ | x·W_izᵀ + h·W_hzᵀ | h·W_hnᵀ ]   for hidden features 0..63
// with x [128,64], h [128,128], w_ih [384,64], w_hh [384,128], no pre-scaling.  Exposed as ctgcn_selftest_umma.
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ x, const float* __restrict__ h,
                                                                const uint8_t* __restrict__ packed, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t su = sbase, sh = sbase + 2 * A_PLANE, sw = sbase + 4 * A_PLANE;
    const uint32_t bar_w = sw + CHUNK_BYTES, bar_d = bar_w + 8, tptr = bar_d + 8;
    const int warp = threadIdx.x >> 5, m = threadIdx.x;
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_d, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tptr, 256);
    for (int kb = 0; kb < 16; ++kb) {
        float f8[8];
        uint4 hi, lo;
        if (kb < 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f8[e] = x[m * 64 + kb * 8 + e];
            split8(f8, hi, lo);
            *reinterpret_cast<uint4*>(smem + kb * (TILE_M * 16) + m * 16) = hi;
            *reinterpret_cast<uint4*>(smem + A_PLANE + kb * (TILE_M * 16) + m * 16) = lo;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) f8[e] = h[m * 128 + kb * 8 + e];
        split8(f8, hi, lo);
        *reinterpret_cast<uint4*>(smem + 2 * A_PLANE + kb * (TILE_M * 16) + m * 16) = hi;
        *reinterpret_cast<uint4*>(smem + 3 * A_PLANE + kb * (TILE_M * 16) + m * 16) = lo;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 4 * A_PLANE + CHUNK_BYTES + 16);
    if (warp == 0) {
        // packed order for d_in = 64: X half0 = chunks 0,1; X half1 = 2,3; H half0 = 4..7
        const int order[6] = {0, 1, 4, 5, 6, 7};
        uint32_t par = 0;
        for (int j = 0; j < 6; ++j) {
            if (elect_one()) {
                mbar_expect_tx(bar_w, CHUNK_BYTES);
                bulk_g2s(sw, packed + (size_t)order[j] * CHUNK_BYTES, CHUNK_BYTES, bar_w);
            }
            __syncwarp();
            mbar_wait(bar_w, par);
            tc_fence_after();
            if (elect_one()) {
                const bool rec = j >= 2;
                const int kc = rec ? j - 2 : j;
                issue_chunk(desc_lo(rec ? sh : su, TILE_M * 16) + kc * (CHUNK_K / 8) * ((TILE_M * 16) >> 4),
                            desc_lo(sw, CHUNK_ROWS * 16), tmem + (rec ? COL_R : COL_IN), !rec && kc == 0, rec && kc == 0);
                umma_commit(bar_d);
            }
            __syncwarp();
            mbar_wait(bar_d, par);     // the single weight buffer is reused: wait for the MMAs that read it
            tc_fence_after();
            par ^= 1;
        }
    }
    __syncthreads();
    tc_fence_after();
    const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
    for (int c = 0; c < 256; c += 8) {
        float v[8];
        tmem_ld8(tl + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) out[m * 256 + c + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

static long long* g_gru_trace = nullptr;
void set_gru_trace(long long* buf) { g_gru_trace = buf; }
// returns 0 = done, <0 = error, 1 = shape not supported by this path
int launch_gru_tc(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* w_ih,
                  const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                  int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (h != H || (d_in != 64 && d_in != 128)) return 1;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(seq) || (srs & 3) || (sss & 3)) return 1;
    if (sc ? ((sc->row_stride & 3) || (sc->col_offset & 3)) : (!al16(y) || (yrs & 3) || (yss & 3))) return 1;
    const int nchunks = 2 * chunks_per_part(d_in) + 2 * chunks_per_part(H);
    const size_t packed_bytes = (size_t)nchunks * CHUNK_BYTES;
    CTGCN_REQUIRE(ws && ws_bytes >= packed_bytes + 4 * H * sizeof(float), "gru_tc: workspace too small");
    uint8_t* packed = (uint8_t*)ws;
    float* bias4 = (float*)(packed + packed_bytes);
    {
        ProfScope prof(PROF_PACK, st);
        const int threads = nchunks * UNITS_PER_PLANE;
        pack_weights_kernel<<<(threads + 255) / 256, 256, 0, st>>>(w_ih, w_hh, b_ih, b_hh, d_in, packed, bias4, 1);
        CTGCN_LAUNCH_OK("pack_weights_kernel");
    }
    // per device (the attribute belongs to the device's context; a process may drive several GPUs) and cheap: set on every call
    int dev = 0, sm_count = 0;
    CTGCN_CUDA_OK(cudaGetDevice(&dev));
    CTGCN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    CTGCN_CUDA_OK(cudaFuncSetAttribute(gru_tc_kernel<CTGCN_GRU_SUM_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    CTGCN_CUDA_OK(cudaFuncSetAttribute(gru_tc_kernel<CTGCN_GRU_EACH_LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    Params p;
    p.seq = seq;
    p.srs = srs;
    p.sss = sss;
    p.n = n;
    p.steps = steps;
    p.d_in = d_in;
    p.packed = packed;
    p.bias4 = bias4;
    p.ln_w = ln_w;
    p.ln_b = ln_b;
    p.eps = eps;
    p.y = y;
    p.yrs = yrs;
    p.yss = yss;
    p.sc = sc ? *sc : RowScatter();
    p.num_tiles = (int)((n + TILE_M - 1) / TILE_M);
    p.trace = g_gru_trace;
    const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
    ProfScope prof(PROF_GRU, st);
    if (mode == CTGCN_GRU_SUM_LN)
        gru_tc_kernel<CTGCN_GRU_SUM_LN><<<grid, THREADS, SMEM_BYTES, st>>>(p);
    else
        gru_tc_kernel<CTGCN_GRU_EACH_LN><<<grid, THREADS, SMEM_BYTES, st>>>(p);
    CTGCN_LAUNCH_OK("gru_tc_kernel");
    return CTGCN_OK;
}

}  // namespace ctgcn

using namespace ctgcn;

// Test hook (see umma_selftest_kernel): out[128,256] from x[128,64], h[128,128], w_ih[384,64], w_hh[384,128].
// workspace ≥ 512 KB of device memory.
extern "C" int ctgcn_selftest_umma(const float* x, const float* h, const float* w_ih, const float* w_hh, float* out,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    const int nchunks = 2 * chunks_per_part(64) + 2 * chunks_per_part(H);
    const size_t need = (size_t)nchunks * CHUNK_BYTES + 4 * H * sizeof(float);
    CTGCN_REQUIRE(x && h && w_ih && w_hh && out && workspace && workspace_bytes >= need,
                  "selftest_umma: bad arguments (workspace needs %zu bytes)", need);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* packed = (uint8_t*)workspace;
    float* bias4 = (float*)(packed + (size_t)nchunks * CHUNK_BYTES);
    pack_weights_kernel<<<(nchunks * UNITS_PER_PLANE + 255) / 256, 256, 0, st>>>(w_ih, w_hh, nullptr, nullptr, 64, packed, bias4, 0);
    CTGCN_LAUNCH_OK("pack_weights_kernel(selftest)");
    const int smem = 4 * A_PLANE + CHUNK_BYTES + 64;
    CTGCN_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_selftest_kernel<<<1, 128, smem, st>>>(x, h, packed, out);
    CTGCN_LAUNCH_OK("umma_selftest_kernel");
    return CTGCN_OK;
}
