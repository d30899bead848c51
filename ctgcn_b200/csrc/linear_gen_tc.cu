// tcgen05 dense layer  y = act(x·Wᵀ + b)  for ARBITRARY widths (d_in ≤ 1024, d_out ≤ 512, both multiples of 4) — the MLP layers of
// the reference's CTGCN-S / CGCN-S configurations (layers.py:95-106 with hidden_dim = 500: 204→500→500→128 at models.py:230 with
// config/facebook.json), which linear_tc.cu (64 / 128-wide, weights resident) does not take and the fp32 kernel runs at the fp32
// FMA peak (0.58 ms per 60 K-row layer at cfg3: 76 % of that configuration's step).
// Numerics: THREE bf16 planes per operand here (a = hi + mid + lo, 24 significant bits) and six MMAs per product
//   hi·hi + hi·mid + mid·hi + hi·lo + mid·mid + lo·hi          (error ≈ 2⁻²⁴ per product: fp32-level)
// The two-plane / three-MMA scheme of the GRU kernels (≈ 1e-5 per layer) passed every CoreDiffusion golden but not the
// element-wise bar on `ctgcn_S_fb_T12` once the 3-layer selu MLP ran through it too (19 of 153 600 elements off by up to 2× the
// tolerance): three 500-wide layers with selu in between amplify the operand rounding.  These layers are a small part of the
// step, so they get the exact scheme.
//
// One persistent CTA per SM, 128-row tiles.  Neither operand is resident: the input runs through two 48 KB slots in slices of 64
// columns (loaders: fp32 rows → bf16 hi/lo planes, columns ≥ d_in are zeros), the packed weights through a 2-stage ring of 48 KB
// chunks (128 output columns × 64 k, zero-padded), consumption order slice-major:
//   for slice s: for chunk c: D[:, 128c .. 128c+127] += A_s · B_{s,c}ᵀ         (4 K-steps × 6 split MMAs, M = 128, N = 128)
// The whole [128 × d_out] accumulator lives in TMEM (≤ 512 columns; two buffers when d_out ≤ 256, so that the epilogue of a tile
// overlaps the MMAs of the next one).
//   warp 0       weight producer (cp.async.bulk + mbarrier tx)
//   warp 1       MMA issuer
//   warps 4-11   loaders (16 rows each)
//   warps 12-15  epilogue: tcgen05.ld → + bias → selu? → 128-bit stores
// The weight stream (d_in·d_out·4 bytes per tile from the L2) bounds this kernel at ≈ half the tensor peak for 512-wide layers;
// pairing CTAs as gru_tc2.cu does would halve it — not done: these layers are 60 K rows in the configurations that have them.
#include "common.cuh"
#include "tc_common.cuh"

namespace ctgcn {
namespace {
using namespace tc;

constexpr int TILE_M = 128, SLICE_K = 64, NCHUNK = 128, STAGES = 2;
constexpr int PLANE = TILE_M * SLICE_K * 2;          // 16 KB: one bf16 plane of a 128 × 64 operand block (A slice or B chunk)
constexpr int BLOCK = 3 * PLANE;                     // hi | mid | lo

constexpr int SM_A = 0;                              // 2 slots
constexpr int SM_B = SM_A + 2 * BLOCK;               // ring
constexpr int SM_BIAS = SM_B + STAGES * BLOCK;       // 512 floats
constexpr int SM_STAGE = SM_BIAS + 512 * 4;          // epilogue staging: 4 warps × 32 rows × 64 B (see linear_tc.cu)
constexpr int SM_BAR = SM_STAGE + 4 * 2048;
enum { A_READY = 0, A_FREE = 2, B_FULL = 4, B_EMPTY = 4 + STAGES, ACC_FULL = 4 + 2 * STAGES, ACC_FREE = 6 + 2 * STAGES, NUM_BARS = 8 + 2 * STAGES };
constexpr int SM_TMEM_PTR = SM_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = SM_TMEM_PTR + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int NUM_LOADER_WARPS = 8, NUM_EPI_WARPS = 4, FIRST_LOADER_WARP = 4, FIRST_EPI_WARP = 12, THREADS = 512;

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Packed weights: chunk (s, c) at (s·nchunks + c)·BLOCK; element (row n of the chunk, k of the slice) of a plane at
// (k/8)·2048 + n·16 + (k%8)·2; rows ≥ d_out and columns ≥ d_in are zeros.
__global__ void pack_gen_kernel(const float* __restrict__ w, int d_in, int d_out, int nslices, int nchunks, uint8_t* __restrict__ packed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int UNITS = NCHUNK * (SLICE_K / 8);     // 16-byte units of one plane of a chunk
    if (t >= nslices * nchunks * UNITS) return;
    const int chunk = t / UNITS, u = t % UNITS;
    const int s = chunk / nchunks, c = chunk % nchunks;
    const int kb = u / NCHUNK, row = u % NCHUNK;
    const int n = c * NCHUNK + row, k0 = s * SLICE_K + kb * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (n < d_out && k0 + i < d_in) ? w[(int64_t)n * d_in + k0 + i] : 0.f;
    uint4 hi, mid, lo;
    split3_8(v, hi, mid, lo);
    uint8_t* dst = packed + (size_t)chunk * BLOCK + kb * (NCHUNK * 16) + row * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + PLANE) = mid;
    *reinterpret_cast<uint4*>(dst + 2 * PLANE) = lo;
}

struct ParamsG {
    const float* x;
    int64_t ldx, n;
    int d_in, d_out, nslices, nchunks;
    const uint8_t* packed;
    const float* bias;
    int act;
    float* y;
    int64_t ldy;
    int num_tiles;
};

__global__ void __launch_bounds__(THREADS, 1) linear_gen_kernel(const ParamsG p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int i) { return sbase + SM_BAR + 8u * i; };
    const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int nbuf = p.nchunks * NCHUNK <= 256 ? 2 : 1;          // accumulator buffers of 256 columns

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(A_READY + b), NUM_LOADER_WARPS);
            mbar_init(bar(A_FREE + b), 1);
            mbar_init(bar(ACC_FULL + b), 1);
            mbar_init(bar(ACC_FREE + b), NUM_EPI_WARPS);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(B_FULL + s), 1);
            mbar_init(bar(B_EMPTY + s), 1);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 512; i += THREADS)
        reinterpret_cast<float*>(smem + SM_BIAS)[i] = (p.bias && i < p.d_out) ? p.bias[i] : 0.f;
    if (warp == 1) tmem_alloc(sbase + SM_TMEM_PTR, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_PTR);

    if (warp == 0) {
        // ===================================================== weight producer: every tile streams the whole packed matrix
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const int total = p.nslices * p.nchunks;
            for (int t = 0; t < my_tiles; ++t) {
                for (int c = 0; c < total; ++c) {
                    mbar_wait(bar(B_EMPTY + stage), phase ^ 1);
                    mbar_expect_tx(bar(B_FULL + stage), BLOCK);
                    bulk_g2s(sbase + SM_B + stage * BLOCK, p.packed + (size_t)c * BLOCK, BLOCK, bar(B_FULL + stage));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, NCHUNK);
        constexpr uint32_t K_STEP = (2 * TILE_M * 16) >> 4, PL = PLANE >> 4;
        uint32_t stage = 0, phase = 0, use = 0;                  // use: running slice counter (slot = use & 1)
        for (int t = 0; t < my_tiles; ++t) {
            const int b = nbuf == 2 ? (t & 1) : 0;
            const uint32_t acc_par = nbuf == 2 ? ((t >> 1) & 1) : (t & 1);
            mbar_wait(bar(ACC_FREE + b), acc_par ^ 1);
            tc_fence_after();
            for (int s = 0; s < p.nslices; ++s, ++use) {
                mbar_wait(bar(A_READY + (use & 1)), (use >> 1) & 1);
                tc_fence_after();
                const uint32_t a0 = desc_lo(sbase + SM_A + (use & 1) * BLOCK, TILE_M * 16);
                for (int c = 0; c < p.nchunks; ++c) {
                    mbar_wait(bar(B_FULL + stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b0 = desc_lo(sbase + SM_B + stage * BLOCK, NCHUNK * 16);
                        const uint32_t d = tmem + b * 256 + c * NCHUNK;
#pragma unroll
                        for (int ks = 0; ks < SLICE_K / 16; ++ks) {
                            const uint64_t ah = desc64(a0 + ks * K_STEP), am = desc64(a0 + PL + ks * K_STEP), al = desc64(a0 + 2 * PL + ks * K_STEP);
                            const uint64_t bh = desc64(b0 + ks * K_STEP), bm = desc64(b0 + PL + ks * K_STEP), bl = desc64(b0 + 2 * PL + ks * K_STEP);
                            umma_bf16(d, al, bh, idesc, (s == 0 && ks == 0) ? 0u : 1u);    // small terms first
                            umma_bf16(d, ah, bl, idesc, 1u);
                            umma_bf16(d, am, bm, idesc, 1u);
                            umma_bf16(d, am, bh, idesc, 1u);
                            umma_bf16(d, ah, bm, idesc, 1u);
                            umma_bf16(d, ah, bh, idesc, 1u);
                        }
                        umma_commit(bar(B_EMPTY + stage));
                    }
                    __syncwarp();
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (elect_one()) umma_commit(bar(A_FREE + (use & 1)));
                __syncwarp();
            }
            if (elect_one()) umma_commit(bar(ACC_FULL + b));
            __syncwarp();
        }
    } else if (warp < FIRST_LOADER_WARP) {
        // idle
    } else if (warp < FIRST_EPI_WARP) {
        // ===================================================== loaders: 16 tile rows per warp, one 64-column slice at a time
        // one 32-row × 32-column block per warp and slice (StageBlock, tc_common.cuh): row group w % 4, column half w / 4
        const int lw = warp - FIRST_LOADER_WARP, g = lw & 3, half = lw >> 2;
        // Software pipeline over the flattened (tile, slice) sequence: the loads of slice k+1 are issued before slice k is converted
        // and stored, so a slice is always in flight (two register blocks, used alternately).
        const int total = my_tiles * p.nslices;
        auto issue = [&](int k, StageBlock& blk) {
            const int t = k / p.nslices, s = k - t * p.nslices;
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + 32 * g;
            const int64_t left = p.n - row0;
            const int c0 = s * SLICE_K + 32 * half;
            blk.load(p.x + row0 * p.ldx + c0, p.ldx, left > 32 ? 32 : (int)left, p.d_in - c0, lane);
        };
        auto stage = [&](int k, const StageBlock& blk) {
            const uint32_t use = (uint32_t)k;                    // running slice counter: slot = use & 1
            mbar_wait(bar(A_FREE + (use & 1)), ((use >> 1) & 1) ^ 1);
            blk.store<3>(smem + SM_A + (use & 1) * BLOCK + (4 * half) * (TILE_M * 16) + (32 * g) * 16, PLANE, TILE_M * 16, lane);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(A_READY + (use & 1)));
        };
        StageBlock ba, bb;
        if (total > 0) issue(0, ba);
        for (int k = 0; k < total; k += 2) {
            if (k + 1 < total) issue(k + 1, bb);
            stage(k, ba);
            if (k + 1 < total) {
                if (k + 2 < total) issue(k + 2, ba);
                stage(k + 1, bb);
            }
        }
    } else {
        // ===================================================== epilogue: thread = tile row
        const int q = warp & 3;
        const float* bias = reinterpret_cast<const float*>(smem + SM_BIAS);
        const uint32_t tmem_lane = tmem + ((uint32_t)(32 * q) << 16);
        for (int t = 0; t < my_tiles; ++t) {
            const int b = nbuf == 2 ? (t & 1) : 0;
            const uint32_t acc_par = nbuf == 2 ? ((t >> 1) & 1) : (t & 1);
            mbar_wait(bar(ACC_FULL + b), acc_par);
            tc_fence_after();
            uint8_t* stage = smem + SM_STAGE + q * 2048;
            const int64_t warp_row0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + 32 * q;
            for (int c = 0; c < p.d_out; c += 16) {
                float v0[8], v1[8];
                tmem_ld8(tmem_lane + b * 256 + c, v0);
                tmem_ld8(tmem_lane + b * 256 + c + 8, v1);
                tmem_ld_wait();
                // raw accumulators row-major through the warp's staging tile (XOR-swizzled 16-byte pieces): 8 rows × 64 B per store
                // instruction instead of 32 rows × 16 B; bias and activation after the read-back (a lane owns 4 fixed columns there)
#pragma unroll
                for (int pc = 0; pc < 4; ++pc) {
                    const float* o = pc < 2 ? v0 + 4 * pc : v1 + 4 * (pc - 2);
                    *reinterpret_cast<float4*>(stage + lane * 64 + ((pc ^ ((lane >> 1) & 3)) << 4)) = make_float4(o[0], o[1], o[2], o[3]);
                }
                __syncwarp();
                const int pc = lane & 3;
                const float4 b4 = *reinterpret_cast<const float4*>(bias + c + 4 * pc);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = 8 * i + (lane >> 2);
                    float4 v = *reinterpret_cast<const float4*>(stage + r * 64 + ((pc ^ ((r >> 1) & 3)) << 4));
                    v.x += b4.x;
                    v.y += b4.y;
                    v.z += b4.z;
                    v.w += b4.w;
                    if (p.act == CTGCN_ACT_SELU) {
                        v.x = selu_fast(v.x);
                        v.y = selu_fast(v.y);
                        v.z = selu_fast(v.z);
                        v.w = selu_fast(v.w);
                    }
                    if (warp_row0 + r < p.n && c + 4 * pc < p.d_out)
                        *reinterpret_cast<float4*>(p.y + (warp_row0 + r) * p.ldy + c + 4 * pc) = v;
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(ACC_FREE + b));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

bool linear_gen_tc_takes(int64_t d_in, int64_t d_out) {
    return d_in >= 8 && d_in <= 1024 && d_out >= 8 && d_out <= 512 && (d_in & 3) == 0 && (d_out & 3) == 0;
}
size_t linear_gen_tc_workspace_bytes(int64_t d_in, int64_t d_out) {
    if (!linear_gen_tc_takes(d_in, d_out)) return 0;
    return (size_t)ceil_div((int)d_in, SLICE_K) * ceil_div((int)d_out, NCHUNK) * BLOCK;
}

// returns 0 = done, <0 = error, 1 = shape not supported by this path
int launch_linear_gen_tc(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* w, const float* b, int64_t d_out, int act,
                         float* y, int64_t ldy, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!linear_gen_tc_takes(d_in, d_out)) return 1;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(x) || !al16(y) || (ldx & 3) || (ldy & 3)) return 1;
    const size_t need = linear_gen_tc_workspace_bytes(d_in, d_out);
    CTGCN_REQUIRE(ws && ws_bytes >= need, "linear_gen_tc: workspace of %zu bytes, need %zu", ws_bytes, need);
    ParamsG p;
    p.x = x;
    p.ldx = ldx;
    p.n = n;
    p.d_in = (int)d_in;
    p.d_out = (int)d_out;
    p.nslices = ceil_div((int)d_in, SLICE_K);
    p.nchunks = ceil_div((int)d_out, NCHUNK);
    p.packed = (const uint8_t*)ws;
    p.bias = b;
    p.act = act;
    p.y = y;
    p.ldy = ldy;
    p.num_tiles = (int)((n + TILE_M - 1) / TILE_M);
    {
        ProfScope prof(PROF_PACK, st);
        const int units = p.nslices * p.nchunks * NCHUNK * (SLICE_K / 8);
        pack_gen_kernel<<<(units + 255) / 256, 256, 0, st>>>(w, p.d_in, p.d_out, p.nslices, p.nchunks, (uint8_t*)ws);
        CTGCN_LAUNCH_OK("pack_gen_kernel");
    }
    int dev = 0, sm_count = 0;
    CTGCN_CUDA_OK(cudaGetDevice(&dev));
    CTGCN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    CTGCN_CUDA_OK(cudaFuncSetAttribute(linear_gen_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
    ProfScope prof(PROF_LINEAR, st);
    linear_gen_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(p);
    CTGCN_LAUNCH_OK("linear_gen_kernel");
    return CTGCN_OK;
}

}  // namespace ctgcn
