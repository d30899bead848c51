// tcgen05 dense layer  y = act(x·Wᵀ + b)  for the 128-wide MLP layers (layers.py:95-106) — same split-bf16 scheme as
// gru_tc.cu (hi·hi + lo·hi + hi·lo, fp32 accumulation in TMEM) so that the 1e-4 parity bar holds.
//
// The layer is HBM-bound (4·(d_in + d_out) bytes per row against 6·d_in·d_out issued flops): one persistent CTA per SM
// keeps the packed weights (≤ 64 KB) resident in shared memory and streams 128-row tiles through a double-buffered
// A operand and two TMEM accumulator buffers:
//   warp 0       MMA issuer (elected lane): M=128, N=d_out, K=16, no-swizzle K-major descriptors
//   warps 4-11   loaders: fp32 rows → bf16 hi/lo planes in core-matrix order; a whole tile (64 KB) is in flight at once
//   warps 12-15  epilogue: tcgen05.ld → + bias → selu? → 128-bit stores
// Shapes: d_in, d_out ∈ {64, 128}; everything else goes to the SIMT kernel in gru_simt.cu.
#include "common.cuh"
#include "tc_common.cuh"

namespace ctgcn {
namespace {
using namespace tc;

constexpr int TILE_M = 128;
constexpr int MAX_D = 128;
constexpr int A_PLANE = TILE_M * MAX_D * 2;   // 32 KB
constexpr int SM_A = 0;                       // 2 buffers × (hi | lo)
constexpr int SM_B = SM_A + 4 * A_PLANE;      // weights hi | lo (plane = d_out·d_in·2 ≤ 32 KB)
constexpr int SM_BIAS = SM_B + 2 * A_PLANE;
constexpr int SM_STAGE = SM_BIAS + MAX_D * 4;  // epilogue staging: 4 warps × 32 rows × 64 B (16 output columns at a time)
constexpr int SM_BAR = SM_STAGE + 4 * 2048;
constexpr int NUM_BARS = 9;                   // a_ready[2] a_free[2] acc_full[2] acc_free[2] w_full
constexpr int SM_TMEM_PTR = SM_BAR + NUM_BARS * 8;
constexpr int SMEM_BYTES = SM_TMEM_PTR + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
constexpr int NUM_LOADER_WARPS = 8, NUM_EPI_WARPS = 4;
constexpr int FIRST_LOADER_WARP = 4, FIRST_EPI_WARP = 12;
constexpr int THREADS = 32 * 16;
enum { BAR_A_READY = 0, BAR_A_FREE = 2, BAR_ACC_FULL = 4, BAR_ACC_FREE = 6, BAR_W_FULL = 8 };

// Weight image: element (n, k) of W [d_out, d_in] at (k/8)·(d_out·16) + n·16 + (k%8)·2, hi plane then lo plane.
__global__ void pack_linear_weights_kernel(const float* __restrict__ w, int d_in, int d_out, uint8_t* __restrict__ packed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int units = d_out * (d_in / 8);
    if (t >= units) return;
    const int kb = t / d_out, n = t % d_out;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = w[(int64_t)n * d_in + kb * 8 + i];
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* dst = packed + kb * (d_out * 16) + n * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + d_out * d_in * 2) = lo;
}

struct Params {
    const float* x;
    int64_t ldx, n;
    int d_in, d_out;
    const uint8_t* packed;
    const float* bias;
    int act;
    float* y;
    int64_t ldy;
    int num_tiles;
};

template <int N>
__device__ __forceinline__ void issue_tile(uint32_t a_lo32, uint32_t b_lo32, uint32_t b_plane16, uint32_t d_tmem, int ksteps) {
    constexpr uint32_t idesc = umma_idesc_bf16(TILE_M, N);
    constexpr uint32_t A_STEP = (2 * TILE_M * 16) >> 4, B_STEP = (2 * N * 16) >> 4, A_LO_PLANE = A_PLANE >> 4;
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t ah = desc64(a_lo32 + ks * A_STEP), al = desc64(a_lo32 + A_LO_PLANE + ks * A_STEP);
        const uint64_t bh = desc64(b_lo32 + ks * B_STEP), bl = desc64(b_lo32 + b_plane16 + ks * B_STEP);
        umma_bf16(d_tmem, ah, bh, idesc, ks == 0 ? 0u : 1u);
        umma_bf16(d_tmem, al, bh, idesc, 1u);
        umma_bf16(d_tmem, ah, bl, idesc, 1u);
    }
}

__global__ void __launch_bounds__(THREADS, 1) linear_tc_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto bar = [&](int i) { return sbase + SM_BAR + 8u * i; };
    const int my_tiles = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t w_plane = (uint32_t)p.d_out * p.d_in * 2;

    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(BAR_A_READY + b), NUM_LOADER_WARPS);
            mbar_init(bar(BAR_A_FREE + b), 1);
            mbar_init(bar(BAR_ACC_FULL + b), 1);
            mbar_init(bar(BAR_ACC_FREE + b), NUM_EPI_WARPS);
        }
        mbar_init(bar(BAR_W_FULL), 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < p.d_out; i += THREADS) reinterpret_cast<float*>(smem + SM_BIAS)[i] = p.bias ? p.bias[i] : 0.f;
    if (warp == 0) tmem_alloc(sbase + SM_TMEM_PTR, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + SM_TMEM_PTR);

    if (warp == 0) {
        // ===================================================== weights (once) + MMA issuer
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (elect_one()) {
            mbar_expect_tx(bar(BAR_W_FULL), 2 * w_plane);
            bulk_g2s(sbase + SM_B, p.packed, 2 * w_plane, bar(BAR_W_FULL));
        }
        __syncwarp();
        mbar_wait(bar(BAR_W_FULL), 0);
        const uint32_t b_desc = desc_lo(sbase + SM_B, p.d_out * 16);
        for (int t = 0; t < my_tiles; ++t) {
            const int b = t & 1;
            const uint32_t par = (t >> 1) & 1;
            mbar_wait(bar(BAR_A_READY + b), par);
            mbar_wait(bar(BAR_ACC_FREE + b), par ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_desc = desc_lo(sbase + SM_A + b * 2 * A_PLANE, TILE_M * 16);
                if (p.d_out == 128) issue_tile<128>(a_desc, b_desc, w_plane >> 4, tmem + b * 128, p.d_in / 16);
                else issue_tile<64>(a_desc, b_desc, w_plane >> 4, tmem + b * 128, p.d_in / 16);
                umma_commit(bar(BAR_A_FREE + b));
                umma_commit(bar(BAR_ACC_FULL + b));
            }
            __syncwarp();
        }
    } else if (warp < FIRST_LOADER_WARP) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else if (warp < FIRST_EPI_WARP) {
        // ===================================================== loaders: 16 tile rows per warp, the whole tile in flight
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        // 32-row × 32-column blocks (StageBlock, tc_common.cuh): warp w takes row group w % 4 and the column spans w / 4, w / 4 + 2
        const int lw = warp - FIRST_LOADER_WARP, g = lw & 3;
        const int nspans = p.d_in / 32, per_warp = nspans / 2;      // 2 (d_in = 128) or 1 (d_in = 64)
        // Software pipeline: the loads of tile t+1 are issued block by block as soon as the block's registers have been stored for
        // tile t, so a whole tile (64 KB per SM) stays in flight while the previous one is converted.
        StageBlock blk[2];
        auto issue = [&](int t, int u) {
            const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + 32 * g;
            const int64_t left = p.n - row0;
            blk[u].load(p.x + row0 * p.ldx + 32 * ((lw >> 2) + 2 * u), p.ldx, left > 32 ? 32 : (int)left, 32, lane);
        };
        if (my_tiles > 0) {
            issue(0, 0);
            if (per_warp > 1) issue(0, 1);
        }
        for (int t = 0; t < my_tiles; ++t) {
            const int b = t & 1;
            uint8_t* a_hi = smem + SM_A + b * 2 * A_PLANE;
            mbar_wait(bar(BAR_A_FREE + b), ((t >> 1) & 1) ^ 1);
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (u < per_warp) {
                    const int sp = (lw >> 2) + 2 * u;
                    blk[u].store(a_hi + (4 * sp) * (TILE_M * 16) + (32 * g) * 16, A_PLANE, TILE_M * 16, lane);
                    if (t + 1 < my_tiles) issue(t + 1, u);
                }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_A_READY + b));
        }
    } else {
        // ===================================================== epilogue: thread = tile row
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        const int q = warp & 3;
        const int m = 32 * q + lane;
        const float* bias = reinterpret_cast<const float*>(smem + SM_BIAS);
        const uint32_t tmem_lane = tmem + ((uint32_t)(32 * q) << 16);
        for (int t = 0; t < my_tiles; ++t) {
            const int b = t & 1;
            mbar_wait(bar(BAR_ACC_FULL + b), (t >> 1) & 1);
            tc_fence_after();
            // Stores go through a 2 KB staging tile per warp: a thread owns a ROW of the accumulator, so direct stores touch 32
            // different 128-byte lines per instruction (16 bytes each); re-read row-major, an instruction covers 8 rows × 64 B.
            // 16-byte pieces are XOR-swizzled by row pair: both the writes (lane = row) and the reads are bank-conflict-free.
            uint8_t* stage = smem + SM_STAGE + q * 2048;
            const int64_t warp_row0 = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * TILE_M + 32 * q;
            for (int c = 0; c < p.d_out; c += 16) {
                float v0[8], v1[8];
                tmem_ld8(tmem_lane + b * 128 + c, v0);
                tmem_ld8(tmem_lane + b * 128 + c + 8, v1);
                tmem_ld_wait();
                // raw accumulators through the staging tile; bias and activation are applied after the transposed read-back, where a
                // lane owns 4 fixed columns of 4 rows: one 128-bit bias load per lane and column group instead of 16 broadcast loads
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
                    if (warp_row0 + r < p.n) *reinterpret_cast<float4*>(p.y + (warp_row0 + r) * p.ldy + c + 4 * pc) = v;
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(BAR_ACC_FREE + b));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// returns 0 = done, <0 = error, 1 = shape not supported by this path.  workspace ≥ 4·d_in·d_out bytes (packed weights).
int launch_linear_tc(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* w, const float* b, int64_t d_out,
                     int act, float* y, int64_t ldy, void* ws, cudaStream_t st) {
    if ((d_in != 64 && d_in != 128) || (d_out != 64 && d_out != 128)) return 1;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(x) || !al16(y) || (ldx & 3) || (ldy & 3)) return 1;
    uint8_t* packed = (uint8_t*)ws;
    {
        ProfScope prof(PROF_PACK, st);
        const int units = (int)(d_out * (d_in / 8));
        pack_linear_weights_kernel<<<(units + 255) / 256, 256, 0, st>>>(w, (int)d_in, (int)d_out, packed);
        CTGCN_LAUNCH_OK("pack_linear_weights_kernel");
    }
    // per device (a process may drive several GPUs; the attribute belongs to the device's context): set on every call
    int dev = 0, sm_count = 0;
    CTGCN_CUDA_OK(cudaGetDevice(&dev));
    CTGCN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    CTGCN_CUDA_OK(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    Params p;
    p.x = x;
    p.ldx = ldx;
    p.n = n;
    p.d_in = (int)d_in;
    p.d_out = (int)d_out;
    p.packed = packed;
    p.bias = b;
    p.act = act;
    p.y = y;
    p.ldy = ldy;
    p.num_tiles = (int)((n + TILE_M - 1) / TILE_M);
    const int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
    ProfScope prof(PROF_LINEAR, st);
    linear_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(p);
    CTGCN_LAUNCH_OK("linear_tc_kernel");
    return CTGCN_OK;
}

}  // namespace ctgcn
