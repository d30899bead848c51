// tcgen05 GRU over a short sequence + Σ/LayerNorm epilogue on CTA PAIRS (cta_group::2) — the compute-bound half of
// CoreDiffusion.forward (layers.py:59-62) and the temporal GRU of CTGCN.forward (models.py:249-250).
//
// Why pairs (profiles/r02_experiments.md): the one-CTA kernel (gru_tc.cu) is bound by the SM's shared-memory / L1 data path
// (128 B/cycle), not by its recurrence chain: per 128-row tile-step the tensor core fetches 960 KB of operands and the weight
// ring is re-written with the full 393 KB packed weight set — ≈ 1.65 MB ≈ 12.9 K cycles of that pipe against 9.2 K cycles of
// MMA work.  With cta_group::2 the two CTAs of a cluster share every weight chunk: each CTA holds HALF of the chunk's B rows,
// the leader issues M = 256 MMAs over both CTAs' 128-row A tiles (accumulators stay per CTA: 128 TMEM lanes each), so per SM
// the B-operand fetch and the ring writes halve (≈ 1.17 MB ≈ 9.1 K cycles per tile-step).
//
// Numerics as in gru_tc.cu: every fp32 operand is split a = hi + lo (two bf16 planes) and each product is three MMAs
// hi·hi + lo·hi + hi·lo accumulated in fp32 in TMEM.  New here: the biases are added by the tensor core — ONE extra K = 16 MMA
// per 64-feature block and step, A = a resident block of ones (k = 0, 1), B = the block's [b_in | b_ir+b_hr | b_iz+b_hz | b_hn]
// as bf16 hi (k = 0) and lo (k = 1), N = 256, opens the accumulation of the whole accumulator set.  Every other MMA is then a
// plain N = 192 accumulate (no "fresh"/"split-first" special cases, which would split B differently across the pair), step 0
// needs no special gate code (W_hn·h + b_hn = b_hn is already in the accumulator), and the gate warps lose their 32 bias adds and
// 8 broadcast shared-memory loads per 8-feature pass (measured harmless to the result: relL2 3.4e-7, r02_experiments.md).
//
// Per CTA (512 threads, one 128-row tile for the WHOLE sequence; both CTAs of a pair run in lock step):
//   warp 0      weight producer: this CTA's half chunks (96 rows × 32 k, hi|lo = 12 KB) through a 6-stage ring (cp.async.bulk)
//   warp 1      leader: MMA issuer.  follower: relays "my half chunk has landed" to the leader (a plain bulk copy can only
//               complete_tx on a barrier of its own CTA)
//   warps 4-7   input loaders: fp32 rows → bf16 hi/lo planes in UMMA core-matrix order
//   warps 8-15  gate math (tcgen05.ld, ex2/rcp sigmoid & tanh), h written back as the next step's A operand, Σh / LayerNorm
// Barriers the leader's MMA warp waits on (U_READY, H_READY, ACC_FREE*, W_PEER*) live in the LEADER's shared memory and count the
// arrivals of both CTAs (the follower's warps arrive remotely: mapa + mbarrier.arrive, CTA-scope release — see tc_common.cuh for
// why not cluster scope); everything the MMAs signal (W_EMPTY*, U_FREE, ACC_FULL*) is committed to both CTAs at once
// (tcgen05.commit … multicast::cluster).  The input loaders pull the NEXT step's rows into the L2 one step ahead (prefetch.global.L2).
// The same code compiles for CG = 1 (no cluster, whole chunks, local barriers): the A/B check of the pairing itself.
//
// Measured (B200, 1 M rows, profiles/r02_experiments.md): core GRU (10 steps, SUM_LN) 4.74 ms, temporal (8 steps, EACH_LN) 4.46-4.70 ms
// against 5.27-5.33 / 5.60 ms for the one-CTA kernel of round 1 and 4.79-5.10 / 5.08-5.26 ms for this kernel without pairing.
// Variants that were built, measured and removed: 16 gate warps (4.73 ms: the gate math of a block is MUFU-paced, ≈ 3.9 K cycles
// with 8 or 16 warps against a floor of 2.56 K), a double-buffered h block 0 (4.77 ms: costs two ring stages), two concurrent
// gate groups + both input parts ahead of "h ready" + early accumulator release (5.08 ms: the sets are only free at the very end
// of a group's gate phase, so the input parts land ON the chain).
//
// Shapes: H = 128, d_in ∈ {32, 64, 96, 128}.  Other shapes: gru_simt.cu.
#include "common.cuh"
#include "tc_common.cuh"

namespace ctgcn {
namespace {

using namespace tc;

constexpr int H = 128, TILE_M = 128, BLK = 64, NB = H / BLK;   // NB blocks of 64 hidden features, one accumulator set each
constexpr int CHUNK_K = 32, GATE_ROWS = 192;
constexpr int A_PLANE = TILE_M * H * 2;                        // 32 KB: one bf16 plane of a 128 × 128 operand tile
constexpr int NUM_LOADER_WARPS = 4;
constexpr int FIRST_LOADER_WARP = 4, FIRST_WORKER_WARP = 8;    // warp 0 producer, warp 1 MMA / relay, warps 2-3 idle
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t COL_IN = 0, COL_R = 64, COL_Z = 128, COL_HN = 192;   // inside a 256-column accumulator set

constexpr int NW = 8;                                           // gate-math warps
constexpr int FIRST_GATHER_WARP = 2, NUM_GATHER_WARPS = 6;      // FUSED only: warps 2-7 (idle warps + the input loaders' warps)
constexpr int G_SLOT = 512, G_GROUP = 5, G_GROUPS = 2;          // gather ring of one warp: row slots of 512 B, cp.async groups of 5
constexpr int G_SLOTS = G_GROUP * G_GROUPS;                     // 10 rows = 5 KB in flight per warp, 30 KB per CTA
constexpr int U_IMAGE = 2 * A_PLANE;                            // one step's U operand image (hi | lo planes) in the scratch
// FUSED (CG = 2 only): the cumulative k-core SpMM of layers.py:41-48 runs INSIDE this kernel — warps 2-7 of every CTA gather the
// per-core sums U of the tile the CTA processes NEXT while the tensor cores work on the current one (see the gather section).
template <int CG, bool FUSED = false>
struct Lay {   // per-CTA shared-memory map (identical in both CTAs of a pair: the MMA descriptors are CTA-relative)
    static_assert(!FUSED || CG == 2, "the fused build needs the shared memory the pairing frees");
    static constexpr int ROWS = GATE_ROWS / CG;                  // weight rows of a chunk held by one CTA
    static constexpr int W_PLANE = ROWS * CHUNK_K * 2;
    static constexpr int W_CHUNK = 2 * W_PLANE;                  // hi | lo
    static constexpr int STAGES = CG == 2 ? (FUSED ? 4 : 6) : 3; // 72 KB of weights in flight (FUSED: 48 KB + the gather ring)
    static constexpr int FOLD_ROWS = 256 / CG;                   // bias rows [in | r | z | hn] of a block held by one CTA
    static constexpr int FOLD_ONES = 2 * TILE_M * 16;            // A block [2 k-blocks][128 rows][8 bf16]
    static constexpr int FOLD_BIAS = 2 * FOLD_ROWS * 16;         // B block of one feature block
    static constexpr int FOLD_BYTES = FOLD_ONES + NB * FOLD_BIAS;
    static constexpr int SM_U = 0;                               // U hi | lo (planes of A_PLANE bytes)
    static constexpr int SM_H = SM_U + 2 * A_PLANE;              // h hi | lo
    static constexpr int SM_W = SM_H + 2 * A_PLANE;              // weight ring
    static constexpr int SM_G = SM_W + STAGES * W_CHUNK;          // gather rings (FUSED): [6 warps][10 slots][512 B]
    static constexpr int G_BYTES = FUSED ? NUM_GATHER_WARPS * G_SLOTS * G_SLOT : 0;
    static constexpr int SM_FOLD = SM_G + G_BYTES;
    static constexpr int SM_LN = SM_FOLD + FOLD_BYTES;           // ln_w | ln_b
    static constexpr int SM_RED = SM_LN + 2 * H * 4;             // [2 buffers][2 feature groups][128 rows] fp32
    static constexpr int SM_ROWPTR = SM_RED + 2 * 2 * TILE_M * 4;   // FUSED: [6 warps][24] int32 row pointers of a warp's ≤ 22 rows
    static constexpr int SM_BAR = SM_ROWPTR + (FUSED ? NUM_GATHER_WARPS * 24 * 4 : 0);
    // U_READY: !FUSED: the loaders of both CTAs have staged U (leader).  FUSED: "the follower's U image has landed" (leader, 1 arrival)
    enum { W_FULL = 0, W_EMPTY = STAGES, W_PEER = 2 * STAGES, U_READY = 3 * STAGES, U_FREE, H_READY, ACC_FULL0, ACC_FULL1,
           ACC_FREE0, ACC_FREE1, TILE_READY0, TILE_READY1, TILE_FREE0, TILE_FREE1, U_FULL, U_READY1, U_FREE1, NUM_BARS };
    static constexpr int SM_TMEM_PTR = SM_BAR + NUM_BARS * 8;
    static constexpr int SMEM_BYTES = SM_TMEM_PTR + 16;
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static constexpr int THREADS = 32 * (FIRST_WORKER_WARP + NW);
    // setmaxnreg budgets (warps 0-3 | loaders 4-7 | gate warps) out of the 512 × 128 registers the CTA starts with
    static constexpr int REG_WG0 = FUSED ? 88 : 56, REG_LOAD = FUSED ? 88 : 112, REG_GATE = 168;   // FUSED: warps 2-7 run the gather
    static_assert(128 * REG_WG0 + 128 * REG_LOAD + 32 * NW * REG_GATE <= THREADS * 128, "register pool");
    // byte offset (from the CTA's shared-memory base) of the hi 16-byte unit holding features f..f+7 (f % 8 == 0) of row m
    static __device__ __forceinline__ uint32_t h_unit(int f, int m) { return SM_H + (f >> 3) * (TILE_M * 16) + m * 16; }
    static constexpr int H_LO = A_PLANE;                         // byte distance hi plane → lo plane
};

// ------------------------------------------------------------------------------------------------ weight packing
// Chunk order: part ∈ {X block0, X block1, H block0, H block1}, then k-chunk of 32.  An X chunk holds the rows [n | r | z] of W_ih
// for the block's 64 hidden features, an H chunk the rows [r | z | n] of W_hh.  Per chunk: rank 0's ROWS rows (hi plane, lo
// plane), then rank 1's; element (row, k) of a plane at (k/8)·ROWS·16 + row·16 + (k%8)·2 (no-swizzle K-major core matrices,
// LBO = ROWS·16, SBO = 128).  r/z rows and biases are multiplied by −log2(e), n rows by 2·log2(e): sigmoid/tanh need a bare ex2.
// Then per rank the bias-fold image: ones block | NB bias blocks (see the header comment).
__host__ __device__ constexpr int chunks_of(int k) { return k / CHUNK_K; }
// weight chunks of one block of the input part: d_in ≤ 128 (a multiple of 32): exact; wider inputs are handled in slices of
// SLICE_K = 64 columns (two chunks), the last one zero-padded
constexpr int SLICE_K = 64;
__host__ __device__ constexpr int x_chunks(int d_in) { return d_in <= H ? d_in / CHUNK_K : (d_in + SLICE_K - 1) / SLICE_K * (SLICE_K / CHUNK_K); }

template <int CG>
__global__ void pack2_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                             const float* __restrict__ b_hh, int d_in, uint8_t* __restrict__ packed, uint8_t* __restrict__ fold) {
    using LY = Lay<CG>;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int UNITS = GATE_ROWS * (CHUNK_K / 8);             // 16-byte units of one plane of a whole chunk (all ranks)
    const int cx = x_chunks(d_in), chh = chunks_of(H);           // X chunks cover d_in padded with zero columns
    const int nchunks = NB * cx + NB * chh;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    // ---- bias-fold images: CG × (256 ones units + NB × 2·FOLD_ROWS bias units)
    constexpr int FOLD_UNITS = LY::FOLD_BYTES / 16;
    if (t < CG * FOLD_UNITS) {
        const int rank = t / FOLD_UNITS, u = t % FOLD_UNITS;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (u < TILE_M) {
            v.x = 0x3f803f80u;                                   // k = 0, 1: bf16 1.0 | 1.0 (k-block 0 of every row)
        } else if (u >= LY::FOLD_ONES / 16) {
            const int b = u - LY::FOLD_ONES / 16, blk = b / (2 * LY::FOLD_ROWS), r = b % (2 * LY::FOLD_ROWS);
            if (r < LY::FOLD_ROWS && b_ih) {                     // k-block 0; k-block 1 stays zero
                const int n = rank * LY::FOLD_ROWS + r, g4 = n / BLK, f = blk * BLK + n % BLK;   // accumulator column order
                float bv;
                if (g4 == 0) bv = b_ih[2 * H + f] * (2.f * kLog2e);                              // b_in
                else if (g4 == 1) bv = (b_ih[f] + b_hh[f]) * -kLog2e;                            // b_ir + b_hr
                else if (g4 == 2) bv = (b_ih[H + f] + b_hh[H + f]) * -kLog2e;                    // b_iz + b_hz
                else bv = b_hh[2 * H + f] * (2.f * kLog2e);                                      // b_hn
                const __nv_bfloat16 bh = __float2bfloat16_rn(bv);
                const __nv_bfloat16 bl = __float2bfloat16_rn(bv - __bfloat162float(bh));
                v.x = (uint32_t)__bfloat16_as_ushort(bh) | ((uint32_t)__bfloat16_as_ushort(bl) << 16);
            }
        }
        *reinterpret_cast<uint4*>(fold + (size_t)rank * LY::FOLD_BYTES + 16 * (size_t)u) = v;
    }
    if (t >= nchunks * UNITS) return;
    const int c = t / UNITS, unit = t % UNITS;
    const bool is_x = c < NB * cx;
    const int cc = is_x ? c : c - NB * cx;
    const int per = is_x ? cx : chh;
    const int ktot = is_x ? d_in : H;
    const int blk = cc / per, kc = cc % per;
    const int kb = unit / GATE_ROWS, row = unit % GATE_ROWS;
    const int g3 = row / BLK, f = row % BLK;
    const int gate = is_x ? (g3 == 0 ? 2 : g3 - 1) : g3;        // X: [n, r, z]   H: [r, z, n]
    const int k0 = kc * CHUNK_K + kb * 8;
    const float* src = (is_x ? w_ih : w_hh) + (int64_t)(gate * H + blk * BLK + f) * ktot + k0;
    const float scale = gate < 2 ? -kLog2e : 2.f * kLog2e;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = k0 + i < ktot ? src[i] * scale : 0.f;
    uint4 hi, lo;
    split8(v, hi, lo);
    const int rank = row / LY::ROWS, rr = row % LY::ROWS;
    uint8_t* dst = packed + (size_t)c * (CG * LY::W_CHUNK) + (size_t)rank * LY::W_CHUNK + kb * (LY::ROWS * 16) + rr * 16;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + LY::W_PLANE) = lo;
}

// ------------------------------------------------------------------------------------------------ kernel
struct Params2 {
    const float* seq;
    int64_t srs, sss, n;
    int steps, d_in;
    const uint8_t* packed;
    const uint8_t* fold;
    const float* ln_w;
    const float* ln_b;
    float eps;
    float* y;
    int64_t yrs, yss;
    RowScatter sc;      // SUM_LN only: rows go to their node slice's buffer (fused snapshot exchange)
    int num_tiles;
    long long* trace;   // optional [32 events][64 steps] clock64 stamps of block 0 (ctgcn_debug_gru_trace), else NULL
    // FUSED: the graph plan (level-tagged union CSR), the layer input x [n, d_in] and the per-CTA scratch for U:
    // [CTA][2 tile slots][steps = K cores][64 KB operand image: bf16 hi plane | lo plane, (k/8)·2048 + row·16 + (k%8)·2].
    // seq / srs / sss are unused then.
    const int32_t* rowptr;
    const int32_t* col;
    const float* val;
    const uint8_t* lvl;
    const float* x;
    int64_t ldx;
    uint8_t* scratch;
};

#define GRU2_TRACE(e, gs)                                                                   \
    do {                                                                                    \
        if (p.trace && blockIdx.x == 0 && (gs) < 64u) p.trace[(e) * 64 + (gs)] = clock64(); \
    } while (0)

// SLICED (d_in > 128, e.g. the 500 → 128 first CoreDiffusion layer of every shipped CTGCN-C config, models.py:228): a 128-row U tile
// of that width does not fit next to h, so the input part runs slice-major — the loaders stage 64 input columns at a time into
// one of two 32 KB slots while the MMAs of BOTH blocks consume the other (accumulating across slices), then the recurrent parts
// follow as usual.  Both accumulator sets are busy from the first slice on, so the input phase of a step does not overlap the
// previous step's gate math: ≈ 70 % of the tensor-core bound at 512 → 128 instead of ≈ 85 %, against 1.6 % for the fp32 kernel.
template <int CG, int MODE, bool FUSED, bool SLICED = false>
__global__ void __launch_bounds__(Lay<CG, FUSED>::THREADS, 1) gru2_kernel(const Params2 p) {
    using LY = Lay<CG, FUSED>;
    static_assert(!FUSED || MODE == CTGCN_GRU_SUM_LN, "the fused build is the CoreDiffusion layer");
    static_assert(!(FUSED && SLICED), "the fused build takes d_in ≤ 128");
    constexpr int SLOT_PLANE = TILE_M * SLICE_K * 2;            // SLICED: one bf16 plane of a 128 × 64 slice (16 KB); slot = hi | lo
    constexpr int THREADS = LY::THREADS;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar0 = sbase + LY::SM_BAR;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    const uint32_t rank = CG == 2 ? cluster_rank() : 0u;
    const int cluster_id = (int)blockIdx.x / CG, nclusters = (int)gridDim.x / CG;
    const int num_groups = (p.num_tiles + CG - 1) / CG;          // tile pairs
    const int my_iters = (num_groups - cluster_id + nclusters - 1) / nclusters;
    const int cpx = x_chunks(p.d_in);
    constexpr int cph = chunks_of(H);
    auto tile_of = [&](int t) { return (int64_t)(cluster_id + t * nclusters) * CG + rank; };

    if (threadIdx.x == 0) {
        for (int s = 0; s < LY::STAGES; ++s) {
            mbar_init(bar(LY::W_FULL + s), 1);
            mbar_init(bar(LY::W_EMPTY + s), 1);
            mbar_init(bar(LY::W_PEER + s), 1);
        }
        mbar_init(bar(LY::U_READY), FUSED ? 1 : CG * NUM_LOADER_WARPS);
        mbar_init(bar(LY::U_FREE), 1);
        mbar_init(bar(LY::U_READY1), CG * NUM_LOADER_WARPS);    // SLICED: second U slot
        mbar_init(bar(LY::U_FREE1), 1);
        mbar_init(bar(LY::H_READY), CG * NW);
        mbar_init(bar(LY::ACC_FULL0), 1);
        mbar_init(bar(LY::ACC_FULL1), 1);
        mbar_init(bar(LY::ACC_FREE0), CG * NW);
        mbar_init(bar(LY::ACC_FREE1), CG * NW);
        if constexpr (FUSED) {
            for (int b = 0; b < 2; ++b) {
                mbar_init(bar(LY::TILE_READY0 + b), NUM_GATHER_WARPS);
                mbar_init(bar(LY::TILE_FREE0 + b), 1);
            }
            mbar_init(bar(LY::U_FULL), 1);
        }
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < H; i += THREADS) {
        reinterpret_cast<float*>(smem + LY::SM_LN)[i] = p.ln_w[i];
        reinterpret_cast<float*>(smem + LY::SM_LN)[H + i] = p.ln_b[i];
    }
    {   // this rank's bias-fold image (read by the tensor core: async proxy)
        const uint4* src = reinterpret_cast<const uint4*>(p.fold + (size_t)rank * LY::FOLD_BYTES);
        uint4* dst = reinterpret_cast<uint4*>(smem + LY::SM_FOLD);
        for (int i = threadIdx.x; i < LY::FOLD_BYTES / 16; i += THREADS) dst[i] = src[i];
        fence_proxy_async();
    }
    if constexpr (CG == 2) {
        cluster_sync_all();                                     // barriers initialised in BOTH CTAs before anyone arrives remotely
        if (warp == 1) tmem_alloc2(sbase + LY::SM_TMEM_PTR, TMEM_COLS);
        tc_fence_before();
        cluster_sync_all();
    } else {
        if (warp == 1) tmem_alloc(sbase + LY::SM_TMEM_PTR, TMEM_COLS);
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + LY::SM_TMEM_PTR);

    // every role branch starts with its warpgroup's setmaxnreg (Lay::REG_*)
    if (warp == 0) {
        // ===================================================== weight producer (this CTA's half of every chunk)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LY::REG_WG0));
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const int nx = NB * cpx;
            const uint8_t* mine = p.packed + (size_t)rank * LY::W_CHUNK;
            for (int t = 0; t < my_iters; ++t) {
                for (int i = 0; i < p.steps; ++i) {
                    auto fetch_chunk = [&](int c) {
                        mbar_wait(bar(LY::W_EMPTY + stage), phase ^ 1);
                        mbar_expect_tx(bar(LY::W_FULL + stage), LY::W_CHUNK);
                        bulk_g2s(sbase + LY::SM_W + stage * LY::W_CHUNK, mine + (size_t)c * (CG * LY::W_CHUNK), LY::W_CHUNK,
                                 bar(LY::W_FULL + stage));
                        if (++stage == LY::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    };
                    if constexpr (SLICED) {
                        // consumption order: per 64-column slice the two X chunks of block 0, then of block 1; then H block0, H block1
                        for (int sl = 0; sl < cpx / 2; ++sl)
                            for (int blk = 0; blk < NB; ++blk)
                                for (int c = 0; c < 2; ++c) fetch_chunk(blk * cpx + 2 * sl + c);
                        if (i > 0)
                            for (int c = nx; c < nx + NB * cph; ++c) fetch_chunk(c);
                        continue;
                    }
                    // consumption order of the MMA issuer: X block0, [H block0], X block1, [H block1]
                    for (int seg = 0; seg < 2 * NB; ++seg) {
                        const bool rec = seg & 1;
                        if (rec && i == 0) continue;
                        const int blk = seg >> 1;
                        const int first = rec ? nx + blk * cph : blk * cpx;
                        const int count = rec ? cph : cpx;
                        for (int c = first; c < first + count; ++c) fetch_chunk(c);
                    }
                }
            }
        }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LY::REG_WG0));
        // FUSED: this warp (in BOTH CTAs) also fetches its CTA's U operand of every step — the 64 KB image the gather warps left
        // in the scratch slot of the tile — with two bulk copies, as soon as the input parts of the previous step have released
        // the U buffer.  gs = global step counter of this CTA.
        auto fetch_u = [&](int t, int i, uint32_t gs) {
            if (i == 0) mbar_wait(bar(LY::TILE_READY0 + (t & 1)), (t >> 1) & 1);      // the gather warps have finished the tile
            mbar_wait(bar(LY::U_FREE), (gs & 1) ^ 1);
            if (lane == 0) {
                asm volatile("fence.proxy.async.global;" ::: "memory");              // generic-proxy stores of the gather warps → bulk copy
                const uint32_t plane = (uint32_t)(p.d_in / 8) * (TILE_M * 16);
                const uint8_t* img = p.scratch + (((size_t)blockIdx.x * 2 + (t & 1)) * p.steps + i) * (size_t)U_IMAGE;
                mbar_expect_tx(bar(LY::U_FULL), 2 * plane);
                bulk_g2s(sbase + LY::SM_U, img, plane, bar(LY::U_FULL));
                bulk_g2s(sbase + LY::SM_U + A_PLANE, img + A_PLANE, plane, bar(LY::U_FULL));
            }
            __syncwarp();
            mbar_wait(bar(LY::U_FULL), gs & 1);
            if (i == p.steps - 1 && lane == 0) mbar_arrive(bar(LY::TILE_FREE0 + (t & 1)));   // the scratch slot may be refilled
        };
        if (CG == 2 && rank != 0) {
            // ===================================================== follower: relay "my half of the chunk has landed"
            uint32_t stage = 0, phase = 0, gs = 0;
            for (int t = 0; t < my_iters; ++t) {
                for (int i = 0; i < p.steps; ++i, ++gs) {
                    if constexpr (FUSED) {
                        fetch_u(t, i, gs);
                        if (lane == 0) mbar_arrive_remote(bar(LY::U_READY), 0);
                        __syncwarp();
                    }
                    const int nch = NB * cpx + (i > 0 ? NB * cph : 0);
                    for (int c = 0; c < nch; ++c) {
                        mbar_wait(bar(LY::W_FULL + stage), phase);
                        if (lane == 0) mbar_arrive_remote(bar(LY::W_PEER + stage), 0);
                        __syncwarp();
                        if (++stage == LY::STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        } else {
            // ===================================================== MMA issuer (all 32 lanes run the control flow, one lane issues)
            constexpr uint32_t idesc192 = umma_idesc_bf16(TILE_M * CG, 192), idesc256 = umma_idesc_bf16(TILE_M * CG, 256);
            constexpr uint32_t A_STEP = (2 * TILE_M * 16) >> 4, B_STEP = (2 * LY::ROWS * 16) >> 4;
            constexpr uint32_t U_LO_PLANE = A_PLANE >> 4, H_LO_PLANE = LY::H_LO >> 4, B_LO_PLANE = LY::W_PLANE >> 4;
            uint32_t stage = 0, phase = 0, gs = 0;
            long long w_wait = 0;                                   // trace: cycles spent waiting for weight chunks in this step
            const uint64_t ones_desc = desc64(desc_lo(sbase + LY::SM_FOLD, TILE_M * 16));
            // biases open the accumulation of the block's whole accumulator set [W_in·x | r | z | W_hn·h]
            auto fold = [&](int blk) {
                if (elect_one())
                    umma_cg<CG>(tmem + (blk & 1) * 256, ones_desc,
                                desc64(desc_lo(sbase + LY::SM_FOLD + LY::FOLD_ONES + blk * LY::FOLD_BIAS, LY::FOLD_ROWS * 16)),
                                idesc256, 0u);
                __syncwarp();
            };
            // the next weight chunk of the ring × the 32 A columns at shared-memory address a_addr (hi plane; lo plane a_lo16·16
            // bytes further): 2 K-steps × 3 split products into the 192 accumulator columns at d
            auto chunk_mmas = [&](uint32_t a_addr, uint32_t a_lo16, uint32_t d) {
                const long long t0 = p.trace ? clock64() : 0;
                mbar_wait(bar(LY::W_FULL + stage), phase);
                if constexpr (CG == 2) mbar_wait_cluster(bar(LY::W_PEER + stage), phase);
                if (p.trace) w_wait += clock64() - t0;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a0 = desc_lo(a_addr, TILE_M * 16);
                    const uint32_t b0 = desc_lo(sbase + LY::SM_W + stage * LY::W_CHUNK, LY::ROWS * 16);
#pragma unroll
                    for (int ks = 0; ks < CHUNK_K / 16; ++ks) {
                        const uint64_t ah = desc64(a0 + ks * A_STEP), al = desc64(a0 + a_lo16 + ks * A_STEP);
                        const uint64_t bh = desc64(b0 + ks * B_STEP), bl = desc64(b0 + B_LO_PLANE + ks * B_STEP);
                        umma_cg<CG>(d, ah, bh, idesc192, 1u);
                        umma_cg<CG>(d, al, bh, idesc192, 1u);
                        umma_cg<CG>(d, ah, bl, idesc192, 1u);
                    }
                    commit_cg<CG>(bar(LY::W_EMPTY + stage));
                }
                __syncwarp();
                if (++stage == LY::STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            // one part = one block (64 hidden features) of the input (A = U) or recurrent (A = h_{i-1}) contribution
            auto run_part = [&](int nchunks, int blk, bool recurrent) {
                const uint32_t d = tmem + (blk & 1) * 256 + (recurrent ? COL_R : COL_IN);
                for (int kc = 0; kc < nchunks; ++kc) {
                    if (!recurrent) chunk_mmas(sbase + LY::SM_U + kc * (CHUNK_K / 8) * (TILE_M * 16), U_LO_PLANE, d);
                    else chunk_mmas(sbase + LY::h_unit(kc * CHUNK_K, 0), H_LO_PLANE, d);
                }
            };
            auto commit = [&](int b) {
                if (elect_one()) commit_cg<CG>(bar(b));
                __syncwarp();
            };
            for (int t = 0; t < my_iters; ++t) {
                for (int i = 0; i < p.steps; ++i, ++gs) {
                    const uint32_t par = gs & 1;
                    if (lane == 0) GRU2_TRACE(0, gs);
                    if constexpr (SLICED) {
                        // slice-major input phase: both accumulator sets are opened, then every 64-column slice of U (staged by
                        // the loaders into alternating slots) feeds the two X chunks of block 0 and of block 1
                        wait_pair<CG>(bar(LY::ACC_FREE0), par ^ 1);
                        wait_pair<CG>(bar(LY::ACC_FREE1), par ^ 1);
                        tc_fence_after();
                        fold(0);
                        fold(1);
                        const int nsl = cpx / 2;
                        for (int sl = 0; sl < nsl; ++sl) {
                            const uint32_t use = (uint32_t)(gs * nsl + sl);      // running slice counter: slot = use & 1, phase = use >> 1
                            wait_pair<CG>(bar((use & 1) ? LY::U_READY1 : LY::U_READY), (use >> 1) & 1);
                            tc_fence_after();
                            const uint32_t slot_addr = sbase + LY::SM_U + (use & 1) * (2 * SLOT_PLANE);
                            for (int blk = 0; blk < NB; ++blk)
                                for (int c = 0; c < 2; ++c)
                                    chunk_mmas(slot_addr + c * (CHUNK_K / 8) * (TILE_M * 16), SLOT_PLANE >> 4, tmem + blk * 256 + COL_IN);
                            commit((use & 1) ? LY::U_FREE1 : LY::U_FREE);
                        }
                        if (lane == 0) GRU2_TRACE(2, gs);
                        if (i > 0) {
                            wait_pair<CG>(bar(LY::H_READY), par ^ 1);
                            tc_fence_after();
                            if (lane == 0) GRU2_TRACE(3, gs);
                            run_part(cph, 0, true);
                        }
                        commit(LY::ACC_FULL0);
                        if (i > 0) run_part(cph, 1, true);
                        commit(LY::ACC_FULL1);
                        if (lane == 0) GRU2_TRACE(7, gs);
                        continue;
                    }
                    if constexpr (FUSED) fetch_u(t, i, gs);
                    wait_pair<CG>(bar(LY::U_READY), par);
                    wait_pair<CG>(bar(LY::ACC_FREE0), par ^ 1);
                    tc_fence_after();
                    if (lane == 0) GRU2_TRACE(1, gs);
                    fold(0);
                    run_part(cpx, 0, false);
                    if (lane == 0) GRU2_TRACE(2, gs);
                    if (i > 0) {
                        // the recurrence h_{i-1} → gates → h_i is the critical chain: block 0's recurrent part goes ahead of
                        // block 1's input part (same chunk order in the producers)
                        wait_pair<CG>(bar(LY::H_READY), par ^ 1);
                        tc_fence_after();
                        if (lane == 0) GRU2_TRACE(3, gs);
                        run_part(cph, 0, true);
                    }
                    commit(LY::ACC_FULL0);
                    if (lane == 0) GRU2_TRACE(4, gs);
                    wait_pair<CG>(bar(LY::ACC_FREE1), par ^ 1);
                    tc_fence_after();
                    if (lane == 0) GRU2_TRACE(5, gs);
                    fold(1);
                    run_part(cpx, 1, false);
                    commit(LY::U_FREE);
                    if (lane == 0) GRU2_TRACE(6, gs);
                    if (i > 0) run_part(cph, 1, true);
                    commit(LY::ACC_FULL1);
                    if (lane == 0) GRU2_TRACE(7, gs);
                    if (lane == 0 && p.trace && blockIdx.x == 0 && gs < 64u) p.trace[15 * 64 + gs] = w_wait;
                    w_wait = 0;
                }
            }
        }
    } else if (warp < FIRST_WORKER_WARP) {
        if constexpr (FUSED) {
            // ===================================================== gather warps (2-7): the cumulative k-core SpMM of the NEXT tile
            // layers.py:41-48:  S_i = S_{i-1} + A_i·x,  U_i = relu(S_i)  over the level-tagged union CSR (spmm.cu has the
            // stand-alone kernel and the derivation).  A warp owns 21-22 consecutive rows of the tile = one contiguous range of CSR
            // entries.  Every entry is ONE feature row of x (≤ 512 B) copied global → shared memory by ONE warp-wide cp.async
            // (LDGSTS: 16 bytes per lane, no destination registers) into a 10-slot ring: 5 KB in flight per warp, 30 KB per CTA,
            // without a single data register — which is what lets an HBM-bound gather run next to the register-hungry gate warps.
            // Entries are handled in batches of 10 (lane j < 10 keeps the metadata of its entry) and groups of 5 (one cp.async group
            // each): after a group is consumed its slots are refilled with the same group of the next batch, so one group is always
            // in flight behind the one being consumed.  A lane only ever reads the 16 bytes it copied itself.
            // Each finished level of a row is stored as the tensor-core operand it will be: bf16 hi / lo planes in core-matrix
            // order (the split the input loaders of the unfused build do), into this CTA's scratch slot (t & 1), image of the
            // level — warp 1 pulls a step's 64 KB image into shared memory with two bulk copies one tile later.  Written and read
            // by the same SM, mostly out of the L2.
            // What was tried first (profiles/r02_experiments.md): one cp.async.bulk per row — the bulk-copy engine keeps only a
            // handful of such small copies in flight per SM; two gather warps with the loop unrolled — 56 KB of code, the kernel
            // 5× slower (instruction cache shared by five warp roles); two warps, compact loop — a warp's serial per-entry control
            // flow takes ≈ 460 cycles per entry, 4× too slow.  Hence six warps and NO unrolling here.
            if (warp < FIRST_LOADER_WARP) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LY::REG_WG0));
            else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LY::REG_LOAD));
            const int gw = warp - FIRST_GATHER_WARP;
            const uint32_t ring = sbase + LY::SM_G + gw * (G_SLOTS * G_SLOT);
            const uint8_t* ring_p = smem + LY::SM_G + gw * (G_SLOTS * G_SLOT);
            int32_t* rp_s = reinterpret_cast<int32_t*>(smem + LY::SM_ROWPTR) + gw * 24;
            const int d = p.d_in, K = p.steps;
            const bool lane_on = 4 * lane < d;                  // this lane's float4 of a feature row exists
            const float* const xlane = p.x + 4 * lane;
            const int grp = lane < G_SLOTS ? lane / G_GROUP : -1;   // the group of this lane's slot (lanes ≥ 10 hold no entry)
            const int lrow0 = gw * TILE_M / NUM_GATHER_WARPS, lrow1 = (gw + 1) * TILE_M / NUM_GATHER_WARPS;   // rows of the tile
            for (int tg = 0; tg < my_iters; ++tg) {
                const int slot = tg & 1;
                if (gw == 0 && lane == 0) GRU2_TRACE(24, (uint32_t)tg);
                if (tg >= 2) mbar_wait(bar(LY::TILE_FREE0 + slot), ((tg >> 1) - 1) & 1);   // warp 1 has fetched all of tile tg − 2
                if (gw == 0 && lane == 0) GRU2_TRACE(25, (uint32_t)tg);
                const int64_t r0 = tile_of(tg) * TILE_M + lrow0;
                int64_t r1 = r0 + (lrow1 - lrow0);
                if (r1 > p.n) r1 = p.n;
                const int nrows = r1 > r0 ? (int)(r1 - r0) : 0;
                // image of level 0, this lane's 16-byte unit of the tile's first row: + level·U_IMAGE + row·16
                uint8_t* const img = p.scratch + ((size_t)blockIdx.x * 2 + slot) * K * (size_t)U_IMAGE + (lane & 1) * A_PLANE +
                                     (lane >> 1) * (TILE_M * 16);
                if (nrows > 0) {
                    __syncwarp();
                    if (lane <= nrows) rp_s[lane] = __ldg(p.rowptr + r0 + lane);
                    __syncwarp();
                    const int E0 = rp_s[0], E1 = rp_s[nrows];
                    const int nbatch = (E1 - E0 + G_SLOTS - 1) / G_SLOTS;
                    // metadata of (batch, lane < 10): entry E0 + 10·batch + lane — column, weight, and level byte | row << 8.  The
                    // row of an entry (binary search in the warp's ≤ 23 row pointers) is found HERE, ten entries at a time, so that
                    // the serial consume loop below has one compare on its fast path.
                    // Two batches of metadata are in flight (n2 → nx → cur): the arrays stream from HBM (≈ 2.4 K cycles measured), one
                    // batch of lead is not enough.  The row search comes BEFORE the loads are issued so that its shared-memory reads
                    // never share a scoreboard with (and wait for) the global loads.
                    int c_nx = 0, l_nx = 0, r_nx = 0, m_cur = 0, c_n2 = 0, l_n2 = 0, r_n2 = 0;
                    float w_nx = 0.f, w_cur = 0.f, w_n2 = 0.f;
                    auto load_meta = [&](int b) {                   // → n2
                        const int e = E0 + G_SLOTS * b + lane;
                        if (b < nbatch && lane < G_SLOTS && e < E1) {
                            int lo = 0, hi = nrows - 1;
#pragma unroll 1
                            while (lo < hi) {               // largest r with rowptr[r] ≤ e (rows without entries are skipped over)
                                const int mid = (lo + hi + 1) >> 1;
                                if (rp_s[mid] <= e) lo = mid;
                                else hi = mid - 1;
                            }
                            r_n2 = lo << 8;
                            c_n2 = __ldg(p.col + e);
                            w_n2 = __ldg(p.val + e);
                            l_n2 = __ldg(p.lvl + e);
                        }
                    };
                    auto shift_meta = [&]() {
                        c_nx = c_n2;
                        w_nx = w_n2;
                        l_nx = l_n2;
                        r_nx = r_n2;
                    };
                    // copy the feature rows of group g of batch b (column indices in *_nx of the group's lanes) into the group's
                    // slots: one cp.async group, possibly empty (the group count per batch stays uniform)
                    auto issue = [&](int b, int g) {
                        const int ebase = E0 + G_SLOTS * b + G_GROUP * g;
                        int cj[G_GROUP];
#pragma unroll
                        for (int i2 = 0; i2 < G_GROUP; ++i2) cj[i2] = __shfl_sync(0xffffffffu, c_nx, G_GROUP * g + i2);
#pragma unroll
                        for (int i2 = 0; i2 < G_GROUP; ++i2)
                            if (ebase + i2 < E1 && lane_on)
                                cp_async16(ring + (G_GROUP * g + i2) * G_SLOT + 16 * lane, xlane + (int64_t)cj[i2] * p.ldx);
                        cp_async_commit();
                        if (grp == g) {
                            w_cur = w_nx;
                            m_cur = l_nx | r_nx;
                        }
                    };
                    float4 P = make_float4(0.f, 0.f, 0.f, 0.f), S = P;
                    int cur = 0, row = 0;
                    uint8_t* urow = img + (size_t)lrow0 * 16;
                    // close levels (S += P; U_level = relu(S) → scratch, split into bf16 hi / lo) and rows until the cursor stands
                    // at (trow, tlev); rows in between (without entries) get K zero levels.  The ONE place that stores.
                    auto flush_to = [&](int trow, int tlev) {
#pragma unroll 1
                        while (true) {
                            const int target = row < trow ? K : tlev;
#pragma unroll 1
                            while (cur < target) {
                                S.x += P.x;
                                S.y += P.y;
                                S.z += P.z;
                                S.w += P.w;
                                // lane pair (2j, 2j+1) holds features 8j .. 8j+7 = one 16-byte unit per plane: the even lane
                                // stores the hi unit, the odd lane the lo unit
                                uint32_t h0, l0, h1, l1;
                                split2(fmaxf(S.x, 0.f), fmaxf(S.y, 0.f), h0, l0);
                                split2(fmaxf(S.z, 0.f), fmaxf(S.w, 0.f), h1, l1);
                                const bool odd = lane & 1;
                                const uint32_t s0 = __shfl_xor_sync(0xffffffffu, odd ? h0 : l0, 1);
                                const uint32_t s1 = __shfl_xor_sync(0xffffffffu, odd ? h1 : l1, 1);
                                if (lane_on)
                                    *reinterpret_cast<uint4*>(urow + (size_t)cur * U_IMAGE) =
                                        odd ? make_uint4(s0, s1, l0, l1) : make_uint4(h0, h1, s0, s1);
                                ++cur;
                            }
                            if (row >= trow) break;
                            ++row;
                            urow += 16;
                            P = S = make_float4(0.f, 0.f, 0.f, 0.f);
                            cur = 0;
                        }
                    };
                    load_meta(0);
                    shift_meta();
                    load_meta(1);
                    if (nbatch > 0) {
#pragma unroll 1
                        for (int g = 0; g < G_GROUPS; ++g) issue(0, g);
                        shift_meta();
                        load_meta(2);
                    }
#pragma unroll 1
                    for (int b = 0; b < nbatch; ++b) {
#pragma unroll 1
                        for (int g = 0; g < G_GROUPS; ++g) {
                            const int ebase = E0 + G_SLOTS * b + G_GROUP * g;
                            cp_async_wait<G_GROUPS - 1>();      // this lane's share of the group has landed
                            int cnt = E1 - ebase;
                            cnt = cnt > G_GROUP ? G_GROUP : cnt;
                            // software pipeline over the group's entries: the next entry's weight / metadata / feature slice are
                            // fetched before the current one is accumulated
                            float wj = 0.f;
                            int mj = 0;
                            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
                            auto fetch = [&](int i2, float& w_o, int& m_o, float4& x_o) {
                                w_o = __shfl_sync(0xffffffffu, w_cur, G_GROUP * g + i2);
                                m_o = __shfl_sync(0xffffffffu, m_cur, G_GROUP * g + i2);
                                if (lane_on) x_o = *reinterpret_cast<const float4*>(ring_p + (G_GROUP * g + i2) * G_SLOT + 16 * lane);
                            };
                            if (cnt > 0) fetch(0, wj, mj, xv);
#pragma unroll 1
                            for (int i2 = 0; i2 < cnt; ++i2) {  // warp-uniform trip count
                                float wn = 0.f;
                                int mn = 0;
                                float4 xn = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (i2 + 1 < cnt) fetch(i2 + 1, wn, mn, xn);
                                const int erow = mj >> 8, elev = mj & 127;
                                if (erow != row || elev != cur) flush_to(erow, elev);      // fast path: same row, same level
                                if (mj & 128) {                     // one-shot entry (present in A_level only): straight into S
                                    S.x = fmaf(wj, xv.x, S.x);
                                    S.y = fmaf(wj, xv.y, S.y);
                                    S.z = fmaf(wj, xv.z, S.z);
                                    S.w = fmaf(wj, xv.w, S.w);
                                } else {
                                    P.x = fmaf(wj, xv.x, P.x);
                                    P.y = fmaf(wj, xv.y, P.y);
                                    P.z = fmaf(wj, xv.z, P.z);
                                    P.w = fmaf(wj, xv.w, P.w);
                                }
                                wj = wn;
                                mj = mn;
                                xv = xn;
                            }
                            // the group's slots are free (every lane has read its own 16 bytes): refill them with the same group
                            // of the next batch; past the end an empty group keeps "one group behind" true
                            if (b + 1 < nbatch) issue(b + 1, g);
                            else cp_async_commit();
                        }
                        shift_meta();                           // batch b + 2 (loaded one iteration ago) becomes "next"
                        load_meta(b + 3);
                    }
                    flush_to(nrows - 1, K);                     // close the last rows (and rows without entries)
                    cp_async_wait<0>();
                }
                // this warp's rows of the tile are in the scratch slot: hand them to warp 1 (bulk copies read them: async proxy)
                asm volatile("fence.proxy.async.global;" ::: "memory");
                __syncwarp();
                if (gw == 0 && lane == 0) GRU2_TRACE(26, (uint32_t)tg);
                if (lane == 0) mbar_arrive(bar(LY::TILE_READY0 + slot));
            }
        } else if (warp < FIRST_LOADER_WARP) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LY::REG_WG0));   // warps 2-3 idle
        } else {
            // ===================================================== input loaders (warps 4-7)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LY::REG_LOAD));
            // A warp owns 32 tile rows.  Per load instruction its lanes cover 8 rows × 4 k-blocks (r = lane%8, c = lane/8):
            // 128 contiguous bytes per row (8 L1 wavefronts instead of 32 for a row-per-lane mapping) and the 16-byte
            // shared-memory stores of one 8-lane phase hit 8 consecutive rows of one k-block (conflict-free).
            const int r8 = lane & 7, c4 = lane >> 3;
            const int row_base = 32 * (warp - FIRST_LOADER_WARP);
            uint8_t* u_hi = smem + LY::SM_U;
            const int nkg = p.d_in / 32;             // k-groups of 4 k-blocks
            const int nit = 4 * nkg;                 // (row-group, k-group) iterations per step: 16 for d_in = 128
            uint32_t gs = 0;
            // U comes from HBM (written by the SpMM, far larger than the L2): the rows of the NEXT step are pulled into the L2 one
            // whole step ahead, so that the loads issued around "U buffer free" see L2 latency instead of DRAM latency
            auto prefetch_step = [&](int t, int i) {
                if (i >= p.steps) {
                    i = 0;
                    if (++t >= my_iters) return;
                }
                const int64_t tile_row0 = tile_of(t) * TILE_M;
                const float* base = p.seq + (int64_t)i * p.sss;
                const int lines_per_row = (p.d_in + 31) / 32;               // 128-byte lines
                for (int l = lane; l < 32 * lines_per_row; l += 32) {       // this warp's 32 rows
                    const int64_t srow = tile_row0 + row_base + l / lines_per_row;
                    if (srow < p.n) prefetch_l2(base + srow * p.srs + (l % lines_per_row) * 32);
                }
            };
            prefetch_step(0, 0);
            if constexpr (SLICED) {
                // 64 input columns at a time into alternating 32 KB slots ([hi plane 8 k-blocks][lo plane]); columns ≥ d_in are zeros
                uint32_t use = 0;
                const int nsl = cpx / 2;
                for (int t = 0; t < my_iters; ++t) {
                    const int64_t tile_row0 = tile_of(t) * TILE_M;
                    for (int i = 0; i < p.steps; ++i) {
                        const float* base = p.seq + (int64_t)i * p.sss;
                        prefetch_step(t, i + 1);
                        for (int sl = 0; sl < nsl; ++sl, ++use) {
                            float4 v[16];
#pragma unroll
                            for (int u = 0; u < 8; ++u) {       // (row-group of 8, k-group of 4 k-blocks): 4 × 2 iterations
                                const int rg = u & 3, kg = u >> 2;
                                const int64_t srow = tile_row0 + row_base + 8 * rg + r8;
                                const int c0 = sl * SLICE_K + (4 * kg + c4) * 8;
                                const float* src = base + srow * p.srs + c0;
                                const bool ok = srow < p.n;
                                v[2 * u] = ok && c0 + 4 <= p.d_in ? __ldg(reinterpret_cast<const float4*>(src)) : make_float4(0.f, 0.f, 0.f, 0.f);
                                v[2 * u + 1] = ok && c0 + 8 <= p.d_in ? __ldg(reinterpret_cast<const float4*>(src + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                            mbar_wait(bar((use & 1) ? LY::U_FREE1 : LY::U_FREE), ((use >> 1) & 1) ^ 1);
                            uint8_t* slot = u_hi + (use & 1) * (2 * SLOT_PLANE);
#pragma unroll
                            for (int u = 0; u < 8; ++u) {
                                const int rg = u & 3, kg = u >> 2;
                                const int m = row_base + 8 * rg + r8, kb = 4 * kg + c4;
                                const float f8[8] = {v[2 * u].x, v[2 * u].y, v[2 * u].z, v[2 * u].w,
                                                     v[2 * u + 1].x, v[2 * u + 1].y, v[2 * u + 1].z, v[2 * u + 1].w};
                                uint4 hi, lo;
                                split8(f8, hi, lo);
                                *reinterpret_cast<uint4*>(slot + kb * (TILE_M * 16) + m * 16) = hi;
                                *reinterpret_cast<uint4*>(slot + SLOT_PLANE + kb * (TILE_M * 16) + m * 16) = lo;
                            }
                            fence_proxy_async();
                            __syncwarp();
                            if (lane == 0) arrive_leader<CG>(bar((use & 1) ? LY::U_READY1 : LY::U_READY), rank);
                        }
                    }
                }
            } else
            for (int t = 0; t < my_iters; ++t) {
                const int64_t tile_row0 = tile_of(t) * TILE_M;
                for (int i = 0; i < p.steps; ++i, ++gs) {
                    const float* base = p.seq + (int64_t)i * p.sss;
                    // A quarter-warp reads 128 contiguous bytes of ONE row (lane = 16-byte piece p8 of row q4 of a 4-row group): 8 L1
                    // data-pipe wavefronts per load instead of 32 for the former "8 rows per quarter" mapping; a lane then holds half
                    // an operand unit and stores it with 64-bit stores.  An "iteration" = two such loads (the lane's float4 pair).
                    float4 v[16];
                    const int q4 = lane >> 3, p8 = lane & 7;
                    auto load_batch = [&](int it0, int cnt) {   // up to 8 iterations = 16 LDG.128 in flight per lane
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const int e = 2 * it0 + u, rgrp = e & 7, span = e >> 3;      // (row group of 4: 8 per warp, 32-column span)
                            const int64_t srow = tile_row0 + row_base + 4 * rgrp + q4;
                            v[u] = (u < 2 * cnt && srow < p.n) ? __ldg(reinterpret_cast<const float4*>(base + srow * p.srs + 32 * span + 4 * p8))
                                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    };
                    auto store_batch = [&](int it0, int cnt) {  // fp32 → bf16 hi/lo planes, 4 k-elements (8 B) per store
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            if (u < 2 * cnt) {
                                const int e = 2 * it0 + u, rgrp = e & 7, span = e >> 3;
                                const int m = row_base + 4 * rgrp + q4, kb = 4 * span + (p8 >> 1);
                                uint2 hi, lo;
                                split2(v[u].x, v[u].y, hi.x, lo.x);
                                split2(v[u].z, v[u].w, hi.y, lo.y);
                                uint8_t* dst = u_hi + kb * (TILE_M * 16) + m * 16 + 8 * (p8 & 1);
                                *reinterpret_cast<uint2*>(dst) = hi;
                                *reinterpret_cast<uint2*>(dst + A_PLANE) = lo;
                            }
                        }
                    };
                    const int first = nit < 8 ? nit : 8;
                    load_batch(0, first);   // global loads are issued BEFORE the buffer is free: their latency is off the loop
                    prefetch_step(t, i + 1);
                    // the input parts of the previous step's MMAs must have released the single U buffer
                    mbar_wait(bar(LY::U_FREE), (gs & 1) ^ 1);
                    if (threadIdx.x == FIRST_LOADER_WARP * 32) GRU2_TRACE(13, gs);
                    store_batch(0, first);
                    for (int it0 = 8; it0 < nit; it0 += 8) {
                        const int cnt = nit - it0 < 8 ? nit - it0 : 8;
                        load_batch(it0, cnt);
                        store_batch(it0, cnt);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) arrive_leader<CG>(bar(LY::U_READY), rank);
                    if (threadIdx.x == FIRST_LOADER_WARP * 32) GRU2_TRACE(14, gs);
                }
            }
        }
    } else {
        // ===================================================== gate warps: gate math, h, Σh, LayerNorm
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LY::REG_GATE));
        constexpr int CHW = NW / 4;          // warps sharing one TMEM lane quarter: partial row sums to exchange
        constexpr int FPT = BLK / CHW;       // features per thread and block: 32
        constexpr int NACC = 2 * FPT;        // features per thread
        const int ww = warp - FIRST_WORKER_WARP;
        const int q = warp & 3;              // TMEM lane quarter this warp may access
        const int ch = ww >> 2;              // which FPT of a block's 64 features this thread owns
        const int m = 32 * q + lane;         // row inside the tile
        const uint32_t tmem_lane = tmem + ((uint32_t)(32 * q) << 16);
        const float* lnw = reinterpret_cast<const float*>(smem + LY::SM_LN);
        float* red = reinterpret_cast<float*>(smem + LY::SM_RED);
        uint32_t gs = 0;
        // feature of this thread's value j (0 ≤ j < NACC)
        auto feat = [&](int j) { return (j / FPT) * BLK + ch * FPT + j % FPT; };

        // row statistics over 128 values: this thread holds NACC of them, CHW−1 other warps of its lane quarter the rest
        auto row_stats = [&](const float (&v)[NACC], float& mean, float& rstd) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NACC; ++j) s += v[j];
            red[ch * TILE_M + m] = s;
            asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
            float tot = 0.f;
#pragma unroll
            for (int c = 0; c < CHW; ++c) tot += red[c * TILE_M + m];     // same order in every thread of the row
            mean = tot * (1.f / H);
            float sq = 0.f;
#pragma unroll
            for (int j = 0; j < NACC; ++j) {
                const float dlt = v[j] - mean;
                sq = fmaf(dlt, dlt, sq);
            }
            red[(CHW + ch) * TILE_M + m] = sq;
            asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
            float totq = 0.f;
#pragma unroll
            for (int c = 0; c < CHW; ++c) totq += red[(CHW + c) * TILE_M + m];
            rstd = rsqrtf(totq * (1.f / H) + p.eps);
        };
        // EACH_LN: LayerNorm(h_s) of this step straight to y (row-per-thread 16-byte stores)
        auto layer_norm_store = [&](const float (&v)[NACC], float* dst_row, bool valid) {
            float mean, rstd;
            row_stats(v, mean, rstd);
            if (valid) {
#pragma unroll
                for (int j4 = 0; j4 < NACC; j4 += 4) {
                    const int f = feat(j4);
                    float4 o;
                    o.x = (v[j4 + 0] - mean) * rstd * lnw[f + 0] + lnw[H + f + 0];
                    o.y = (v[j4 + 1] - mean) * rstd * lnw[f + 1] + lnw[H + f + 1];
                    o.z = (v[j4 + 2] - mean) * rstd * lnw[f + 2] + lnw[H + f + 2];
                    o.w = (v[j4 + 3] - mean) * rstd * lnw[f + 3] + lnw[H + f + 3];
                    *reinterpret_cast<float4*>(dst_row + f) = o;
                }
            }
        };
        // SUM_LN result of a tile: normalised rows are staged in the (now idle) h buffer with a 16-byte XOR swizzle and
        // written out one whole 512-byte row per warp instruction — to y, or straight into the owning node slice's
        // (peer) buffer: NVLink wants full-line stores, not 32 scattered 16-byte pieces per instruction.
        auto layer_norm_store_rows = [&](const float (&v)[NACC], int64_t tile_row0) {
            float mean, rstd;
            row_stats(v, mean, rstd);
            float* stage = reinterpret_cast<float*>(smem + LY::SM_H);   // 64 KB: [128 rows][32 chunks of 4 floats], chunk c of row r at c ^ (r & 31)
#pragma unroll
            for (int j4 = 0; j4 < NACC; j4 += 4) {
                const int f = feat(j4);
                float4 o;
                o.x = (v[j4 + 0] - mean) * rstd * lnw[f + 0] + lnw[H + f + 0];
                o.y = (v[j4 + 1] - mean) * rstd * lnw[f + 1] + lnw[H + f + 1];
                o.z = (v[j4 + 2] - mean) * rstd * lnw[f + 2] + lnw[H + f + 2];
                o.w = (v[j4 + 3] - mean) * rstd * lnw[f + 3] + lnw[H + f + 3];
                *reinterpret_cast<float4*>(stage + m * H + (((f >> 2) ^ (m & 31)) << 2)) = o;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
            for (int rr = 0; rr < TILE_M / NW; ++rr) {
                const int r = ww * (TILE_M / NW) + rr;
                const int64_t grow = tile_row0 + r;
                if (grow >= p.n) break;   // warp-uniform
                const float4 o = *reinterpret_cast<const float4*>(stage + r * H + ((lane ^ (r & 31)) << 2));
                float* dst = p.sc.slices ? p.sc.row_ptr(grow) : p.y + grow * p.yrs;
                *reinterpret_cast<float4*>(dst + 4 * lane) = o;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");   // the staging area is h again from the next tile's first step on
        };
        auto put_h8 = [&](const float (&f8)[8], int f) {
            uint4 hi, lo;
            split8(f8, hi, lo);
            uint8_t* dst = smem + LY::h_unit(f, m);
            *reinterpret_cast<uint4*>(dst) = hi;
            *reinterpret_cast<uint4*>(dst + LY::H_LO) = lo;
        };

        for (int t = 0; t < my_iters; ++t) {
            const int64_t row = tile_of(t) * TILE_M + m;
            const bool valid = row < p.n;
            float acc_out[NACC];   // Σ_s h_s (SUM_LN) / h_s of the current step (EACH_LN) for this thread's features
#pragma unroll
            for (int j = 0; j < NACC; ++j) acc_out[j] = 0.f;

            for (int i = 0; i < p.steps; ++i, ++gs) {
                const uint32_t par = gs & 1;
                float h0[FPT];   // block 0 of h_i, published only when no MMA reads h_{i-1} any more
#pragma unroll
                for (int hf = 0; hf < NB; ++hf) {
                    if (lane == 0 && q == 0 && ch == 0) GRU2_TRACE(8 + 2 * hf, gs);
                    mbar_wait(bar(LY::ACC_FULL0 + hf), par);
                    tc_fence_after();
                    if (lane == 0 && q == 0 && ch == 0) GRU2_TRACE(9 + 2 * hf, gs);
#pragma unroll
                    for (int sub = 0; sub < FPT / 8; ++sub) {
                        const int f0 = hf * BLK + ch * FPT + sub * 8;           // first of 8 features (one 16-byte operand unit)
                        const uint32_t col = hf * 256 + ch * FPT + sub * 8;     // + gate block
                        float hold[8], hn8[8];
                        if (i > 0) {
                            const uint8_t* src = smem + LY::h_unit(f0, m);
                            const uint4 hi = *reinterpret_cast<const uint4*>(src);
                            const uint4 lo = *reinterpret_cast<const uint4*>(src + LY::H_LO);
                            join8(hi, lo, hold);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) hold[j] = 0.f;
                        }
                        float ea[8], eb[8], gi[8], gh[8];
                        tmem_ld8(tmem_lane + col + COL_R, ea);
                        tmem_ld8(tmem_lane + col + COL_Z, eb);
                        tmem_ld8(tmem_lane + col + COL_IN, gi);
                        tmem_ld8(tmem_lane + col + COL_HN, gh);                 // step 0: b_hn alone (bias fold, no recurrent part)
                        tmem_ld_wait();
                        gate_math<8>(ea, eb, gi, gh, hold, hn8);
                        if (hf == 0) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) h0[sub * 8 + j] = hn8[j];
                        } else {
                            // every MMA that reads h_{i-1} has completed (acc1 full): h may be overwritten in place.  The held-back
                            // block 0 is published piecewise here so that its ALU work hides under the MUFU latency of this pass.
                            const float f8[8] = {h0[sub * 8], h0[sub * 8 + 1], h0[sub * 8 + 2], h0[sub * 8 + 3],
                                                 h0[sub * 8 + 4], h0[sub * 8 + 5], h0[sub * 8 + 6], h0[sub * 8 + 7]};
                            put_h8(f8, ch * FPT + sub * 8);
                            put_h8(hn8, f0);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int a = hf * FPT + sub * 8 + j;
                            if (MODE == CTGCN_GRU_SUM_LN) acc_out[a] += hn8[j];
                            else acc_out[a] = hn8[j];
                        }
                        if (lane == 0 && q == 0 && ch == 0 && sub < 4) GRU2_TRACE(16 + hf * 4 + sub, gs);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_leader<CG>(bar(LY::ACC_FREE0 + hf), rank);
                }
                // this warp's share of h_i is complete in shared memory (all shares: the next step's recurrent MMAs may read it)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) arrive_leader<CG>(bar(LY::H_READY), rank);
                if (lane == 0 && q == 0 && ch == 0) GRU2_TRACE(12, gs);
                if (MODE == CTGCN_GRU_EACH_LN) layer_norm_store(acc_out, p.y + row * p.yrs + (int64_t)i * p.yss, valid);
            }
            if (MODE == CTGCN_GRU_SUM_LN) layer_norm_store_rows(acc_out, row - m);
        }
    }

    tc_fence_before();
    if constexpr (CG == 2) {
        cluster_sync_all();      // neither CTA may exit (or free TMEM) while the pair's MMAs can still touch its memory
        if (warp == 1) tmem_dealloc2(tmem, TMEM_COLS);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
    }
}

// CTAs of a launch over n rows: one cluster of CG per CG tiles, at most one CTA per SM
int grid_ctas(int64_t n, int cg, int sm_count) {
    const int64_t groups = ((n + TILE_M - 1) / TILE_M + cg - 1) / cg;
    const int64_t max_clusters = sm_count / cg;
    return (int)(groups < max_clusters ? groups : max_clusters) * cg;
}
// per-CTA scratch of the fused build: [CTAs][2 tile slots][k][64 KB operand image]
size_t fused_scratch_bytes(int ctas, int k, int) { return (size_t)ctas * 2 * k * U_IMAGE; }

template <int CG>
size_t packed_bytes(int d_in) {
    return (size_t)(NB * x_chunks(d_in) + NB * chunks_of(H)) * CG * Lay<CG>::W_CHUNK;
}

template <int CG, bool FUSED = false, bool SLICED = false>
int launch_variant(const Params2& p0, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int mode, void* ws,
                   size_t ws_bytes, cudaStream_t st) {
    using LY = Lay<CG, FUSED>;
    Params2 p = p0;
    const size_t pb = packed_bytes<CG>(p.d_in), fb = (size_t)CG * LY::FOLD_BYTES;
    int dev = 0, sm_count = 0;
    CTGCN_CUDA_OK(cudaGetDevice(&dev));
    CTGCN_CUDA_OK(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    const size_t sb = FUSED ? fused_scratch_bytes(grid_ctas(p.n, CG, sm_count), p.steps, p.d_in) : 0;
    CTGCN_REQUIRE(ws && ws_bytes >= align_up(pb + fb, 256) + sb, "gru_tc2: workspace too small (%zu < %zu)", ws_bytes,
                  align_up(pb + fb, 256) + sb);
    uint8_t* packed = (uint8_t*)ws;
    uint8_t* fold = packed + pb;
    if (FUSED) p.scratch = (uint8_t*)ws + align_up(pb + fb, 256);
    {
        ProfScope prof(PROF_PACK, st);
        const int nchunks = NB * x_chunks(p.d_in) + NB * chunks_of(H);
        int threads = nchunks * GATE_ROWS * (CHUNK_K / 8);
        const int fold_threads = CG * LY::FOLD_BYTES / 16;
        if (fold_threads > threads) threads = fold_threads;
        pack2_kernel<CG><<<(threads + 255) / 256, 256, 0, st>>>(w_ih, w_hh, b_ih, b_hh, p.d_in, packed, fold);
        CTGCN_LAUNCH_OK("pack2_kernel");
    }
    p.packed = packed;
    p.fold = fold;
    void (*kern)(const Params2);
    if constexpr (FUSED) {
        kern = gru2_kernel<CG, CTGCN_GRU_SUM_LN, true>;
    } else if constexpr (SLICED) {
        CTGCN_REQUIRE(mode == CTGCN_GRU_SUM_LN, "gru_tc2: inputs wider than %d are handled for the core-axis GRU only", H);
        kern = gru2_kernel<CG, CTGCN_GRU_SUM_LN, false, true>;
    } else {
        kern = mode == CTGCN_GRU_SUM_LN ? gru2_kernel<CG, CTGCN_GRU_SUM_LN, false> : gru2_kernel<CG, CTGCN_GRU_EACH_LN, false>;
    }
    CTGCN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LY::SMEM_BYTES));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid_ctas(p.n, CG, sm_count));
    cfg.blockDim = dim3(LY::THREADS);
    cfg.dynamicSmemBytes = LY::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ProfScope prof(PROF_GRU, st);
    CTGCN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, p));
    CTGCN_LAUNCH_OK("gru2_kernel");
    return CTGCN_OK;
}

}  // namespace

static long long* g_gru2_trace = nullptr;
void set_gru2_trace(long long* buf) { g_gru2_trace = buf; }

size_t gru_tc2_workspace_bytes(int d_in) {
    // the larger of the two builds (whole chunks: same weight bytes; fold images: 2 × 12 KB vs 20 KB)
    return align_up(packed_bytes<2>(d_in) + 2 * (size_t)Lay<2>::FOLD_BYTES + (size_t)Lay<1>::FOLD_BYTES, 256);
}
bool gru_tc2_takes(int d_in, int h) { return h == H && d_in >= 32 && (d_in <= H ? d_in % 32 == 0 : (d_in <= 1024 && d_in % 4 == 0)); }

// returns 0 = done, <0 = error, 1 = shape not supported by this path.  cg: CTAs per MMA (2 = CTA pairs, 1 = the A/B build without pairing)
int launch_gru_tc2(int cg, const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* w_ih,
                   const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                   int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* sc, void* ws, size_t ws_bytes, cudaStream_t st) {
    const bool sliced = d_in > H;                                // slice-major input phase (core-axis GRU, pairs only)
    if (h != H || d_in < 32 || (sliced ? (d_in > 1024 || (d_in & 3) || cg != 2 || mode != CTGCN_GRU_SUM_LN) : (d_in % 32) != 0)) return 1;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(seq) || (srs & 3) || (sss & 3)) return 1;
    if (sc ? ((sc->row_stride & 3) || (sc->col_offset & 3)) : (!al16(y) || (yrs & 3) || (yss & 3))) return 1;
    Params2 p;
    p.seq = seq;
    p.srs = srs;
    p.sss = sss;
    p.n = n;
    p.steps = steps;
    p.d_in = d_in;
    p.packed = nullptr;
    p.fold = nullptr;
    p.ln_w = ln_w;
    p.ln_b = ln_b;
    p.eps = eps;
    p.y = y;
    p.yrs = yrs;
    p.yss = yss;
    p.sc = sc ? *sc : RowScatter();
    p.num_tiles = (int)((n + TILE_M - 1) / TILE_M);
    p.trace = g_gru2_trace;
    p.rowptr = nullptr;
    p.col = nullptr;
    p.val = nullptr;
    p.lvl = nullptr;
    p.x = nullptr;
    p.ldx = 0;
    p.scratch = nullptr;
    if (sliced) return launch_variant<2, false, true>(p, w_ih, w_hh, b_ih, b_hh, mode, ws, ws_bytes, st);
    if (cg == 2) return launch_variant<2>(p, w_ih, w_hh, b_ih, b_hh, mode, ws, ws_bytes, st);
    return launch_variant<1>(p, w_ih, w_hh, b_ih, b_hh, mode, ws, ws_bytes, st);
}

// CoreDiffusion.forward (layers.py:38-63) in ONE launch: the cumulative SpMM runs inside the GRU kernel (gather warps).
// Workspace: core_diffusion_fused_workspace_bytes.  returns 0 = done, <0 = error, 1 = shape / graph not supported by this path
size_t core_diffusion_fused_workspace_bytes(int64_t n, int k, int d_in) {
    int dev = 0, sm_count = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        sm_count = 256;
    return align_up(packed_bytes<2>(d_in) + 2 * (size_t)Lay<2, true>::FOLD_BYTES, 256) +
           fused_scratch_bytes(grid_ctas(n, 2, sm_count), k, d_in);
}

int launch_core_diffusion_fused(const ctgcn_plan* plan, int64_t row0, int64_t rows, const float* x, int64_t ldx, int d_in, int h,
                                const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* ln_w,
                                const float* ln_b, float eps, float* y, int64_t yrs, const RowScatter* sc, void* ws, size_t ws_bytes,
                                cudaStream_t st) {
    if (h != H || d_in < 32 || d_in > 128 || (d_in % 32) || plan->k < 1 || plan->k > 64) return 1;
    if (row0 != 0 || rows != plan->n_rows) return 1;             // row chunks: the two-kernel path
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(x) || (ldx & 3)) return 1;
    if (sc ? ((sc->row_stride & 3) || (sc->col_offset & 3)) : (!al16(y) || (yrs & 3))) return 1;
    // a warp walks its 64 rows' entries serially: hub rows (power-law graphs) would stall the tile — two-kernel path + row splitting
    if (plan->max_row_entries > 2048) return 1;
    Params2 p;
    p.seq = nullptr;
    p.srs = p.sss = 0;
    p.n = plan->n_rows;
    p.steps = plan->k;
    p.d_in = d_in;
    p.packed = nullptr;
    p.fold = nullptr;
    p.ln_w = ln_w;
    p.ln_b = ln_b;
    p.eps = eps;
    p.y = y;
    p.yrs = yrs;
    p.yss = 0;
    p.sc = sc ? *sc : RowScatter();
    p.num_tiles = (int)((plan->n_rows + TILE_M - 1) / TILE_M);
    p.trace = g_gru2_trace;
    p.rowptr = plan->rowptr;
    p.col = plan->col;
    p.val = plan->val;
    p.lvl = plan->lvl;
    p.x = x;
    p.ldx = ldx;
    p.scratch = nullptr;
    return launch_variant<2, true>(p, w_ih, w_hh, b_ih, b_hh, CTGCN_GRU_SUM_LN, ws, ws_bytes, st);
}

}  // namespace ctgcn
