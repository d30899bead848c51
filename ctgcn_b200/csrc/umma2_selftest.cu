// Building-block self test (green on a B200 since round 2): one GRU half-step through `tcgen05.mma.cta_group::2`.
//
// profiles/r02_gru_design.md, step 2: a CTA pair shares every weight chunk — each CTA holds HALF of the B rows, the leader issues
// M = 256 MMAs over both CTAs' 128-row A tiles, accumulators stay per CTA (128 TMEM lanes each).  This kernel exercises exactly
// the mechanisms the pipelined kernel will need, on the arithmetic of umma_selftest_kernel (gru_tc.cu):
//   out[256 × 256] = [ x·W_inᵀ | x·W_irᵀ + h·W_hrᵀ | x·W_izᵀ + h·W_hzᵀ | h·W_hnᵀ ]   hidden features 0..63,
//   x [256, 64], h [256, 128]; rows 0..127 belong to CTA 0, rows 128..255 to CTA 1.
// Mechanisms: cluster (2,1,1) launch; tcgen05.alloc/dealloc.cta_group::2 by the same warp of both CTAs; per-CTA half chunks
// fetched with cp.async.bulk onto a LOCAL mbarrier and relayed to the leader with a remote mbarrier arrive (a non-tensor bulk
// copy cannot complete_tx on the peer's barrier); M = 256 instruction descriptors; the recurrent part as an N = 128 (r|z,
// accumulate) and an N = 64 (W_hn·h, fresh) stream, because the N = 192 / 128 / 64 forms split B differently across the pair;
// tcgen05.commit…multicast::cluster to both CTAs.
//
// Packed weight layout ("pair chunks", built by pack_pair_kernel): chunk order as in gru_tc.cu (X half0, X half1, H half0, H half1
// by K = 32); per chunk CTA 0's 12 KB then CTA 1's 12 KB; per CTA bf16 hi plane (6 KB) then lo plane.
//   X chunk: the CTA's 96 of the 192 rows [n | r | z]:  (k/8)·1536 + row·16 + (k%8)·2
//   H chunk: 64 rows of ONE gate (CTA 0: r, CTA 1: z) at (k/8)·1024 + row·16, then 32 rows of W_hn (features 32·rank …) at
//            4096 + (k/8)·512 + row·16.
#include "common.cuh"
#include "tc_common.cuh"

namespace ctgcn {
namespace {

using namespace tc;

constexpr int H = 128, TILE_M = 128, CHUNK_K = 32;
constexpr int A_PLANE = TILE_M * H * 2;          // 32 KB: one bf16 plane of a 128 × 128 A tile
constexpr int HALF_PLANE = 96 * CHUNK_K * 2;     // 6 KB
constexpr int HALF_BYTES = 2 * HALF_PLANE;       // 12 KB: what one CTA holds of a chunk
constexpr int PAIR_CHUNK = 2 * HALF_BYTES;

__global__ void pack_pair_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, int d_in,
                                 uint8_t* __restrict__ packed) {
    const int cx = d_in / CHUNK_K, chh = H / CHUNK_K;
    const int nchunks = 2 * cx + 2 * chh;
    constexpr int UNITS = 2 * 96 * (CHUNK_K / 8);           // 16-byte units of one plane of a pair chunk (both CTAs)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nchunks * UNITS) return;
    const int c = t / UNITS, u = t % UNITS;
    const int rank = u / (96 * 4), v = u % (96 * 4);
    const bool is_x = c < 2 * cx;
    const int cc = is_x ? c : c - 2 * cx, per = is_x ? cx : chh, ktot = is_x ? d_in : H;
    const int half = cc / per, kc = cc % per;
    int gate, f, kb;
    size_t off;
    if (is_x) {                                             // 96 rows of [n | r | z]
        kb = v / 96;
        const int row = 96 * rank + v % 96, blk = row / 64;
        gate = blk == 0 ? 2 : blk - 1;
        f = row % 64;
        off = (size_t)kb * 1536 + (v % 96) * 16;
    } else if (v < 64 * 4) {                                // 64 rows of r (CTA 0) or z (CTA 1)
        kb = v / 64;
        gate = rank;
        f = v % 64;
        off = (size_t)kb * 1024 + f * 16;
    } else {                                                // 32 rows of W_hn
        const int w = v - 64 * 4;
        kb = w / 32;
        gate = 2;
        f = 32 * rank + w % 32;
        off = 4096 + (size_t)kb * 512 + (w % 32) * 16;
    }
    const float* src = (is_x ? w_ih : w_hh) + (int64_t)(gate * H + half * 64 + f) * ktot + kc * CHUNK_K + kb * 8;
    float x8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x8[i] = src[i];
    uint4 hi, lo;
    split8(x8, hi, lo);
    uint8_t* dst = packed + (size_t)c * PAIR_CHUNK + (size_t)rank * HALF_BYTES + off;
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + HALF_PLANE) = lo;
}

// split products hi·hi + lo·hi + hi·lo of one K = 16 step; a_lo32 / b_lo32: descriptor low words of the hi planes
__device__ __forceinline__ void issue3(uint32_t d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, bool fresh) {
    constexpr uint32_t A_LO = A_PLANE >> 4, B_LO = HALF_PLANE >> 4;
    umma2_bf16(d, desc64(a_lo32), desc64(b_lo32), idesc, fresh ? 0u : 1u);
    umma2_bf16(d, desc64(a_lo32 + A_LO), desc64(b_lo32), idesc, 1u);
    umma2_bf16(d, desc64(a_lo32), desc64(b_lo32 + B_LO), idesc, 1u);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    umma2_selftest_kernel(const float* __restrict__ x, const float* __restrict__ h, const uint8_t* __restrict__ packed,
                          float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t su = sbase, sh = sbase + 2 * A_PLANE, sw = sbase + 4 * A_PLANE;
    const uint32_t bar_w = sw + HALF_BYTES, bar_peer = bar_w + 8, bar_d = bar_peer + 8, tptr = bar_d + 8;
    const int warp = threadIdx.x >> 5, m = threadIdx.x;
    const uint32_t rank = cluster_rank();
    const int row0 = 128 * (int)rank;
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_peer, 1);
        mbar_init(bar_d, 1);
        fence_barrier_init();
    }
    for (int kb = 0; kb < 16; ++kb) {                       // this CTA's A tiles: x [128, 64] and h [128, 128] as bf16 hi|lo planes
        float f8[8];
        uint4 hi, lo;
        if (kb < 8) {
#pragma unroll
            for (int e = 0; e < 8; ++e) f8[e] = x[(row0 + m) * 64 + kb * 8 + e];
            split8(f8, hi, lo);
            *reinterpret_cast<uint4*>(smem + kb * (TILE_M * 16) + m * 16) = hi;
            *reinterpret_cast<uint4*>(smem + A_PLANE + kb * (TILE_M * 16) + m * 16) = lo;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) f8[e] = h[(row0 + m) * 128 + kb * 8 + e];
        split8(f8, hi, lo);
        *reinterpret_cast<uint4*>(smem + 2 * A_PLANE + kb * (TILE_M * 16) + m * 16) = hi;
        *reinterpret_cast<uint4*>(smem + 3 * A_PLANE + kb * (TILE_M * 16) + m * 16) = lo;
    }
    fence_proxy_async();
    cluster_sync_all();                                     // barriers initialised and A tiles staged in BOTH CTAs
    if (warp == 0) tmem_alloc2(tptr, 256);                  // the same warp of both CTAs
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 4 * A_PLANE + HALF_BYTES + 24);

    if (warp == 0) {
        // packed order for d_in = 64: X half0 = chunks 0,1; X half1 = 2,3; H half0 = 4..7
        const int order[6] = {0, 1, 4, 5, 6, 7};
        constexpr uint32_t i192 = umma_idesc_bf16(256, 192), i128 = umma_idesc_bf16(256, 128), i64 = umma_idesc_bf16(256, 64);
        uint32_t par = 0;
        for (int j = 0; j < 6; ++j) {
            if (elect_one()) {                              // every CTA fetches ITS half of the chunk
                mbar_expect_tx(bar_w, HALF_BYTES);
                bulk_g2s(sw, packed + (size_t)order[j] * PAIR_CHUNK + (size_t)rank * HALF_BYTES, HALF_BYTES, bar_w);
            }
            __syncwarp();
            mbar_wait(bar_w, par);
            if (rank == 1) {
                if (elect_one()) mbar_arrive_remote(bar_peer, 0);   // relay: the follower's half has landed
                __syncwarp();
            } else {
                mbar_wait_cluster(bar_peer, par);
                tc_fence_after();
                if (elect_one()) {
                    const bool rec = j >= 2;
                    const int kc = rec ? j - 2 : j;
                    const uint32_t a = desc_lo(rec ? sh : su, TILE_M * 16) + kc * (CHUNK_K / 8) * ((TILE_M * 16) >> 4);
#pragma unroll
                    for (int ks = 0; ks < CHUNK_K / 16; ++ks) {
                        const uint32_t a_ks = a + ks * ((2 * TILE_M * 16) >> 4);
                        if (!rec) {
                            issue3(tmem, a_ks, desc_lo(sw, 1536) + ks * ((2 * 1536) >> 4), i192, kc == 0 && ks == 0);
                        } else {
                            issue3(tmem + 64, a_ks, desc_lo(sw, 1024) + ks * ((2 * 1024) >> 4), i128, false);
                            issue3(tmem + 192, a_ks, desc_lo(sw + 4096, 512) + ks * ((2 * 512) >> 4), i64, kc == 0 && ks == 0);
                        }
                    }
                    umma2_commit(bar_d);                    // both CTAs: the MMAs that read this stage are done
                }
                __syncwarp();
            }
            mbar_wait(bar_d, par);                          // the single weight buffer of each CTA is reused
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
        for (int j = 0; j < 8; ++j) out[(row0 + m) * 256 + c + j] = v[j];
    }
    tc_fence_before();
    cluster_sync_all();                                     // both CTAs are done with TMEM before either frees it
    if (warp == 0) tmem_dealloc2(tmem, 256);
}

}  // namespace
}  // namespace ctgcn

using namespace ctgcn;

// Test hook: out[256,256] from x[256,64], h[256,128], w_ih[384,64], w_hh[384,128] (see the kernel).  workspace ≥ 512 KB.
extern "C" int ctgcn_selftest_umma_pair(const float* x, const float* h, const float* w_ih, const float* w_hh, float* out,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    const int nchunks = 2 * (64 / CHUNK_K) + 2 * (H / CHUNK_K);
    const size_t need = (size_t)nchunks * PAIR_CHUNK;
    CTGCN_REQUIRE(x && h && w_ih && w_hh && out && workspace && workspace_bytes >= need,
                  "selftest_umma_pair: bad arguments (workspace needs %zu bytes)", need);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* packed = (uint8_t*)workspace;
    constexpr int UNITS = 2 * 96 * (CHUNK_K / 8);
    pack_pair_kernel<<<(nchunks * UNITS + 255) / 256, 256, 0, st>>>(w_ih, w_hh, 64, packed);
    CTGCN_LAUNCH_OK("pack_pair_kernel");
    const int smem = 4 * A_PLANE + HALF_BYTES + 64;
    CTGCN_CUDA_OK(cudaFuncSetAttribute(umma2_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma2_selftest_kernel<<<2, 128, smem, st>>>(x, h, packed, out);
    CTGCN_LAUNCH_OK("umma2_selftest_kernel");
    return CTGCN_OK;
}
