// tcgen05 / TMEM / mbarrier / bulk-copy PTX wrappers and the split-bf16 helpers shared by the tensor-core kernels
// (gru_tc.cu, linear_tc.cu).  sm_100a only.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace ctgcn {
namespace tc {

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    // the suspend-time hint lets the hardware park the warp instead of spinning through issue slots
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    return ok;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// K-major, no-swizzle ("interleaved") operands: 8-row × 16-byte core matrices; LBO = byte distance between the two
// 8-element K halves of one K=16 MMA (= rows·16 in the layouts used here), SBO = 128 B between consecutive 8-row
// groups.  The 64-bit descriptor's high word is constant (SBO, version 1); the low word = (address >> 4) | (LBO >> 4) << 16.
constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo) { return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; }
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred;
}
// No "memory" clobber on the loads on purpose: TMEM is not memory the compiler knows about, and without the clobber the
// loads of the next 8-feature pass can be hoisted above the shared-memory stores of the current one (tmem_ld_wait keeps
// its clobber and orders the consumers).
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// Ampere-style asynchronous copy global → shared memory, 16 bytes per thread, cached at the L2 only (LDGSTS)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------ CTA pairs (cta_group::2)
// Verified in isolation by ctgcn_selftest_umma_pair (umma2_selftest.cu); used by the pipelined kernel in gru_tc2.cu.
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster.  Default semantics (release at CTA scope),
// the form CUTLASS's ClusterBarrier::arrive(cta_id) uses: what is published before it is SHARED MEMORY of the arriving CTA
// (coherent by construction; made visible to the async proxy by fence.proxy.async), so no cluster-scope fence is needed — the
// .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR and the matching .acquire.cluster wait to CCTL.IVALL (L1 invalidate),
// which measured +1.4 K cycles on every hand-off of the pair kernel (profiles/r02_experiments.md).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// wait on a LOCAL barrier whose arrivals come from the peer CTA (same wait as for local arrivals, see above)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z)
        : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar) {   // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}

// CG-generic forms (CG = CTAs per MMA): used by gru_tc2.cu and gru_wide_tc.cu
template <int CG>
__device__ __forceinline__ void umma_cg(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 2) umma2_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
    else umma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
}
template <int CG>
__device__ __forceinline__ void commit_cg(uint32_t bar) {
    if constexpr (CG == 2) umma2_commit(bar);
    else umma_commit(bar);
}
// arrive on a barrier that lives in the pair's LEADER (rank 0)
template <int CG>
__device__ __forceinline__ void arrive_leader(uint32_t bar, uint32_t rank) {
    if (CG == 2 && rank != 0) mbar_arrive_remote(bar, 0);
    else mbar_arrive(bar);
}
// leader-side wait on a barrier with arrivals from both CTAs
template <int CG>
__device__ __forceinline__ void wait_pair(uint32_t bar, uint32_t parity) {
    if constexpr (CG == 2) mbar_wait_cluster(bar, parity);
    else mbar_wait(bar, parity);
}


// ------------------------------------------------------------------------------------------------ bf16 split
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    // packed conversions (F2FP, ALU pipe) instead of scalar F2F (XU pipe, shared with the MUFU gate math)
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    split2(v[0], v[1], hi.x, lo.x);
    split2(v[2], v[3], hi.y, lo.y);
    split2(v[4], v[5], hi.z, lo.z);
    split2(v[6], v[7], hi.w, lo.w);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void join8(const uint4& hi, const uint4& lo, float (&v)[8]) {
    v[0] = bf_lo(hi.x) + bf_lo(lo.x);
    v[1] = bf_hi(hi.x) + bf_hi(lo.x);
    v[2] = bf_lo(hi.y) + bf_lo(lo.y);
    v[3] = bf_hi(hi.y) + bf_hi(lo.y);
    v[4] = bf_lo(hi.z) + bf_lo(lo.z);
    v[5] = bf_hi(hi.z) + bf_hi(lo.z);
    v[6] = bf_lo(hi.w) + bf_lo(lo.w);
    v[7] = bf_hi(hi.w) + bf_hi(lo.w);
}


// a ≈ hi + mid + lo with bf16 planes (packed pairs): 24 significant bits
__device__ __forceinline__ void split3_2(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);
    hi = *reinterpret_cast<const uint32_t*>(&h2);
    const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
    const __nv_bfloat162 m2 = __floats2bfloat162_rn(ra, rb);
    mid = *reinterpret_cast<const uint32_t*>(&m2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(ra - __uint_as_float(mid << 16), rb - __uint_as_float(mid & 0xffff0000u));
    lo = *reinterpret_cast<const uint32_t*>(&l2);
}
__device__ __forceinline__ void split3_8(const float (&v)[8], uint4& hi, uint4& mid, uint4& lo) {
    split3_2(v[0], v[1], hi.x, mid.x, lo.x);
    split3_2(v[2], v[3], hi.y, mid.y, lo.y);
    split3_2(v[4], v[5], hi.z, mid.z, lo.z);
    split3_2(v[6], v[7], hi.w, mid.w, lo.w);
}

// ------------------------------------------------------------------------------------------------ operand staging (loader warps)
// One warp turns a block of 32 rows × 32 fp32 columns of a row-major matrix into the bf16 hi | lo operand image
// (unit (row, kb) of 8 consecutive k = 16 bytes at kb·(rows·16) + row·16 of a plane).
// Why not "lane = (row, 8 columns)": a 128-bit load / store is processed per QUARTER-warp, and every distinct line a quarter touches
// is a wavefront of the L1 data pipe — that mapping reads 8 rows per quarter = 32 wavefronts per instruction, and ncu showed the
// LSU data pipe (l1tex__data_pipe_lsu_wavefronts) at 84 % in linear_tc_kernel (56 % with this mapping).  Used by the dense-layer
// kernels only: in the GRU kernels the LSU pipe is not what binds (48 % in gru2_kernel) and the extra shuffles / selects sit on the
// loaders' critical path — gru2_kernel measured 4.6 → 5.1 ms with this block and with a lighter 8-row variant (r02_experiments.md).
// Here a quarter reads 128 contiguous bytes of ONE row (2 wavefronts):
//   load (b, i), b < 4, i < 2: quarter q = lane/8 reads row 8q + 2b + i, lane piece p = lane%8 → columns 4p … 4p+3;
//   lane pairs swap halves (4 shuffles per b): the even lane now owns unit (row 8q + 2b, kb = p/2), the odd one (row 8q + 2b + 1, kb);
//   the four units of a lane are rotated by kb (two select stages) so that store j writes row 8q + 2·((j + kb) mod 4) + parity:
//   a quarter's 8 lanes hit 8 different rows mod 8 = all 32 banks once (conflict-free 128-bit stores).
struct StageBlock {
    float4 a[4][2];
    // src: first element of the block (row 0, column 0); rows_ok: rows of the block that exist (≤ 0: none); cols_ok: columns that exist
    __device__ __forceinline__ void load(const float* __restrict__ src, int64_t ld, int rows_ok, int cols_ok, int lane) {
        const int q = lane >> 3, p = lane & 7;
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int row = 8 * q + 2 * b + i;
                a[b][i] = (row < rows_ok && 4 * p + 4 <= cols_ok) ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)row * ld + 4 * p))
                                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
            }
    }
    // planes: hi plane base of the block's row 0 / k-block 0 (byte pointer); next plane at +plane_bytes; kb_stride = rows_of_operand·16
    // PLANES = 2: hi | lo (16 significant bits), 3: hi | mid | lo (24)
    template <int PLANES = 2>
    __device__ __forceinline__ void store(uint8_t* hi_base, uint32_t plane_bytes, uint32_t kb_stride, int lane) const {
        const int q = lane >> 3, m = (lane & 7) >> 1, par = lane & 1;
        float v[4][8];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const float4 send = par ? a[b][0] : a[b][1];
            float4 recv;
            recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
            recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
            recv.z = __shfl_xor_sync(0xffffffffu, send.z, 1);
            recv.w = __shfl_xor_sync(0xffffffffu, send.w, 1);
            const float4 lo4 = par ? recv : a[b][0], hi4 = par ? a[b][1] : recv;   // columns 8kb … 8kb+3 | 8kb+4 … 8kb+7 of the lane's row
            v[b][0] = lo4.x; v[b][1] = lo4.y; v[b][2] = lo4.z; v[b][3] = lo4.w;
            v[b][4] = hi4.x; v[b][5] = hi4.y; v[b][6] = hi4.z; v[b][7] = hi4.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float r1a = (m & 1) ? v[(j + 1) & 3][e] : v[j][e];
                const float r1b = (m & 1) ? v[(j + 3) & 3][e] : v[(j + 2) & 3][e];
                w[e] = (m & 2) ? r1b : r1a;                                  // = v[(j + m) & 3][e]
            }
            const int row = 8 * q + 2 * ((j + m) & 3) + par;
            uint8_t* dst = hi_base + (uint32_t)m * kb_stride + row * 16;
            if constexpr (PLANES == 2) {
                uint4 hi, lo;
                split8(w, hi, lo);
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + plane_bytes) = lo;
            } else {
                uint4 hi, mid, lo;
                split3_8(w, hi, mid, lo);
                *reinterpret_cast<uint4*>(dst) = hi;
                *reinterpret_cast<uint4*>(dst + plane_bytes) = mid;
                *reinterpret_cast<uint4*>(dst + 2 * plane_bytes) = lo;
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ GRU gate math
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

// fp32 gate math of W pre-activation columns (complete: weights and biases, pre-scaled) → h_new.  Written stage by stage over
// the W features so that the dependent chains (ex2 → rcp → ex2 → rcp) are interleaved.
// sigmoid(a) = 1/(1 + 2^a'), tanh(s) = 1 − 2/(1 + 2^s'); one reciprocal serves r and z.
template <int W>
__device__ __forceinline__ void gate_math(float (&ea)[W], float (&eb)[W], float (&gi)[W], const float (&gh)[W], const float (&hold)[W],
                                          float (&hn)[W]) {
    float zz[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
        ea[j] = ex2_approx(ea[j]);
        eb[j] = ex2_approx(eb[j]);
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
        ea[j] = 1.f + fminf(ea[j], 1e18f);     // clamped so that the product below stays finite
        eb[j] = 1.f + fminf(eb[j], 1e18f);
    }
#pragma unroll
    for (int j = 0; j < W; ++j) hn[j] = rcp_approx(ea[j] * eb[j]);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        zz[j] = hn[j] * ea[j];                                          // z
        gi[j] = fmaf(hn[j] * eb[j], gh[j], gi[j]);                      // W_in x + b_in + r ⊙ (W_hn h + b_hn)
    }
#pragma unroll
    for (int j = 0; j < W; ++j) gi[j] = ex2_approx(gi[j]);
#pragma unroll
    for (int j = 0; j < W; ++j) gi[j] = rcp_approx(1.f + gi[j]);
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float nn = fmaf(-2.f, gi[j], 1.f);                        // tanh
        hn[j] = fmaf(zz[j], hold[j] - nn, nn);                          // (1 − z) n + z h
    }
}

// selu for the tensor-core dense-layer epilogues (4 epilogue warps per SM: the libm expm1f cost 0.44 ms of a 0.74 ms launch at
// 1 M × 128 → 128).  expm1(v), v ≤ 0: ex2.approx − 1 below −1/8 (absolute error ≈ 2⁻²² of 1 → relative ≤ 2e-6 there), a degree-4
// Taylor polynomial above (truncation v⁵/120 → relative ≤ 2e-6).
__device__ __forceinline__ float selu_fast(float v) {
    constexpr float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
    const float e = ex2_approx(v * 1.4426950408889634f) - 1.f;
    const float pl = v * fmaf(v, fmaf(v, fmaf(v, 1.f / 24.f, 1.f / 6.f), 0.5f), 1.f);
    return scale * (v > 0.f ? v : alpha * (v > -0.125f ? pl : e));
}

}  // namespace tc
}  // namespace ctgcn
