// Device-side negative sampling + fused negative-sampling loss.
//
// Reference: metrics.NegativeSamplingLoss (metrics.py:18-93).  Per snapshot and batch of nodes:
//   sampling (metrics.py:68-93, a Python loop with random.sample per node): a node with at most neg_num walk co-occurrence
//     neighbours keeps all of them (stored order), any other node neg_num of them drawn uniformly WITHOUT replacement;
//     neg_num negatives are neg_num distinct POSITIONS of the frequency-expanded node list, drawn once per snapshot;
//   loss (metrics.py:55-61): with S = number of (node, positive) samples,
//     mean_s softplus(−<e_node, e_pos>) + Q · mean_s softplus(<e_node, Σ_j e_neg_j>)        (BCEWithLogits, mean reduction).
// Here the draw is one thread per batch node running Floyd's subset sampling on a counter-based generator (no host loop, no
// host↔device traffic), and the loss / its gradient are warp-per-node gather kernels: nothing of size S·D is materialised
// (the reference builds e[node_indices], e[pos_indices] and a [S, neg_num] score matrix).  HBM-bound gathers like the SpMM.
#include "common.cuh"

namespace ctgcn {
namespace {

constexpr int MAX_NEG = CTGCN_MAX_NEG;
constexpr int WARPS = 8;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform integer in [0, bound): draw number `draw` of stream `stream` (multiply-high mapping: bias < bound / 2^64)
__device__ __forceinline__ int64_t rnd_below(uint64_t seed, uint64_t stream, uint32_t draw, int64_t bound) {
    const uint64_t r = mix64(mix64(seed ^ (stream * 0x9E3779B97F4A7C15ull)) + 0xD1B54A32D192ED03ull * (draw + 1));
    return (int64_t)__umul64hi(r, (uint64_t)bound);
}
// Floyd: a uniformly random m-subset of {0 … n-1}, m ≤ MAX_NEG < n, in chosen[0..m)
__device__ __forceinline__ void floyd(int64_t n, int m, uint64_t seed, uint64_t stream, int64_t* chosen) {
    int c = 0;
    for (int64_t j = n - m; j < n; ++j, ++c) {
        int64_t t = rnd_below(seed, stream, (uint32_t)c, j + 1);
        for (int i = 0; i < c; ++i)
            if (chosen[i] == t) {
                t = j;
                break;
            }
        chosen[c] = t;
    }
}

__global__ void neg_sample_kernel(const int64_t* __restrict__ pair_ptr, const int32_t* __restrict__ pair_idx, int64_t n_nodes,
                                  const int32_t* __restrict__ freq, int64_t freq_len, const int64_t* __restrict__ batch,
                                  int64_t n_batch, int neg_num, uint64_t seed, int32_t* __restrict__ pos,
                                  int32_t* __restrict__ count, int32_t* __restrict__ neg) {
    const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t chosen[MAX_NEG];
    if (b == n_batch) {                                   // the snapshot's negatives: distinct positions of the frequency list
        if (freq_len == neg_num) {
            for (int i = 0; i < neg_num; ++i) neg[i] = freq[i];
        } else {
            floyd(freq_len, neg_num, seed, ~0ull, chosen);
            for (int i = 0; i < neg_num; ++i) neg[i] = freq[chosen[i]];
        }
        return;
    }
    if (b > n_batch) return;
    const int64_t node = batch[b];
    int32_t* out = pos + b * neg_num;
    int64_t deg = 0, start = 0;
    if (node >= 0 && node < n_nodes) {
        start = pair_ptr[node];
        deg = pair_ptr[node + 1] - start;
    }
    int kept;
    if (deg <= neg_num) {
        kept = (int)deg;
        for (int i = 0; i < kept; ++i) out[i] = pair_idx[start + i];
    } else {
        kept = neg_num;
        floyd(deg, neg_num, seed, (uint64_t)b, chosen);
        for (int i = 0; i < kept; ++i) out[i] = pair_idx[start + chosen[i]];
    }
    for (int i = kept; i < neg_num; ++i) out[i] = -1;
    count[b] = kept;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float dot_rows(const float* __restrict__ a, const float* __restrict__ b, int d, int lane) {
    float s = 0.f;
    for (int i = lane; i < d; i += 32) s = fmaf(__ldg(a + i), __ldg(b + i), s);
    return warp_sum(s);
}
__device__ __forceinline__ float softplusf(float v) { return fmaxf(v, 0.f) + log1pf(expf(-fabsf(v))); }
__device__ __forceinline__ float sigmoidf(float v) { return 1.f / (1.f + expf(-v)); }

// negsum[d] = Σ_j e[neg_j, d]   (fixed order over j)
__global__ void neg_sum_kernel(const float* __restrict__ emb, int64_t ld, int d, const int32_t* __restrict__ neg, int neg_num,
                               float* __restrict__ negsum) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int j = 0; j < neg_num; ++j) s += emb[(int64_t)neg[j] * ld + i];
        negsum[i] = s;
    }
}

// warp per batch node: partial[b] = (Σ_s softplus(−pos_score_s), count_b · softplus(neg_score_b))
__global__ void __launch_bounds__(WARPS * 32)
    neg_loss_fwd_kernel(const float* __restrict__ emb, int64_t ld, int d, const int64_t* __restrict__ batch,
                        int64_t n_batch, const int32_t* __restrict__ pos, const int32_t* __restrict__ count, int neg_num,
                        const float* __restrict__ negsum, float2* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x * (int64_t)WARPS + (threadIdx.x >> 5);
    if (b >= n_batch) return;
    const int cnt = count[b];
    float pp = 0.f, pn = 0.f;
    if (cnt > 0) {
        const float* en = emb + batch[b] * ld;
        pn = cnt * softplusf(dot_rows(en, negsum, d, lane));
        for (int s = 0; s < cnt; ++s) pp += softplusf(-dot_rows(en, emb + (int64_t)pos[b * neg_num + s] * ld, d, lane));
    }
    if (lane == 0) partial[b] = make_float2(pp, pn);
}

// one block, fixed summation order: loss = (Σ_b p.x + Q·Σ_b p.y) / S,  stat = {S, 0}
__global__ void neg_loss_reduce_kernel(const float2* __restrict__ partial, const int32_t* __restrict__ count, int64_t n_batch,
                                       float q, float* __restrict__ loss, double* __restrict__ stat) {
    __shared__ double sp[256], sn[256], sc[256];
    double a = 0.0, c = 0.0, n = 0.0;
    for (int64_t b = threadIdx.x; b < n_batch; b += 256) {
        a += partial[b].x;
        c += partial[b].y;
        n += count[b];
    }
    sp[threadIdx.x] = a, sn[threadIdx.x] = c, sc[threadIdx.x] = n;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            sp[threadIdx.x] += sp[threadIdx.x + o];
            sn[threadIdx.x] += sn[threadIdx.x + o];
            sc[threadIdx.x] += sc[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stat[0] = sc[0];
        stat[1] = 0.0;
        loss[0] = sc[0] > 0.0 ? (float)((sp[0] + (double)q * sn[0]) / sc[0]) : 0.f;
    }
}

// warp per batch node: the gradient of the loss above, accumulated into grad (+=) with atomics (a node can be the batch
// node of one warp and a positive of others); gneg[d] collects Σ_b gn_b · e_node_b for the negatives.
__global__ void __launch_bounds__(WARPS * 32)
    neg_loss_bwd_kernel(const float* __restrict__ emb, int64_t ld, int d, const int64_t* __restrict__ batch, int64_t n_batch,
                        const int32_t* __restrict__ pos, const int32_t* __restrict__ count, int neg_num,
                        const float* __restrict__ negsum, float q, const float* __restrict__ grad_loss,
                        const double* __restrict__ stat, float* __restrict__ grad, int64_t ldg, float* __restrict__ gneg) {
    const int lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x * (int64_t)WARPS + (threadIdx.x >> 5);
    if (b >= n_batch) return;
    const int cnt = count[b];
    if (cnt == 0) return;
    const float inv = grad_loss[0] / (float)stat[0];
    const int64_t node = batch[b];
    const float* en = emb + node * ld;
    float* gnode = grad + node * ldg;
    const float gn = q * cnt * sigmoidf(dot_rows(en, negsum, d, lane)) * inv;
    for (int i = lane; i < d; i += 32) {
        atomicAdd(gnode + i, gn * negsum[i]);
        atomicAdd(gneg + i, gn * en[i]);
    }
    for (int s = 0; s < cnt; ++s) {
        const int64_t p = pos[b * neg_num + s];
        const float* ep = emb + p * ld;
        const float gp = -sigmoidf(-dot_rows(en, ep, d, lane)) * inv;
        float* gpos = grad + p * ldg;
        for (int i = lane; i < d; i += 32) {
            atomicAdd(gnode + i, gp * ep[i]);
            atomicAdd(gpos + i, gp * en[i]);
        }
    }
}

__global__ void neg_grad_scatter_kernel(const int32_t* __restrict__ neg, int neg_num, int d, const float* __restrict__ gneg,
                                        const double* __restrict__ stat, float* __restrict__ grad, int64_t ldg) {
    if (stat[0] <= 0.0) return;
    const int j = blockIdx.x;
    for (int i = threadIdx.x; i < d; i += blockDim.x) atomicAdd(grad + (int64_t)neg[j] * ldg + i, gneg[i]);
}

struct LossWs {
    float* negsum;
    float* gneg;
    double* stat;
    float2* partial;
};
size_t ws_bytes(int64_t n_batch, int d) { return 2 * align_up((size_t)d * sizeof(float), 256) + 256 + (size_t)n_batch * sizeof(float2); }
LossWs carve(void* ws, int d) {
    char* p = (char*)ws;
    const size_t v = align_up((size_t)d * sizeof(float), 256);
    return {(float*)p, (float*)(p + v), (double*)(p + 2 * v), (float2*)(p + 2 * v + 256)};
}

}  // namespace
}  // namespace ctgcn

using namespace ctgcn;

extern "C" int ctgcn_neg_sample(const int64_t* pair_ptr, const int32_t* pair_idx, int64_t n_nodes, const int32_t* freq,
                                int64_t freq_len, const int64_t* batch, int64_t n_batch, int neg_num, uint64_t seed,
                                int32_t* pos, int32_t* count, int32_t* neg, void* stream) {
    CTGCN_REQUIRE(pair_ptr && freq && neg && (n_batch <= 0 || (pair_idx && batch && pos && count)), "neg_sample: NULL argument");
    CTGCN_REQUIRE(neg_num >= 1 && neg_num <= MAX_NEG, "neg_sample: neg_num=%d outside [1,%d]", neg_num, MAX_NEG);
    CTGCN_REQUIRE(n_nodes > 0 && n_batch >= 0, "neg_sample: bad sizes");
    CTGCN_REQUIRE(freq_len >= neg_num, "neg_sample: the frequency list holds %lld entries, fewer than neg_num=%d",
                  (long long)freq_len, neg_num);   // random.sample raises ValueError here (metrics.py:88)
    const unsigned blocks = (unsigned)((n_batch + 1 + 127) / 128);
    neg_sample_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(pair_ptr, pair_idx, n_nodes, freq, freq_len, batch, n_batch, neg_num,
                                                               seed, pos, count, neg);
    CTGCN_LAUNCH_OK("neg_sample_kernel");
    return CTGCN_OK;
}

extern "C" size_t ctgcn_neg_loss_workspace_bytes(int64_t n_batch, int d) {
    if (n_batch < 0 || d <= 0) return 0;
    return ws_bytes(n_batch, d);
}

static int loss_args_ok(const float* emb, int64_t ld, int64_t n_nodes, int d, const int64_t* batch, int64_t n_batch,
                        const int32_t* pos, const int32_t* count, const int32_t* neg, int neg_num, void* ws, size_t ws_b) {
    CTGCN_REQUIRE(emb && neg && (n_batch <= 0 || (batch && pos && count)), "neg_loss: NULL argument");   // empty tensors have no address
    CTGCN_REQUIRE(d >= 1 && ld >= d && n_nodes > 0 && n_batch >= 0, "neg_loss: bad sizes");
    CTGCN_REQUIRE(neg_num >= 1 && neg_num <= MAX_NEG, "neg_loss: neg_num=%d outside [1,%d]", neg_num, MAX_NEG);
    if (!ws || ws_b < ws_bytes(n_batch, d)) {
        set_error("neg_loss: workspace of %zu bytes, need %zu", ws_b, ws_bytes(n_batch, d));
        return CTGCN_ENOMEM;
    }
    return CTGCN_OK;
}

extern "C" int ctgcn_neg_loss_fwd(const float* emb, int64_t ld, int64_t n_nodes, int d, const int64_t* batch, int64_t n_batch,
                                  const int32_t* pos, const int32_t* count, const int32_t* neg, int neg_num, float q,
                                  float* loss, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = loss_args_ok(emb, ld, n_nodes, d, batch, n_batch, pos, count, neg, neg_num, workspace, workspace_bytes);
    if (rc) return rc;
    CTGCN_REQUIRE(loss, "neg_loss_fwd: NULL loss");
    cudaStream_t st = (cudaStream_t)stream;
    const LossWs w = carve(workspace, d);
    neg_sum_kernel<<<(d + 255) / 256, 256, 0, st>>>(emb, ld, d, neg, neg_num, w.negsum);
    CTGCN_LAUNCH_OK("neg_sum_kernel");
    if (n_batch > 0) {
        neg_loss_fwd_kernel<<<(unsigned)((n_batch + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(emb, ld, d, batch, n_batch, pos, count,
                                                                                          neg_num, w.negsum, w.partial);
        CTGCN_LAUNCH_OK("neg_loss_fwd_kernel");
    }
    neg_loss_reduce_kernel<<<1, 256, 0, st>>>(w.partial, count, n_batch, q, loss, w.stat);
    CTGCN_LAUNCH_OK("neg_loss_reduce_kernel");
    return CTGCN_OK;
}

// workspace: the one ctgcn_neg_loss_fwd filled for the same arguments (negsum and the sample count are read from it)
extern "C" int ctgcn_neg_loss_bwd(const float* emb, int64_t ld, int64_t n_nodes, int d, const int64_t* batch, int64_t n_batch,
                                  const int32_t* pos, const int32_t* count, const int32_t* neg, int neg_num, float q,
                                  const float* grad_loss, float* grad_emb, int64_t ldg, void* workspace, size_t workspace_bytes,
                                  void* stream) {
    int rc = loss_args_ok(emb, ld, n_nodes, d, batch, n_batch, pos, count, neg, neg_num, workspace, workspace_bytes);
    if (rc) return rc;
    CTGCN_REQUIRE(grad_loss && grad_emb && ldg >= d, "neg_loss_bwd: bad gradient arguments");
    if (n_batch == 0) return CTGCN_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const LossWs w = carve(workspace, d);
    CTGCN_CUDA_OK(cudaMemsetAsync(w.gneg, 0, (size_t)d * sizeof(float), st));
    neg_loss_bwd_kernel<<<(unsigned)((n_batch + WARPS - 1) / WARPS), WARPS * 32, 0, st>>>(emb, ld, d, batch, n_batch, pos, count,
                                                                                      neg_num, w.negsum, q, grad_loss, w.stat,
                                                                                      grad_emb, ldg, w.gneg);
    CTGCN_LAUNCH_OK("neg_loss_bwd_kernel");
    neg_grad_scatter_kernel<<<neg_num, 128, 0, st>>>(neg, neg_num, d, w.gneg, w.stat, grad_emb, ldg);
    CTGCN_LAUNCH_OK("neg_grad_scatter_kernel");
    return CTGCN_OK;
}
