// fp32 CUDA-core GRU over a short sequence + Σ/LayerNorm epilogue, and the dense Linear(+selu) layer.
//
// This is the any-shape path (and the on-device numerical reference for the tcgen05 path in
// gru_tc.cu): plain fp32 FMAs, register-tiled, weights streamed through shared memory.
//
// Reference arithmetic:
//   layers.py:59-62   output,_ = GRU(hx); output.sum(dim=1); LayerNorm      (mode SUM_LN, sequence = cores)
//   models.py:249-250 out,_ = GRU(hx); LayerNorm(out)                       (mode EACH_LN, sequence = snapshots)
//   layers.py:97-105  nn.Linear (+ F.selu)
// GRU cell (PyTorch packing r,z,n):  r = σ(W_ir x + b_ir + W_hr h + b_hr), z likewise,
//   n = tanh(W_in x + b_in + r ⊙ (W_hn h + b_hn)),  h' = (1 − z) ⊙ n + z ⊙ h,  h_0 = 0.
//
// A CTA owns R = 8·RM consecutive nodes for the WHOLE sequence: h and Σh never leave shared memory,
// only the LayerNorm result is written (N·H floats per call instead of the reference's
// [N,K,H] GRU output + [N,H] sum + [N,H] norm round trips).
#include "common.cuh"

namespace ctgcn {
namespace {

constexpr int THREADS = 256;
constexpr int KC = 8;     // k rows per weight chunk
constexpr int FBW = 128;  // features per feature block (4 per thread, stride 32)

__device__ __forceinline__ float sigmoid_f(float v) { return 1.f / (1.f + expf(-v)); }

template <int RM>
struct Tile {
    static constexpr int R = 8 * RM;   // rows per CTA
    static constexpr int RP = R + 4;   // padded row count of the transposed tiles (keeps float4 alignment)
};

// acc[g][r][j] += Σ_k A[k][row r] · Wt[k][gate g][feature j]   for one K-range, weights streamed in chunks.
// A: transposed tile in shared memory (As[k*RP + row]).  Wt: global, k-major [ktot][ng*hout].
// gate g of this part is accumulated into acc[gmap[g]].
template <int RM, int NG, int GA, int GB, int GC, int GD = 3>
__device__ __forceinline__ void gemm_part(float (&acc)[4][RM][4], const float* __restrict__ As, int ktot,
                                          const float* __restrict__ Wt, int hout, int fb, float* __restrict__ ws) {
    constexpr int RP = Tile<RM>::RP;
    constexpr int gmap[4] = {GA, GB, GC, GD};
    constexpr int CHUNK = KC * NG * FBW;
    constexpr int PER_THREAD = CHUNK / THREADS;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int ldw = NG * hout;
    float wreg[PER_THREAD];
    auto prefetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) {
            const int idx = tid + THREADS * i;
            const int kk = idx / (NG * FBW), rem = idx % (NG * FBW);
            const int g = rem / FBW, f = fb * FBW + rem % FBW;
            wreg[i] = (k0 + kk < ktot && f < hout) ? __ldg(Wt + (int64_t)(k0 + kk) * ldw + g * hout + f) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < PER_THREAD; ++i) ws[buf * CHUNK + tid + THREADS * i] = wreg[i];
    };
    const int nchunks = (ktot + KC - 1) / KC;
    prefetch(0);
    stash(0);
    __syncthreads();
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) prefetch((c + 1) * KC);
        const float* wb = ws + (c & 1) * CHUNK;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float a[RM];
            const float* ap = As + (c * KC + kk) * RP + ty * RM;
#pragma unroll
            for (int r4 = 0; r4 < RM; r4 += 4) {
                const float4 t = *reinterpret_cast<const float4*>(ap + r4);
                a[r4] = t.x;
                a[r4 + 1] = t.y;
                a[r4 + 2] = t.z;
                a[r4 + 3] = t.w;
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w = wb[(kk * NG + g) * FBW + tx + 32 * j];
#pragma unroll
                    for (int r = 0; r < RM; ++r) acc[gmap[g]][r][j] = fmaf(a[r], w, acc[gmap[g]][r][j]);
                }
            }
        }
        if (c + 1 < nchunks) stash((c + 1) & 1);
        __syncthreads();
    }
}

// LayerNorm of R rows held transposed in shared memory (buf[f*RP + row]); warp `ty` handles its RM rows.
template <int RM>
__device__ __forceinline__ void layer_norm_rows(const float* __restrict__ buf, int h, const float* __restrict__ ln_w,
                                                const float* __restrict__ ln_b, float eps, int64_t row0, int64_t n,
                                                float* __restrict__ y, int64_t yrs, const RowScatter& sc) {
    constexpr int RP = Tile<RM>::RP;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = 0; r < RM; ++r) {
        const int row = ty * RM + r;
        if (row0 + row >= n) break;  // warp-uniform
        float s = 0.f;
        for (int f = tx; f < h; f += 32) s += buf[f * RP + row];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)h;
        float q = 0.f;
        for (int f = tx; f < h; f += 32) {
            const float dlt = buf[f * RP + row] - mean;
            q = fmaf(dlt, dlt, q);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q / (float)h + eps);
        float* yr = sc.slices ? sc.row_ptr(row0 + row) : y + (row0 + row) * yrs;
        for (int f = tx; f < h; f += 32) yr[f] = (buf[f * RP + row] - mean) * rstd * __ldg(ln_w + f) + __ldg(ln_b + f);
    }
}

template <int RM>
__global__ void __launch_bounds__(THREADS)
    gru_seq_kernel(const float* __restrict__ seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h,
                   const float* __restrict__ wt_ih, const float* __restrict__ wt_hh, const float* __restrict__ b_ih,
                   const float* __restrict__ b_hh, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                   float eps, int mode, float* __restrict__ y, int64_t yrs, int64_t yss, const RowScatter sc) {
    constexpr int R = Tile<RM>::R, RP = Tile<RM>::RP;
    extern __shared__ __align__(16) float smem[];
    const int d_pad = (d_in + KC - 1) / KC * KC, h_pad = (h + KC - 1) / KC * KC;
    float* xs = smem;                     // [d_pad][RP]
    float* hs = xs + d_pad * RP;          // [2][h_pad][RP]
    float* os = hs + 2 * h_pad * RP;      // [h_pad][RP]   Σ_s h_s (SUM_LN only)
    float* ws = os + h_pad * RP;          // [2][KC*3*FBW]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int64_t row0 = blockIdx.x * (int64_t)R;

    for (int i = tid; i < (d_pad + 3 * h_pad) * RP; i += THREADS) smem[i] = 0.f;
    __syncthreads();

    const int nfb = (h + FBW - 1) / FBW;
    int cur = 0;
    for (int s = 0; s < steps; ++s) {
        // stage the step's input tile, transposed: xs[k][row]
        for (int rr = ty; rr < R; rr += THREADS / 32) {
            const int64_t row = row0 + rr;
            const float* src = seq + row * srs + (int64_t)s * sss;
            for (int k = tx; k < d_in; k += 32) xs[k * RP + rr] = row < n ? __ldg(src + k) : 0.f;
        }
        __syncthreads();
        const float* hcur = hs + cur * h_pad * RP;
        float* hnxt = hs + (cur ^ 1) * h_pad * RP;
        for (int fb = 0; fb < nfb; ++fb) {
            float acc[4][RM][4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int r = 0; r < RM; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[g][r][j] = 0.f;
            gemm_part<RM, 3, 0, 1, 2>(acc, xs, d_in, wt_ih, h, fb, ws);
            if (s > 0) gemm_part<RM, 3, 0, 1, 3>(acc, hcur, h, wt_hh, h, fb, ws);  // h_0 = 0
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = fb * FBW + tx + 32 * j;
                if (f < h) {
                    const float bir = b_ih ? __ldg(b_ih + f) : 0.f, biz = b_ih ? __ldg(b_ih + h + f) : 0.f,
                                bin = b_ih ? __ldg(b_ih + 2 * h + f) : 0.f;
                    const float bhr = b_hh ? __ldg(b_hh + f) : 0.f, bhz = b_hh ? __ldg(b_hh + h + f) : 0.f,
                                bhn = b_hh ? __ldg(b_hh + 2 * h + f) : 0.f;
#pragma unroll
                    for (int r = 0; r < RM; ++r) {
                        const int row = ty * RM + r;
                        const float rg = sigmoid_f(acc[0][r][j] + bir + bhr);
                        const float zg = sigmoid_f(acc[1][r][j] + biz + bhz);
                        const float ng = tanhf(acc[2][r][j] + bin + rg * (acc[3][r][j] + bhn));
                        const float hold = hcur[f * RP + row];
                        const float hnew = (1.f - zg) * ng + zg * hold;
                        hnxt[f * RP + row] = hnew;
                        if (mode == CTGCN_GRU_SUM_LN) os[f * RP + row] += hnew;
                    }
                }
            }
        }
        __syncthreads();
        if (mode == CTGCN_GRU_EACH_LN) {
            layer_norm_rows<RM>(hnxt, h, ln_w, ln_b, eps, row0, n, y + (int64_t)s * yss, yrs, RowScatter());
        }
        cur ^= 1;
    }
    if (mode == CTGCN_GRU_SUM_LN) layer_norm_rows<RM>(os, h, ln_w, ln_b, eps, row0, n, y, yrs, sc);
}

// LSTM flavour of the same sequence kernel (layers.py:27-28 / models.py:234-235, rnn_type='LSTM'): nn.LSTM(num_layers=1,
// batch_first=True), h_0 = c_0 = 0, PyTorch packing [i; f; g; o]:
//   i = σ(W_ii x + b_ii + W_hi h + b_hi), f, o likewise, g = tanh(W_ig x + b_ig + W_hg h + b_hg),
//   c' = f ⊙ c + i ⊙ g,  h' = o ⊙ tanh(c').
// Same tiling as gru_seq_kernel; the cell state lives in shared memory next to h (each element is owned by one thread).
template <int RM>
__global__ void __launch_bounds__(THREADS)
    lstm_seq_kernel(const float* __restrict__ seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h,
                    const float* __restrict__ wt_ih, const float* __restrict__ wt_hh, const float* __restrict__ b_ih,
                    const float* __restrict__ b_hh, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                    float eps, int mode, float* __restrict__ y, int64_t yrs, int64_t yss, const RowScatter sc) {
    constexpr int R = Tile<RM>::R, RP = Tile<RM>::RP;
    extern __shared__ __align__(16) float smem[];
    const int d_pad = (d_in + KC - 1) / KC * KC, h_pad = (h + KC - 1) / KC * KC;
    float* xs = smem;                     // [d_pad][RP]
    float* hs = xs + d_pad * RP;          // [2][h_pad][RP]
    float* os = hs + 2 * h_pad * RP;      // [h_pad][RP]   Σ_s h_s (SUM_LN only)
    float* cs = os + h_pad * RP;          // [h_pad][RP]   cell state
    float* ws = cs + h_pad * RP;          // [2][KC*4*FBW]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int64_t row0 = blockIdx.x * (int64_t)R;

    for (int i = tid; i < (d_pad + 4 * h_pad) * RP; i += THREADS) smem[i] = 0.f;
    __syncthreads();

    const int nfb = (h + FBW - 1) / FBW;
    int cur = 0;
    for (int s = 0; s < steps; ++s) {
        for (int rr = ty; rr < R; rr += THREADS / 32) {
            const int64_t row = row0 + rr;
            const float* src = seq + row * srs + (int64_t)s * sss;
            for (int k = tx; k < d_in; k += 32) xs[k * RP + rr] = row < n ? __ldg(src + k) : 0.f;
        }
        __syncthreads();
        const float* hcur = hs + cur * h_pad * RP;
        float* hnxt = hs + (cur ^ 1) * h_pad * RP;
        for (int fb = 0; fb < nfb; ++fb) {
            float acc[4][RM][4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int r = 0; r < RM; ++r)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[g][r][j] = 0.f;
            gemm_part<RM, 4, 0, 1, 2, 3>(acc, xs, d_in, wt_ih, h, fb, ws);
            if (s > 0) gemm_part<RM, 4, 0, 1, 2, 3>(acc, hcur, h, wt_hh, h, fb, ws);  // h_0 = 0
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int f = fb * FBW + tx + 32 * j;
                if (f < h) {
                    float bg[4];
#pragma unroll
                    for (int g = 0; g < 4; ++g)
                        bg[g] = (b_ih ? __ldg(b_ih + g * h + f) : 0.f) + (b_hh ? __ldg(b_hh + g * h + f) : 0.f);
#pragma unroll
                    for (int r = 0; r < RM; ++r) {
                        const int row = ty * RM + r;
                        const float ig = sigmoid_f(acc[0][r][j] + bg[0]);
                        const float fg = sigmoid_f(acc[1][r][j] + bg[1]);
                        const float gg = tanhf(acc[2][r][j] + bg[2]);
                        const float og = sigmoid_f(acc[3][r][j] + bg[3]);
                        const float cnew = fg * cs[f * RP + row] + ig * gg;
                        cs[f * RP + row] = cnew;
                        const float hnew = og * tanhf(cnew);
                        hnxt[f * RP + row] = hnew;
                        if (mode == CTGCN_GRU_SUM_LN) os[f * RP + row] += hnew;
                    }
                }
            }
        }
        __syncthreads();
        if (mode == CTGCN_GRU_EACH_LN) {
            layer_norm_rows<RM>(hnxt, h, ln_w, ln_b, eps, row0, n, y + (int64_t)s * yss, yrs, RowScatter());
        }
        cur ^= 1;
    }
    if (mode == CTGCN_GRU_SUM_LN) layer_norm_rows<RM>(os, h, ln_w, ln_b, eps, row0, n, y, yrs, sc);
}

// y[n, d_out] = act(x · Wᵀ + b);  wt is k-major [d_in][d_out].  CTA: 64 rows × 128 features, k chunks of 16.
constexpr int LKC = 16;
__global__ void __launch_bounds__(THREADS)
    linear_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int64_t d_in, const float* __restrict__ wt,
                  const float* __restrict__ bias, int64_t d_out, int act, float* __restrict__ y, int64_t ldy) {
    constexpr int RM = 8, R = 64, RP = 68;
    __shared__ __align__(16) float as[2][LKC * RP];
    __shared__ __align__(16) float ws[2][LKC * FBW];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int64_t row0 = blockIdx.x * (int64_t)R;
    const int64_t f0 = blockIdx.y * (int64_t)FBW;
    float acc[RM][4];
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[r][j] = 0.f;
    float areg[4], wreg[8];
    auto prefetch = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {  // 64 rows × 16 k
            const int idx = tid + THREADS * i;
            const int rr = idx / LKC, kk = idx % LKC;
            areg[i] = (row0 + rr < n && k0 + kk < d_in) ? __ldg(x + (row0 + rr) * ldx + k0 + kk) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // 16 k × 128 features
            const int idx = tid + THREADS * i;
            const int kk = idx / FBW, c = idx % FBW;
            wreg[i] = (k0 + kk < d_in && f0 + c < d_out) ? __ldg(wt + (k0 + kk) * d_out + f0 + c) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + THREADS * i;
            as[buf][(idx % LKC) * RP + idx / LKC] = areg[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ws[buf][tid + THREADS * i] = wreg[i];
    };
    const int64_t nchunks = (d_in + LKC - 1) / LKC;
    prefetch(0);
    stash(0);
    __syncthreads();
    for (int64_t c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) prefetch((c + 1) * LKC);
        const float* ab = as[c & 1];
        const float* wb = ws[c & 1];
#pragma unroll
        for (int kk = 0; kk < LKC; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(ab + kk * RP + ty * RM);
            const float4 a1 = *reinterpret_cast<const float4*>(ab + kk * RP + ty * RM + 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float w = wb[kk * FBW + tx + 32 * j];
#pragma unroll
                for (int r = 0; r < RM; ++r) acc[r][j] = fmaf(a[r], w, acc[r][j]);
            }
        }
        if (c + 1 < nchunks) stash((c + 1) & 1);
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t f = f0 + tx + 32 * j;
        if (f >= d_out) continue;
        const float b = bias ? __ldg(bias + f) : 0.f;
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            const int64_t row = row0 + ty * RM + r;
            if (row < n) {
                float v = acc[r][j] + b;
                if (act == CTGCN_ACT_SELU) {
                    const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
                    v = scale * (v > 0.f ? v : alpha * expm1f(v));
                }
                y[row * ldy + f] = v;
            }
        }
    }
}

template <int RM>
size_t gru_smem_bytes(int d_in, int h) {
    const int d_pad = (d_in + KC - 1) / KC * KC, h_pad = (h + KC - 1) / KC * KC;
    return ((size_t)(d_pad + 3 * h_pad) * Tile<RM>::RP + 2 * KC * 3 * FBW) * sizeof(float);
}

template <int RM>
int launch_gru_t(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* wt_ih,
                 const float* wt_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                 int mode, float* y, int64_t yrs, int64_t yss, const RowScatter& sc, cudaStream_t st) {
    const size_t smem = gru_smem_bytes<RM>(d_in, h);
    CTGCN_CUDA_OK(cudaFuncSetAttribute(gru_seq_kernel<RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((n + Tile<RM>::R - 1) / Tile<RM>::R);
    gru_seq_kernel<RM><<<blocks, THREADS, smem, st>>>(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b,
                                                      eps, mode, y, yrs, yss, sc);
    CTGCN_LAUNCH_OK("gru_seq_kernel");
    return CTGCN_OK;
}

template <int RM>
size_t lstm_smem_bytes(int d_in, int h) {
    const int d_pad = (d_in + KC - 1) / KC * KC, h_pad = (h + KC - 1) / KC * KC;
    return ((size_t)(d_pad + 4 * h_pad) * Tile<RM>::RP + 2 * KC * 4 * FBW) * sizeof(float);
}

template <int RM>
int launch_lstm_t(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* wt_ih,
                  const float* wt_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                  int mode, float* y, int64_t yrs, int64_t yss, const RowScatter& sc, cudaStream_t st) {
    const size_t smem = lstm_smem_bytes<RM>(d_in, h);
    CTGCN_CUDA_OK(cudaFuncSetAttribute(lstm_seq_kernel<RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned blocks = (unsigned)((n + Tile<RM>::R - 1) / Tile<RM>::R);
    lstm_seq_kernel<RM><<<blocks, THREADS, smem, st>>>(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b,
                                                       eps, mode, y, yrs, yss, sc);
    CTGCN_LAUNCH_OK("lstm_seq_kernel");
    return CTGCN_OK;
}

}  // namespace

int launch_lstm_simt(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* wt_ih,
                     const float* wt_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                     int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* scp, cudaStream_t st) {
    constexpr size_t kMaxSmem = 227 * 1024;
    const RowScatter sc = scp ? *scp : RowScatter();
    ProfScope prof(PROF_GRU, st);
    if (lstm_smem_bytes<8>(d_in, h) <= kMaxSmem)
        return launch_lstm_t<8>(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, st);
    CTGCN_REQUIRE(lstm_smem_bytes<4>(d_in, h) <= kMaxSmem, "lstm: d_in=%d, h=%d needs more than 227 KB of shared memory", d_in, h);
    return launch_lstm_t<4>(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, st);
}

int launch_gru_simt(const float* seq, int64_t srs, int64_t sss, int64_t n, int steps, int d_in, int h, const float* wt_ih,
                    const float* wt_hh, const float* b_ih, const float* b_hh, const float* ln_w, const float* ln_b, float eps,
                    int mode, float* y, int64_t yrs, int64_t yss, const RowScatter* scp, cudaStream_t st) {
    constexpr size_t kMaxSmem = 227 * 1024;
    const RowScatter sc = scp ? *scp : RowScatter();
    ProfScope prof(PROF_GRU, st);
    if (gru_smem_bytes<8>(d_in, h) <= kMaxSmem)
        return launch_gru_t<8>(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, st);
    CTGCN_REQUIRE(gru_smem_bytes<4>(d_in, h) <= kMaxSmem, "gru: d_in=%d, h=%d needs more than 227 KB of shared memory", d_in, h);
    return launch_gru_t<4>(seq, srs, sss, n, steps, d_in, h, wt_ih, wt_hh, b_ih, b_hh, ln_w, ln_b, eps, mode, y, yrs, yss, sc, st);
}

int launch_linear_simt(const float* x, int64_t ldx, int64_t n, int64_t d_in, const float* wt, const float* b, int64_t d_out,
                       int act, float* y, int64_t ldy, cudaStream_t st) {
    dim3 grid((unsigned)((n + 63) / 64), (unsigned)((d_out + FBW - 1) / FBW));
    CTGCN_REQUIRE(grid.y <= 65535, "linear: d_out too large");
    ProfScope prof(PROF_LINEAR, st);
    linear_kernel<<<grid, THREADS, 0, st>>>(x, ldx, n, d_in, wt, b, d_out, act, y, ldy);
    CTGCN_LAUNCH_OK("linear_kernel");
    return CTGCN_OK;
}

}  // namespace ctgcn
