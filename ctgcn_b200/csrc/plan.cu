// Graph-plan builder: K uncoalesced COO matrices of one snapshot → one level-tagged union CSR.
//
// Replaces what torch.sparse.mm re-does on every call in the reference (layers.py:41-45 on the
// adj_list built by helper.py:51-82 / utils.py:89-95): coalescing + format conversion, here done
// ONCE per adj_list and merged over the K nested k-core matrices so that the hot kernel reads every
// distinct edge once.  One-off work: uses CUB radix sort / scan from the CUDA
// toolkit for the sorting, hand-written kernels for key construction, suffix detection and CSR
// finalisation.  Not on the timed path.
#include <cub/cub.cuh>

#include <stdlib.h>

#include "common.cuh"

namespace ctgcn {
namespace {

constexpr int LEVEL_BITS = 7;

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <class T>
    T* as() {
        return reinterpret_cast<T*>(p);
    }
};

__global__ void make_keys_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                 const float* __restrict__ vals, int64_t nnz, int level, int64_t n_rows,
                                 int64_t n_cols, uint64_t* __restrict__ keys, float* __restrict__ vout,
                                 int* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    int64_t r = rows[i], c = cols[i];
    if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) {
        *bad = 1;
        r = 0;
        c = 0;
    }
    keys[i] = ((uint64_t)(r * n_cols + c) << LEVEL_BITS) | (uint64_t)level;
    vout[i] = vals[i];
}

// One thread per sorted element; the first element of every (row,col) group does the group's work.
// A group holds ≤ K distinct levels (plus duplicates).  The longest suffix of the core list on which the
// entry is present with one value becomes a "nested" entry at the suffix's first level; every other
// present level becomes a one-shot entry (exact: no subtraction, any list is representable).
template <bool WRITE>
__global__ void group_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ vals, int64_t total, int k,
                             int64_t n_cols, int64_t* __restrict__ counts, const int64_t* __restrict__ offsets,
                             uint64_t* __restrict__ key2, uint64_t* __restrict__ pay,
                             unsigned long long* __restrict__ stat /* [0]=coalesced, [1]=oneshot */) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= total) return;
    const uint64_t g = keys[j] >> LEVEL_BITS;
    if (j > 0 && (keys[j - 1] >> LEVEL_BITS) == g) {
        if (!WRITE) counts[j] = 0;
        return;
    }
    float vs[CTGCN_MAX_CORES];
    uint64_t mask = 0;
    for (int64_t jj = j; jj < total; ++jj) {
        const uint64_t kk = keys[jj];
        if ((kk >> LEVEL_BITS) != g) break;
        const int lev = (int)(kk & ((1u << LEVEL_BITS) - 1));
        const uint64_t bit = 1ull << lev;
        if (mask & bit) {
            vs[lev] += vals[jj];
        } else {
            vs[lev] = vals[jj];
            mask |= bit;
        }
    }
    int f = k;
    float v = 0.f;
    if ((mask >> (k - 1)) & 1ull) {
        f = k - 1;
        v = vs[k - 1];
        while (f > 0 && ((mask >> (f - 1)) & 1ull) && vs[f - 1] == v) --f;
    }
    const uint64_t below = (f >= 64) ? mask : (mask & ((1ull << f) - 1ull));
    const int n_one = __popcll(below);
    const int n_out = (f < k ? 1 : 0) + n_one;
    if (!WRITE) {
        counts[j] = n_out;
        atomicAdd(&stat[0], (unsigned long long)__popcll(mask));
        if (n_one) atomicAdd(&stat[1], (unsigned long long)n_one);
        return;
    }
    int64_t o = offsets[j];
    const uint64_t row = g / (uint64_t)n_cols, col = g % (uint64_t)n_cols;
    if (f < k) {
        key2[o] = (((row << LEVEL_BITS) | (uint64_t)f) << 1);
        pay[o] = (col << 32) | (uint64_t)__float_as_uint(v);
        ++o;
    }
    uint64_t m = below;
    while (m) {
        const int lev = __ffsll((long long)m) - 1;
        m &= m - 1;
        key2[o] = (((row << LEVEL_BITS) | (uint64_t)lev) << 1) | 1ull;
        pay[o] = (col << 32) | (uint64_t)__float_as_uint(vs[lev]);
        ++o;
    }
}

__global__ void finalize_kernel(const uint64_t* __restrict__ key2, const uint64_t* __restrict__ pay, int64_t entries,
                                int64_t n_rows, int32_t* __restrict__ rowptr, int32_t* __restrict__ col,
                                float* __restrict__ val, uint8_t* __restrict__ lvl) {
    int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= entries) return;
    const uint64_t k2 = key2[j];
    const uint64_t p = pay[j];
    col[j] = (int32_t)(p >> 32);
    val[j] = __uint_as_float((uint32_t)(p & 0xffffffffu));
    lvl[j] = (uint8_t)(((k2 >> 1) & 127u) | ((k2 & 1ull) << 7));
    const int64_t row = (int64_t)(k2 >> (LEVEL_BITS + 1));
    const int64_t prev = (j == 0) ? -1 : (int64_t)(key2[j - 1] >> (LEVEL_BITS + 1));
    for (int64_t r = prev + 1; r <= row; ++r) rowptr[r] = (int32_t)j;
    if (j == entries - 1)
        for (int64_t r = row + 1; r <= n_rows; ++r) rowptr[r] = (int32_t)entries;
}

__global__ void validate_csr_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                    const uint8_t* __restrict__ lvl, int64_t n_rows, int64_t n_cols, int k,
                                    int64_t entries, int* __restrict__ bad, unsigned long long* __restrict__ n_one) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const int32_t s = rowptr[r], e = rowptr[r + 1];
    if (s > e || s < 0 || e > entries || (r == 0 && s != 0) || (r == n_rows - 1 && e != entries)) {
        *bad = 1;
        return;
    }
    int prev = 0;
    unsigned long long ones = 0;
    for (int32_t j = s; j < e; ++j) {
        const int lev = lvl[j] & 127;
        if (lev < prev || lev >= k || col[j] < 0 || col[j] >= n_cols) *bad = 1;
        prev = lev;
        ones += lvl[j] >> 7;
    }
    if (ones) atomicAdd(n_one, ones);
}

static int bits_for(uint64_t v) {  // number of bits needed to represent values in [0, v)
    int b = 1;
    while (b < 64 && (1ull << b) < v) ++b;
    return b;
}

__global__ void max_row_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows, int* __restrict__ out) {
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int v = r < n_rows ? rowptr[r + 1] - rowptr[r] : 0;
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}

// Rows above the hub threshold → segments (see ctgcn_plan in common.cuh).  One-off host pass over the row pointers.
// Threshold: measured at cfg5s (power-law 1 M / 10 M, 256-d, K = 20) 8192 / 4096 / 2048 / 1024 / 512 entries per segment give
// 8.70 / 6.07 / 4.79 / 4.04 / 3.75 ms per launch, so the smallest of those whose partial-sum scratch stays below 512 MB is taken.
static int build_hub_split(ctgcn_plan* p, cudaStream_t st) {
    if (p->max_row_entries <= ctgcn_plan::HUB_THRESHOLD_MIN) return CTGCN_OK;
    std::vector<int32_t> rp((size_t)p->n_rows + 1);
    CTGCN_CUDA_OK(cudaMemcpyAsync(rp.data(), p->rowptr, rp.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CTGCN_CUDA_OK(cudaStreamSynchronize(st));
    int T = 0;
    for (int cand = ctgcn_plan::HUB_THRESHOLD_MIN; cand < p->max_row_entries; cand *= 2) {
        size_t segs = 0;
        for (int64_t r = 0; r < p->n_rows; ++r) {
            const int32_t len = rp[r + 1] - rp[r];
            if (len > cand) segs += (size_t)(len + cand - 1) / cand;
        }
        if (segs * (size_t)p->k * ctgcn_plan::HUB_DMAX * sizeof(float) <= ((size_t)512 << 20)) {
            T = cand;
            break;
        }
    }
    if (!T) return CTGCN_OK;   // dense-ish input: keep the one-warp-per-row pass
    p->hub_threshold = T;
    std::vector<int32_t> row_end(rp.begin() + 1, rp.end()), seg_start, seg_end;
    p->h_hub_rows.clear();
    p->h_hub_seg_ptr.assign(1, 0);
    for (int64_t r = 0; r < p->n_rows; ++r) {
        const int32_t a = rp[r], b = rp[r + 1];
        if (b - a <= T) continue;
        row_end[r] = a;
        for (int32_t s = a; s < b; s += T) {
            seg_start.push_back(s);
            seg_end.push_back(s + T < b ? s + T : b);
        }
        p->h_hub_rows.push_back((int32_t)r);
        p->h_hub_seg_ptr.push_back((int32_t)seg_start.size());
    }
    const size_t n_seg = seg_start.size(), n_hub = p->h_hub_rows.size();
    const size_t scratch = n_seg * (size_t)p->k * ctgcn_plan::HUB_DMAX * sizeof(float);
    auto up = [&](int32_t** dst, const std::vector<int32_t>& v) -> cudaError_t {
        cudaError_t e = cudaMalloc(dst, v.size() * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemcpyAsync(*dst, v.data(), v.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        return e;
    };
    CTGCN_CUDA_OK(up(&p->row_end, row_end));
    CTGCN_CUDA_OK(up(&p->seg_start, seg_start));
    CTGCN_CUDA_OK(up(&p->seg_end, seg_end));
    CTGCN_CUDA_OK(up(&p->hub_rows, p->h_hub_rows));
    CTGCN_CUDA_OK(up(&p->hub_seg_ptr, p->h_hub_seg_ptr));
    CTGCN_CUDA_OK(cudaMalloc(&p->hub_scratch, scratch));
    CTGCN_CUDA_OK(cudaStreamSynchronize(st));   // the host vectors go away
    p->bytes += (size_t)p->n_rows * 4 + n_seg * 8 + n_hub * 8 + 4 + scratch;
    return CTGCN_OK;
}

// longest row of the finished CSR → p->max_row_entries (synchronises the stream)
static int measure_rows(ctgcn_plan* p, cudaStream_t st) {
    p->max_row_entries = 0;
    if (p->n_rows == 0 || p->entries == 0) return CTGCN_OK;
    int* d = nullptr;
    int h = 0;
    CTGCN_CUDA_OK(cudaMalloc(&d, sizeof(int)));
    cudaError_t e = cudaMemsetAsync(d, 0, sizeof(int), st);
    if (e == cudaSuccess) {
        max_row_kernel<<<(unsigned)((p->n_rows + 255) / 256), 256, 0, st>>>(p->rowptr, p->n_rows, d);
        e = cudaGetLastError();
        count_launch();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) {
        set_error("plan: CUDA error while measuring rows: %s", cudaGetErrorString(e));
        return CTGCN_ECUDA;
    }
    p->max_row_entries = h;
    return build_hub_split(p, st);
}

static int alloc_plan_arrays(ctgcn_plan* p) {
    const size_t e = (size_t)(p->entries > 0 ? p->entries : 1);
    // +16 entries of slack so that 128-bit vector loads of col/val/lvl never leave the allocation
    CTGCN_CUDA_OK(cudaMalloc(&p->rowptr, (p->n_rows + 1) * sizeof(int32_t)));
    CTGCN_CUDA_OK(cudaMalloc(&p->col, (e + 16) * sizeof(int32_t)));
    CTGCN_CUDA_OK(cudaMalloc(&p->val, (e + 16) * sizeof(float)));
    CTGCN_CUDA_OK(cudaMalloc(&p->lvl, (e + 16) * sizeof(uint8_t)));
    p->bytes = (p->n_rows + 1) * 4 + (e + 16) * 9;
    return CTGCN_OK;
}

}  // namespace
}  // namespace ctgcn

using namespace ctgcn;

extern "C" int ctgcn_plan_create_coo(int64_t n_rows, int64_t n_cols, int k, const int64_t* const* rows,
                                     const int64_t* const* cols, const float* const* vals, const int64_t* nnz,
                                     int on_device, void* stream, ctgcn_plan** out) {
    CTGCN_REQUIRE(out != nullptr, "plan_create_coo: out is NULL");
    *out = nullptr;
    CTGCN_REQUIRE(n_rows > 0 && n_cols > 0, "plan_create_coo: empty shape %lld x %lld", (long long)n_rows, (long long)n_cols);
    CTGCN_REQUIRE(k >= 1 && k <= CTGCN_MAX_CORES, "plan_create_coo: k=%d outside [1,%d]", k, CTGCN_MAX_CORES);
    CTGCN_REQUIRE(n_rows < (1ll << 31) && n_cols < (1ll << 31), "plan_create_coo: shape exceeds int32 indices");
    const int key_bits = bits_for((uint64_t)n_rows * (uint64_t)n_cols) + LEVEL_BITS;
    CTGCN_REQUIRE(key_bits <= 64, "plan_create_coo: n_rows*n_cols too large for 64-bit sort keys");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total = 0;
    for (int i = 0; i < k; ++i) {
        CTGCN_REQUIRE(nnz[i] >= 0, "plan_create_coo: negative nnz");
        total += nnz[i];
    }
    ctgcn_plan* p = new ctgcn_plan();
    p->n_rows = n_rows;
    p->n_cols = n_cols;
    p->k = k;
    p->nnz_raw_sum = total;
    cudaGetDevice(&p->device);
    auto fail = [&](int code) {
        ctgcn_plan_destroy(p);
        return code;
    };

    if (total == 0) {
        p->entries = 0;
        int rc = alloc_plan_arrays(p);
        if (rc) return fail(rc);
        if (cudaMemsetAsync(p->rowptr, 0, (n_rows + 1) * sizeof(int32_t), st) != cudaSuccess) return fail(CTGCN_ECUDA);
        cudaStreamSynchronize(st);
        *out = p;
        return CTGCN_OK;
    }

    DevBuf keys, keys_alt, v, v_alt, counts, offsets, flags, tmp_r, tmp_c, tmp_v, cub_tmp;
#define PLAN_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e));  \
            return fail(_e == cudaErrorMemoryAllocation ? CTGCN_ENOMEM : CTGCN_ECUDA);               \
        }                                                                                            \
    } while (0)

    PLAN_CUDA(keys.alloc(total * 8));
    PLAN_CUDA(keys_alt.alloc(total * 8));
    PLAN_CUDA(v.alloc(total * 4));
    PLAN_CUDA(v_alt.alloc(total * 4));
    PLAN_CUDA(flags.alloc(64));
    PLAN_CUDA(cudaMemsetAsync(flags.p, 0, 64, st));
    int* d_bad = flags.as<int>();
    unsigned long long* d_stat = reinterpret_cast<unsigned long long*>(flags.as<char>() + 16);

    int64_t max_nnz = 0;
    for (int i = 0; i < k; ++i) max_nnz = nnz[i] > max_nnz ? nnz[i] : max_nnz;
    if (!on_device) {
        PLAN_CUDA(tmp_r.alloc(max_nnz * 8));
        PLAN_CUDA(tmp_c.alloc(max_nnz * 8));
        PLAN_CUDA(tmp_v.alloc(max_nnz * 4));
    }
    int64_t off = 0;
    for (int i = 0; i < k; ++i) {
        if (nnz[i] == 0) continue;
        const int64_t *r = rows[i], *c = cols[i];
        const float* vv = vals[i];
        if (!on_device) {
            PLAN_CUDA(cudaMemcpyAsync(tmp_r.p, r, nnz[i] * 8, cudaMemcpyHostToDevice, st));
            PLAN_CUDA(cudaMemcpyAsync(tmp_c.p, c, nnz[i] * 8, cudaMemcpyHostToDevice, st));
            PLAN_CUDA(cudaMemcpyAsync(tmp_v.p, vv, nnz[i] * 4, cudaMemcpyHostToDevice, st));
            r = tmp_r.as<int64_t>();
            c = tmp_c.as<int64_t>();
            vv = tmp_v.as<float>();
        }
        const int threads = 256;
        const int64_t blocks = (nnz[i] + threads - 1) / threads;
        make_keys_kernel<<<(unsigned)blocks, threads, 0, st>>>(r, c, vv, nnz[i], i, n_rows, n_cols,
                                                               keys.as<uint64_t>() + off, v.as<float>() + off, d_bad);
        PLAN_CUDA(cudaGetLastError());
        count_launch();
        if (!on_device) PLAN_CUDA(cudaStreamSynchronize(st));  // tmp buffers are reused by the next matrix
        off += nnz[i];
    }

    // sort by (row, col, level); stable, so duplicates keep their input order
    size_t tmp_bytes = 0;
    PLAN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.as<uint64_t>(), keys_alt.as<uint64_t>(),
                                              v.as<float>(), v_alt.as<float>(), total, 0, key_bits, st));
    PLAN_CUDA(cub_tmp.alloc(tmp_bytes));
    PLAN_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, keys.as<uint64_t>(), keys_alt.as<uint64_t>(),
                                              v.as<float>(), v_alt.as<float>(), total, 0, key_bits, st));
    count_launch();
    const uint64_t* skeys = keys_alt.as<uint64_t>();
    const float* svals = v_alt.as<float>();

    PLAN_CUDA(counts.alloc(total * 8));
    PLAN_CUDA(offsets.alloc(total * 8));
    const int threads = 128;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    group_kernel<false><<<blocks, threads, 0, st>>>(skeys, svals, total, k, n_cols, counts.as<int64_t>(), nullptr,
                                                    nullptr, nullptr, d_stat);
    PLAN_CUDA(cudaGetLastError());
    count_launch();
    size_t scan_bytes = 0;
    PLAN_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, counts.as<int64_t>(), offsets.as<int64_t>(), total, st));
    if (scan_bytes > tmp_bytes) {
        cudaFree(cub_tmp.p);
        cub_tmp.p = nullptr;
        PLAN_CUDA(cub_tmp.alloc(scan_bytes));
    }
    PLAN_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, scan_bytes, counts.as<int64_t>(), offsets.as<int64_t>(), total, st));
    count_launch();
    int64_t last_off = 0, last_cnt = 0;
    int h_bad = 0;
    unsigned long long h_stat[2] = {0, 0};
    PLAN_CUDA(cudaMemcpyAsync(&last_off, offsets.as<int64_t>() + total - 1, 8, cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(cudaMemcpyAsync(&last_cnt, counts.as<int64_t>() + total - 1, 8, cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(cudaMemcpyAsync(&h_bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(cudaMemcpyAsync(h_stat, d_stat, 16, cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(cudaStreamSynchronize(st));
    if (h_bad) {
        set_error("plan_create_coo: index out of range for shape %lld x %lld", (long long)n_rows, (long long)n_cols);
        return fail(CTGCN_EINVAL);
    }
    const int64_t entries = last_off + last_cnt;
    if (entries >= (1ll << 31)) {
        set_error("plan_create_coo: %lld union entries exceed int32 row pointers", (long long)entries);
        return fail(CTGCN_EINVAL);
    }
    p->entries = entries;
    p->nnz_coalesced = (int64_t)h_stat[0];
    p->n_oneshot = (int64_t)h_stat[1];

    // the first sort's input buffers are free now: reuse them for (key2, payload)
    uint64_t* key2 = keys.as<uint64_t>();
    DevBuf pay, key2_alt, pay_alt;
    PLAN_CUDA(pay.alloc(entries * 8));
    group_kernel<true><<<blocks, threads, 0, st>>>(skeys, svals, total, k, n_cols, nullptr, offsets.as<int64_t>(), key2,
                                                   pay.as<uint64_t>(), nullptr);
    PLAN_CUDA(cudaGetLastError());
    count_launch();
    PLAN_CUDA(key2_alt.alloc(entries * 8));
    PLAN_CUDA(pay_alt.alloc(entries * 8));
    const int key2_bits = bits_for((uint64_t)n_rows) + LEVEL_BITS + 1;
    size_t tmp2 = 0;
    PLAN_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp2, key2, key2_alt.as<uint64_t>(), pay.as<uint64_t>(),
                                              pay_alt.as<uint64_t>(), entries, 0, key2_bits, st));
    if (tmp2 > (scan_bytes > tmp_bytes ? scan_bytes : tmp_bytes)) {
        cudaFree(cub_tmp.p);
        cub_tmp.p = nullptr;
        PLAN_CUDA(cub_tmp.alloc(tmp2));
    }
    PLAN_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp2, key2, key2_alt.as<uint64_t>(), pay.as<uint64_t>(),
                                              pay_alt.as<uint64_t>(), entries, 0, key2_bits, st));
    count_launch();
    {
        int rc = alloc_plan_arrays(p);
        if (rc) return fail(rc);
    }
    const unsigned fblocks = (unsigned)((entries + 255) / 256);
    finalize_kernel<<<fblocks, 256, 0, st>>>(key2_alt.as<uint64_t>(), pay_alt.as<uint64_t>(), entries, n_rows, p->rowptr,
                                             p->col, p->val, p->lvl);
    PLAN_CUDA(cudaGetLastError());
    count_launch();
    PLAN_CUDA(cudaStreamSynchronize(st));
#undef PLAN_CUDA
    {
        int rc = measure_rows(p, st);
        if (rc) return fail(rc);
    }
    *out = p;
    return CTGCN_OK;
}

extern "C" int ctgcn_plan_create_csr(int64_t n_rows, int64_t n_cols, int k, const int32_t* rowptr, const int32_t* col,
                                     const float* val, const uint8_t* level, int64_t nnz_raw_sum, int on_device,
                                     void* stream, ctgcn_plan** out) {
    CTGCN_REQUIRE(out != nullptr, "plan_create_csr: out is NULL");
    *out = nullptr;
    CTGCN_REQUIRE(n_rows > 0 && n_cols > 0 && n_rows < (1ll << 31) && n_cols < (1ll << 31), "plan_create_csr: bad shape");
    CTGCN_REQUIRE(k >= 1 && k <= CTGCN_MAX_CORES, "plan_create_csr: k=%d outside [1,%d]", k, CTGCN_MAX_CORES);
    cudaStream_t st = (cudaStream_t)stream;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    int32_t last = 0;
    if (on_device) {
        CTGCN_CUDA_OK(cudaMemcpyAsync(&last, rowptr + n_rows, 4, cudaMemcpyDeviceToHost, st));
        CTGCN_CUDA_OK(cudaStreamSynchronize(st));
    } else {
        last = rowptr[n_rows];
    }
    CTGCN_REQUIRE(last >= 0, "plan_create_csr: negative entry count");
    ctgcn_plan* p = new ctgcn_plan();
    p->n_rows = n_rows;
    p->n_cols = n_cols;
    p->k = k;
    p->entries = last;
    p->nnz_raw_sum = nnz_raw_sum;
    p->nnz_coalesced = nnz_raw_sum;
    cudaGetDevice(&p->device);
    int rc = alloc_plan_arrays(p);
    if (rc) {
        ctgcn_plan_destroy(p);
        return rc;
    }
    DevBuf flags;
    int h_bad = 0;
    unsigned long long h_one = 0;
    cudaError_t e = flags.alloc(32);
    if (e == cudaSuccess) e = cudaMemsetAsync(flags.p, 0, 32, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->rowptr, rowptr, (n_rows + 1) * 4, kind, st);
    if (e == cudaSuccess && last) e = cudaMemcpyAsync(p->col, col, (size_t)last * 4, kind, st);
    if (e == cudaSuccess && last) e = cudaMemcpyAsync(p->val, val, (size_t)last * 4, kind, st);
    if (e == cudaSuccess && last) e = cudaMemcpyAsync(p->lvl, level, (size_t)last, kind, st);
    if (e == cudaSuccess) {
        validate_csr_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(
            p->rowptr, p->col, p->lvl, n_rows, n_cols, k, last, flags.as<int>(),
            reinterpret_cast<unsigned long long*>(flags.as<char>() + 8));
        e = cudaGetLastError();
        count_launch();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_bad, flags.p, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_one, flags.as<char>() + 8, 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        set_error("plan_create_csr: CUDA error: %s", cudaGetErrorString(e));
        ctgcn_plan_destroy(p);
        return CTGCN_ECUDA;
    }
    if (h_bad) {
        set_error("plan_create_csr: malformed CSR (row pointers, column range, level range or level order)");
        ctgcn_plan_destroy(p);
        return CTGCN_EINVAL;
    }
    p->n_oneshot = (int64_t)h_one;
    {
        int rc = measure_rows(p, st);
        if (rc) {
            ctgcn_plan_destroy(p);
            return rc;
        }
    }
    *out = p;
    return CTGCN_OK;
}

extern "C" int ctgcn_plan_destroy(ctgcn_plan* p) {
    if (!p) return CTGCN_OK;
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != p->device) cudaSetDevice(p->device);
    if (p->rowptr) cudaFree(p->rowptr);
    if (p->col) cudaFree(p->col);
    if (p->val) cudaFree(p->val);
    if (p->lvl) cudaFree(p->lvl);
    for (void* q : {(void*)p->row_end, (void*)p->seg_start, (void*)p->seg_end, (void*)p->hub_rows, (void*)p->hub_seg_ptr,
                    (void*)p->hub_scratch})
        if (q) cudaFree(q);
    if (cur != p->device) cudaSetDevice(cur);
    delete p;
    return CTGCN_OK;
}

extern "C" int ctgcn_plan_stats(const ctgcn_plan* p, int64_t stats[8]) {
    CTGCN_REQUIRE(p && stats, "plan_stats: NULL argument");
    stats[0] = p->n_rows;
    stats[1] = p->n_cols;
    stats[2] = p->k;
    stats[3] = p->entries;
    stats[4] = p->nnz_raw_sum;
    stats[5] = p->nnz_coalesced;
    stats[6] = p->n_oneshot;
    stats[7] = (int64_t)p->bytes;
    return CTGCN_OK;
}

extern "C" int ctgcn_plan_arrays(const ctgcn_plan* p, int32_t* rowptr, int32_t* col, float* val, uint8_t* level,
                                 void* stream) {
    CTGCN_REQUIRE(p, "plan_arrays: NULL plan");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t e = (size_t)p->entries;
    if (rowptr) CTGCN_CUDA_OK(cudaMemcpyAsync(rowptr, p->rowptr, (p->n_rows + 1) * 4, cudaMemcpyDeviceToDevice, st));
    if (col && e) CTGCN_CUDA_OK(cudaMemcpyAsync(col, p->col, e * 4, cudaMemcpyDeviceToDevice, st));
    if (val && e) CTGCN_CUDA_OK(cudaMemcpyAsync(val, p->val, e * 4, cudaMemcpyDeviceToDevice, st));
    if (level && e) CTGCN_CUDA_OK(cudaMemcpyAsync(level, p->lvl, e, cudaMemcpyDeviceToDevice, st));
    return CTGCN_OK;
}
