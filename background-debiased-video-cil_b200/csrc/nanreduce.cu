// NaN-masked temporal median / mean over float32 frames -- the reduction of the reference's simulated-
// camera-motion ("type C") background extraction:
//
//     ave_frame = np.nanmedian(transform_frames, axis=0).astype(np.uint8)     cil_tools/extract_background.py:95-96
//     ave_frame = np.nanmean(transform_frames, axis=0).astype(np.uint8)       :97-98
//
// with `frame[frame == 0] = np.nan` (:91) folded in as `zero_is_missing`.  The frames are
// RandomResizedCrop(100) outputs (T x 100 x 100 x 3 float32, T <= max_frames = 500): 30,000 columns of at
// most a few hundred values; one thread owns one column.
//
//  * median, T <= 512 (nan_median_planes_kernel): the same bit-plane rank select as the uint8 median
//    (median_ldsm.cuh), on 32-bit keys.  A thread reads its column 32 rows at a time straight from global
//    memory (coalesced across the block, one register per value), turns the floats into order-preserving
//    uint32 keys, builds the validity word of the 32 rows, and bit-transposes the 32 x 32 block in registers
//    (5 butterfly stages; the 16- and 8-bit stages are byte permutes).  The 32 plane words go to shared
//    memory ([plane][thread], conflict-free); a column of T rows is ceil(T/32) such groups.  Then 32 passes,
//    MSB first: ones += popc(alive & plane) (POPC + IMAD), alive &= plane ^ keep -- 2 LOP3 per word and pass
//    for the lower middle, 2 more for the upper middle (it shares the lower one's state until they part, then
//    it is the minimum of its own set).  The number of valid values, hence the ranks and whether two middles
//    exist, differs per column.  Result f32((a + b) / 2): numpy sums the two middles in float32, divides by 2;
//  * median, longer columns (nan_reduce_kernel): keys staged in shared memory or a global scratch buffer,
//    missing values as the maximum key, one count of the column per key bit and rank;
//  * mean: float32 running sum in frame order (np.nansum adds frame by frame along the strided axis),
//    then f32(f64(sum) / f64(n)) (numpy divides by an int64 count);
//  * n == 0 gives NaN; the uint8 cast is numpy's on x86-64: truncation, 0 for NaN.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "bgd_common.cuh"

namespace bgd {
namespace {

constexpr int kNanThreads = 64;

__device__ __forceinline__ uint32_t key_of(float x)
{
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float value_of(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// k-th smallest (0-based) of the T keys of this thread's column
template <bool SMEM>
__device__ __forceinline__ uint32_t select_kth(const uint32_t *keys, int64_t stride, int T, int k)
{
    uint32_t prefix = 0u, mask = 0u;
#pragma unroll 1
    for (int b = 31; b >= 0; --b) {
        const uint32_t bit = 1u << b;
        int zeros = 0;                                   // keys matching the prefix whose bit b is 0
        for (int t = 0; t < T; ++t) {
            const uint32_t v = keys[(int64_t)t * stride];
            zeros += ((v & (mask | bit)) == prefix) ? 1 : 0;
        }
        if (k >= zeros) {
            k -= zeros;
            prefix |= bit;
        }
        mask |= bit;
    }
    return prefix;
}

// blockIdx.y = video: rows row0[v] .. row0[v] + T[v] - 1 of `frames`, output row v (row0 == nullptr: one video of T_one rows)
template <bool SMEM>
__global__ void __launch_bounds__(kNanThreads) nan_reduce_kernel(const float *__restrict__ frames, int T_one,
                                                                 const int64_t *__restrict__ row0, const int32_t *__restrict__ Tv,
                                                                 int64_t N, int avg_method, int zero_is_missing,
                                                                 uint8_t *__restrict__ out_u8, float *__restrict__ out_f32,
                                                                 uint32_t *__restrict__ scratch)
{
    extern __shared__ uint32_t s_keys[];
    const int64_t n = (int64_t)blockIdx.x * kNanThreads + threadIdx.x;
    if (n >= N) return;
    const int64_t v = blockIdx.y;
    const int T = row0 ? Tv[v] : T_one;
    const int64_t first = row0 ? row0[v] : 0;
    frames += first * N;
    if (out_u8) out_u8 += v * N;
    if (out_f32) out_f32 += v * N;
    uint32_t *keys = SMEM ? s_keys + threadIdx.x : scratch + first * N + n;
    const int64_t stride = SMEM ? kNanThreads : N;

    int valid = 0;
    float sum = 0.0f;
    for (int t = 0; t < T; ++t) {
        const float x = frames[(int64_t)t * N + n];
        const bool miss = isnan(x) || (zero_is_missing && x == 0.0f);
        valid += miss ? 0 : 1;
        if (avg_method != 0) sum = __fadd_rn(sum, miss ? 0.0f : x);
        else keys[(int64_t)t * stride] = miss ? 0xFFFFFFFFu : key_of(x);
    }
    float res;
    if (valid == 0) {
        res = __uint_as_float(0x7FC00000u);
    } else if (avg_method != 0) {
        res = (float)((double)sum / (double)valid);
    } else {
        const float a = value_of(select_kth<SMEM>(keys, stride, T, (valid - 1) >> 1));
        const float b = (valid & 1) ? a : value_of(select_kth<SMEM>(keys, stride, T, valid >> 1));
        res = __fdiv_rn(__fadd_rn(a, b), 2.0f);
    }
    if (out_f32) out_f32[n] = res;
    if (out_u8) out_u8[n] = isnan(res) ? (uint8_t)0 : (uint8_t)(int)res;      // trunc toward zero; domain 0 <= v < 256
}


// ---- bit-plane median on 32-bit keys ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bsel(uint32_t a, uint32_t b, uint32_t mask)   // (a & mask) | (b & ~mask)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
}
// w[r] bit b  ->  w[b] bit r
__device__ __forceinline__ void bit_transpose32(uint32_t (&w)[32])
{
#pragma unroll
    for (int k = 0; k < 16; ++k) {                       // 16-bit halves: byte permutes
        const uint32_t t = w[k], u = w[k + 16];
        w[k] = __byte_perm(t, u, 0x5410);
        w[k + 16] = __byte_perm(t, u, 0x7632);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {                       // bytes
        const int k = (q & 7) | ((q & 8) << 1);
        const uint32_t t = w[k], u = w[k + 8];
        w[k] = __byte_perm(t, u, 0x6240);
        w[k + 8] = __byte_perm(t, u, 0x7351);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = (q & 3) | ((q & 12) << 1);
        const uint32_t t = w[k], u = w[k + 4];
        w[k] = bsel(t, u << 4, 0x0F0F0F0Fu);
        w[k + 4] = bsel(t >> 4, u, 0x0F0F0F0Fu);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k = (q & 1) | ((q & 14) << 1);
        const uint32_t t = w[k], u = w[k + 2];
        w[k] = bsel(t, u << 2, 0x33333333u);
        w[k + 2] = bsel(t >> 2, u, 0x33333333u);
    }
#pragma unroll
    for (int k = 0; k < 32; k += 2) {
        const uint32_t t = w[k], u = w[k + 1];
        w[k] = bsel(t, u << 1, 0x55555555u);
        w[k + 1] = bsel(t >> 1, u, 0x55555555u);
    }
}
__device__ __forceinline__ int popc_acc32(uint32_t x, int acc, uint32_t one)     // acc + popc(x), the add on the FMA pipe
{
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(__popc(x)), "r"(one), "r"(acc));
    return r;
}
__device__ __forceinline__ int imad32(int a, int b, int c)      // a * b + c issued as IMAD; the callers pass run-time operands
{
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ int isel32(int a, int b, int mask)
{
    int d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
}

// NW = capacity in 32-row groups (T <= 32 NW); planes[(k * 32 + b) * kNanThreads + thread] in dynamic shared memory
template <int NW>
__global__ void __launch_bounds__(kNanThreads) nan_median_planes_kernel(const float *__restrict__ frames, const int64_t *__restrict__ row0,
                                                                        const int32_t *__restrict__ Tv, const int32_t *__restrict__ slot,
                                                                        int64_t N, int zero_is_missing,
                                                                        uint8_t *__restrict__ out_u8, float *__restrict__ out_f32,
                                                                        uint32_t one)
{
    extern __shared__ uint32_t s_planes[];
    const int64_t n = (int64_t)blockIdx.x * kNanThreads + threadIdx.x;
    if (n >= N) return;
    const int64_t v = blockIdx.y;
    const int T = Tv[v];
    const float *col = frames + row0[v] * N + n;
    uint32_t *planes = s_planes + threadIdx.x;
    const int nw = (T + 31) >> 5;                        // groups this video uses (uniform over the block)

    uint32_t alive[NW], alive2[NW];
    int nvalid = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
        alive[k] = 0u;
        if (k < nw) {
            // 32 rows of the column, raw float bits; rows past the video's end read as NaN (= missing)
            uint32_t w[32];
            const int rows = T - k * 32;                 // >= 1; all 32 rows exist unless this is the last group
            const float *p = col + (int64_t)k * 32 * N;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                w[r] = r < rows ? __float_as_uint(*p) : 0x7FC00000u;
                p += N;
            }
            bit_transpose32(w);                          // w[b] = bit b of the 32 rows
            // validity and the order-preserving key, on the planes (32 rows per operation):
            //   NaN = exponent all ones and mantissa non-zero; +-0 = no magnitude bit set;
            //   key = bits ^ (sign ? 0xFFFFFFFF : 0x80000000)  <=>  K[31] = ~S, K[b] = U[b] ^ S
            uint32_t mant = 0u, exp_or = 0u, exp_and = 0xFFFFFFFFu;
#pragma unroll
            for (int b = 0; b < 23; ++b) mant |= w[b];
#pragma unroll
            for (int b = 23; b < 31; ++b) {
                exp_or |= w[b];
                exp_and &= w[b];
            }
            uint32_t valid = ~(exp_and & mant);
            if (zero_is_missing) valid &= mant | exp_or;
            const uint32_t sgn = w[31];
#pragma unroll
            for (int b = 0; b < 31; ++b) planes[(k * 32 + b) * kNanThreads] = w[b] ^ sgn;
            planes[(k * 32 + 31) * kNanThreads] = ~sgn;
            alive[k] = valid;
            nvalid += __popc(valid);
        }
        alive2[k] = alive[k];
    }

    float res;
    if (nvalid == 0) {
        res = __uint_as_float(0x7FC00000u);
    } else {
        // Two independent rank selects, lower middle (rank (n-1)/2) and upper middle (rank n/2; the same element when the
        // count is odd).  State of a select: rc = rank - (number of alive rows), always negative; with `ones` alive rows
        // having the bit set, d = rc + ones = rank - zeros decides the bit (0 iff d < 0), rc becomes d for bit 0 and stays
        // for bit 1, i.e. rc - ones * m with m = d >> 31.  Counts, state and result bits are IMADs (FMA pipe): per word and
        // pass the LOP3 pipe sees the 4 mask operations only, per pass the two sign shifts.  acc collects sum_b m_b 2^b
        // with m_b = -1 where the bit is 0, so the key is 0xFFFFFFFF + acc.
        int rc = ((nvalid - 1) >> 1) - nvalid, rc2 = (nvalid >> 1) - nvalid;
        int lo_acc = 0, hi_acc = 0;
        const int NEG1 = -(int)one, TWO = (int)one << 1;
        const uint32_t *pl = planes + 31 * kNanThreads;   // plane b of group 0; group k is k * 32 planes further
#pragma unroll 4
        for (int b = 31; b >= 0; --b, pl -= kNanThreads) {
            uint32_t P[NW];
            int n1 = 0, n2 = 0;                          // minus the number of alive rows with the bit set
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                P[k] = k < nw ? pl[k * 32 * kNanThreads] : 0u;
                n1 = imad32(__popc(alive[k] & P[k]), NEG1, n1);
                n2 = imad32(__popc(alive2[k] & P[k]), NEG1, n2);
            }
            const int m0 = imad32(n1, NEG1, rc) >> 31;   // all-ones: the bit is 0
            const int m2 = imad32(n2, NEG1, rc2) >> 31;
            rc = imad32(n1, m0, rc);
            rc2 = imad32(n2, m2, rc2);
            lo_acc = imad32(lo_acc, TWO, m0);
            hi_acc = imad32(hi_acc, TWO, m2);
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                alive[k] &= P[k] ^ (uint32_t)m0;
                alive2[k] &= P[k] ^ (uint32_t)m2;
            }
        }
        const float a = value_of(0xFFFFFFFFu + (uint32_t)lo_acc);
        const float c = value_of(0xFFFFFFFFu + (uint32_t)hi_acc);
        res = __fdiv_rn(__fadd_rn(a, c), 2.0f);
    }
    const int64_t o = (int64_t)slot[v] * N + n;
    if (out_f32) out_f32[o] = res;
    if (out_u8) out_u8[o] = isnan(res) ? (uint8_t)0 : (uint8_t)(int)res;
}

template <int NW>
int launch_planes(const float *d_frames, const int64_t *d_row0, const int32_t *d_T, const int32_t *d_slot, int64_t V, int64_t N,
                  int zero_is_missing, uint8_t *d_out_u8, float *d_out_f32, cudaStream_t stream)
{
    const size_t smem = (size_t)NW * 32 * kNanThreads * sizeof(uint32_t);
    auto kern = nan_median_planes_kernel<NW>;
    BGD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BGD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const dim3 grid((unsigned)((N + kNanThreads - 1) / kNanThreads), (unsigned)V);
    kern<<<grid, kNanThreads, smem, stream>>>(d_frames, d_row0, d_T, d_slot, N, zero_is_missing, d_out_u8, d_out_f32, 1u);
    count_launch();
    BGD_CUDA_TRY(cudaGetLastError());
    return BGD_OK;
}

constexpr int kPlaneClasses[] = {1, 2, 3, 4, 5, 6, 8, 12, 16};        // instantiated capacities (32-row groups)
int plane_class(int64_t T)
{
    const int nw = (int)((T + 31) / 32);
    for (int c : kPlaneClasses)
        if (nw <= c) return c;
    return 0;
}

int launch_planes_any(int cls, const float *d_frames, const int64_t *d_row0, const int32_t *d_T, const int32_t *d_slot, int64_t V,
                      int64_t N, int zero_is_missing, uint8_t *d_out_u8, float *d_out_f32, cudaStream_t stream)
{
#define BGD_NW_CASE(NWV) if (cls == NWV) return launch_planes<NWV>(d_frames, d_row0, d_T, d_slot, V, N, zero_is_missing, d_out_u8, d_out_f32, stream)
    BGD_NW_CASE(1); BGD_NW_CASE(2); BGD_NW_CASE(3); BGD_NW_CASE(4); BGD_NW_CASE(5); BGD_NW_CASE(6); BGD_NW_CASE(8);
    BGD_NW_CASE(12); BGD_NW_CASE(16);
#undef BGD_NW_CASE
    return fail(BGD_ERR_UNSUPPORTED, "nan_temporal_reduce: no kernel for %d plane groups", cls);
}

}  // namespace

int launch_nan_reduce(const float *d_frames, int64_t T, int64_t N, int avg_method, int zero_is_missing, uint8_t *d_out_u8,
                      float *d_out_f32, cudaStream_t stream)
{
    if (T < 0 || N < 0) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: negative size");
    if (T == 0) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: no frames");     // reference: nanmedian of [] raises
    if (N == 0) return BGD_OK;
    if (!d_frames || (!d_out_u8 && !d_out_f32)) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: null pointer");
    if (avg_method != 0 && avg_method != 1) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: avg_method must be 0 (median) or 1 (mean)");
    const int64_t offs[2] = {0, T};
    return launch_nan_reduce_varlen(d_frames, offs, 1, N, avg_method, zero_is_missing, d_out_u8, d_out_f32, stream);
}

int launch_nan_reduce_varlen(const float *d_frames, const int64_t *h_offsets, int64_t V, int64_t N, int avg_method,
                             int zero_is_missing, uint8_t *d_out_u8, float *d_out_f32, cudaStream_t stream)
{
    if (V < 0 || N < 0) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: negative size");
    if (V == 0) return BGD_OK;
    if (!h_offsets) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: null offsets");
    int64_t T_max = 0;
    for (int64_t v = 0; v < V; ++v) {
        const int64_t T = h_offsets[v + 1] - h_offsets[v];
        if (T <= 0) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: video %lld has no frames", (long long)v);   // reference: nanmedian of [] raises
        T_max = std::max(T_max, T);
    }
    if (N == 0) return BGD_OK;
    if (!d_frames || (!d_out_u8 && !d_out_f32)) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: null pointer");
    if (avg_method != 0 && avg_method != 1) return fail(BGD_ERR_INVALID, "nan_temporal_reduce: avg_method must be 0 (median) or 1 (mean)");
    if (T_max >= ((int64_t)1 << 31)) return fail(BGD_ERR_UNSUPPORTED, "nan_temporal_reduce: more than 2^31 frames");
    if (V > 65535) return fail(BGD_ERR_UNSUPPORTED, "nan_temporal_reduce: more than 65535 videos per call");
    DeviceProps dp;
    if (int rc = current_device_props(&dp)) return rc;

    static const bool no_planes = getenv("BGD_NAN_NO_PLANES") != nullptr;     // differential testing of the two median kernels
    const bool use_planes = avg_method == 0 && T_max <= 512 && !no_planes;

    // tables row0[V] | T[V] | slot[V]; for the bit-plane kernel the videos are grouped by plane-group capacity so that
    // every launch sizes its shared memory (= CTAs per SM) for its own videos
    Workspace &ws = thread_workspace();
    if (int rc = ws.acquire((size_t)V * 16)) return rc;
    int64_t *h_row0 = static_cast<int64_t *>(ws.h_pinned);
    int32_t *h_T = reinterpret_cast<int32_t *>(h_row0 + V);
    int32_t *h_slot = h_T + V;
    std::vector<int64_t> order(V);
    for (int64_t v = 0; v < V; ++v) order[v] = v;
    if (use_planes)
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
            return plane_class(h_offsets[a + 1] - h_offsets[a]) < plane_class(h_offsets[b + 1] - h_offsets[b]);
        });
    for (int64_t i = 0; i < V; ++i) {
        const int64_t v = order[i];
        h_row0[i] = h_offsets[v];
        h_T[i] = (int32_t)(h_offsets[v + 1] - h_offsets[v]);
        h_slot[i] = (int32_t)v;
    }
    BGD_CUDA_TRY(cudaMemcpyAsync(ws.d_ptr, ws.h_pinned, (size_t)V * 16, cudaMemcpyHostToDevice, stream));
    const int64_t *d_row0 = static_cast<const int64_t *>(ws.d_ptr);
    const int32_t *d_T = reinterpret_cast<const int32_t *>(d_row0 + V);
    const int32_t *d_slot = d_T + V;

    const dim3 grid((unsigned)((N + kNanThreads - 1) / kNanThreads), (unsigned)V);
    const size_t smem = (size_t)T_max * kNanThreads * sizeof(uint32_t);
    int rc = BGD_OK;
    if (use_planes) {
        for (int64_t i = 0; i < V && rc == BGD_OK;) {
            const int cls = plane_class(h_T[i]);
            int64_t j = i;
            while (j < V && plane_class(h_T[j]) == cls) ++j;
            rc = launch_planes_any(cls, d_frames, d_row0 + i, d_T + i, d_slot + i, j - i, N, zero_is_missing, d_out_u8, d_out_f32, stream);
            i = j;
        }
    } else if (avg_method == 1 || smem <= (size_t)dp.smem_optin) {
        const size_t dyn = avg_method == 1 ? 0 : smem;
        cudaError_t e = cudaFuncSetAttribute(nan_reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) rc = fail(BGD_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        else nan_reduce_kernel<true><<<grid, kNanThreads, dyn, stream>>>(d_frames, 0, d_row0, d_T, N, avg_method, zero_is_missing,
                                                                          d_out_u8, d_out_f32, nullptr);
    } else {                                             // very long columns: keys in a global scratch buffer
        uint32_t *scratch = nullptr;
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&scratch), (size_t)h_offsets[V] * N * sizeof(uint32_t), stream);
        if (e != cudaSuccess) rc = fail(BGD_ERR_CUDA, "cudaMallocAsync failed: %s", cudaGetErrorString(e));
        else {
            nan_reduce_kernel<false><<<grid, kNanThreads, 0, stream>>>(d_frames, 0, d_row0, d_T, N, avg_method, zero_is_missing,
                                                                        d_out_u8, d_out_f32, scratch);
            cudaFreeAsync(scratch, stream);
        }
    }
    if (rc == BGD_OK && !use_planes) {
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(BGD_ERR_CUDA, "nan_reduce_kernel launch failed: %s", cudaGetErrorString(e));
    }
    const int rc2 = ws.release(stream);
    return rc ? rc : rc2;
}

}  // namespace bgd
