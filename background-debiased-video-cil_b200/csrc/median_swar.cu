// Temporal median, generic variant (BGD_MEDIAN_SWAR).
//
// Replaces np.median(frames, axis=0).astype(uint8) of cil_tools/extract_background.py:73.
// One thread owns 4 adjacent byte columns (or 1 when N or the base pointer is not 4-aligned)
// and finds the lower middle order statistic by an 8-step binary search on the value, counting
// `v >= candidate` with byte-SIMD compares; one more sweep yields the upper middle statistic for
// even T.  Every sweep re-reads the T rows from global memory (L2 for the tile sizes that run
// concurrently), so this variant is the any-shape fallback and the differential-test partner of
// the bit-sliced kernel, not the fast path.
#include "bgd_common.cuh"

namespace bgd {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads)
median_swar_vec4_kernel(const uint8_t *__restrict__ frames, const int64_t *__restrict__ row0,
                        const int32_t *__restrict__ Tv, int64_t N, uint8_t *__restrict__ out)
{
    const int64_t v = blockIdx.y;
    const int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x;   // word (4 columns) index
    const int64_t nwords = N >> 2;
    if (w >= nwords) return;
    const int T = Tv[v];
    const uint32_t *base = reinterpret_cast<const uint32_t *>(frames + row0[v] * N) + w;
    const int64_t stride = nwords;
    const uint32_t need_lo = (uint32_t)(T - (T - 1) / 2);   // #(v >= x) needed for s[(T-1)/2] >= x

    uint32_t lo = 0;
#pragma unroll 1
    for (int bit = 7; bit >= 0; --bit) {
        const uint32_t cand = lo | (0x01010101u << bit);
        uint32_t c02 = 0, c13 = 0;                           // 16-bit counters: bytes (0,2), (1,3)
#pragma unroll 4
        for (int t = 0; t < T; ++t) {
            const uint32_t x = __ldg(base + (int64_t)t * stride);
            const uint32_t m = __vcmpgeu4(x, cand);          // 0xFF where byte >= candidate
            c02 += m & 0x00010001u;
            c13 += (m >> 8) & 0x00010001u;
        }
        uint32_t acc = 0;
        acc |= ((c02 & 0xFFFFu) >= need_lo) ? 0x000000FFu : 0u;
        acc |= ((c13 & 0xFFFFu) >= need_lo) ? 0x0000FF00u : 0u;
        acc |= ((c02 >> 16) >= need_lo) ? 0x00FF0000u : 0u;
        acc |= ((c13 >> 16) >= need_lo) ? 0xFF000000u : 0u;
        lo = (lo & ~acc) | (cand & acc);
    }

    uint32_t res = lo;
    if ((T & 1) == 0) {
        // upper middle: s[T/2] == lo if #(v <= lo) >= T/2 + 1, else the smallest value above lo
        uint32_t g02 = 0, g13 = 0, mn = 0xFFFFFFFFu;
#pragma unroll 4
        for (int t = 0; t < T; ++t) {
            const uint32_t x = __ldg(base + (int64_t)t * stride);
            const uint32_t m = __vcmpgtu4(x, lo);
            g02 += m & 0x00010001u;
            g13 += (m >> 8) & 0x00010001u;
            mn = __vminu4(mn, x | ~m);
        }
        const uint32_t need_le = (uint32_t)(T / 2 + 1);
        uint32_t same = 0;
        same |= ((uint32_t)T - (g02 & 0xFFFFu) >= need_le) ? 0x000000FFu : 0u;
        same |= ((uint32_t)T - (g13 & 0xFFFFu) >= need_le) ? 0x0000FF00u : 0u;
        same |= ((uint32_t)T - (g02 >> 16) >= need_le) ? 0x00FF0000u : 0u;
        same |= ((uint32_t)T - (g13 >> 16) >= need_le) ? 0xFF000000u : 0u;
        const uint32_t hi = (lo & same) | (mn & ~same);
        res = __vhaddu4(lo, hi);                             // per-byte floor((lo + hi) / 2)
    }
    reinterpret_cast<uint32_t *>(out + v * N)[w] = res;
}

__global__ void __launch_bounds__(kThreads)
median_swar_scalar_kernel(const uint8_t *__restrict__ frames, const int64_t *__restrict__ row0,
                          const int32_t *__restrict__ Tv, int64_t N, uint8_t *__restrict__ out)
{
    const int64_t v = blockIdx.y;
    const int64_t n = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (n >= N) return;
    const int T = Tv[v];
    const uint8_t *base = frames + row0[v] * N + n;
    const int need_lo = T - (T - 1) / 2;
    int lo = 0;
#pragma unroll 1
    for (int bit = 7; bit >= 0; --bit) {
        const int cand = lo | (1 << bit);
        int cnt = 0;
        for (int t = 0; t < T; ++t) cnt += (int)__ldg(base + (int64_t)t * N) >= cand;
        if (cnt >= need_lo) lo = cand;
    }
    int res = lo;
    if ((T & 1) == 0) {
        int gt = 0, mn = 255;
        for (int t = 0; t < T; ++t) {
            const int x = __ldg(base + (int64_t)t * N);
            if (x > lo) { ++gt; mn = min(mn, x); }
        }
        const int hi = (T - gt >= T / 2 + 1) ? lo : mn;
        res = (lo + hi) >> 1;
    }
    out[v * N + n] = (uint8_t)res;
}

}  // namespace

int launch_median_swar(const uint8_t *d_frames, const int64_t *d_row0, const int32_t *d_T, int64_t V,
                       int64_t N, uint8_t *d_out, int T_max, cudaStream_t stream)
{
    if (V == 0 || N == 0) return BGD_OK;
    if (T_max > 65535) return fail(BGD_ERR_UNSUPPORTED, "median: more than 65535 frames per video");
    if (V > 65535) return fail(BGD_ERR_INVALID, "median (generic variant): more than 65535 videos per call");
    const bool vec4 = (N % 4 == 0) && (reinterpret_cast<uintptr_t>(d_frames) % 4 == 0) &&
                      (reinterpret_cast<uintptr_t>(d_out) % 4 == 0);
    if (vec4) {
        dim3 grid((unsigned)((N / 4 + kThreads - 1) / kThreads), (unsigned)V);
        median_swar_vec4_kernel<<<grid, kThreads, 0, stream>>>(d_frames, d_row0, d_T, N, d_out);
    } else {
        dim3 grid((unsigned)((N + kThreads - 1) / kThreads), (unsigned)V);
        median_swar_scalar_kernel<<<grid, kThreads, 0, stream>>>(d_frames, d_row0, d_T, N, d_out);
    }
    count_launch();
    BGD_CUDA_TRY(cudaGetLastError());
    return BGD_OK;
}

}  // namespace bgd
