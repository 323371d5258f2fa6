// Temporal median, product variant (BGD_MEDIAN_BITSLICED): TMA-staged, register-resident,
// bit-sliced radix select.
//
// Replaces  np.median(frames, axis=0).astype(np.uint8)   (cil_tools/extract_background.py:73,
// libs/loader/comix_loader.py:161) for many videos per launch; result is bit-exact:
//     out[n] = (s[(T-1)/2] + s[T/2]) >> 1,   s = sort(frames[:, n]).
//
// Data movement (HBM once): a CTA owns a tile of [T rows] x [256*S byte columns] of one video.
// Warp 0 brings the tile into shared memory with one TMA bulk copy (cp.async.bulk, SASS UBLKCP)
// per row, completion on an mbarrier.  Every thread then pulls its R rows x 32 columns into
// registers with two conflict-free 16-byte shared loads per row and never touches the tile
// again, so the bulk copies of the CTA's NEXT tile are issued immediately and land while the
// 8 select passes run.  The select passes use registers only (plus ~100 bytes/thread of shared
// memory for the cross-thread count), and the single output row is written with 16-byte stores.
//
// Arithmetic: the 8 x 32 bits of a thread's 32 byte-columns of one row are transposed (3 stages
// of masked shifts between 8 registers) into 8 bit-plane words, bit x of plane b = bit b of
// column x.  From there one 32-bit logic instruction works on 32 columns at once:
//   pass b = 7..0 (MSB first), per row:   e = alive ? plane_b : sticky        (1 LOP3)
//   count of e over the T rows            carry-save adders, 2 LOP3 per row
//   bit b of the answer                   [count >= T - k]  (k = rank sought, 0-based)
//   alive &= ~(plane_b ^ bit),  sticky = e                                    (1 LOP3)
// "sticky" keeps an eliminated element voting 1 if it is above the answer's prefix and 0 if
// below, so the threshold T - k is the same in every pass.  For even T the second rank (T/2)
// shares the first's state until the pass in which the two disagree; after that it is the
// minimum of its alive set, which needs an OR instead of a count (2 LOP3 per row).
// The T rows of a column group are spread over J = 4*JC threads; their partial counts (4 bit
// planes each) meet in shared memory once per pass, where one warp adds them, decides the bit
// for every column group and publishes it (2 CTA barriers per pass).
#include <algorithm>
#include <map>
#include <vector>

#include "bgd_common.cuh"

namespace bgd {
namespace {

constexpr int kMaxR = 12;                 // rows per thread (template parameter range 1..12)
constexpr int kStripBytes = 256;          // columns per strip: 8 column-group lanes x 32 bytes
constexpr int kCountPlanes = 11;          // counts up to 2047 rows

// ---- PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// ---- bit-plane arithmetic -----------------------------------------------------------------
__device__ __forceinline__ void full_add(uint32_t &acc, uint32_t x, uint32_t y, uint32_t &carry)
{
    const uint32_t s = acc ^ x ^ y;
    carry = (acc & x) | (y & (acc ^ x));
    acc = s;
}
__device__ __forceinline__ void half_add(uint32_t &acc, uint32_t x, uint32_t &carry)
{
    carry = acc & x;
    acc ^= x;
}

// Adds N words of weight 2^L into the bit-sliced counter c[0..NPL).  Carries out of plane NPL-1
// are dropped: callers size NPL so that they are zero.
template <int NPL, int L, int N>
__device__ __forceinline__ void csa_add(uint32_t (&c)[NPL], const uint32_t (&x)[N])
{
    if constexpr (L < NPL && N > 0) {
        constexpr int NC = (N + 1) / 2;
        uint32_t carry[NC];
#pragma unroll
        for (int i = 0; i + 1 < N; i += 2) full_add(c[L], x[i], x[i + 1], carry[i / 2]);
        if constexpr (N & 1) half_add(c[L], x[N - 1], carry[NC - 1]);
        csa_add<NPL, L + 1, NC>(c, carry);
    }
}

// 8x8 bit-matrix transpose across 8 registers (self-inverse).  In: w[k] = 4 bytes (columns
// 4k..4k+3 of a 32-column group).  Out: w[b] = bit plane b, bit (8*y + k) <-> column 4k + y.
__device__ __forceinline__ void bit_transpose8(uint32_t (&w)[8])
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t = w[k], u = w[k + 4];
        w[k] = (t & 0x0F0F0F0Fu) | ((u << 4) & 0xF0F0F0F0u);
        w[k + 4] = ((t >> 4) & 0x0F0F0F0Fu) | (u & 0xF0F0F0F0u);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = (q & 1) | ((q & 2) << 1);          // 0, 1, 4, 5
        const uint32_t t = w[k], u = w[k + 2];
        w[k] = (t & 0x33333333u) | ((u << 2) & 0xCCCCCCCCu);
        w[k + 2] = ((t >> 2) & 0x33333333u) | (u & 0xCCCCCCCCu);
    }
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const uint32_t t = w[k], u = w[k + 1];
        w[k] = (t & 0x55555555u) | ((u << 1) & 0xAAAAAAAAu);
        w[k + 1] = ((t >> 1) & 0x55555555u) | (u & 0xAAAAAAAAu);
    }
}

// [count >= K] for a bit-sliced count (LSB plane first) and a CTA-uniform K.
__device__ __forceinline__ uint32_t count_ge(const uint32_t (&n)[kCountPlanes], uint32_t K)
{
    uint32_t ge = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < kCountPlanes; ++b) ge = ((K >> b) & 1u) ? (n[b] & ge) : (n[b] | ge);
    return ge;
}

// ---- kernel -------------------------------------------------------------------------------
struct KParams {
    const uint8_t *frames;
    uint8_t *out;
    const int64_t *vid_row0;     // [n_videos] first row of each video of this launch
    const int32_t *vid_T;        // [n_videos] rows
    const int64_t *vid_out;      // [n_videos] output slot
    int64_t N;
    int64_t num_tiles;
    int32_t tiles_per_video;
    int32_t S;                   // strips (256-byte column blocks) per tile
    int32_t JC;                  // row-chunk warps per strip; J = 4 * JC row chunks
    int32_t tile_rows_cap;       // smem rows reserved per tile (max T of the launch)
};

template <int R>
struct LaunchBounds {
    static constexpr int kMaxThreads = R <= 6 ? 512 : 384;
};

template <int R, bool EVEN>
__device__ __forceinline__ void select_and_store(const KParams &prm, uint32_t (&P)[R][8], const int T,
                                                 const int row_base, const int g, const int j, const int G,
                                                 const int J, uint4 *red, uint32_t *red_any, uint32_t *dec,
                                                 uint8_t *out_row, const int width)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t alive[R], sticky[R], alive2[EVEN ? R : 1];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        alive[i] = (row_base + i < T) ? 0xFFFFFFFFu : 0u;
        sticky[i] = 0u;
        if (EVEN) alive2[i] = alive[i];
    }
    const uint32_t K1 = (uint32_t)(T - (T - 1) / 2);      // votes needed for rank (T-1)/2
    const uint32_t K2 = (uint32_t)(T - T / 2);            // votes needed for rank T/2

    // reducer state (meaningful in warp 0 only)
    const int H = 32 / G > 0 ? 32 / G : 1;                // J-split factor inside the reducer warp
    const int rg = lane % G, rh = lane / G;
    uint32_t diverged = 0u, lo[8], hi[8];

#pragma unroll
    for (int b = 7; b >= 0; --b) {
        const int par = b & 1;
        uint32_t e[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            e[i] = (alive[i] & P[i][b]) | (~alive[i] & sticky[i]);
            sticky[i] = e[i];
        }
        uint32_t c[4] = {0u, 0u, 0u, 0u};
        csa_add<4, 0, R>(c, e);
        red[(par * J + j) * G + g] = make_uint4(c[0], c[1], c[2], c[3]);
        if (EVEN) {
            uint32_t any = 0u;
#pragma unroll
            for (int i = 0; i < R; ++i) any |= alive2[i] & ~P[i][b];
            red_any[(par * J + j) * G + g] = any;
        }
        __syncthreads();

        if (warp == 0) {
            uint32_t n[kCountPlanes];
#pragma unroll
            for (int q = 0; q < kCountPlanes; ++q) n[q] = 0u;
            uint32_t any = 0u;
            if (rh < H) {
                for (int jj = rh; jj < J; jj += 2 * H) {
                    const uint4 u = red[(par * J + jj) * G + rg];
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (jj + H < J) v = red[(par * J + jj + H) * G + rg];
                    const uint32_t x0[2] = {u.x, v.x}, x1[2] = {u.y, v.y}, x2[2] = {u.z, v.z}, x3[2] = {u.w, v.w};
                    csa_add<kCountPlanes, 0, 2>(n, x0);
                    csa_add<kCountPlanes, 1, 2>(n, x1);
                    csa_add<kCountPlanes, 2, 2>(n, x2);
                    csa_add<kCountPlanes, 3, 2>(n, x3);
                    if (EVEN) {
                        any |= red_any[(par * J + jj) * G + rg];
                        if (jj + H < J) any |= red_any[(par * J + jj + H) * G + rg];
                    }
                }
            }
            // combine the H partial sums of each column group (lanes rg, rg+G, ...)
            for (int off = G; off < 32; off <<= 1) {
                uint32_t carry = 0u;
#pragma unroll
                for (int q = 0; q < kCountPlanes; ++q) {
                    const uint32_t o = __shfl_down_sync(0xffffffffu, n[q], off);
                    const uint32_t s = n[q] ^ o ^ carry;
                    carry = (n[q] & o) | (carry & (n[q] ^ o));
                    n[q] = s;
                }
                if (EVEN) any |= __shfl_down_sync(0xffffffffu, any, off);
            }
            const uint32_t c1 = count_ge(n, K1);
            uint32_t c2 = c1;
            if (EVEN) {
                const uint32_t c2_shared = count_ge(n, K2);
                c2 = (diverged & ~any) | (~diverged & c2_shared);
                diverged |= c1 ^ c2;
            }
            lo[b] = c1;
            hi[b] = c2;
            if (lane < G) {
                dec[(par * 2 + 0) * G + lane] = c1;
                if (EVEN) dec[(par * 2 + 1) * G + lane] = c2;
            }
        }
        __syncthreads();

        const uint32_t C1 = dec[(par * 2 + 0) * G + g];
#pragma unroll
        for (int i = 0; i < R; ++i) alive[i] &= ~(P[i][b] ^ C1);
        if (EVEN) {
            const uint32_t C2 = dec[(par * 2 + 1) * G + g];
#pragma unroll
            for (int i = 0; i < R; ++i) alive2[i] &= ~(P[i][b] ^ C2);
        }
    }

    // warp 0, lanes < G: lo/hi planes -> floor((lo + hi) / 2) -> bytes -> global
    if (warp == 0 && lane < G) {
        uint32_t res[8];
        if (EVEN) {
            uint32_t carry = 0u, s[9];
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                s[b] = lo[b] ^ hi[b] ^ carry;
                carry = (lo[b] & hi[b]) | (carry & (lo[b] ^ hi[b]));
            }
            s[8] = carry;
#pragma unroll
            for (int b = 0; b < 8; ++b) res[b] = s[b + 1];
        } else {
#pragma unroll
            for (int b = 0; b < 8; ++b) res[b] = lo[b];
        }
        bit_transpose8(res);
        const int offA = (lane >> 3) * kStripBytes + (lane & 7) * 16;
        const int offB = offA + 128;
        if (offA < width) *reinterpret_cast<uint4 *>(out_row + offA) = make_uint4(res[0], res[1], res[2], res[3]);
        if (offB < width) *reinterpret_cast<uint4 *>(out_row + offB) = make_uint4(res[4], res[5], res[6], res[7]);
    }
}

template <int R>
__global__ void __launch_bounds__(LaunchBounds<R>::kMaxThreads) median_bitsliced_kernel(const KParams prm)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int S = prm.S, JC = prm.JC;
    const int G = 8 * S, J = 4 * JC;
    const int tile_bytes_w = S * kStripBytes;                         // tile width in bytes
    uint8_t *buf = smem;                                              // [tile_rows_cap][tile_bytes_w]
    size_t off = (size_t)prm.tile_rows_cap * tile_bytes_w;
    uint4 *red = reinterpret_cast<uint4 *>(smem + off);               // [2][J][G]
    off += (size_t)2 * J * G * sizeof(uint4);
    uint32_t *red_any = reinterpret_cast<uint32_t *>(smem + off);     // [2][J][G]
    off += (size_t)2 * J * G * sizeof(uint32_t);
    uint32_t *dec = reinterpret_cast<uint32_t *>(smem + off);         // [2][2][G]
    off += (size_t)4 * G * sizeof(uint32_t);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + ((off + 7) & ~(size_t)7));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int strip = warp % S, jc = warp / S;
    const int g = strip * 8 + (lane & 7);
    const int j = jc * 4 + (lane >> 3);
    const int row_base = j * R;
    const int offA = strip * kStripBytes + (lane & 7) * 16;

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const uint64_t policy = policy_evict_first();
    auto issue_tile = [&](int64_t tile) {                            // warp 0 only
        const int64_t vid = tile / prm.tiles_per_video;
        const int ct = (int)(tile - vid * prm.tiles_per_video);
        const int T = prm.vid_T[vid];
        const int64_t col0 = (int64_t)ct * tile_bytes_w;
        const int width = (int)min((int64_t)tile_bytes_w, prm.N - col0);
        const uint8_t *src = prm.frames + prm.vid_row0[vid] * prm.N + col0;
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)T * (uint32_t)width);
        __syncwarp();
        for (int r = lane; r < T; r += 32)
            bulk_g2s(buf + (size_t)r * tile_bytes_w, src + (int64_t)r * prm.N, (uint32_t)width, bar, policy);
    };

    int64_t tile = blockIdx.x;
    uint32_t phase = 0;
    if (tile < prm.num_tiles && warp == 0) issue_tile(tile);

    for (; tile < prm.num_tiles; tile += gridDim.x) {
        const int64_t vid = tile / prm.tiles_per_video;
        const int ct = (int)(tile - vid * prm.tiles_per_video);
        const int T = prm.vid_T[vid];
        const int64_t col0 = (int64_t)ct * tile_bytes_w;
        const int width = (int)min((int64_t)tile_bytes_w, prm.N - col0);

        mbar_wait(bar, phase);
        phase ^= 1u;

        uint32_t P[R][8];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            const int r = row_base + i;
            if (r < T) {
                const uint4 a = *reinterpret_cast<const uint4 *>(buf + (size_t)r * tile_bytes_w + offA);
                const uint4 b = *reinterpret_cast<const uint4 *>(buf + (size_t)r * tile_bytes_w + offA + 128);
                P[i][0] = a.x; P[i][1] = a.y; P[i][2] = a.z; P[i][3] = a.w;
                P[i][4] = b.x; P[i][5] = b.y; P[i][6] = b.z; P[i][7] = b.w;
                bit_transpose8(P[i]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) P[i][q] = 0u;
            }
        }
        __syncthreads();                       // every thread has its rows: the tile buffer is free

        const int64_t next = tile + gridDim.x;
        if (warp == 0 && next < prm.num_tiles) {
            fence_proxy_async();
            issue_tile(next);
        }

        uint8_t *out_row = prm.out + prm.vid_out[vid] * prm.N + col0;
        if (T & 1) select_and_store<R, false>(prm, P, T, row_base, g, j, G, J, red, red_any, dec, out_row, width);
        else       select_and_store<R, true>(prm, P, T, row_base, g, j, G, J, red, red_any, dec, out_row, width);
    }
}

// ---- host side: plan + launch ----------------------------------------------------------------
struct ClassKey {
    int R, JC;
    bool operator<(const ClassKey &o) const { return R != o.R ? R < o.R : JC < o.JC; }
};

struct Tuning {
    int target_r = 8;          // preferred rows per thread
    int target_threads = 256;  // preferred CTA size
    int ctas_per_sm = 2;       // shared-memory budget divisor
};

Tuning read_tuning()
{
    Tuning t;
    if (const char *s = getenv("BGD_MEDIAN_TARGET_R")) t.target_r = std::max(1, std::min(kMaxR, atoi(s)));
    if (const char *s = getenv("BGD_MEDIAN_TARGET_THREADS")) t.target_threads = std::max(32, atoi(s));
    if (const char *s = getenv("BGD_MEDIAN_CTAS_PER_SM")) t.ctas_per_sm = std::max(1, atoi(s));
    return t;
}

template <int R>
int max_threads_for() { return LaunchBounds<R>::kMaxThreads; }

int max_threads_of(int R)
{
    switch (R) {
#define C(r) case r: return max_threads_for<r>();
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
    }
    return 0;
}

// Rows per thread R and row-chunk warps JC for a video of T frames: 4*JC*R >= T with the least
// padding at (or above) the preferred R, and a CTA (32*JC threads per strip) the kernel can launch.
ClassKey classify(int T, const Tuning &tn)
{
    for (int r = std::min(tn.target_r, kMaxR); r <= kMaxR; ++r) {
        const int JC = std::max(1, (T + 4 * r - 1) / (4 * r));
        const int R = std::max(1, (T + 4 * JC - 1) / (4 * JC));
        if (R <= kMaxR && 32 * JC <= max_threads_of(R)) return ClassKey{R, JC};
    }
    const int JC = std::max(1, (T + 4 * kMaxR - 1) / (4 * kMaxR));
    return ClassKey{std::min(kMaxR, std::max(1, (T + 4 * JC - 1) / (4 * JC))), JC};
}

size_t smem_bytes_for(int rows_cap, int S, int JC)
{
    const int G = 8 * S, J = 4 * JC;
    size_t b = (size_t)rows_cap * S * kStripBytes;
    b += (size_t)2 * J * G * 16 + (size_t)2 * J * G * 4 + (size_t)4 * G * 4;
    b = (b + 7) & ~(size_t)7;
    return b + 16;
}

template <int R>
int launch_r(const KParams &prm, int grid, int threads, size_t smem, cudaStream_t stream)
{
    static thread_local size_t configured[64] = {0};
    int dev = 0;
    BGD_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && configured[dev] < smem) {
        BGD_CUDA_TRY(cudaFuncSetAttribute(median_bitsliced_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
        configured[dev] = smem;
    } else if (dev >= 64) {
        BGD_CUDA_TRY(cudaFuncSetAttribute(median_bitsliced_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    }
    median_bitsliced_kernel<R><<<grid, threads, smem, stream>>>(prm);
    count_launch();
    BGD_CUDA_TRY(cudaGetLastError());
    return BGD_OK;
}

template <int R>
int occupancy_r(int threads, size_t smem, int *blocks)
{
    BGD_CUDA_TRY(cudaFuncSetAttribute(median_bitsliced_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BGD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, median_bitsliced_kernel<R>, threads, smem));
    return BGD_OK;
}

}  // namespace

bool median_bitsliced_supports(int64_t T_max, int64_t N)
{
    if (N <= 0 || N % 16 != 0) return false;
    if (T_max < 1) return false;
    DeviceProps dp;
    if (current_device_props(&dp) != BGD_OK) return false;
    const Tuning tn = read_tuning();
    const ClassKey k = classify((int)std::min<int64_t>(T_max, 1 << 20), tn);
    if (T_max > (int64_t)4 * k.JC * k.R) return false;
    if (32 * k.JC > max_threads_of(k.R)) return false;
    if (T_max >= (1 << kCountPlanes)) return false;
    return smem_bytes_for((int)T_max, 1, k.JC) <= (size_t)dp.smem_optin;
}

int median_bitsliced_varlen(const uint8_t *d_frames, const int64_t *h_offsets, int64_t V, int64_t N,
                            uint8_t *d_out, cudaStream_t stream)
{
    if (V == 0 || N == 0) return BGD_OK;
    DeviceProps dp;
    if (int rc = current_device_props(&dp)) return rc;
    const Tuning tn = read_tuning();

    // bucket the videos by (R, JC)
    std::map<ClassKey, std::vector<int64_t>> classes;
    for (int64_t v = 0; v < V; ++v) {
        const int64_t T = h_offsets[v + 1] - h_offsets[v];
        classes[classify((int)T, tn)].push_back(v);
    }

    // one table upload for all classes: row0[V] | out[V] | T[V], in class order
    Workspace &ws = thread_workspace();
    const size_t tbl_bytes = (size_t)V * (8 + 8 + 4);
    if (int rc = ws.acquire(tbl_bytes)) return rc;
    int64_t *h_row0 = static_cast<int64_t *>(ws.h_pinned);
    int64_t *h_out = h_row0 + V;
    int32_t *h_T = reinterpret_cast<int32_t *>(h_out + V);
    const int64_t *d_row0 = static_cast<const int64_t *>(ws.d_ptr);
    const int64_t *d_outi = d_row0 + V;
    const int32_t *d_T = reinterpret_cast<const int32_t *>(d_outi + V);
    {
        int64_t pos = 0;
        for (auto &kv : classes)
            for (int64_t v : kv.second) {
                h_row0[pos] = h_offsets[v];
                h_out[pos] = v;
                h_T[pos] = (int32_t)(h_offsets[v + 1] - h_offsets[v]);
                ++pos;
            }
    }
    BGD_CUDA_TRY(cudaMemcpyAsync(ws.d_ptr, ws.h_pinned, tbl_bytes, cudaMemcpyHostToDevice, stream));

    int64_t pos = 0;
    int rc = BGD_OK;
    for (auto &kv : classes) {
        const ClassKey key = kv.first;
        const int64_t nv = (int64_t)kv.second.size();
        int T_cap = 0;
        for (int64_t i = 0; i < nv; ++i) T_cap = std::max(T_cap, (int)h_T[pos + i]);

        // tile width: as many strips as the thread and shared-memory budgets allow
        const int max_thr = max_threads_of(key.R);
        // S in {1, 2, 4}: the reducer warp maps G = 8*S column groups onto 32 lanes
        int S = 4;
        const size_t budget = (size_t)dp.smem_optin / tn.ctas_per_sm - 1024;
        while (S > 1 && (32 * S * key.JC > std::max(tn.target_threads, 32 * key.JC) || 32 * S * key.JC > max_thr ||
                         smem_bytes_for(T_cap, S, key.JC) > budget || (int64_t)(S / 2) * kStripBytes >= N))
            S /= 2;
        const size_t smem = smem_bytes_for(T_cap, S, key.JC);
        if (smem > (size_t)dp.smem_optin || 32 * S * key.JC > max_thr) {
            rc = fail(BGD_ERR_UNSUPPORTED, "median (bit-sliced): T=%d does not fit one CTA", T_cap);
            break;
        }
        const int threads = 32 * S * key.JC;

        KParams prm{};
        prm.frames = d_frames;
        prm.out = d_out;
        prm.vid_row0 = d_row0 + pos;
        prm.vid_T = d_T + pos;
        prm.vid_out = d_outi + pos;
        prm.N = N;
        prm.tiles_per_video = (int32_t)((N + (int64_t)S * kStripBytes - 1) / ((int64_t)S * kStripBytes));
        prm.num_tiles = nv * prm.tiles_per_video;
        prm.S = S;
        prm.JC = key.JC;
        prm.tile_rows_cap = T_cap;

        int blocks_per_sm = 1;
        switch (key.R) {
#define C(r) case r: rc = occupancy_r<r>(threads, smem, &blocks_per_sm); break;
            C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
        }
        if (rc) break;
        if (blocks_per_sm < 1) { rc = fail(BGD_ERR_CUDA, "median (bit-sliced): kernel does not fit (R=%d)", key.R); break; }
        const int grid = (int)std::min<int64_t>(prm.num_tiles, (int64_t)dp.sm_count * blocks_per_sm);
        switch (key.R) {
#define C(r) case r: rc = launch_r<r>(prm, grid, threads, smem, stream); break;
            C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
        }
        if (rc) break;
        pos += nv;
    }
    const int rc2 = ws.release(stream);
    return rc ? rc : rc2;
}

}  // namespace bgd
