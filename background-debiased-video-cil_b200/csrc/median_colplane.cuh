// Temporal median, thread-per-column bit-plane select (BGD_MEDIAN_COLPLANE) -- the AUTO path.
//
// Replaces  np.median(frames, axis=0).astype(np.uint8)   (cil_tools/extract_background.py:73,
// libs/loader/comix_loader.py:161); bit-exact:  out[n] = (s[(T-1)/2] + s[T/2]) >> 1.
//
// One thread owns C adjacent byte columns (C = 1, 2 or 4) for ALL T rows of a video, so the
// 8-pass radix select needs no communication between threads at all.
//
//  * HBM once: a CTA's tile ([T rows] x [C * blockDim bytes]) arrives in shared memory through 2-D
//    TMA tensor copies (boxes of 256 bytes x 2^k rows, SASS UTMALDG) signalled on an mbarrier; as
//    soon as every thread has moved its columns into registers the copies of the CTA's next tile
//    are issued, so they land while the select runs on registers.
//  * Bit planes across rows: 8 rows x C columns (x 4/C row groups) are packed into 8 registers and
//    transposed (3 stages of masked shifts) so that register b holds bit b of 32/C rows of each of
//    the thread's C columns.  A video of T rows is NW = ceil(T*C/32) such blocks: 8*NW registers.
//  * Pass b = 7..0: z = alive & plane_b (1 LOP3 per word), the NW words are compressed with
//    carry-save adders (~2 LOP3 per word) into <=5 count planes, and a masked POPC per plane and
//    column gives the number of alive ones.  Classic MSB-first rank select per column then decides
//    the bit (k < zeros ? 0 : 1) and alive &= ~(plane_b ^ bit) (1 LOP3 per word).
//  * Even T: the second rank (T/2) shares the first's state until the pass in which they disagree;
//    after that it is the minimum of its own alive set, which needs an OR over the words instead
//    of a count.
#pragma once

#include <cuda.h>

#include "bgd_common.cuh"

namespace bgd {
namespace colplane {

constexpr int kStripBytes = 256;          // TMA box width
constexpr int kNumMaps = 9;               // boxes of 2^0 .. 2^8 rows
constexpr int kMaxThreads = 256;

struct alignas(64) CParams {
    CUtensorMap maps[kNumMaps];  // maps[k]: frames as [rows][N] uint8, box 256 bytes x 2^k rows
    uint8_t *out;
    const int64_t *vid_row0;     // [n_videos] first row of each video (relative to the map's base)
    const int32_t *vid_T;        // [n_videos]
    const int64_t *vid_out;      // [n_videos] output slot
    int64_t N;
    int64_t num_tiles;
    int32_t tiles_per_video;
    int32_t rows_cap;            // smem rows per strip (= NW * 32 / C)
};

// host entry: launches the <C, NW, EVEN> instantiation
int launch_c1(int NW, bool even, const CParams &prm, int threads, int sm_count, size_t smem, cudaStream_t stream);
int launch_c2(int NW, bool even, const CParams &prm, int threads, int sm_count, size_t smem, cudaStream_t stream);
int launch_c4(int NW, bool even, const CParams &prm, int threads, int sm_count, size_t smem, cudaStream_t stream);

#ifdef __CUDACC__
// ---- PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Waits for the tile; a copy that never completes (it cannot, short of a corrupted tensor map) traps after
// ~20 s instead of hanging the GPU: the host then sees a launch failure, not a dead device.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;
    uint64_t t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!mbar_try_wait(bar, parity)) {
        uint64_t t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int col, int row, uint64_t *bar,
                                            uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst_smem)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(row), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// (a & mask) | (b & ~mask) in one LOP3
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t mask)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
}

// x >> S.  LOP3/SHF share one half-rate pipe (64 lanes/clk/SM) and the kernel is bound by it.  The
// FMA-pipe alternative, a multiply-high by 2^(32-S) (IMAD.HI), measured 32 lanes/clk/SM on B200
// (profiles/r1_microbench_instruction_throughput.txt) and made the kernel ~1% slower, so SHF stays.
template <int S>
__device__ __forceinline__ uint32_t shr(uint32_t x)
{
#ifdef BGD_SHR_ON_FMA
    return __umulhi(x, 1u << (32 - S));
#else
    return x >> S;
#endif
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
// Vertical count of N one-bit words: planes[L..] receive the binary digits of the per-bit-position
// sum.  Each weight is reduced to one word with 3:2 compressors ((N-1)/2 full adders, at most one
// half adder); the carries form the next weight's inputs.
template <int NPL, int L, int N>
__device__ __forceinline__ void csa_count(uint32_t (&planes)[NPL], const uint32_t (&x)[N])
{
    static_assert(N >= 1, "csa_count needs at least one word");
    if constexpr (L < NPL) {
        constexpr int NFA = (N - 1) / 2, NHA = (N - 1) % 2, NC = NFA + NHA;
        uint32_t acc = x[0];
        uint32_t carry[NC > 0 ? NC : 1];
#pragma unroll
        for (int i = 0; i < NFA; ++i) full_add(acc, x[1 + 2 * i], x[2 + 2 * i], carry[i]);
        if constexpr (NHA) half_add(acc, x[N - 1], carry[NFA]);
        planes[L] = acc;
        if constexpr (NC > 0) {
            csa_count<NPL, L + 1, (NC > 0 ? NC : 1)>(planes, carry);
        } else {
#pragma unroll
            for (int l = L + 1; l < NPL; ++l) planes[l] = 0u;
        }
    }
}
// 8x8 bit-matrix transpose across 8 registers: w[m] byte y bit b  ->  w[b] byte y bit m.
__device__ __forceinline__ void bit_transpose8(uint32_t (&w)[8])
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t = w[k], u = w[k + 4];
        w[k] = bitsel(t, u << 4, 0x0F0F0F0Fu);
        w[k + 4] = bitsel(shr<4>(t), u, 0x0F0F0F0Fu);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = (q & 1) | ((q & 2) << 1);          // 0, 1, 4, 5
        const uint32_t t = w[k], u = w[k + 2];
        w[k] = bitsel(t, u << 2, 0x33333333u);
        w[k + 2] = bitsel(shr<2>(t), u, 0x33333333u);
    }
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const uint32_t t = w[k], u = w[k + 1];
        w[k] = bitsel(t, u << 1, 0x55555555u);
        w[k + 1] = bitsel(shr<1>(t), u, 0x55555555u);
    }
}

template <int C> __host__ __device__ constexpr uint32_t col_mask(int c)
{
    return C == 1 ? 0xFFFFFFFFu : (C == 2 ? (c == 0 ? 0x00FF00FFu : 0xFF00FF00u) : (0xFFu << (8 * c)));
}
__host__ __device__ constexpr int bits_for(int n) { return n < 2 ? 1 : (n < 4 ? 2 : (n < 8 ? 3 : (n < 16 ? 4 : 5))); }

// ---- kernel -------------------------------------------------------------------------------
// blockDim.x * C must be a multiple of 256.
template <int C, int NW, bool EVEN>
__global__ void __launch_bounds__(kMaxThreads) median_colplane_kernel(const __grid_constant__ CParams prm)
{
    constexpr int CB = C == 1 ? 0 : (C == 2 ? 1 : 2);   // log2(C)
    constexpr int RPW = 32 / C;                          // rows per plane word
    constexpr int NPC = bits_for(NW);                    // count planes for NW one-bit words
    extern __shared__ __align__(128) uint8_t smem[];
    const int rows_cap = prm.rows_cap;
    const int tile_w = C * (int)blockDim.x;
    const int n_strips = tile_w / kStripBytes;
    uint8_t *buf = smem;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)n_strips * rows_cap * kStripBytes);

    const int x = C * (int)threadIdx.x;                  // first byte column of this thread in the tile
    const uint8_t *my = buf + (size_t)(x >> 8) * rows_cap * kStripBytes + (x & 255);

    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint64_t policy = 0;
    if (threadIdx.x == 0) policy = policy_evict_first();
    auto issue_tile = [&](int64_t tile) {                // thread 0 only
        const int64_t vid = tile / prm.tiles_per_video;
        const int ct = (int)(tile - vid * prm.tiles_per_video);
        const int T = prm.vid_T[vid];
        const int64_t row0 = prm.vid_row0[vid];
        mbar_arrive_expect_tx(bar, (uint32_t)T * (uint32_t)tile_w);
        for (int s = 0; s < n_strips; ++s) {
            const int col = ct * tile_w + s * kStripBytes;
            uint8_t *dst = buf + (size_t)s * rows_cap * kStripBytes;
            int r = 0;
            while (T - r >= 256) {
                tma_load_2d(dst + (size_t)r * kStripBytes, &prm.maps[8], col, (int)(row0 + r), bar, policy);
                r += 256;
            }
#pragma unroll
            for (int k = 7; k >= 0; --k)
                if ((T - r) & (1 << k)) {
                    tma_load_2d(dst + (size_t)r * kStripBytes, &prm.maps[k], col, (int)(row0 + r), bar, policy);
                    r += 1 << k;
                }
        }
    };

    int64_t tile = blockIdx.x;
    uint32_t phase = 0;
    if (threadIdx.x == 0 && tile < prm.num_tiles) issue_tile(tile);

    for (; tile < prm.num_tiles; tile += gridDim.x) {
        const int64_t vid = tile / prm.tiles_per_video;
        const int ct = (int)(tile - vid * prm.tiles_per_video);
        const int T = prm.vid_T[vid];

        mbar_wait(bar, phase);
        phase ^= 1u;

        // ---- shared memory -> registers, transposed to bit planes --------------------------
        uint32_t P[NW][8];
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const uint8_t *blk = my + (size_t)k * RPW * kStripBytes;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if constexpr (C == 4) {
                    P[k][m] = *reinterpret_cast<const uint32_t *>(blk + m * kStripBytes);
                } else if constexpr (C == 2) {
                    const uint32_t h0 = *reinterpret_cast<const uint16_t *>(blk + m * kStripBytes);
                    const uint32_t h1 = *reinterpret_cast<const uint16_t *>(blk + (8 + m) * kStripBytes);
                    P[k][m] = __byte_perm(h0, h1, 0x5410);
                } else {
                    const uint32_t b0 = blk[m * kStripBytes], b1 = blk[(8 + m) * kStripBytes];
                    const uint32_t b2 = blk[(16 + m) * kStripBytes], b3 = blk[(24 + m) * kStripBytes];
                    P[k][m] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
                }
            }
            bit_transpose8(P[k]);
        }
        __syncthreads();                                 // every thread has its columns: buffer is free
        {
            const int64_t next = tile + gridDim.x;
            if (threadIdx.x == 0 && next < prm.num_tiles) {
                fence_proxy_async();
                issue_tile(next);
            }
        }

        // ---- alive masks: bit (byte y, bit m) of word k is row k*RPW + (y >> CB)*8 + m ---------
        uint32_t alive[NW], alive2[EVEN ? NW : 1];
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int n = T - k * RPW;
            uint32_t mask = 0u;
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                int mm = n - 8 * (y >> CB);
                mm = mm < 0 ? 0 : (mm > 8 ? 8 : mm);
                mask |= ((1u << mm) - 1u) << (8 * y);
            }
            alive[k] = mask;
            if (EVEN) alive2[k] = mask;
        }

        // ---- 8-pass MSB-first rank select, per column ----------------------------------------
        int rank[C], cnt[C], lo[C], hi[C];
        bool diverged[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            rank[c] = (T - 1) / 2;                       // 0-based rank of the lower middle element
            cnt[c] = T;
            lo[c] = hi[c] = 0;
            diverged[c] = false;
        }
#pragma unroll
        for (int b = 7; b >= 0; --b) {
            uint32_t z[NW];
#pragma unroll
            for (int k = 0; k < NW; ++k) z[k] = alive[k] & P[k][b];
            uint32_t cs[NPC];
            csa_count<NPC, 0, NW>(cs, z);
            uint32_t any0 = 0u;                          // rows of the second rank's set whose bit is 0
            if (EVEN) {
#pragma unroll
                for (int k = 0; k < NW; ++k) any0 |= alive2[k] & ~P[k][b];
            }
            uint32_t keep1 = 0u, keep2 = 0u;             // column masks where the chosen bit is 1
            int ones_all = 0;                            // over all C columns (C == 2: second column = all - first)
            if constexpr (C == 2) {
#pragma unroll
                for (int q = 0; q < NPC; ++q) ones_all += __popc(cs[q]) << q;
            }
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const uint32_t M = col_mask<C>(c);
                int ones = 0;
                if (C == 2 && c == 1) {
                    ones = ones_all;                     // ones_all has had column 0 subtracted below
                } else {
#pragma unroll
                    for (int q = 0; q < NPC; ++q) ones += __popc(C == 1 ? cs[q] : (cs[q] & M)) << q;
                    if (C == 2) ones_all -= ones;
                }
                const int zeros = cnt[c] - ones;
#ifndef BGD_RANK_ARITH
                const bool take0 = rank[c] < zeros;
                bool take0_2 = take0;
                if (EVEN) {
                    take0_2 = diverged[c] ? ((any0 & M) != 0u) : (rank[c] + 1 < zeros);
                    diverged[c] = diverged[c] || (take0_2 != take0);
                }
                cnt[c] = take0 ? zeros : ones;
                rank[c] = take0 ? rank[c] : rank[c] - zeros;
                lo[c] |= take0 ? 0 : (1 << b);
                keep1 |= take0 ? 0u : M;
                if (EVEN) {
                    hi[c] |= take0_2 ? 0 : (1 << b);
                    keep2 |= take0_2 ? 0u : M;
                }
#else
                // the same decisions as arithmetic on 0/1 integers (multiplies on the FMA pipe); measured
                // 3-8% slower than the select form above on B200, kept for experiments only
                const int d = rank[c] - zeros;                       // < 0  <=>  the bit is 0
                const uint32_t one = 1u - ((uint32_t)d >> 31);       // chosen bit of the lower middle
                cnt[c] = zeros + (int)one * (ones - zeros);
                rank[c] = rank[c] - (int)one * zeros;
                lo[c] += (int)(one << b);
                keep1 += one * M;
                if (EVEN) {
                    const uint32_t shared = 1u - ((uint32_t)(d + 1) >> 31);               // rank + 1 >= zeros
                    const uint32_t alone = ((any0 & M) == 0u) ? 1u : 0u;                  // min of its own set
                    const uint32_t dv = diverged[c] ? 1u : 0u;
                    const uint32_t two = dv * alone + (1u - dv) * shared;
                    diverged[c] = diverged[c] || (two != one);
                    hi[c] += (int)(two << b);
                    keep2 += two * M;
                }
#endif
            }
#pragma unroll
            for (int k = 0; k < NW; ++k) alive[k] &= ~(P[k][b] ^ keep1);
            if (EVEN) {
#pragma unroll
                for (int k = 0; k < NW; ++k) alive2[k] &= ~(P[k][b] ^ keep2);
            }
        }

        // ---- store C bytes -----------------------------------------------------------------------
        const int64_t col0 = (int64_t)ct * tile_w + x;
        if (col0 < prm.N) {                              // N % 16 == 0 and x % C == 0: all C bytes are inside
            uint32_t packed = 0u;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int v = EVEN ? ((lo[c] + hi[c]) >> 1) : lo[c];
                packed |= (uint32_t)v << (8 * c);
            }
            uint8_t *dst = prm.out + prm.vid_out[vid] * prm.N + col0;
            if constexpr (C == 4) *reinterpret_cast<uint32_t *>(dst) = packed;
            else if constexpr (C == 2) *reinterpret_cast<uint16_t *>(dst) = (uint16_t)packed;
            else *dst = (uint8_t)packed;
        }
    }
}

template <int C, int NW, bool EVEN>
int launch_one(const CParams &prm, int threads, int sm_count, size_t smem, cudaStream_t stream)
{
    auto kern = median_colplane_kernel<C, NW, EVEN>;
    BGD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BGD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int blocks_per_sm = 0;
    BGD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem));
    if (blocks_per_sm < 1)
        return fail(BGD_ERR_CUDA, "median (column-plane): kernel C=%d NW=%d does not fit an SM (%d threads, %zu B smem)",
                    C, NW, threads, smem);
    const int grid = (int)(prm.num_tiles < (int64_t)sm_count * blocks_per_sm ? prm.num_tiles
                                                                            : (int64_t)sm_count * blocks_per_sm);
    kern<<<grid, threads, smem, stream>>>(prm);
    count_launch();
    BGD_CUDA_TRY(cudaGetLastError());
    return BGD_OK;
}

template <int C, int NW>
int launch_parity(bool even, const CParams &prm, int threads, int sm_count, size_t smem, cudaStream_t stream)
{
    return even ? launch_one<C, NW, true>(prm, threads, sm_count, smem, stream)
                : launch_one<C, NW, false>(prm, threads, sm_count, smem, stream);
}
#endif  // __CUDACC__

}  // namespace colplane
}  // namespace bgd
