// Temporal median, product variant (BGD_MEDIAN_BITSLICED): TMA-staged, register-resident,
// bit-sliced radix select.
//
// Replaces  np.median(frames, axis=0).astype(np.uint8)   (cil_tools/extract_background.py:73,
// libs/loader/comix_loader.py:161) for many videos per launch; result is bit-exact:
//     out[n] = (s[(T-1)/2] + s[T/2]) >> 1,   s = sort(frames[:, n]).
//
// Data movement (HBM once).  A CTA owns a tile of [T rows] x [256*S byte columns] of one video.
// Its service warp brings the tile into shared memory with 2-D TMA tensor copies
// (cp.async.bulk.tensor, SASS UTMALDG): T is decomposed into boxes of 256/128/.../1 rows, so a
// tile costs at most a handful of instructions and reads exactly T rows.  Every worker thread
// then pulls its R rows x 32 columns into registers with two conflict-free 16-byte shared loads
// per row and never touches the tile again, so the copies of the CTA's NEXT tile are issued as
// soon as the last worker has its rows and land while the 8 select passes run.  The passes use
// registers only (plus ~100 bytes/thread of shared memory for the cross-thread count), and the
// single output row leaves with 16-byte stores.
//
// Arithmetic.  The 8 x 32 bits of a thread's 32 byte-columns of one row are transposed (3 stages
// of masked shifts between 8 registers) into 8 bit-plane words, bit x of plane b = bit b of
// column x.  From there one 32-bit logic instruction works on 32 columns at once:
//   pass b = 7..0 (MSB first), per row:   e = alive ? plane_b : sticky        (1 LOP3)
//   count of e over the T rows            carry-save adders, ~2 LOP3 per row
//   bit b of the answer                   [count >= T - k]  (k = rank sought, 0-based)
//   alive &= ~(plane_b ^ bit),  sticky = e                                    (1 LOP3)
// "sticky" keeps an eliminated element voting 1 if it is above the answer's prefix and 0 if
// below, so the threshold T - k is the same in every pass.  For even T the second rank (T/2)
// shares the first's state until the pass in which the two disagree; after that it is the
// minimum of its alive set, which needs an OR instead of a count (2 LOP3 per row).
//
// Roles.  Worker warps hold the rows (J = 4*JC row chunks per column group).  Once per pass their
// partial counts (4 bit planes each) meet in shared memory, where the service warp adds them with
// a column-compression adder, decides the bit for every column group and publishes it; the two
// sides hand over with named barriers (workers: arrive 1 / sync 2, service: sync 1 / arrive 2).
#include <cuda.h>

#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#include "bgd_common.cuh"

namespace bgd {
namespace {

constexpr int kMaxR = 12;                 // rows per thread (template parameter range 1..12)
constexpr int kStripBytes = 256;          // columns per strip: 8 column-group lanes x 32 bytes
constexpr int kNumMaps = 9;               // TMA boxes of 2^0 .. 2^8 rows
constexpr int kBarPartials = 1, kBarDecision = 2, kBarTileFree = 3;

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
// 2-D TMA tensor copy global -> shared: box (256 bytes x 2^k rows) at (col, row).
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, int col, int row,
                                            uint64_t *bar, uint64_t policy)
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
__device__ __forceinline__ void bar_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// (a & mask) | (b & ~mask) in one LOP3
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b, uint32_t mask)
{
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
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

// Column-compression adder: four 4-bit numbers (bit planes x/y/z/w of q[0..3]) -> 6-bit sum.
__device__ __forceinline__ void add_4x4(const uint4 (&q)[4], uint32_t (&s)[6])
{
    uint32_t c1a, c1b, c2a, c2b, c2c, c3a, c3b, c3c, c4a, c4b, c4c, c5, t, u;
    t = q[0].x; full_add(t, q[1].x, q[2].x, c1a); half_add(t, q[3].x, c1b); s[0] = t;
    t = q[0].y; full_add(t, q[1].y, q[2].y, c2a); u = q[3].y; full_add(u, c1a, c1b, c2b); half_add(t, u, c2c); s[1] = t;
    t = q[0].z; full_add(t, q[1].z, q[2].z, c3a); u = q[3].z; full_add(u, c2a, c2b, c3b); full_add(t, u, c2c, c3c); s[2] = t;
    t = q[0].w; full_add(t, q[1].w, q[2].w, c4a); u = q[3].w; full_add(u, c3a, c3b, c4b); full_add(t, u, c3c, c4c); s[3] = t;
    t = c4a; full_add(t, c4b, c4c, c5); s[4] = t;
    s[5] = c5;
}

// tot += s (s has NS planes), ripple carry, NPL-plane result.
template <int NPL, int NS>
__device__ __forceinline__ void ripple_add(uint32_t (&tot)[NPL], const uint32_t (&s)[NS])
{
    uint32_t carry = 0u;
#pragma unroll
    for (int b = 0; b < NPL; ++b) {
        const uint32_t x = b < NS ? s[b] : 0u;
        const uint32_t sum = tot[b] ^ x ^ carry;
        carry = (tot[b] & x) | (carry & (tot[b] ^ x));
        tot[b] = sum;
    }
}

// 8x8 bit-matrix transpose across 8 registers (self-inverse).  In: w[k] = 4 bytes (columns
// 4k..4k+3 of a 32-column group).  Out: w[b] = bit plane b, bit (8*y + k) <-> column 4k + y.
__device__ __forceinline__ void bit_transpose8(uint32_t (&w)[8])
{
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t t = w[k], u = w[k + 4];
        w[k] = bitsel(t, u << 4, 0x0F0F0F0Fu);
        w[k + 4] = bitsel(t >> 4, u, 0x0F0F0F0Fu);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = (q & 1) | ((q & 2) << 1);          // 0, 1, 4, 5
        const uint32_t t = w[k], u = w[k + 2];
        w[k] = bitsel(t, u << 2, 0x33333333u);
        w[k + 2] = bitsel(t >> 2, u, 0x33333333u);
    }
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
        const uint32_t t = w[k], u = w[k + 1];
        w[k] = bitsel(t, u << 1, 0x55555555u);
        w[k + 1] = bitsel(t >> 1, u, 0x55555555u);
    }
}

// [count >= K] for a bit-sliced count (LSB plane first) and a CTA-uniform K.
template <int NPL>
__device__ __forceinline__ uint32_t count_ge(const uint32_t (&n)[NPL], uint32_t K)
{
    uint32_t ge = 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < NPL; ++b) ge = ((K >> b) & 1u) ? (n[b] & ge) : (n[b] | ge);
    return ge;
}

// ---- kernel -------------------------------------------------------------------------------
struct alignas(64) KParams {
    CUtensorMap maps[kNumMaps];  // maps[k]: frames as [rows][N] uint8, box 256 bytes x 2^k rows
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

struct Smem {
    uint8_t *buf;        // [tile_rows_cap][S][256]   strip-major per row block, see tile layout below
    uint4 *red;          // [2][J][G]   partial counts
    uint32_t *red_any;   // [2][J][G]
    uint32_t *dec;       // [2][2][G]   decisions of the current pass
    uint32_t *ans;       // [2][8][G]   lo / hi planes (service warp only)
    uint64_t *bar;
};

__device__ __forceinline__ Smem carve(uint8_t *smem, int rows_cap, int S, int JC)
{
    const int G = 8 * S, J = 4 * JC;
    Smem m;
    m.buf = smem;
    size_t off = (size_t)rows_cap * S * kStripBytes;
    m.red = reinterpret_cast<uint4 *>(smem + off);
    off += (size_t)2 * J * G * sizeof(uint4);
    m.red_any = reinterpret_cast<uint32_t *>(smem + off);
    off += (size_t)2 * J * G * sizeof(uint32_t);
    m.dec = reinterpret_cast<uint32_t *>(smem + off);
    off += (size_t)4 * G * sizeof(uint32_t);
    m.ans = reinterpret_cast<uint32_t *>(smem + off);
    off += (size_t)16 * G * sizeof(uint32_t);
    m.bar = reinterpret_cast<uint64_t *>(smem + ((off + 7) & ~(size_t)7));
    return m;
}

size_t smem_bytes_for(int rows_cap, int S, int JC)
{
    const int G = 8 * S, J = 4 * JC;
    size_t b = (size_t)rows_cap * S * kStripBytes;
    b += (size_t)2 * J * G * 16 + (size_t)2 * J * G * 4 + (size_t)4 * G * 4 + (size_t)16 * G * 4;
    b = (b + 7) & ~(size_t)7;
    return b + 16;
}

// The landing buffer holds the tile strip by strip: strip s occupies rows_cap*256 contiguous
// bytes, row r of strip s at (s*rows_cap + r)*256.  Each TMA box is 256 bytes wide, so one strip
// of one row block is one box.

template <int R, bool EVEN>
struct Threads {
    // register budget: planes 8R + alive/sticky(/alive2) (2 or 3)R + ~40 working registers
    static constexpr int kRegs = 8 * R + (EVEN ? 3 : 2) * R + 32;
    static constexpr int kFit = (65536 / kRegs) / 128 * 128;      // ptxas budgets registers per 128 threads
    static constexpr int kLaunch = kFit > 1024 ? 1024 : (kFit < 128 ? 128 : kFit);
};

template <int R, bool EVEN, int NPL>
__global__ void __launch_bounds__(Threads<R, EVEN>::kLaunch)
median_bitsliced_kernel(const __grid_constant__ KParams prm)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int S = prm.S, JC = prm.JC;
    const int G = 8 * S, J = 4 * JC;
    const int rows_cap = prm.tile_rows_cap;
    const Smem sm = carve(smem_raw, rows_cap, S, JC);
    const int n_workers = 32 * S * JC;
    const int n_all = n_workers + 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_service = warp == S * JC;
    const int tile_w = S * kStripBytes;

    if (threadIdx.x == 0) {
        mbar_init(sm.bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (is_service) {
        // ================= service warp: TMA producer + reducer + output =================
        const uint64_t policy = policy_evict_first();
        auto issue_tile = [&](int64_t tile) {
            const int64_t vid = tile / prm.tiles_per_video;
            const int ct = (int)(tile - vid * prm.tiles_per_video);
            const int T = prm.vid_T[vid];
            const int64_t row0 = prm.vid_row0[vid];
            if (lane == 0) {
                mbar_arrive_expect_tx(sm.bar, (uint32_t)T * (uint32_t)tile_w);
                for (int s = 0; s < S; ++s) {
                    const int col = ct * tile_w + s * kStripBytes;
                    uint8_t *dst = sm.buf + (size_t)s * rows_cap * kStripBytes;
                    int r = 0;
                    while (T - r >= 256) {
                        tma_load_2d(dst + (size_t)r * kStripBytes, &prm.maps[8], col, (int)(row0 + r), sm.bar, policy);
                        r += 256;
                    }
#pragma unroll
                    for (int k = 7; k >= 0; --k)
                        if ((T - r) & (1 << k)) {
                            tma_load_2d(dst + (size_t)r * kStripBytes, &prm.maps[k], col, (int)(row0 + r), sm.bar, policy);
                            r += 1 << k;
                        }
                }
            }
            __syncwarp();
        };

        const int H = 32 / G;                             // lanes per column group in the reducer
        const int rg = lane % G, rh = lane / G;
        int64_t tile = blockIdx.x;
        if (tile < prm.num_tiles) issue_tile(tile);
        for (; tile < prm.num_tiles; tile += gridDim.x) {
            const int64_t vid = tile / prm.tiles_per_video;
            const int ct = (int)(tile - vid * prm.tiles_per_video);
            const int T = prm.vid_T[vid];
            const uint32_t K1 = (uint32_t)(T - (T - 1) / 2);      // votes needed for rank (T-1)/2
            const uint32_t K2 = (uint32_t)(T - T / 2);            // votes needed for rank T/2

            bar_sync(kBarTileFree, n_all);                        // workers hold their rows
            const int64_t next = tile + gridDim.x;
            if (next < prm.num_tiles) {
                fence_proxy_async();
                issue_tile(next);
            }

            uint32_t diverged = 0u;
#pragma unroll 1
            for (int b = 7; b >= 0; --b) {
                const int par = b & 1;
                bar_sync(kBarPartials, n_all);
                uint32_t n[NPL];
#pragma unroll
                for (int q = 0; q < NPL; ++q) n[q] = 0u;
                uint32_t any = 0u;
                for (int j0 = rh; j0 < J; j0 += 4 * H) {
                    uint4 q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int jj = j0 + u * H;
                        q[u] = jj < J ? sm.red[(par * J + jj) * G + rg] : make_uint4(0u, 0u, 0u, 0u);
                        if (EVEN && jj < J) any |= sm.red_any[(par * J + jj) * G + rg];
                    }
                    uint32_t s6[6];
                    add_4x4(q, s6);
                    ripple_add<NPL, 6>(n, s6);
                }
                for (int off = G; off < 32; off <<= 1) {          // lanes rg, rg+G, ... -> lane rg
                    uint32_t o[NPL];
#pragma unroll
                    for (int q = 0; q < NPL; ++q) o[q] = __shfl_down_sync(0xffffffffu, n[q], off);
                    ripple_add<NPL, NPL>(n, o);
                    if (EVEN) any |= __shfl_down_sync(0xffffffffu, any, off);
                }
                const uint32_t c1 = count_ge<NPL>(n, K1);
                uint32_t c2 = c1;
                if (EVEN) {
                    const uint32_t c2_shared = count_ge<NPL>(n, K2);
                    c2 = (diverged & ~any) | (~diverged & c2_shared);
                    diverged |= c1 ^ c2;
                }
                if (lane < G) {
                    sm.dec[(par * 2 + 0) * G + lane] = c1;
                    sm.ans[(0 * 8 + b) * G + lane] = c1;
                    if (EVEN) {
                        sm.dec[(par * 2 + 1) * G + lane] = c2;
                        sm.ans[(1 * 8 + b) * G + lane] = c2;
                    }
                }
                bar_arrive(kBarDecision, n_all);
            }

            // lo/hi planes -> floor((lo + hi) / 2) -> bytes -> global (lanes < G, one column group each)
            if (lane < G) {
                uint32_t res[8];
                if (EVEN) {
                    uint32_t carry = 0u, s[9];
#pragma unroll
                    for (int b = 0; b < 8; ++b) {
                        const uint32_t lo = sm.ans[(0 * 8 + b) * G + lane], hi = sm.ans[(1 * 8 + b) * G + lane];
                        s[b] = lo ^ hi ^ carry;
                        carry = (lo & hi) | (carry & (lo ^ hi));
                    }
                    s[8] = carry;
#pragma unroll
                    for (int b = 0; b < 8; ++b) res[b] = s[b + 1];
                } else {
#pragma unroll
                    for (int b = 0; b < 8; ++b) res[b] = sm.ans[(0 * 8 + b) * G + lane];
                }
                bit_transpose8(res);
                const int64_t col0 = (int64_t)ct * tile_w;
                const int width = (int)min((int64_t)tile_w, prm.N - col0);
                uint8_t *out_row = prm.out + prm.vid_out[vid] * prm.N + col0;
                const int offA = (lane >> 3) * kStripBytes + (lane & 7) * 16;
                const int offB = offA + 128;
                if (offA < width) *reinterpret_cast<uint4 *>(out_row + offA) = make_uint4(res[0], res[1], res[2], res[3]);
                if (offB < width) *reinterpret_cast<uint4 *>(out_row + offB) = make_uint4(res[4], res[5], res[6], res[7]);
            }
            __syncwarp();
        }
        return;
    }

    // ================= worker warps =================
    const int strip = warp % S, jc = warp / S;
    const int g = strip * 8 + (lane & 7);
    const int j = jc * 4 + (lane >> 3);
    const int row_base = j * R;
    const uint8_t *my_rows = sm.buf + ((size_t)strip * rows_cap + row_base) * kStripBytes + (lane & 7) * 16;

    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile < prm.num_tiles; tile += gridDim.x) {
        const int64_t vid = tile / prm.tiles_per_video;
        const int T = prm.vid_T[vid];

        mbar_wait(sm.bar, phase);
        phase ^= 1u;

        uint32_t P[R][8];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            if (row_base + i < T) {
                const uint4 a = *reinterpret_cast<const uint4 *>(my_rows + (size_t)i * kStripBytes);
                const uint4 b = *reinterpret_cast<const uint4 *>(my_rows + (size_t)i * kStripBytes + 128);
                P[i][0] = a.x; P[i][1] = a.y; P[i][2] = a.z; P[i][3] = a.w;
                P[i][4] = b.x; P[i][5] = b.y; P[i][6] = b.z; P[i][7] = b.w;
                bit_transpose8(P[i]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) P[i][q] = 0u;
            }
        }
        bar_arrive(kBarTileFree, n_all);           // this thread no longer needs the tile buffer

        uint32_t alive[R], sticky[R], alive2[EVEN ? R : 1];
#pragma unroll
        for (int i = 0; i < R; ++i) {
            alive[i] = (row_base + i < T) ? 0xFFFFFFFFu : 0u;
            sticky[i] = 0u;
            if (EVEN) alive2[i] = alive[i];
        }

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
            sm.red[(par * J + j) * G + g] = make_uint4(c[0], c[1], c[2], c[3]);
            if (EVEN) {
                uint32_t any = 0u;
#pragma unroll
                for (int i = 0; i < R; ++i) any |= alive2[i] & ~P[i][b];
                sm.red_any[(par * J + j) * G + g] = any;
            }
            bar_arrive(kBarPartials, n_all);
            bar_sync(kBarDecision, n_all);
            const uint32_t C1 = sm.dec[(par * 2 + 0) * G + g];
#pragma unroll
            for (int i = 0; i < R; ++i) alive[i] &= ~(P[i][b] ^ C1);
            if (EVEN) {
                const uint32_t C2 = sm.dec[(par * 2 + 1) * G + g];
#pragma unroll
                for (int i = 0; i < R; ++i) alive2[i] &= ~(P[i][b] ^ C2);
            }
        }
    }
}

// ---- host side: plan + launch ----------------------------------------------------------------
struct ClassKey {
    int R, JC, even, npl;
    bool operator<(const ClassKey &o) const
    {
        return std::tie(R, JC, even, npl) < std::tie(o.R, o.JC, o.even, o.npl);
    }
};

struct Tuning {
    int target_r = 8;          // preferred rows per thread
    int target_threads = 256;  // preferred worker threads per CTA
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

int npl_for(int T) { return T < 64 ? 6 : (T < 256 ? 8 : 10); }

template <int R, bool EVEN>
constexpr int max_threads_v() { return Threads<R, EVEN>::kLaunch; }

int max_threads_of(int R, bool even)
{
    switch (R) {
#define C(r) case r: return even ? max_threads_v<r, true>() : max_threads_v<r, false>();
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
    }
    return 0;
}

// Rows per thread R and row-chunk warps JC for a video of T frames: 4*JC*R >= T with the least
// padding at (or above) the preferred R, and a CTA (32*JC worker threads per strip + the service
// warp) the kernel can launch.
ClassKey classify(int T, const Tuning &tn)
{
    const bool even = (T & 1) == 0;
    const int npl = npl_for(T);
    for (int r = std::min(tn.target_r, kMaxR); r <= kMaxR; ++r) {
        const int JC = std::max(1, (T + 4 * r - 1) / (4 * r));
        const int R = std::max(1, (T + 4 * JC - 1) / (4 * JC));
        if (R <= kMaxR && 32 * JC + 32 <= max_threads_of(R, even)) return ClassKey{R, JC, even, npl};
    }
    const int JC = std::max(1, (T + 4 * kMaxR - 1) / (4 * kMaxR));
    return ClassKey{std::min(kMaxR, std::max(1, (T + 4 * JC - 1) / (4 * JC))), JC, even, npl};
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int get_encode_fn(EncodeTiledFn *out)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        BGD_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p)
            return fail(BGD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    *out = fn;
    return BGD_OK;
}

template <int R, bool EVEN, int NPL>
int launch_k(const KParams &prm, int sm_count, int threads, size_t smem, cudaStream_t stream)
{
    auto kern = median_bitsliced_kernel<R, EVEN, NPL>;
    BGD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks_per_sm = 0;
    BGD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, threads, smem));
    if (blocks_per_sm < 1)
        return fail(BGD_ERR_CUDA, "median (bit-sliced): kernel R=%d does not fit an SM (%d threads, %zu B smem)", R,
                    threads, smem);
    const int grid = (int)std::min<int64_t>(prm.num_tiles, (int64_t)sm_count * blocks_per_sm);
    kern<<<grid, threads, smem, stream>>>(prm);
    count_launch();
    BGD_CUDA_TRY(cudaGetLastError());
    return BGD_OK;
}

template <int R>
int launch_r(const ClassKey &k, const KParams &prm, int sm_count, int threads, size_t smem, cudaStream_t stream)
{
    if (k.even) {
        if (k.npl == 6) return launch_k<R, true, 6>(prm, sm_count, threads, smem, stream);
        if (k.npl == 8) return launch_k<R, true, 8>(prm, sm_count, threads, smem, stream);
        return launch_k<R, true, 10>(prm, sm_count, threads, smem, stream);
    }
    if (k.npl == 6) return launch_k<R, false, 6>(prm, sm_count, threads, smem, stream);
    if (k.npl == 8) return launch_k<R, false, 8>(prm, sm_count, threads, smem, stream);
    return launch_k<R, false, 10>(prm, sm_count, threads, smem, stream);
}

}  // namespace

bool median_bitsliced_supports(int64_t T_max, int64_t N)
{
    if (N <= 0 || N % 16 != 0 || N >= ((int64_t)1 << 32)) return false;
    if (T_max < 1 || T_max >= 1024) return false;
    DeviceProps dp;
    if (current_device_props(&dp) != BGD_OK) return false;
    const Tuning tn = read_tuning();
    // both parities up to T_max must fit (a batch mixes them)
    for (int T = (int)std::max<int64_t>(1, T_max - 1); T <= T_max; ++T) {
        const ClassKey k = classify(T, tn);
        if (T > 4 * k.JC * k.R) return false;
        if (32 * k.JC + 32 > max_threads_of(k.R, k.even)) return false;
        if (smem_bytes_for(T, 1, k.JC) > (size_t)dp.smem_optin) return false;
    }
    return true;
}

int median_bitsliced_varlen(const uint8_t *d_frames, const int64_t *h_offsets, int64_t V, int64_t N,
                            uint8_t *d_out, cudaStream_t stream)
{
    if (V == 0 || N == 0) return BGD_OK;
    DeviceProps dp;
    if (int rc = current_device_props(&dp)) return rc;
    const Tuning tn = read_tuning();
    EncodeTiledFn encode = nullptr;
    if (int rc = get_encode_fn(&encode)) return rc;

    // the frames buffer as a 2-D uint8 tensor [rows][N]; one map per box height 2^k
    const int64_t row_lo = h_offsets[0], row_hi = h_offsets[V];
    if (row_hi - row_lo >= ((int64_t)1 << 31)) return fail(BGD_ERR_UNSUPPORTED, "median: more than 2^31 rows per call");
    KParams prm{};
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)(row_hi - row_lo)};
        const cuuint64_t gstride[1] = {(cuuint64_t)N};
        const cuuint32_t estride[2] = {1, 1};
        void *base = const_cast<uint8_t *>(d_frames) + row_lo * N;
        for (int k = 0; k < kNumMaps; ++k) {
            const cuuint32_t box[2] = {(cuuint32_t)kStripBytes, (cuuint32_t)1 << k};
            const CUresult r = encode(&prm.maps[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estride,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(BGD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for box of %d rows", (int)r, 1 << k);
        }
    }

    // bucket the videos by kernel configuration
    std::map<ClassKey, std::vector<int64_t>> classes;
    for (int64_t v = 0; v < V; ++v) classes[classify((int)(h_offsets[v + 1] - h_offsets[v]), tn)].push_back(v);

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
                h_row0[pos] = h_offsets[v] - row_lo;          // row index inside the tensor map
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

        // S in {1, 2, 4}: the reducer maps G = 8*S column groups onto the 32 lanes of the service warp
        const int max_thr = max_threads_of(key.R, key.even != 0);
        const size_t budget = (size_t)dp.smem_optin / tn.ctas_per_sm - 1024;
        int S = 4;
        while (S > 1 && (32 * S * key.JC > std::max(tn.target_threads, 32 * key.JC) || 32 * S * key.JC + 32 > max_thr ||
                         smem_bytes_for(T_cap, S, key.JC) > budget || (int64_t)(S / 2) * kStripBytes >= N))
            S /= 2;
        const size_t smem = smem_bytes_for(T_cap, S, key.JC);
        const int threads = 32 * S * key.JC + 32;
        if (smem > (size_t)dp.smem_optin || threads > max_thr) {
            rc = fail(BGD_ERR_UNSUPPORTED, "median (bit-sliced): T=%d does not fit one CTA", T_cap);
            break;
        }

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

        switch (key.R) {
#define C(r) case r: rc = launch_r<r>(key, prm, dp.sm_count, threads, smem, stream); break;
            C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
            default: rc = fail(BGD_ERR_UNSUPPORTED, "median (bit-sliced): R=%d", key.R);
        }
        if (rc) break;
        pos += nv;
    }
    const int rc2 = ws.release(stream);
    return rc ? rc : rc2;
}

}  // namespace bgd
