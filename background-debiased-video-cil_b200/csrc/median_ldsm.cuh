// Temporal median, transposing-load bit-plane select (BGD_MEDIAN_LDSM) -- the AUTO path for T <= 512.
//
// Replaces  np.median(frames, axis=0).astype(np.uint8)   (cil_tools/extract_background.py:73,
// libs/loader/comix_loader.py:161); bit-exact:  out[n] = (s[(T-1)/2] + s[T/2]) >> 1.
//
// Same idea as median_colplane.cuh (one thread owns whole byte columns, so the 8-pass radix
// select needs no communication), rebuilt around what the instruction-throughput probes
// (profiles/r1_microbench_instruction_throughput.txt, r1_microbench2_*.txt) say about B200:
//
//  * LOP3/SHF/PRMT/SEL share one 64-lane/clk pipe and the column-plane kernel saturates it;
//    POPC has its own 16-lane/clk pipe and IMAD runs on the FMA pipe, both idle there.
//  * ldmatrix.m16n16.trans.b8 (SASS LDSM.8.MT1616) hands a lane 4 consecutive *rows* of one byte
//    column in one register -- the shared-memory read and the byte gather the column-plane kernel
//    spends 0.5 LDS + 0.25 PRMT per byte on become 1/16 instruction per byte.
//
// So: a CTA tile is [T rows] x [128 or 256 bytes], staged as 128-byte-wide strips by 2-D TMA tensor
// copies (boxes of 128 B x 2^k rows, 128-byte swizzle, one mbarrier per buffer; a ring of 1-4
// buffers, as many as fit).  The swizzle (16-byte chunk index ^= row % 8) is what makes the
// transposing loads bank-conflict free: 122.7 B/clk/SM measured against 32 B/clk/SM for an
// unswizzled 256-byte pitch.  Two warps share a strip; lane (g = lane / 4, q = lane % 4) owns byte
// columns 16 chunk(q) + g and + 8 of it for ALL T rows, chunk(q) = {0,4,1,5}[q] for the even warp and
// {2,6,3,7}[q] for the odd one.  One ldmatrix.x2 gives it 8 rows of both; 8 registers (32 rows) of a
// column are bit-transposed in place (3 stages of masked shifts) into 8 plane words: word b holds
// bit b of 32 rows.  A column of T rows is NH = ceil(T / 16) half groups: NH / 2 full groups per
// column plus, when NH is odd, one group that the lane's two columns share (16 rows each; the
// columns' alive masks keep them apart), so the work follows T in steps of 16 rows.  For
// 256 < T <= 512 twice the warps read a strip and each lane keeps one of its two columns.
//  * Per pass b = 7..0 and word:  z = alive & plane_b (LOP3), POPC(z) accumulated by IMAD (neither
//    on the LOP3 pipe), and alive &= plane_b ^ keep (LOP3).  The select state is one integer per
//    column, rc = rank - alive count; the result bits accumulate by IMAD.
//  * Even T: the second rank (T/2) shares the first's state until the pass in which they
//    disagree; after that it is the minimum of its own alive set (an OR over the words instead of
//    a count).
//  * As soon as every lane has its columns in registers the CTA's next tile is requested (the warps
//    take turns at it), so the TMA copies land while the select runs on registers; several CTAs per
//    SM overlap the rest.
#pragma once

#include "median_colplane.cuh"

namespace bgd {
namespace ldsm {

constexpr int kStripW = 128;              // bytes per TMA strip (= the 128-byte swizzle span); 2 warps per strip
constexpr int kMaxStrips = 2;             // a CTA tile is 1 or 2 strips wide
constexpr int kMaxNH = 16;                // T <= 256: at most 16 half groups (of 16 rows) per column
#ifndef BGD_LDSM_TWOSEL_MAX
#define BGD_LDSM_TWOSEL_MAX 4             // even T: two independent selects up to this many plane words per column
#endif
#ifndef BGD_LDSM_TREE
#define BGD_LDSM_TREE 0                   // 1: partial sums / partial ORs in two chains (A/B: profiles/r2_ab_tree.txt)
#endif
#ifndef BGD_LDSM_IMADSTATE_MAX
#define BGD_LDSM_IMADSTATE_MAX 4          // select state updated by IMAD (not LOP3) up to this many plane words per column
#endif

struct alignas(64) LParams {
    // frames as [rows][N] uint8, SWIZZLE_128B.  Two-column kernels (NH <= 16): maps[j] has a box of 128 bytes x the
    // j-th frame count of the launch's class (see issue_tile); one-column kernels: maps[k] has a box of 128 bytes x 2^k rows
    CUtensorMap maps[colplane::kNumMaps];
    uint8_t *out;
    const int64_t *vid_row0;     // [n_videos] first row of each video (relative to the map's base)
    const int32_t *vid_T;        // [n_videos]
    const int64_t *vid_out;      // [n_videos] output slot
    const uint32_t *vid_mask;    // [n_videos] alive mask of a column's last plane word (rows beyond T are dead)
    int64_t N;
    int64_t num_tiles;
    int32_t tiles_per_video;
    int32_t rows_cap;            // smem rows per strip (= NH * 16)
    uint32_t one;                // 1, opaque to the compiler: keeps count accumulation on IMAD
    int32_t strips;              // strips per CTA tile (1 or 2): tile width = 128 * strips, block = 64 * strips threads
    int32_t stages;              // tile buffers per CTA (ring); tile i of a CTA lives in buffer i % stages
    int32_t max_blocks_per_sm;   // 0 = as many as fit
    int32_t l2_policy;           // L2 hint of the tensor copies: 0 = evict_first, 1 = evict_normal, 2 = evict_last
};

int launch(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream);
int launch_q0(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream);   // NH 1..6
int launch_q1(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream);   // NH 7..10
int launch_q2(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream);   // NH 11..13
int launch_q3(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream);   // NH 14..16
int launch_q4(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream);   // NH 18..24 (one column per lane)
int launch_q5(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream);   // NH 26..32 (one column per lane)

#ifdef __CUDACC__
using colplane::bit_transpose8;
using colplane::fence_mbar_init;
using colplane::fence_proxy_async;
using colplane::mbar_arrive_expect_tx;
using colplane::mbar_init;
using colplane::mbar_wait;
using colplane::policy_evict_first;
using colplane::smem_u32;
using colplane::tma_load_2d;

// r0/r2: rows +0..3 / +4..7 of the lane's first column, r1/r3: the same rows of its second column
__device__ __forceinline__ void ldsm_x2_trans_b8(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t addr)
{
    asm volatile("ldmatrix.sync.aligned.m16n16.x2.trans.shared.b8 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr)
                 : "memory");
}
// acc + popc(x), the add issued as IMAD (FMA pipe): `one` is 1 at run time
__device__ __forceinline__ int popc_acc(uint32_t x, int acc, uint32_t one)
{
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(__popc(x)), "r"(one), "r"(acc));
    return r;
}
// a * b + c issued as IMAD (FMA pipe); the callers pass run-time operands the compiler cannot fold
__device__ __forceinline__ int imad(int a, int b, int c)
{
    int r;
    asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// (a & mask) | (b & ~mask)
__device__ __forceinline__ int isel(int a, int b, int mask)
{
    int d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
    return d;
}

// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

// NH <= 16 (T <= 256): a lane keeps both byte columns its ldmatrix fragments carry (COLS = 2, 2 warps per
// strip).  NH = 18, 20, .. 32 (256 < T <= 512): 8 .. 16 full groups of one column are all the registers
// hold, so twice the warps read the strip and each keeps one of the two columns (COLS = 1).
template <int NH> __host__ __device__ constexpr int cols_of() { return NH <= kMaxNH ? 2 : 1; }
template <int NH, int STRIPS> __host__ __device__ constexpr int threads_of() { return 64 * STRIPS * (3 - cols_of<NH>()); }
template <int NH, int STRIPS> __host__ __device__ constexpr int min_blocks()
{
    return NH <= 12 ? 8 / STRIPS : (NH <= kMaxNH ? 6 / STRIPS : (NH <= 24 ? 3 : 2) / STRIPS);
}

template <int NH, bool EVEN, int STRIPS>
__global__ void __launch_bounds__(threads_of<NH, STRIPS>(), min_blocks<NH, STRIPS>()) median_ldsm_kernel(const __grid_constant__ LParams prm)
{
    constexpr int COLS = cols_of<NH>();
    static_assert(COLS == 2 || (NH % 2 == 0 && NH <= 32), "one-column mode takes whole 32-row groups, T <= 512");
    constexpr int kTileW = STRIPS * kStripW;
    constexpr int FW = NH / 2;                           // full 32-row groups per column
    constexpr bool HALF = (NH & 1) != 0;                 // plus one group shared by the lane's two columns
    constexpr int NWC = FW + (HALF ? 1 : 0);             // plane words a column's select walks
    extern __shared__ uint8_t smem_raw[];
    // the 128-byte swizzle pattern repeats every 1024 bytes of shared-memory address: align the tile to it
    uint8_t *buf = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int stages = prm.stages;
    const uint32_t stage_bytes = (uint32_t)prm.rows_cap * kTileW;
    uint64_t *bar = reinterpret_cast<uint64_t *>(buf + (size_t)stages * stage_bytes);   // bar[s]: buffer s is full
    // desc[s] = {T | columns of the tile inside the frame << 16, alive mask of the last word, output address of the tile}:
    // written by the lane that requests tile s, read by every lane once the tile has landed -- the consumers keep no
    // (video, column tile) cursor of their own and touch no table in global memory
    uint4 *desc = reinterpret_cast<uint4 *>(bar + 8);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t one = prm.one;

    const uint32_t strip_bytes = (uint32_t)prm.rows_cap * kStripW;
    auto chunk_of = [&](int q) { return ((q & 1) << 2) + (q >> 1) + ((warp & 1) << 1); };
    // ldmatrix address role: lane supplies row r8 = 4 * (lane / 16) + lane % 4 (mod 8) of the 16-byte
    // chunk quarter (lane % 16) / 4 reads; chunk c of row r sits at 16 * (c ^ (r % 8))
    const int r8 = ((lane >> 4) << 2) | (lane & 3);
    const int strip_id = COLS == 2 ? warp >> 1 : warp >> 2;
    const bool second = COLS == 1 && ((warp >> 1) & 1);  // one-column mode: this warp keeps the fragments' second column
    const uint32_t ld_base = smem_u32(buf) + (uint32_t)strip_id * strip_bytes +
                             (uint32_t)(r8 * kStripW + ((chunk_of((lane & 15) >> 2) ^ r8) << 4));
    // data role: columns of the tile owned by this lane
    const int colA = strip_id * kStripW + chunk_of(lane & 3) * 16 + (lane >> 2) + (second ? 8 : 0), colB = colA + 8;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(bar + s, 1);
        fence_mbar_init();
    }
    __syncthreads();

    // tiles are numbered video-major; a CTA walks tile, tile + grid, ...: (video, column tile) advance by a
    // fixed step with one carry, no division in the loop (num_tiles < 2^31 is checked on the host)
    const int tpv = prm.tiles_per_video, num_tiles = (int)prm.num_tiles;
    const int step_vid = (int)gridDim.x / tpv, step_ct = (int)gridDim.x - step_vid * tpv;
    auto advance = [&](int &t, int &v, int &c) {
        t += (int)gridDim.x;
        v += step_vid;
        c += step_ct;
        if (c >= tpv) {
            c -= tpv;
            ++v;
        }
    };

    // The warps of the CTA take turns at requesting tiles (one elected lane issues the TMA copies), so the
    // ~150 instructions of a request are spread evenly instead of making warp 0 the one the others wait for.
    constexpr int kWarps = threads_of<NH, STRIPS>() / 32;
    const bool elected = elect_one();
    // L2 priority of the tensor copies.  Every frame byte is used once, but evict_first is the wrong hint: when frame rows
    // do not start on 128-byte boundaries (N % 128 != 0, e.g. 240x427x3) neighbouring strips share the sectors at their
    // common edge, and a line that is dropped at once is fetched twice from HBM (+14..38 % DRAM reads, -17 % speed on the
    // Sth-Sth-v2 shape); and even for aligned rows the default priority is 1.5-5 % faster (profiles/r2_sweep_l2_policy_*.txt).
    const uint64_t policy = prm.l2_policy == 0 ? policy_evict_first() : (prm.l2_policy == 2 ? colplane::policy_evict_last() : colplane::policy_evict_normal());
    // Called by all lanes of warp 0 (the table reads are broadcast with SHFL so that ptxas sees warp-uniform
    // TMA operands and moves them to uniform registers without a per-value loop); the producer lane issues.
    auto issue_tile = [&](int vid, int ct, int slot) {
        uint64_t *bar_s = bar + slot;
        uint8_t *buf_s = buf + (size_t)slot * stage_bytes;
        const int T = __shfl_sync(0xFFFFFFFFu, prm.vid_T[vid], 0);
        const int row0 = __shfl_sync(0xFFFFFFFFu, (int)prm.vid_row0[vid], 0);
        if (!elected) return;
        {
            const int64_t col0 = (int64_t)ct * kTileW;
            const int64_t left = prm.N - col0;
            const uint64_t dst = reinterpret_cast<uint64_t>(prm.out + prm.vid_out[vid] * prm.N + col0);
            desc[slot] = make_uint4((uint32_t)T | ((uint32_t)(left < kTileW ? left : kTileW) << 16), __ldg(prm.vid_mask + vid),
                                    (uint32_t)dst, (uint32_t)(dst >> 32));
        }
        mbar_arrive_expect_tx(bar_s, (uint32_t)T * (uint32_t)kTileW);
        if constexpr (COLS == 2) {
            // the frame counts of a launch are 16 (NH - 1) + 1 (+ 1 for even T) + 2 j, j = 0..7: maps[j] has a box of
            // exactly that many rows, so a strip is ONE tensor copy whatever T is
            const CUtensorMap *map = &prm.maps[(T - (16 * (NH - 1) + 1 + (EVEN ? 1 : 0))) >> 1];
#pragma unroll
            for (int s = 0; s < kTileW / kStripW; ++s)
                tma_load_2d(buf_s + (size_t)s * strip_bytes, map, ct * kTileW + s * kStripW, row0, bar_s, policy);
        } else {
#pragma unroll
            for (int s = 0; s < kTileW / kStripW; ++s) {
                const int col = ct * kTileW + s * kStripW;
                uint8_t *dst = buf_s + (size_t)s * strip_bytes;
                int r = 0;
                while (T - r >= 256) {                       // T = 512 needs two of them
                    tma_load_2d(dst + (size_t)r * kStripW, &prm.maps[8], col, (int)(row0 + r), bar_s, policy);
                    r += 256;
                }
#pragma unroll
                for (int k = 7; k >= 0; --k)
                    if ((T - r) & (1 << k)) {
                        tma_load_2d(dst + (size_t)r * kStripW, &prm.maps[k], col, (int)(row0 + r), bar_s, policy);
                        r += 1 << k;
                    }
            }
        }
    };

    int tile = blockIdx.x;
    int ptile = tile, pvid = tile / tpv, pct = tile - pvid * tpv;      // the producer's cursor: `stages` tiles ahead
    for (int s = 0; s < stages; ++s) {
        if (warp == (s & (kWarps - 1)) && ptile < num_tiles) issue_tile(pvid, pct, s);
        advance(ptile, pvid, pct);
    }

    int slot = 0;
    int turn = stages;                                   // whose turn it is to request a tile (continues the prologue's rotation)
    uint32_t phases = 0;                                 // bit s: parity to wait for on bar[s]
    for (; tile < num_tiles; tile += (int)gridDim.x) {
        mbar_wait(bar + slot, (phases >> slot) & 1u);
        phases ^= 1u << slot;
        const uint4 d0 = desc[slot];
        const int T = (int)(d0.x & 0xFFFFu);
        const uint32_t last_mask = d0.y;
        const int cols_valid = (int)(d0.x >> 16);
        uint8_t *dst = reinterpret_cast<uint8_t *>(((uint64_t)d0.w << 32) | d0.z);
        const uint32_t ld_slot = ld_base + (uint32_t)slot * stage_bytes;

        // ---- shared memory -> registers (transposing loads), then 8x8 bit transposes -----------
        // P[c][k][m] before the transpose: rows 32 k + 4 m .. + 3 of column c, one per byte
        // PS[m]: rows 32 FW + 4 m .. + 3 of column 0 (m < 4) / rows 32 FW + 4 (m - 4) .. + 3 of column 1 (m >= 4)
        uint32_t P[COLS][FW > 0 ? FW : 1][8], PS[8];
#pragma unroll
        for (int k = 0; k < FW; ++k) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if constexpr (COLS == 2) {
                    ldsm_x2_trans_b8(P[0][k][2 * j], P[1][k][2 * j], P[0][k][2 * j + 1], P[1][k][2 * j + 1],
                                     ld_slot + (uint32_t)((k * 32 + j * 8) * kStripW));
                } else {
                    uint32_t a0, b0, a1, b1;
                    ldsm_x2_trans_b8(a0, b0, a1, b1, ld_slot + (uint32_t)((k * 32 + j * 8) * kStripW));
                    P[0][k][2 * j] = second ? b0 : a0;
                    P[0][k][2 * j + 1] = second ? b1 : a1;
                }
            }
        }
        if (HALF) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                ldsm_x2_trans_b8(PS[2 * j], PS[4 + 2 * j], PS[2 * j + 1], PS[4 + 2 * j + 1],
                                 ld_slot + (uint32_t)((FW * 32 + j * 8) * kStripW));
        }
        __syncthreads();                                 // every lane has its columns: buffer is free
        {
            if (warp == (turn & (kWarps - 1)) && ptile < num_tiles) {
                if (elected) fence_proxy_async();
                issue_tile(pvid, pct, slot);
            }
            advance(ptile, pvid, pct);
            ++turn;
            slot = slot + 1 == stages ? 0 : slot + 1;
        }
#pragma unroll
        for (int k = 0; k < FW; ++k) {
#pragma unroll
            for (int c = 0; c < COLS; ++c) bit_transpose8(P[c][k]);
        }
        if (HALF) bit_transpose8(PS);

        // ---- alive mask of a column's last word (from the per-video table): bit 8 y + m is row 32 (NWC-1) + 4 m + y;
        //      in the shared group column 0 owns bits m < 4 of every byte and column 1 (mask << 4) the bits m >= 4 -------

        // ---- 8-pass MSB-first rank select, per column ------------------------------------------------
        // State of a select: rc = rank - (number of alive rows), always negative, where rank is the 0-based rank of the
        // wanted element among the alive rows.  With `ones` alive rows having the bit set, d = rc + ones = rank - zeros
        // decides the bit (0 iff d < 0); the new rc is d for bit 0 (the zeros stay alive) and rc for bit 1 (rank and the
        // alive count both drop by zeros), i.e. rc - ones * m0 with m0 = d >> 31: rank itself never has to be kept and
        // the update is an IMAD.  acc collects  sum_b m_b 2^b  with m_b = -1 where the bit is 0, so the value is 255 + acc.
        // Everything but the two LOP3 per word and the sign shift runs on the FMA and POPC pipes; the LOP3 pipe is the
        // one this kernel fills.  Even T runs the select twice, for ranks T/2 - 1 and T/2: two independent chains cost
        // the LOP3 pipe 4 per word and pass like a shared count with a second set, but no shared-state bookkeeping.
        int med[2] = {0, 0};
        const int ONE = (int)one, NEG1 = -(int)one, TWO = (int)one << 1;
        // Few plane words per column (short videos): the per-pass bookkeeping weighs as much as the word work, so it is
        // kept off the LOP3 pipe entirely -- even T runs two independent selects (ranks T/2 - 1 and T/2: one more POPC
        // and IMAD per word, no shared-state logic) and rc is updated by IMAD.  Many words per column: issue slots are
        // what is short, so the two middles share one count until they part and the upper one then follows the
        // minimum of its own set (an OR over the words instead of a count).
        constexpr bool kTwoSelects = EVEN && NWC <= BGD_LDSM_TWOSEL_MAX;
        constexpr bool kImadState = NWC <= BGD_LDSM_IMADSTATE_MAX;
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            auto plane = [&](int k, int b) -> uint32_t { return (HALF && k == FW) ? PS[b] : P[c][k < FW ? k : 0][b]; };
            uint32_t alive[NWC], alive2[EVEN ? NWC : 1];
#pragma unroll
            for (int k = 0; k < NWC; ++k) {
                alive[k] = k == NWC - 1 ? (HALF ? last_mask << (4 * c) : last_mask) : 0xFFFFFFFFu;
                if (EVEN) alive2[k] = alive[k];
            }
            int rc = ((T - 1) >> 1) - T, rc2 = (T >> 1) - T;
            int lo_acc = 0, hi_acc = 0;
            int diverged = 0;                            // shared count: all-ones once the two middles sit in different sets
#pragma unroll
            for (int b = 7; b >= 0; --b) {
                int d, m0;
                if constexpr (kImadState) {
                    int nones = 0;                       // minus the number of alive rows with the bit set
#pragma unroll
                    for (int k = 0; k < NWC; ++k) nones = imad(__popc(alive[k] & plane(k, b)), NEG1, nones);
                    d = imad(nones, NEG1, rc);
                    m0 = d >> 31;                        // all-ones: the bit is 0
                    rc = imad(nones, m0, rc);
                } else {
#if BGD_LDSM_TREE
                    // two partial sums: the IMAD chain of a pass is half as long (4 warps per scheduler do not hide it all)
                    d = rc;
                    int d1 = 0;
#pragma unroll
                    for (int k = 0; k < NWC; ++k) {
                        if (k & 1) d1 = popc_acc(alive[k] & plane(k, b), d1, one);
                        else d = popc_acc(alive[k] & plane(k, b), d, one);
                    }
                    d = imad(d1, ONE, d);
#else
                    d = rc;
#pragma unroll
                    for (int k = 0; k < NWC; ++k) d = popc_acc(alive[k] & plane(k, b), d, one);
#endif
                    m0 = d >> 31;
                    rc = isel(d, rc, m0);
                }
                lo_acc = imad(lo_acc, TWO, m0);
                if constexpr (kTwoSelects) {
                    int nones2 = 0;
#pragma unroll
                    for (int k = 0; k < NWC; ++k) nones2 = imad(__popc(alive2[k] & plane(k, b)), NEG1, nones2);
                    const int m2 = imad(nones2, NEG1, rc2) >> 31;
                    rc2 = imad(nones2, m2, rc2);
                    hi_acc = imad(hi_acc, TWO, m2);
#pragma unroll
                    for (int k = 0; k < NWC; ++k) alive2[k] &= plane(k, b) ^ (uint32_t)m2;
                } else if constexpr (EVEN) {
                    uint32_t any0 = 0u;                  // rows of the upper middle's set whose bit is 0
#if BGD_LDSM_TREE
                    uint32_t any1 = 0u;
#pragma unroll
                    for (int k = 0; k < NWC; ++k) {
                        if (k & 1) any1 |= alive2[k] & ~plane(k, b);
                        else any0 |= alive2[k] & ~plane(k, b);
                    }
                    any0 |= any1;
#else
#pragma unroll
                    for (int k = 0; k < NWC; ++k) any0 |= alive2[k] & ~plane(k, b);
#endif
                    // shared state: the upper middle has rank + 1, its bit is 0 iff d + 1 < 0; own state: iff any0
                    const int m0_shared = imad(ONE, ONE, d) >> 31;
                    const int m0_own = imad(__popc(any0), NEG1, 0) >> 31;
                    const int m2 = isel(m0_own, m0_shared, diverged);
                    diverged |= m2 ^ m0;
                    hi_acc = imad(hi_acc, TWO, m2);
#pragma unroll
                    for (int k = 0; k < NWC; ++k) alive2[k] &= plane(k, b) ^ (uint32_t)m2;
                }
#pragma unroll
                for (int k = 0; k < NWC; ++k) alive[k] &= plane(k, b) ^ (uint32_t)m0;
            }
            const int lo = 255 + lo_acc, hi = 255 + hi_acc;
            med[c] = EVEN ? ((lo + hi) >> 1) : lo;
        }

        // ---- store ---------------------------------------------------------------------------------------
        if (colA < cols_valid) dst[colA] = (uint8_t)med[0];
        if (COLS == 2 && colB < cols_valid) dst[colB] = (uint8_t)med[1];
    }
}

template <int NH, bool EVEN, int STRIPS>
int launch_strips(const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    constexpr int kThreads = threads_of<NH, STRIPS>();
    auto kern = median_ldsm_kernel<NH, EVEN, STRIPS>;
    BGD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BGD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    int blocks_per_sm = 0;
    BGD_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, kThreads, smem));
    if (blocks_per_sm < 1)
        return fail(BGD_ERR_CUDA, "median (ldsm): kernel NH=%d does not fit an SM (%zu B smem)", NH, smem);
    if (prm.max_blocks_per_sm > 0 && blocks_per_sm > prm.max_blocks_per_sm) blocks_per_sm = prm.max_blocks_per_sm;
    const int64_t cap = (int64_t)sm_count * blocks_per_sm;
    const int grid = (int)(prm.num_tiles < cap ? prm.num_tiles : cap);
    kern<<<grid, kThreads, smem, stream>>>(prm);
    count_launch();
    BGD_CUDA_TRY(cudaGetLastError());
    return BGD_OK;
}

template <int NH, bool EVEN>
int launch_one(const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    return prm.strips == 2 ? launch_strips<NH, EVEN, 2>(prm, sm_count, smem, stream)
                           : launch_strips<NH, EVEN, 1>(prm, sm_count, smem, stream);
}

template <int NH>
int launch_parity(bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    return even ? launch_one<NH, true>(prm, sm_count, smem, stream) : launch_one<NH, false>(prm, sm_count, smem, stream);
}
#endif  // __CUDACC__

}  // namespace ldsm
}  // namespace bgd
