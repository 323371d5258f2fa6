/*
 * CPU oracle (plain C) for the temporal-median background extraction path.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 * Built by oracle/Makefile into oracle/_build/libmedian_oracle.so and loaded by
 * tests/, __graft_entry__.smoke() and bench.py's CPU legs through ctypes.
 *
 * Restates   cil_tools/extract_background.py:73
 *                median_frame = np.median(frames, axis=0).astype(dtype=np.uint8)
 * (and the identical expression at libs/loader/comix_loader.py:161) as exact integer
 * arithmetic:  out[n] = (s[(T-1)/2] + s[T/2]) >> 1,  s = the T values of column n sorted.
 * Parity: pinned through tests/test_oracle_median.py (this file vs. the NumPy restatement
 * vs. the reference's own outputs stored in tests/golden/).
 *
 * The method is a 256-bin counting sort per column -- deliberately a different algorithm
 * from the CUDA kernel's bit-sliced radix select, so the two can check each other.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* frames: [T][N] uint8 row-major; out: [N] uint8.  Returns 0, or -1 on bad arguments. */
int oracle_temporal_median_u8(const uint8_t *frames, int64_t T, int64_t N, uint8_t *out)
{
    if (!frames || !out || T <= 0 || N < 0) return -1;
    const int64_t k_lo = (T - 1) / 2, k_hi = T / 2;
    enum { BLK = 64 };                       /* columns per block: keeps the histograms in L1 */
    uint32_t hist[BLK][256];
    for (int64_t n0 = 0; n0 < N; n0 += BLK) {
        const int64_t nb = (N - n0 < BLK) ? (N - n0) : BLK;
        memset(hist, 0, sizeof(hist));
        for (int64_t t = 0; t < T; ++t) {
            const uint8_t *row = frames + t * N + n0;
            for (int64_t j = 0; j < nb; ++j) hist[j][row[j]]++;
        }
        for (int64_t j = 0; j < nb; ++j) {
            int64_t acc = 0;
            int lo = -1, hi = -1;
            for (int v = 0; v < 256; ++v) {
                acc += hist[j][v];
                if (lo < 0 && acc > k_lo) lo = v;
                if (acc > k_hi) { hi = v; break; }
            }
            out[n0 + j] = (uint8_t)((lo + hi) >> 1);
        }
    }
    return 0;
}

/* Videos concatenated along T: video v owns rows offsets[v] .. offsets[v+1]-1; out: [V][N]. */
int oracle_temporal_median_varlen_u8(const uint8_t *frames, const int64_t *offsets, int64_t V,
                                     int64_t N, uint8_t *out)
{
    if (!offsets || V < 0) return -1;
    for (int64_t v = 0; v < V; ++v) {
        const int64_t t0 = offsets[v], t1 = offsets[v + 1];
        int rc = oracle_temporal_median_u8(frames + t0 * N, t1 - t0, N, out + v * N);
        if (rc) return rc;
    }
    return 0;
}

/*
 * BG-mix blend, restating libs/loader/comix_loader.py:138-145 for one clip:
 *     blend = imgs * (1 - alpha) + bg.view(1,C,H,W) * alpha
 * with imgs = mmaction-normalised foreground (a per-channel 256-entry table built by the
 * caller with the cv2 arithmetic, see oracle/bgmix_oracle.py) and
 * bg = (bg_raw - mean32) / std32 in fp32 (torchvision Normalize, comix_loader.py:74).
 * Every operation is a separately rounded fp32 operation (no FMA): compile with
 * -ffp-contract=off (oracle/Makefile does).
 *
 *   fg      [T][H][W][3] uint8 (RGB, HWC)
 *   bg_crop [3][H][W] fp32, values as produced by Resize+RandomCrop (not yet normalised)
 *   lut     [3][256] fp32
 *   out     [T][3][H][W] fp32
 */
int oracle_bgmix_clip_f32(const uint8_t *fg, const float *bg_crop, const float *lut,
                          const float *bg_mean, const float *bg_std, float one_minus_alpha,
                          float alpha, int apply, int64_t T, int64_t H, int64_t W, float *out)
{
    if (!fg || !lut || !out) return -1;
    const int64_t HW = H * W;
    for (int64_t t = 0; t < T; ++t)
        for (int c = 0; c < 3; ++c)
            for (int64_t p = 0; p < HW; ++p) {
                volatile float f = lut[c * 256 + fg[(t * HW + p) * 3 + c]];
                if (apply) {
                    volatile float b = (bg_crop[c * HW + p] - bg_mean[c]) / bg_std[c];
                    volatile float a = f * one_minus_alpha;
                    volatile float g = b * alpha;
                    f = a + g;
                }
                out[(t * 3 + c) * HW + p] = f;
            }
    return 0;
}
