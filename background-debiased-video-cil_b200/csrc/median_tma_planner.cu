// Host side of the two TMA-staged medians (median_ldsm.cuh for T <= 512, median_colplane.cuh above): bucket
// the videos of a call by kernel configuration (kernel family, plane groups, parity), build the TMA tensor
// maps, upload one table of (first row, frame count, output slot) per video and launch one kernel per bucket.
#include <algorithm>
#include <map>
#include <tuple>
#include <vector>

#include "median_colplane.cuh"
#include "median_ldsm.cuh"

namespace bgd {

namespace {

using colplane::CParams;
using colplane::kNumMaps;
using colplane::kStripBytes;

struct Key {
    int C, NW, even;   // C == 0: the transposing-load kernel (median_ldsm.cuh); NW then counts half groups of 16 rows
    bool operator<(const Key &o) const { return std::tie(C, NW, even) < std::tie(o.C, o.NW, o.even); }
};

struct Tuning {
    int t_c4 = 96;     // T <= t_c4: 4 columns per thread (8 rows per plane word)
    int t_c2 = 240;    // T <= t_c2: 2 columns per thread (16 rows per plane word)
    int threads_c1 = 256, threads_c2 = 128, threads_c4 = 128;
    int ldsm_strips = 0;   // 128-byte strips per CTA tile of the transposing-load kernel; 0 = by parity (see below)
    int ldsm_stages = 0;   // tile buffers per CTA; 0 = as many (up to 4) as fit beside the target CTA count
    int ldsm_blocks = 0;   // CTAs per SM; 0 = register limit (8 / strips for T <= 192, 6 / strips above)
    int l2_policy = -1;    // -1 = evict_normal (see median_ldsm.cuh), else 0 / 1 / 2
    int l2_promo = -1;     // -1 = 256 B for aligned rows, 128 B otherwise; else 0, 64, 128, 256
};

Tuning read_tuning()
{
    Tuning t;
    if (const char *s = getenv("BGD_COL_T_C4")) t.t_c4 = atoi(s);
    if (const char *s = getenv("BGD_COL_T_C2")) t.t_c2 = atoi(s);
    if (const char *s = getenv("BGD_LDSM_STRIPS")) { const int v = atoi(s); t.ldsm_strips = v >= 2 ? 2 : (v == 1 ? 1 : 0); }
    if (const char *s = getenv("BGD_LDSM_STAGES")) t.ldsm_stages = std::max(0, std::min(8, atoi(s)));
    if (const char *s = getenv("BGD_LDSM_BLOCKS")) t.ldsm_blocks = std::max(0, atoi(s));
    if (const char *s = getenv("BGD_TMA_POLICY")) t.l2_policy = std::max(-1, std::min(2, atoi(s)));
    if (const char *s = getenv("BGD_TMA_L2PROMO")) t.l2_promo = atoi(s);
    if (const char *s = getenv("BGD_COL_THREADS_C2")) t.threads_c2 = atoi(s) >= 256 ? 256 : 128;
    if (const char *s = getenv("BGD_COL_THREADS_C4")) { const int v = atoi(s); t.threads_c4 = v >= 256 ? 256 : (v >= 192 ? 192 : (v >= 128 ? 128 : 64)); }
    return t;
}

bool classify(int T, const Tuning &tn, bool use_ldsm, Key *k)
{
    if (use_ldsm && T <= 16 * ldsm::kMaxNH) {
        *k = Key{0, (T + 15) / 16, (T & 1) == 0};
        return true;
    }
    if (use_ldsm && T <= 512) {                       // one column per lane: whole 32-row groups
        *k = Key{0, 2 * ((T + 31) / 32), (T & 1) == 0};
        return true;
    }
    int C = T <= tn.t_c4 ? 4 : (T <= tn.t_c2 ? 2 : 1);
    for (; C >= 1; C /= 2) {
        const int rpw = 32 / C;
        const int NW = (T + rpw - 1) / rpw;
        if (NW <= (C == 1 ? 17 : 15)) {
            *k = Key{C, NW, (T & 1) == 0};
            return true;
        }
    }
    return false;
}

// Alive mask of the last plane word of a column in the transposing-load kernel (median_ldsm.cuh): NH half groups of 16
// rows, NWC = ceil(NH / 2) words; bit 8 y + m of a word is row 32 (NWC - 1) + 4 m + y of the column; when NH is odd the
// last word is shared by the lane's two columns, 16 rows (m < 4) each, and the kernel shifts the mask by 4 for the second.
uint32_t ldsm_last_mask(int T, int NH)
{
    const bool half = (NH & 1) != 0;
    const int nwc = NH / 2 + (half ? 1 : 0);
    const int n = T - 32 * (nwc - 1), cap = half ? 4 : 8;
    uint32_t mask = 0u;
    for (int y = 0; y < 4; ++y) {
        int cnt = (n - y + 3) >> 2;
        cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
        mask |= ((1u << cnt) - 1u) << (8 * y);
    }
    return mask;
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

}  // namespace

bool median_tma_supports(int64_t T_max, int64_t N)
{
    if (N <= 0 || N % 16 != 0 || N >= ((int64_t)1 << 31)) return false;
    if (T_max < 1 || T_max > 544) return false;
    Key k;
    return classify((int)T_max, read_tuning(), false, &k);
}

int median_tma_varlen(const uint8_t *d_frames, const int64_t *h_offsets, int64_t V, int64_t N, uint8_t *d_out,
                           bool use_ldsm, const int64_t *subset, int64_t n_subset, cudaStream_t stream)
{
    if (V == 0 || N == 0) return BGD_OK;
    if (subset && n_subset == 0) return BGD_OK;
    const int64_t n_vid = subset ? n_subset : V;
    DeviceProps dp;
    if (int rc = current_device_props(&dp)) return rc;
    const Tuning tn = read_tuning();
    EncodeTiledFn encode = nullptr;
    if (int rc = get_encode_fn(&encode)) return rc;

    const int64_t row_lo = h_offsets[0], row_hi = h_offsets[V];
    if (row_hi - row_lo >= ((int64_t)1 << 31)) return fail(BGD_ERR_UNSUPPORTED, "median: more than 2^31 rows per call");
    std::map<Key, std::vector<int64_t>> classes;
    bool any_ldsm = false, any_col = false;
    for (int64_t i = 0; i < n_vid; ++i) {
        const int64_t v = subset ? subset[i] : i;
        Key k;
        if (!classify((int)(h_offsets[v + 1] - h_offsets[v]), tn, use_ldsm, &k))
            return fail(BGD_ERR_UNSUPPORTED, "median (column-plane): video %lld has too many frames", (long long)v);
        classes[k].push_back(v);
        (k.C == 0 ? any_ldsm : any_col) = true;
    }

    CParams prm{};
    ldsm::LParams lprm{};
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)(row_hi - row_lo)};
        const cuuint64_t gstride[1] = {(cuuint64_t)N};
        const cuuint32_t estride[2] = {1, 1};
        void *base = const_cast<uint8_t *>(d_frames) + row_lo * N;
        for (int k = 0; k < kNumMaps; ++k) {
            if (any_col) {
                const cuuint32_t box[2] = {(cuuint32_t)kStripBytes, (cuuint32_t)1 << k};
                const CUresult r = encode(&prm.maps[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, gdim, gstride, box, estride,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS)
                    return fail(BGD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a box of %d rows", (int)r, 1 << k);
            }
        }
    }
    // swizzled maps of the transposing-load kernel: built per launch below (the box height depends on the class)
    const cuuint64_t l_gdim[2] = {(cuuint64_t)N, (cuuint64_t)(row_hi - row_lo)};
    const cuuint64_t l_gstride[1] = {(cuuint64_t)N};
    const cuuint32_t l_estride[2] = {1, 1};
    void *l_base = const_cast<uint8_t *>(d_frames) + row_lo * N;
    const bool rows_aligned = N % 128 == 0 && reinterpret_cast<uintptr_t>(l_base) % 128 == 0;
    const int promo_bytes = tn.l2_promo >= 0 ? tn.l2_promo : (rows_aligned ? 256 : 128);
    const CUtensorMapL2promotion promo = promo_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                         : (promo_bytes >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                            : (promo_bytes >= 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE));
    auto ldsm_map = [&](CUtensorMap *out, int box_rows) -> int {
        const cuuint32_t box[2] = {(cuuint32_t)ldsm::kStripW, (cuuint32_t)box_rows};
        const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, l_base, l_gdim, l_gstride, box, l_estride,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return fail(BGD_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a swizzled box of %d rows", (int)r, box_rows);
        return BGD_OK;
    };
    (void)any_ldsm;

    // one table upload for all classes: row0[V] | out[V] | T[V], in class order
    Workspace &ws = thread_workspace();
    const size_t tbl_bytes = (size_t)n_vid * (8 + 8 + 4 + 4);
    if (int rc = ws.acquire(tbl_bytes)) return rc;
    int64_t *h_row0 = static_cast<int64_t *>(ws.h_pinned);
    int64_t *h_out = h_row0 + n_vid;
    int32_t *h_T = reinterpret_cast<int32_t *>(h_out + n_vid);
    uint32_t *h_mask = reinterpret_cast<uint32_t *>(h_T + n_vid);
    const int64_t *d_row0 = static_cast<const int64_t *>(ws.d_ptr);
    const int64_t *d_outi = d_row0 + n_vid;
    const int32_t *d_T = reinterpret_cast<const int32_t *>(d_outi + n_vid);
    const uint32_t *d_mask = reinterpret_cast<const uint32_t *>(d_T + n_vid);
    {
        int64_t pos = 0;
        for (auto &kv : classes)
            for (int64_t v : kv.second) {
                h_row0[pos] = h_offsets[v] - row_lo;
                h_out[pos] = v;
                h_T[pos] = (int32_t)(h_offsets[v + 1] - h_offsets[v]);
                h_mask[pos] = kv.first.C == 0 ? ldsm_last_mask(h_T[pos], kv.first.NW) : 0u;
                ++pos;
            }
    }
    BGD_CUDA_TRY(cudaMemcpyAsync(ws.d_ptr, ws.h_pinned, tbl_bytes, cudaMemcpyHostToDevice, stream));

    int64_t pos = 0;
    int rc = BGD_OK;
    for (auto &kv : classes) {
        const Key key = kv.first;
        const int64_t nv = (int64_t)kv.second.size();
        if (key.C == 0) {
            lprm.out = d_out;
            lprm.vid_row0 = d_row0 + pos;
            lprm.vid_T = d_T + pos;
            lprm.vid_out = d_outi + pos;
            lprm.vid_mask = d_mask + pos;
            lprm.N = N;
            // odd videos of more than 160 frames run close to the HBM limit and gain from wider rows per copy (+8 % at T = 181);
            // even T (bound by the LOP3 pipe) and short videos gain from the finer-grained CTAs (+3..6 %)
            // (profiles/r1_sweep_ldsm_strips.txt)
            lprm.strips = tn.ldsm_strips > 0 ? tn.ldsm_strips : ((!key.even && key.NW >= 11) ? 2 : 1);
            if (key.NW > ldsm::kMaxNH) lprm.strips = 1;   // one-column mode is instantiated for one strip
            const int tile_w = lprm.strips * ldsm::kStripW;
            lprm.tiles_per_video = (int32_t)((N + tile_w - 1) / tile_w);
            lprm.num_tiles = nv * lprm.tiles_per_video;
            if (lprm.num_tiles >= ((int64_t)1 << 31)) {
                rc = fail(BGD_ERR_UNSUPPORTED, "median (ldsm): %lld tiles in one call; split the batch", (long long)lprm.num_tiles);
                break;
            }
            lprm.rows_cap = key.NW * 16;
            lprm.one = 1u;
            lprm.l2_policy = tn.l2_policy >= 0 ? tn.l2_policy : 1;
            if (key.NW <= ldsm::kMaxNH) {
                // one box of exactly T rows per frame count present in this class: T = 16 (NW - 1) + 1 (+ 1) + 2 j
                const int t_lo = 16 * (key.NW - 1) + 1 + (key.even ? 1 : 0);
                bool present[8] = {false, false, false, false, false, false, false, false};
                for (int64_t i = 0; i < nv; ++i) present[(h_T[pos + i] - t_lo) >> 1] = true;
                for (int j = 0; j < 8 && rc == BGD_OK; ++j)
                    if (present[j]) rc = ldsm_map(&lprm.maps[j], t_lo + 2 * j);
            } else {
                for (int k = 0; k < kNumMaps && rc == BGD_OK; ++k) rc = ldsm_map(&lprm.maps[k], 1 << k);
            }
            if (rc) break;
            // buffers + mbarriers + slack to align the buffers to the 1024-byte swizzle atom; as many stages as
            // fit beside the CTA count the registers allow (1 KB per CTA is reserved by the driver)
            const size_t tile_bytes = (size_t)lprm.rows_cap * tile_w;
            // CTAs per SM: what the registers allow (8 x 64 threads up to 192 frames, then 6, 3, 2); short videos
            // need few registers and little shared memory and gain 3-5 % from up to 16 (profiles/r1_sweep_ldsm_blocks.txt)
            const int blocks = tn.ldsm_blocks > 0 ? tn.ldsm_blocks
                               : (key.NW <= 6 ? 16 : (key.NW <= 12 ? 8 : (key.NW <= ldsm::kMaxNH ? 6 : (key.NW <= 24 ? 3 : 2)))) / lprm.strips;
            const size_t per_block = (size_t)dp.smem_per_sm / blocks - 1024;
            int stages = tn.ldsm_stages > 0 ? tn.ldsm_stages : (int)((per_block - 1024 - 64 - 128) / tile_bytes);
            stages = std::max(1, std::min(stages, tn.ldsm_stages > 0 ? 8 : 4));
            lprm.stages = stages;
            lprm.max_blocks_per_sm = blocks;
            const size_t smem = (size_t)stages * tile_bytes + 64 + 128 + 1024;       // tiles, 8 mbarriers, 8 tile descriptors of 16 bytes, alignment slack
            rc = ldsm::launch(key.NW, key.even != 0, lprm, dp.sm_count, smem, stream);
            if (rc) break;
            pos += nv;
            continue;
        }
        const int threads = key.C == 1 ? tn.threads_c1 : (key.C == 2 ? tn.threads_c2 : tn.threads_c4);
        const int tile_w = key.C * threads;
        const int rows_cap = key.NW * (32 / key.C);
        const size_t smem = (size_t)rows_cap * tile_w + 16;
        if (smem > (size_t)dp.smem_optin) {
            rc = fail(BGD_ERR_UNSUPPORTED, "median (column-plane): tile of %d rows x %d bytes exceeds shared memory", rows_cap, tile_w);
            break;
        }
        prm.out = d_out;
        prm.vid_row0 = d_row0 + pos;
        prm.vid_T = d_T + pos;
        prm.vid_out = d_outi + pos;
        prm.N = N;
        prm.tiles_per_video = (int32_t)((N + tile_w - 1) / tile_w);
        prm.num_tiles = nv * prm.tiles_per_video;
        prm.rows_cap = rows_cap;
        if (key.C == 1) rc = colplane::launch_c1(key.NW, key.even != 0, prm, threads, dp.sm_count, smem, stream);
        else if (key.C == 2) rc = colplane::launch_c2(key.NW, key.even != 0, prm, threads, dp.sm_count, smem, stream);
        else rc = colplane::launch_c4(key.NW, key.even != 0, prm, threads, dp.sm_count, smem, stream);
        if (rc) break;
        pos += nv;
    }
    const int rc2 = ws.release(stream);
    return rc ? rc : rc2;
}

}  // namespace bgd
