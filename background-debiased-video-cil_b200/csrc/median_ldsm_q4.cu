// Instantiations of the transposing-load median kernel, one column per lane: NH = 18, 20, .. 24 half groups
// (9..12 full 32-row groups, 256 < T <= 512).
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

template <int NH>
static int launch_wide(bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    return even ? launch_strips<NH, true, 1>(prm, sm_count, smem, stream) : launch_strips<NH, false, 1>(prm, sm_count, smem, stream);
}

int launch_q4(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NH) {
        case 18: return launch_wide<18>(even, prm, sm_count, smem, stream);
        case 20: return launch_wide<20>(even, prm, sm_count, smem, stream);
        case 22: return launch_wide<22>(even, prm, sm_count, smem, stream);
        case 24: return launch_wide<24>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NH=%d out of range", NH);
}

}  // namespace ldsm
}  // namespace bgd
