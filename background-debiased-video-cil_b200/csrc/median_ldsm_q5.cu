// Instantiations of the transposing-load median kernel, one column per lane: NH = 26, 28, .. 32 half groups
// (13..16 full 32-row groups, 256 < T <= 512).
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

template <int NH>
static int launch_wide(bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    return even ? launch_strips<NH, true, 1>(prm, sm_count, smem, stream) : launch_strips<NH, false, 1>(prm, sm_count, smem, stream);
}

int launch_q5(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NH) {
        case 26: return launch_wide<26>(even, prm, sm_count, smem, stream);
        case 28: return launch_wide<28>(even, prm, sm_count, smem, stream);
        case 30: return launch_wide<30>(even, prm, sm_count, smem, stream);
        case 32: return launch_wide<32>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NH=%d out of range", NH);
}

}  // namespace ldsm
}  // namespace bgd
