// Instantiations of the transposing-load median kernel, NH = 14..16 half groups (16 rows) per column.
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

int launch_q3(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NH) {
        case 14: return launch_parity<14>(even, prm, sm_count, smem, stream);
        case 15: return launch_parity<15>(even, prm, sm_count, smem, stream);
        case 16: return launch_parity<16>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NH=%d out of range", NH);
}

}  // namespace ldsm
}  // namespace bgd
