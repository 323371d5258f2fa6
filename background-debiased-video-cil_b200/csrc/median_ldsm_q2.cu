// Instantiations of the transposing-load median kernel, NH = 11..13 half groups (16 rows) per column.
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

int launch_q2(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NH) {
        case 11: return launch_parity<11>(even, prm, sm_count, smem, stream);
        case 12: return launch_parity<12>(even, prm, sm_count, smem, stream);
        case 13: return launch_parity<13>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NH=%d out of range", NH);
}

}  // namespace ldsm
}  // namespace bgd
