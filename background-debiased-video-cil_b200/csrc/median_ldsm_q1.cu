// Instantiations of the transposing-load median kernel, NH = 7..10 half groups (16 rows) per column.
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

int launch_q1(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NH) {
        case 7: return launch_parity<7>(even, prm, sm_count, smem, stream);
        case 8: return launch_parity<8>(even, prm, sm_count, smem, stream);
        case 9: return launch_parity<9>(even, prm, sm_count, smem, stream);
        case 10: return launch_parity<10>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NH=%d out of range", NH);
}

}  // namespace ldsm
}  // namespace bgd
