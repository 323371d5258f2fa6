// Instantiations of the transposing-load median kernel, NH = 1..6 half groups (16 rows) per column.
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

int launch_q0(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NH) {
        case 1: return launch_parity<1>(even, prm, sm_count, smem, stream);
        case 2: return launch_parity<2>(even, prm, sm_count, smem, stream);
        case 3: return launch_parity<3>(even, prm, sm_count, smem, stream);
        case 4: return launch_parity<4>(even, prm, sm_count, smem, stream);
        case 5: return launch_parity<5>(even, prm, sm_count, smem, stream);
        case 6: return launch_parity<6>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NH=%d out of range", NH);
}

int launch(int NH, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    if (NH <= 6) return launch_q0(NH, even, prm, sm_count, smem, stream);
    if (NH <= 10) return launch_q1(NH, even, prm, sm_count, smem, stream);
    if (NH <= 13) return launch_q2(NH, even, prm, sm_count, smem, stream);
    if (NH <= 16) return launch_q3(NH, even, prm, sm_count, smem, stream);
    if (NH <= 24) return launch_q4(NH, even, prm, sm_count, smem, stream);
    return launch_q5(NH, even, prm, sm_count, smem, stream);
}

}  // namespace ldsm
}  // namespace bgd
