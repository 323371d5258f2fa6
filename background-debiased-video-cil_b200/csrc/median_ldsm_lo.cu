// Instantiations of the transposing-load median kernel, NW = 1..4 plane-word groups per column.
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

int launch_lo(int NW, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NW) {
        case 1: return launch_parity<1>(even, prm, sm_count, smem, stream);
        case 2: return launch_parity<2>(even, prm, sm_count, smem, stream);
        case 3: return launch_parity<3>(even, prm, sm_count, smem, stream);
        case 4: return launch_parity<4>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NW=%d out of range", NW);
}

int launch(int NW, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    if (NW <= 4) return launch_lo(NW, even, prm, sm_count, smem, stream);
    if (NW <= 6) return launch_mid(NW, even, prm, sm_count, smem, stream);
    return launch_hi(NW, even, prm, sm_count, smem, stream);
}

}  // namespace ldsm
}  // namespace bgd
