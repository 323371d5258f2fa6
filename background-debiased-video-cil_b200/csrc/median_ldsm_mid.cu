// Instantiations of the transposing-load median kernel, NW = 5..6 plane-word groups per column.
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

int launch_mid(int NW, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NW) {
        case 5: return launch_parity<5>(even, prm, sm_count, smem, stream);
        case 6: return launch_parity<6>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NW=%d out of range", NW);
}

}  // namespace ldsm
}  // namespace bgd
