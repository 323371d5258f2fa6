// Instantiations of the transposing-load median kernel, NW = 7..8 plane-word groups per column.
#include "median_ldsm.cuh"

namespace bgd {
namespace ldsm {

int launch_hi(int NW, bool even, const LParams &prm, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NW) {
        case 7: return launch_parity<7>(even, prm, sm_count, smem, stream);
        case 8: return launch_parity<8>(even, prm, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (ldsm): NW=%d out of range", NW);
}

}  // namespace ldsm
}  // namespace bgd
