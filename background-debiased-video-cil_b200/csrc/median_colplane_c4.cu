// Instantiations of the column-plane median kernel for C = 4 byte column(s) per thread.
#include "median_colplane.cuh"

namespace bgd {
namespace colplane {

int launch_c4(int NW, bool even, const CParams &prm, int threads, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NW) {
        case 1: return launch_parity<4, 1>(even, prm, threads, sm_count, smem, stream);
        case 2: return launch_parity<4, 2>(even, prm, threads, sm_count, smem, stream);
        case 3: return launch_parity<4, 3>(even, prm, threads, sm_count, smem, stream);
        case 4: return launch_parity<4, 4>(even, prm, threads, sm_count, smem, stream);
        case 5: return launch_parity<4, 5>(even, prm, threads, sm_count, smem, stream);
        case 6: return launch_parity<4, 6>(even, prm, threads, sm_count, smem, stream);
        case 7: return launch_parity<4, 7>(even, prm, threads, sm_count, smem, stream);
        case 8: return launch_parity<4, 8>(even, prm, threads, sm_count, smem, stream);
        case 9: return launch_parity<4, 9>(even, prm, threads, sm_count, smem, stream);
        case 10: return launch_parity<4, 10>(even, prm, threads, sm_count, smem, stream);
        case 11: return launch_parity<4, 11>(even, prm, threads, sm_count, smem, stream);
        case 12: return launch_parity<4, 12>(even, prm, threads, sm_count, smem, stream);
        case 13: return launch_parity<4, 13>(even, prm, threads, sm_count, smem, stream);
        case 14: return launch_parity<4, 14>(even, prm, threads, sm_count, smem, stream);
        case 15: return launch_parity<4, 15>(even, prm, threads, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (column-plane): NW=%d out of range for C=4", NW);
}

}  // namespace colplane
}  // namespace bgd
