// Instantiations of the column-plane median kernel for C = 1 byte column(s) per thread.
#include "median_colplane.cuh"

namespace bgd {
namespace colplane {

int launch_c1(int NW, bool even, const CParams &prm, int threads, int sm_count, size_t smem, cudaStream_t stream)
{
    switch (NW) {
        case 1: return launch_parity<1, 1>(even, prm, threads, sm_count, smem, stream);
        case 2: return launch_parity<1, 2>(even, prm, threads, sm_count, smem, stream);
        case 3: return launch_parity<1, 3>(even, prm, threads, sm_count, smem, stream);
        case 4: return launch_parity<1, 4>(even, prm, threads, sm_count, smem, stream);
        case 5: return launch_parity<1, 5>(even, prm, threads, sm_count, smem, stream);
        case 6: return launch_parity<1, 6>(even, prm, threads, sm_count, smem, stream);
        case 7: return launch_parity<1, 7>(even, prm, threads, sm_count, smem, stream);
        case 8: return launch_parity<1, 8>(even, prm, threads, sm_count, smem, stream);
        case 9: return launch_parity<1, 9>(even, prm, threads, sm_count, smem, stream);
        case 10: return launch_parity<1, 10>(even, prm, threads, sm_count, smem, stream);
        case 11: return launch_parity<1, 11>(even, prm, threads, sm_count, smem, stream);
        case 12: return launch_parity<1, 12>(even, prm, threads, sm_count, smem, stream);
        case 13: return launch_parity<1, 13>(even, prm, threads, sm_count, smem, stream);
        case 14: return launch_parity<1, 14>(even, prm, threads, sm_count, smem, stream);
        case 15: return launch_parity<1, 15>(even, prm, threads, sm_count, smem, stream);
        case 16: return launch_parity<1, 16>(even, prm, threads, sm_count, smem, stream);
        case 17: return launch_parity<1, 17>(even, prm, threads, sm_count, smem, stream);
    }
    return fail(BGD_ERR_UNSUPPORTED, "median (column-plane): NW=%d out of range for C=1", NW);
}

}  // namespace colplane
}  // namespace bgd
