// ActorCutMix blend:
//     actor_cut_mix = actor_img * actor_mask + scene_img * (1 - actor_mask)       libs/loader/actor_cut_mix_loader.py:143-148
//     foreground_ratio = sum_t human_mask[t][:, :, 0].sum() / (T * w * h)          :154-163
// on uint8 [T][H][W][3] frames and 0/1 masks.  Every operation is numpy's uint8 arithmetic (wrapping), kept exact
// for any mask value.  Pure streaming work: 3 bytes read and 1 written per element, 16 bytes per thread and access.
#include <algorithm>
#include <initializer_list>

#include "bgd_common.cuh"

namespace bgd {
namespace {

__device__ __forceinline__ uint32_t mix4(uint32_t a, uint32_t m, uint32_t s)
{
    if ((m & 0xFEFEFEFEu) == 0u) {                       // 0/1 masks (what the box pipeline builds): a byte select
        const uint32_t sel = m * 0xFFu;                  // 0x01 -> 0xFF per byte, no carries
        return (a & sel) | (s & ~sel);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t ab = (a >> (8 * i)) & 0xFF, mb = (m >> (8 * i)) & 0xFF, sb = (s >> (8 * i)) & 0xFF;
        const uint32_t v = (((ab * mb) & 0xFF) + ((sb * ((1u - mb) & 0xFF)) & 0xFF)) & 0xFF;
        r |= v << (8 * i);
    }
    return r;
}

#ifndef BGD_CUTMIX_UNROLL
#define BGD_CUTMIX_UNROLL 2
#endif
constexpr int kCutmixUnroll = BGD_CUTMIX_UNROLL;

// n16 16-byte vectors, then a scalar tail; mask_sum accumulates the mask bytes of channel 0 (element index % 3 == 0)
__global__ void __launch_bounds__(256) cutmix_kernel(const uint8_t *__restrict__ actor, const uint8_t *__restrict__ mask,
                                                     const uint8_t *__restrict__ scene, uint8_t *__restrict__ out, int64_t n,
                                                     unsigned long long *__restrict__ mask_sum)
{
    const int64_t n16 = n / 16;
    unsigned long long local = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto sum_channel0 = [&](const uint4 &m, int64_t i) {
        // channel-0 bytes of this vector: element index = 16 i + b with (16 i + b) % 3 == 0, i.e. b % 3 == (3 - i % 3) % 3;
        // one byte-selector word per 4 bytes and phase, summed with dp4a
        const int ph = (int)(i % 3);
        const uint32_t s0 = ph == 0 ? 0x01000001u : (ph == 1 ? 0x00010000u : 0x00000100u);   // bytes 0..3
        const uint32_t s1 = ph == 0 ? 0x00010000u : (ph == 1 ? 0x00000100u : 0x01000001u);   // bytes 4..7
        const uint32_t s2 = ph == 0 ? 0x00000100u : (ph == 1 ? 0x01000001u : 0x00010000u);   // bytes 8..11
        uint32_t acc = __dp4a(m.x, s0, 0u);
        acc = __dp4a(m.y, s1, acc);
        acc = __dp4a(m.z, s2, acc);
        acc = __dp4a(m.w, s0, acc);                      // bytes 12..15 repeat the pattern of bytes 0..3
        local += acc;
    };
    auto mix16 = [](const uint4 &a, const uint4 &m, const uint4 &s) {
        return make_uint4(mix4(a.x, m.x, s.x), mix4(a.y, m.y, s.y), mix4(a.z, m.z, s.z), mix4(a.w, m.w, s.w));
    };
    const uint4 *va = reinterpret_cast<const uint4 *>(actor), *vm = reinterpret_cast<const uint4 *>(mask),
                *vs = reinterpret_cast<const uint4 *>(scene);
    uint4 *vo = reinterpret_cast<uint4 *>(out);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kCutmixUnroll - 1) * stride < n16; i += kCutmixUnroll * stride) {     // kCutmixUnroll vectors of each stream in flight
        uint4 a[kCutmixUnroll], m[kCutmixUnroll], s[kCutmixUnroll];
#pragma unroll
        for (int u = 0; u < kCutmixUnroll; ++u) {
            a[u] = __ldcs(va + i + u * stride); m[u] = __ldcs(vm + i + u * stride); s[u] = __ldcs(vs + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < kCutmixUnroll; ++u) {
            __stcs(vo + i + u * stride, mix16(a[u], m[u], s[u]));
            if (mask_sum) sum_channel0(m[u], i + u * stride);
        }
    }
    for (; i < n16; i += stride) {
        const uint4 a = __ldcs(va + i), m = __ldcs(vm + i), s = __ldcs(vs + i);
        __stcs(vo + i, mix16(a, m, s));
        if (mask_sum) sum_channel0(m, i);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n - n16 * 16)) {          // tail (< 16 bytes)
        const int64_t j = n16 * 16 + threadIdx.x;
        out[j] = (uint8_t)mix4(actor[j], mask[j], scene[j]);
        if (mask_sum && j % 3 == 0) local += mask[j];
    }
    if (mask_sum) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(mask_sum, local);
    }
}

}  // namespace

int launch_cutmix(const uint8_t *d_actor, const uint8_t *d_mask, const uint8_t *d_scene, int64_t n, uint8_t *d_out,
                  unsigned long long *d_mask_sum, cudaStream_t stream)
{
    if (n < 0) return fail(BGD_ERR_INVALID, "actor_cut_mix: negative size");
    if (d_mask_sum) BGD_CUDA_TRY(cudaMemsetAsync(d_mask_sum, 0, sizeof(unsigned long long), stream));
    if (n == 0) return BGD_OK;
    if (!d_actor || !d_mask || !d_scene || !d_out) return fail(BGD_ERR_INVALID, "actor_cut_mix: null pointer");
    for (const void *p : {(const void *)d_actor, (const void *)d_mask, (const void *)d_scene, (const void *)d_out})
        if (reinterpret_cast<uintptr_t>(p) % 16) return fail(BGD_ERR_INVALID, "actor_cut_mix: buffers must be 16-byte aligned");
    DeviceProps dp;
    if (int rc = current_device_props(&dp)) return rc;
    const int64_t want = (n / 16 + 255) / 256;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)dp.sm_count * 8));
    cutmix_kernel<<<grid, 256, 0, stream>>>(d_actor, d_mask, d_scene, d_out, n, d_mask_sum);
    count_launch();
    BGD_CUDA_TRY(cudaGetLastError());
    return BGD_OK;
}

}  // namespace bgd
