// Instruction-throughput probe for sm_100a: how many thread-operations per clock per SM the
// integer instructions the median kernels are built from actually sustain.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

template <int OP>
__device__ __forceinline__ void step(uint32_t (&a)[CHAINS], uint32_t b, uint32_t c)
{
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (OP == 1) asm volatile("shf.r.clamp.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c & 7));
        if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (OP == 3) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (OP == 4) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (OP == 6) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
        if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
        if (OP == 8) asm volatile("shl.b32 %0, %0, 1;" : "+r"(a[i]));
        if (OP == 9) asm volatile("vmax2.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (OP == 10) {   // mix: 2 lop3 + 1 shl + 1 shr  (one transpose pair)
            uint32_t t;
            asm volatile("shl.b32 %0, %1, 4;" : "=r"(t) : "r"(a[i]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE4;" : "+r"(a[i]) : "r"(t), "r"(b));
            asm volatile("shr.u32 %0, %1, 4;" : "=r"(t) : "r"(a[i]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xD8;" : "+r"(a[i]) : "r"(t), "r"(c));
        }
        if (OP == 11) {   // mix: lop3 + mad (alu + fma pipes)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
        if (OP == 13) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(0x10000000u));
        if (OP == 14) {   // transpose pair with both shifts on the fma pipe: 2 lop3 + mul.lo + mul.hi
            uint32_t t;
            asm volatile("mul.lo.u32 %0, %1, 16;" : "=r"(t) : "r"(a[i]));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xE4;" : "+r"(a[i]) : "r"(t), "r"(b));
            asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(a[i]), "r"(0x10000000u));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0xD8;" : "+r"(a[i]) : "r"(t), "r"(c));
        }
        if (OP == 16) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        if (OP == 12) {   // mix: lop3 + vabsdiff4
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
        }
    }
}

template <int OP>
__global__ void probe(uint32_t *out, uint32_t b, uint32_t c, long long *cycles)
{
    uint32_t a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 2654435761u + i;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) step<OP>(a, b, c);
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int ops_per_step, int sms, uint32_t *d_out, long long *d_cyc)
{
    const int threads = 1024, blocks = sms;   // one full CTA per SM: 32 warps
    probe<OP><<<blocks, threads>>>(d_out, 0x9e3779b9u, 0x01010101u, d_cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(d_out, 0x9e3779b9u, 0x01010101u, d_cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc[1024];
    cudaMemcpy(cyc, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; ++i) mean += cyc[i]; mean /= blocks;
    const double ops = (double)threads * CHAINS * ITERS * ops_per_step;
    printf("%-28s %8.1f thread-ops/clk/SM   (%.0f cycles, %.3f ms, %.2f GHz eff)\n", name, ops / mean, mean, ms,
           mean / (ms * 1e6));
}

// copy bandwidth probes: plain 16-byte loads, and TMA bulk copies of `row_bytes` into smem
__global__ void read_ldg(const uint4 *__restrict__ src, size_t n, uint32_t *sink)
{
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldcs(src + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void read_bulk(const uint8_t *__restrict__ src, size_t total, int row_bytes, int rows, uint32_t *sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint8_t *buf[2] = {smem + 128, smem + 128 + (size_t)row_bytes * rows};
    const size_t tile = (size_t)row_bytes * rows;
    const size_t ntiles = total / tile;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    uint32_t phase[2] = {0, 0};
    uint32_t acc = 0;
    auto issue = [&](size_t t, int slot) {
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + slot)), "r"((uint32_t)tile));
            __syncwarp();
            for (int r = threadIdx.x; r < rows; r += 32)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(buf[slot] + (size_t)r * row_bytes)), "l"(src + t * tile + (size_t)r * row_bytes),
                             "r"((uint32_t)row_bytes), "r"(s32(bar + slot)) : "memory");
        }
    };
    size_t t = blockIdx.x;
    int slot = 0;
    if (t < ntiles) issue(t, 0);
    for (; t < ntiles; t += gridDim.x, slot ^= 1) {
        const size_t nxt = t + gridDim.x;
        if (nxt < ntiles) issue(nxt, slot ^ 1);
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n"
                     ::"r"(s32(bar + slot)), "r"(phase[slot]) : "memory");
        phase[slot] ^= 1;
        acc ^= reinterpret_cast<uint32_t *>(buf[slot])[threadIdx.x];
        __syncthreads();
    }
    if (acc == 0x12345678u) *sink = acc;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, cc %d.%d\n", p.name, sms, p.major, p.minor);
    uint32_t *d_out; long long *d_cyc;
    cudaMalloc(&d_out, sizeof(uint32_t) * 1024 * sms);
    cudaMalloc(&d_cyc, sizeof(long long) * sms);
    run<0>("lop3", 1, sms, d_out, d_cyc);
    run<1>("shf", 1, sms, d_out, d_cyc);
    run<2>("imad", 1, sms, d_out, d_cyc);
    run<3>("vabsdiff4.acc", 1, sms, d_out, d_cyc);
    run<4>("dp4a", 1, sms, d_out, d_cyc);
    run<5>("prmt", 1, sms, d_out, d_cyc);
    run<6>("popc", 1, sms, d_out, d_cyc);
    run<7>("iadd", 1, sms, d_out, d_cyc);
    run<8>("shl imm", 1, sms, d_out, d_cyc);
    run<9>("vmax2 (u16x2)", 1, sms, d_out, d_cyc);
    run<10>("transpose pair (4 ops)", 4, sms, d_out, d_cyc);
    run<11>("lop3+imad", 2, sms, d_out, d_cyc);
    run<12>("lop3+vabsdiff4", 2, sms, d_out, d_cyc);
    run<13>("mul.hi (imad.hi)", 1, sms, d_out, d_cyc);
    run<14>("transpose pair, shifts on fma", 4, sms, d_out, d_cyc);
    run<16>("mad.hi", 1, sms, d_out, d_cyc);

    if (getenv("MB_NO_BW")) return 0;
    // bandwidth
    const size_t bytes = (size_t)8 << 30;
    uint8_t *d_src; cudaMalloc(&d_src, bytes); cudaMemset(d_src, 1, bytes);
    uint32_t *d_sink; cudaMalloc(&d_sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        read_ldg<<<sms * 8, 512>>>(reinterpret_cast<const uint4 *>(d_src), bytes / 16, d_sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep) printf("read LDG.128                  %8.1f GB/s\n", bytes / (ms * 1e6));
    }
    const int cfgs[][3] = {{512, 96, 2}, {256, 192, 2}, {1024, 48, 2}, {512, 180, 1}, {2048, 24, 2}, {4096, 12, 2}};
    for (auto &c : cfgs) {
        const int row_bytes = c[0], rows = c[1], ctas = c[2];
        const size_t smem = 128 + 2 * (size_t)row_bytes * rows;
        cudaFuncSetAttribute(read_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            read_bulk<<<sms * ctas, 128, smem>>>(d_src, bytes, row_bytes, rows, d_sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("read TMA bulk %4d B x %3d rows, %d CTA/SM  %8.1f GB/s  (%s)\n", row_bytes, rows, ctas,
                            bytes / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
