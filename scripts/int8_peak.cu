// int8_peak.cu -- measured u8 x u8 -> s32 tensor-core peak of one B200 (SURVEY 8d asks for it: MEASURED_PEAKS.json has
// no int8 figure).  One elected thread per CTA issues back-to-back tcgen05.mma.cta_group::1.kind::i8 instructions
// (M = 128, N in {64, 256}, K = 32 bytes each) on operands that already sit in shared memory (K-major, 128-byte
// swizzle, the same descriptors as fpm_corr_mma_kernel), accumulating in TMEM; nothing else runs, so the rate is the
// tensor pipe's own issue/operand limit.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
// Prints one JSON line: TOPS (2*M*N*K ops per instruction) per shape and CTAs/SM.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include "../fastest_image_pattern_matching_b200/csrc/fpm_mma.cuh"

template <int N>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, unsigned long long* sink)
{
    using namespace fpm_ptx;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (base_u32 - smem_u32(smem_raw));
    constexpr int A_BYTES = 128 * 128, B_BYTES = N * 128;
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + A_BYTES + B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); fence_barrier_init(); }
    fence_proxy_async();
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t IDESC = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 0 && lane == 0) {
        const uint64_t da = smem_desc_sw128(base_u32), db = smem_desc_sw128(base_u32 + A_BYTES);
        for (int it = 0; it < iters; it++) {
            // two independent accumulators, four K = 32 slices of the 128-byte swizzle atom each
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                mma_i8(tmem_base, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC, 1u);
                mma_i8(tmem_base + 256, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), IDESC, 1u);
            }
        }
        tc_commit(smem_u32(bar));
        mbar_wait(smem_u32(bar), 0);
        tc_fence_after();
    }
    __syncthreads();
    if (warp == 0) {
        uint32_t v[16];
        tmem_ld16(tmem_base, v);
        tmem_ld_wait();
        if (lane == 0 && sink) sink[blockIdx.x] = v[0];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int N>
double run(int ctas, int iters, unsigned long long* sink)
{
    const size_t smem = 128 * 128 + N * 128 + 1024 + 64;
    cudaFuncSetAttribute(peak_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    peak_kernel<N><<<ctas, 128, smem>>>(iters / 8, sink);
    cudaDeviceSynchronize();
    double best = 0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(a);
        peak_kernel<N><<<ctas, 128, smem>>>(iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double ops = 2.0 * 128 * N * 32 * 8.0 * iters * ctas;
        best = std::max(best, ops / (ms * 1e-3) / 1e12);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return best;
}

int main(int argc, char** argv)
{
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, dev);
    const int sms = p.multiProcessorCount;
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    unsigned long long* sink = nullptr;
    cudaMalloc(&sink, sizeof(unsigned long long) * sms);
    const double t256 = run<256>(sms, iters, sink);
    const double t64 = run<64>(sms, iters * 2, sink);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("{\"int8_tops_m128n256\": %.1f, \"int8_tops_m128n64\": %.1f, \"sms\": %d, \"sm_clock_mhz_max\": %.0f, "
           "\"how\": \"tcgen05.mma.cta_group::1.kind::i8 u8xu8->s32, M128 K32, operands resident in smem (SW128), 1 CTA/SM, best of 5\"}\n",
           t256, t64, sms, clk / 1000.0);
    return (t256 > 0 && t64 > 0) ? 0 : 1;
}
