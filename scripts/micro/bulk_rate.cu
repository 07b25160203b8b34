// Micro-benchmark: how many small cp.async.bulk (global -> shared) copies per second does an SM sustain?
// Each warp: 32 lanes issue one copy of BYTES each from pseudo-random 16-B aligned addresses of an
// L2-resident buffer, one mbarrier per warp, repeat.  Compare with plain 8-byte gathers (LDG.64).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES>
__global__ void __launch_bounds__(256) bulk_kernel(const double *src, size_t n_elems, int iters, double *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = smem + warp * (32 * 96 + 16);
    uint64_t *bar = reinterpret_cast<uint64_t *>(buf + 32 * 96);
    if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    __syncwarp();
    uint32_t rng = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
    double acc = 0.0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
        rng = rng * 1664525u + 1013904223u;
        const size_t idx = ((size_t)(rng >> 4) % (n_elems - 16)) & ~(size_t)1;          // 16-B aligned
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(32 * BYTES) : "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(buf + lane * 96)), "l"(src + idx), "r"(BYTES), "r"(smem_u32(bar)) : "memory");
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        acc += *reinterpret_cast<double *>(buf + lane * 96);
        __syncwarp();
    }
    if (acc == 1.2345) out[0] = acc;
}

__global__ void __launch_bounds__(256) ldg_kernel(const double *src, size_t n_elems, int iters, int per_iter, double *out) {
    uint32_t rng = (blockIdx.x * 256 + threadIdx.x) * 2654435761u + 12345u;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        rng = rng * 1664525u + 1013904223u;
        // a warp reads per_iter runs of 8 consecutive doubles (like rows of the cutoff sphere)
        uint32_t r = __shfl_sync(0xffffffffu, rng, lane & ~7);
        const size_t idx = ((size_t)(r >> 4) % (n_elems - 16)) + (lane & 7);
        for (int k = 0; k < per_iter; ++k) acc += __ldg(src + ((idx + (size_t)k * 4099) % (n_elems - 16)));
    }
    if (acc == 1.2345) out[0] = acc;
}

int main() {
    const size_t n = 4u << 20;    // 32 MB of doubles: L2-resident
    double *src, *out;
    cudaMalloc(&src, n * 8); cudaMalloc(&out, 8);
    cudaMemset(src, 0, n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 2, iters = 20000;
    const size_t smem = 8 * (32 * 96 + 16);
    auto run = [&](auto kernel, const char *name, int bytes) {
        kernel<<<grid, 256, smem>>>(src, n, 100, out);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        kernel<<<grid, 256, smem>>>(src, n, iters, out);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double copies = (double)grid * 256 * iters;
        printf("%s bytes=%d: %.3f ms, %.3e copies/s, %.2f copies/cycle/SM (1.965 GHz), %.1f GB/s  err=%s\n", name, bytes, ms,
               copies / (ms * 1e-3), copies / (ms * 1e-3) / 148 / 1.965e9, copies * bytes / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    };
    run(bulk_kernel<16>, "bulk", 16);
    run(bulk_kernel<32>, "bulk", 32);
    run(bulk_kernel<64>, "bulk", 64);
    run(bulk_kernel<80>, "bulk", 80);
    {
        ldg_kernel<<<grid, 256>>>(src, n, 100, 8, out);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        ldg_kernel<<<grid, 256>>>(src, n, iters, 8, out);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double rows = (double)grid * 8 * iters * 8 * 4;    // warps x iters x loads x 4 runs of 8 doubles per load
        printf("ldg rows of 64 B: %.3f ms, %.3e rows/s, %.2f rows/cycle/SM\n", ms, rows / (ms * 1e-3), rows / (ms * 1e-3) / 148 / 1.965e9);
    }
    return 0;
}
