// Micro-benchmark: does a line-aligned "brick" layout of the Boltzmann weights cut the DRAM traffic of the cutoff-4
// sphere gather when the weights do NOT fit the L2 (BASELINE configs[4]: 256^3 x 64 realisations)?
//
// hfx_run_kernel with 32 trajectories per warp streams 7.9 KB per event from DRAM (profiles/r01g): a sector miss fills
// the whole 128-byte line and the row-major sphere touches 62 lines = 3.3x its 2 394 algorithmic bytes.  Round 1
// ESTIMATED that 4 x 4 x 1 bricks (16 doubles = one 128-byte line) would touch 44 lines; this measures it.
//
//   row   : w[(z * Y + y) * X + x], linear offsets per sphere entry (the layout of the product)
//   brick : w[(((z * Y/4 + y/4) * X/4 + x/4) << 4) + (y & 3) * 4 + (x & 3)]
// Every warp gathers the 256 sphere entries (8 per lane, dx fastest) of a fresh random interior site per iteration from
// an array far larger than the L2 (N_REAL x 256^3 doubles), next sphere's loads in flight while this one's are consumed.
// Run under `ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum` for the bytes; prints spheres/s itself.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/brick_layout scripts/micro/brick_layout.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <vector>

constexpr int L = 256, R = 4, N_REAL = 16;

__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <bool kBrick>
__global__ void __launch_bounds__(256) gather_kernel(const double *w, const int *off /* [256] packed dx | dy << 8 | dz << 16, +128 each */,
                                                     const int *lin /* [256] row-major linear offsets */, int iters, double *out) {
    const int lane = threadIdx.x & 31;
    int o[8], li[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { o[k] = off[32 * k + lane]; li[k] = lin[32 * k + lane]; }
    uint32_t rng = (blockIdx.x * 256 + (threadIdx.x & ~31)) * 2654435761u + 12345u;   // warp-uniform
    auto addr = [&](int k, int x, int y, int z, size_t real) -> size_t {
        if (!kBrick) return real + (((size_t)z * L + y) * L + x) + li[k];
        const int xx = x + (o[k] & 255) - 128, yy = y + ((o[k] >> 8) & 255) - 128, zz = z + ((o[k] >> 16) & 255) - 128;
        return real + ((((size_t)zz * (L / 4) + (yy >> 2)) * (L / 4) + (xx >> 2)) << 4) + ((yy & 3) << 2) + (xx & 3);
    };
    auto site = [&](int &x, int &y, int &z, size_t &real) {
        x = R + (int)(lcg(rng) % (L - 2 * R)); y = R + (int)(lcg(rng) % (L - 2 * R)); z = R + (int)(lcg(rng) % (L - 2 * R));
        real = (size_t)(lcg(rng) % N_REAL) * L * L * L;
    };
    double cur[8], nxt[8], acc = 0.0;
    int x, y, z;
    size_t real;
    site(x, y, z, real);
#pragma unroll
    for (int k = 0; k < 8; ++k) cur[k] = __ldg(w + addr(k, x, y, z, real));
    for (int it = 0; it < iters; ++it) {
        site(x, y, z, real);
#pragma unroll
        for (int k = 0; k < 8; ++k) nxt[k] = __ldg(w + addr(k, x, y, z, real));
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc += cur[k]; cur[k] = nxt[k]; }
    }
    if (acc == 1.2345) out[0] = acc;
}

int main() {
    const size_t n = (size_t)N_REAL * L * L * L;
    double *w, *out;
    if (cudaMalloc(&w, n * 8) != cudaSuccess) { printf("cudaMalloc of %zu MB failed\n", n * 8 >> 20); return 1; }
    cudaMalloc(&out, 8);
    cudaMemset(w, 0, n * 8);
    std::vector<int> off, lin;
    for (int dz = -R; dz <= R; ++dz)
        for (int dy = -R; dy <= R; ++dy)
            for (int dx = -R; dx <= R; ++dx) {
                const int d2 = dx * dx + dy * dy + dz * dz;
                if (d2 == 0 || d2 > R * R) continue;
                off.push_back((dx + 128) | ((dy + 128) << 8) | ((dz + 128) << 16));
                lin.push_back((dz * L + dy) * L + dx);
            }
    // lines of 128 bytes a sphere touches on average, both layouts (host count over random sites)
    {
        double rows = 0, bricks = 0;
        const int trials = 2000;
        uint32_t s = 7;
        for (int t = 0; t < trials; ++t) {
            auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 8) % (L - 2 * R)) + R; };
            const int x = rnd(), y = rnd(), z = rnd();
            std::vector<size_t> a, b;
            for (size_t k = 0; k < off.size(); ++k) {
                const int xx = x + (off[k] & 255) - 128, yy = y + ((off[k] >> 8) & 255) - 128, zz = z + ((off[k] >> 16) & 255) - 128;
                a.push_back(((((size_t)zz * L + yy) * L + xx) * 8) >> 7);
                b.push_back((((((size_t)zz * (L / 4) + (yy >> 2)) * (L / 4) + (xx >> 2)) << 4) * 8) >> 7);
            }
            auto uniq = [](std::vector<size_t> &v) { std::sort(v.begin(), v.end()); return (double)(std::unique(v.begin(), v.end()) - v.begin()); };
            rows += uniq(a); bricks += uniq(b);
        }
        printf("128-byte lines per cutoff-4 sphere (256 sites, 2048 B): row-major %.1f, 4x4x1 bricks %.1f\n", rows / trials, bricks / trials);
    }
    int *d_off, *d_lin;
    cudaMalloc(&d_off, 256 * 4); cudaMalloc(&d_lin, 256 * 4);
    cudaMemcpy(d_off, off.data(), 256 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_lin, lin.data(), 256 * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 8, iters = 4000;
    for (int brick = 0; brick < 2; ++brick) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (brick) gather_kernel<true><<<grid, 256>>>(w, d_off, d_lin, iters, out);
            else gather_kernel<false><<<grid, 256>>>(w, d_off, d_lin, iters, out);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double spheres = (double)grid * 8 * iters;
        printf("%-5s: %.3f ms, %.3e spheres/s (err: %s)\n", brick ? "brick" : "row", ms, spheres / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
