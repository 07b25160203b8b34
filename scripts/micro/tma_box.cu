// Micro-benchmark for the north_star's design point "stage each exciton's local lattice neighbourhood in shared
// memory (TMA-loaded tiles)": ONE 3-D tensor-map box per exciton instead of 49 row copies.
//
// Weights of a 128^3 lattice padded by R = 4 on every side (periodic copies in x, y would go there; zeros in z):
// a (128 + 8)^3 fp64 array, 20 MB, L2-resident like the real one.  Per "event" a warp needs the 256 weights of the
// cutoff-4 sphere around a random site.
//   tma  : lane 0 issues cp.async.bulk.tensor.3d (box 10 x 9 x 9 doubles = 6480 B: the inner extent must be a
//          multiple of 16 B) behind an mbarrier, DEPTH boxes in flight per warp; when a box has landed every lane
//          reads its 8 sphere entries from shared memory (256 = 8 x 32, dx fastest) and folds them into a sum.
//   ldg  : the shape of hfx_run_kernel v3: every lane issues its 8 gathers straight from global memory
//          (register-resident linear offsets, LDG.E.64.CONSTANT), next exciton's loads in flight while this one's
//          are consumed.
// Prints boxes (= events) per second for both; hfx_run_kernel needs ~300 further warp instructions per event, so
// these are upper bounds for the data movement alone.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/tma_box scripts/micro/tma_box.cu
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

constexpr int R = 4, L = 128, LP = L + 2 * R;       // padded edge
constexpr int BX = 10, BY = 9, BZ = 9;              // box (BX * 8 B must be a multiple of 16 B)
constexpr int BOX_BYTES = BX * BY * BZ * 8;         // 6480
constexpr int BOX_PITCH = 6528;                     // 128-byte aligned slots
constexpr int WARPS = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int DEPTH>
__global__ void __launch_bounds__(WARPS * 32) tma_kernel(const __grid_constant__ CUtensorMap map, const int *box_off /* [256] */,
                                                          int iters, double *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *buf = smem + (size_t)warp * DEPTH * BOX_PITCH;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (size_t)WARPS * DEPTH * BOX_PITCH) + warp * DEPTH;
    if (lane == 0)
        for (int d = 0; d < DEPTH; ++d) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + d)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    int off[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) off[k] = box_off[32 * k + lane];
    uint32_t rng = (blockIdx.x * WARPS * 32 + warp * 32) * 2654435761u + 12345u;   // warp-uniform
    auto issue = [&](int d) {
        const int x = (int)(lcg(rng) % L), y = (int)(lcg(rng) % L), z = (int)(lcg(rng) % L);   // box origin = site - R + R (padding)
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + d)), "r"(BOX_BYTES) : "memory");
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(smem_u32(buf + d * BOX_PITCH)), "l"(&map), "r"(x & ~1), "r"(y), "r"(z), "r"(smem_u32(bar + d)) : "memory");
        }
    };
    for (int d = 0; d < DEPTH; ++d) issue(d);
    double acc = 0.0;
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
        const int d = it % DEPTH;
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(bar + d)), "r"((phase >> d) & 1u) : "memory");
        phase ^= 1u << d;
        const double *box = reinterpret_cast<const double *>(buf + d * BOX_PITCH);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += box[off[k]];
        __syncwarp();                                       // everyone has read the slot before it is refilled
        if (it + DEPTH < iters) issue(d);
    }
    if (acc == 1.2345) out[0] = acc;
}

__global__ void __launch_bounds__(256) ldg_kernel(const double *w, const int *lin /* [256] */, int iters, double *out) {
    const int lane = threadIdx.x & 31;
    int off[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) off[k] = lin[32 * k + lane];
    uint32_t rng = (blockIdx.x * 256 + (threadIdx.x & ~31)) * 2654435761u + 12345u;
    auto site = [&]() {
        const int x = (int)(lcg(rng) % L) + R, y = (int)(lcg(rng) % L) + R, z = (int)(lcg(rng) % L) + R;
        return ((size_t)z * LP + y) * LP + x;
    };
    double cur[8], nxt[8], acc = 0.0;
    size_t s = site();
#pragma unroll
    for (int k = 0; k < 8; ++k) cur[k] = __ldg(w + s + off[k]);
    for (int it = 0; it < iters; ++it) {
        s = site();
#pragma unroll
        for (int k = 0; k < 8; ++k) nxt[k] = __ldg(w + s + off[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc += cur[k]; cur[k] = nxt[k]; }
    }
    if (acc == 1.2345) out[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int DEPTH>
static void run_tma(const CUtensorMap &map, const int *d_box, int iters, double *out, cudaEvent_t e0, cudaEvent_t e1, int sms) {
    const size_t smem = (size_t)WARPS * DEPTH * BOX_PITCH + WARPS * DEPTH * 8 + 128;
    cudaFuncSetAttribute(tma_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tma_kernel<DEPTH>, WARPS * 32, smem);
    const int grid = sms * per_sm;
    tma_kernel<DEPTH><<<grid, WARPS * 32, smem>>>(map, d_box, 200, out);
    cudaEventRecord(e0);
    tma_kernel<DEPTH><<<grid, WARPS * 32, smem>>>(map, d_box, iters, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double boxes = (double)grid * WARPS * iters;
    printf("tma  depth=%d: %d CTAs/SM x %d warps (%zu B smem/CTA), %.3f ms, %.3e boxes/s, %.1f cycles/box/SM, %.0f GB/s landed (err: %s)\n",
           DEPTH, per_sm, WARPS, smem, ms, boxes / (ms * 1e-3), 1.965e9 * sms / (boxes / (ms * 1e-3)), boxes * BOX_BYTES / (ms * 1e-3) / 1e9,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    const size_t n = (size_t)LP * LP * LP;
    double *w, *out;
    cudaMalloc(&w, n * 8); cudaMalloc(&out, 8);
    cudaMemset(w, 0, n * 8);
    // cutoff-4 sphere: 256 offsets, dz-major, dx fastest (the table order of hf_x_configure)
    std::vector<int> box, lin;
    for (int dz = -R; dz <= R; ++dz)
        for (int dy = -R; dy <= R; ++dy)
            for (int dx = -R; dx <= R; ++dx) {
                const int d2 = dx * dx + dy * dy + dz * dz;
                if (d2 == 0 || d2 > R * R) continue;
                box.push_back(((dz + R) * BY + (dy + R)) * BX + (dx + R));
                lin.push_back((dz * LP + dy) * LP + dx);
            }
    printf("sphere: %zu sites; box %dx%dx%d = %d B\n", box.size(), BX, BY, BZ, BOX_BYTES);
    int *d_box, *d_lin;
    cudaMalloc(&d_box, 256 * 4); cudaMalloc(&d_lin, 256 * 4);
    cudaMemcpy(d_box, box.data(), 256 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_lin, lin.data(), 256 * 4, cudaMemcpyHostToDevice);

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &q);
    if (!encode) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
    CUtensorMap map;
    const cuuint64_t dims[3] = {LP, LP, LP}, strides[2] = {(cuuint64_t)LP * 8, (cuuint64_t)LP * LP * 8};
    const cuuint32_t boxd[3] = {BX, BY, BZ}, estr[3] = {1, 1, 1};
    const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, w, dims, strides, boxd, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)rc); return 1; }

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    run_tma<1>(map, d_box, iters, out, e0, e1, sms);
    run_tma<2>(map, d_box, iters, out, e0, e1, sms);
    run_tma<4>(map, d_box, iters, out, e0, e1, sms);
    for (int per_sm = 2; per_sm <= 8; per_sm *= 2) {
        const int grid = sms * per_sm;
        ldg_kernel<<<grid, 256>>>(w, d_lin, 200, out);
        cudaEventRecord(e0);
        ldg_kernel<<<grid, 256>>>(w, d_lin, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double boxes = (double)grid * 8 * iters;
        printf("ldg  %d CTAs/SM x 8 warps: %.3f ms, %.3e spheres/s, %.1f cycles/sphere/SM, %.0f GB/s gathered (err: %s)\n", per_sm, ms,
               boxes / (ms * 1e-3), 1.965e9 * sms / (boxes / (ms * 1e-3)), boxes * 256 * 8 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
