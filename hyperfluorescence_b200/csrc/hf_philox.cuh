// Philox4x32-10 and the uniform-draw contract on the device.
//
// The reference has no counter-based RNG (it uses unseeded MT19937: reseau.py:113,
// molecule.py:72); parity is defined by injecting uniforms through random.Random.random()
// (SURVEY.md §8c).  This header is the device side of that contract:
//
//   key     = (seed_lo, seed_hi)
//   counter = (b_lo, b_hi, ident, stream),  b = d >> 1   (one block yields two uniforms)
//   (lo,hi) = d even ? (w0, w1) : (w2, w3)
//   u       = (((hi << 32) | lo) >> 11) * 2^-53          53-bit, like CPython's random()
//
//   stream 0 TRAJ    ident = trajectory   Lattice._seed after lattice creation (reseau.py:219,
//                                          225, 284, 316, 455, 467)
//   stream 1 SPIN    ident = trajectory   Molecule.generate_exciton (molecule.py:92)
//   stream 2 SPECIES ident = realisation  the species shuffle (reseau.py:164-168)
//   stream 3 ENERGY  ident = realisation  d = 4*site + j, the four gauss() calls of a site
//                                          (molecule.py:180-183)
#pragma once
#include <stdint.h>

#define HF_STREAM_TRAJ 0u
#define HF_STREAM_SPIN 1u
#define HF_STREAM_SPECIES 2u
#define HF_STREAM_ENERGY 3u

struct HfPhiloxBlock { uint32_t w0, w1, w2, w3; };

// Block b of stream (stream, ident): the two uniforms d = 2b and d = 2b + 1.
__device__ __forceinline__ HfPhiloxBlock hf_philox_block(uint32_t k0, uint32_t k1, uint32_t stream,
                                                         uint32_t ident, uint64_t b) {
    uint32_t c0 = (uint32_t)b, c1 = (uint32_t)(b >> 32), c2 = ident, c3 = stream;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    HfPhiloxBlock out = {c0, c1, c2, c3};
    return out;
}

// The same block with the ten round keys precomputed (rk[2r] = k0 + r * 0x9E3779B9, rk[2r+1] = k1 + r * 0xBB67AE85):
// when rk lives in the kernel's parameter block the xor takes it straight from the constant bank and the twenty key
// bumps per block disappear (16 of the 113 instructions of a block in hf_run_c1_kernel, ncu r02a).
__device__ __forceinline__ HfPhiloxBlock hf_philox_block_rk(const uint32_t (&rk)[20], uint32_t stream, uint32_t ident, uint64_t b) {
    uint32_t c0 = (uint32_t)b, c1 = (uint32_t)(b >> 32), c2 = ident, c3 = stream;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ rk[2 * r];
        const uint32_t n2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    HfPhiloxBlock out = {c0, c1, c2, c3};
    return out;
}
__host__ __device__ inline void hf_philox_round_keys(uint32_t k0, uint32_t k1, uint32_t (&rk)[20]) {
    for (int r = 0; r < 10; ++r) { rk[2 * r] = k0 + (uint32_t)r * 0x9E3779B9u; rk[2 * r + 1] = k1 + (uint32_t)r * 0xBB67AE85u; }
}

__device__ __forceinline__ double hf_block_uniform(const HfPhiloxBlock &blk, uint64_t d) {
    const uint64_t w = (d & 1) ? (((uint64_t)blk.w3 << 32) | blk.w2) : (((uint64_t)blk.w1 << 32) | blk.w0);
    return (double)(w >> 11) * 0x1p-53;
}

__device__ __forceinline__ double hf_philox_uniform(uint32_t k0, uint32_t k1, uint32_t stream,
                                                    uint32_t ident, uint64_t d) {
    return hf_block_uniform(hf_philox_block(k0, k1, stream, ident, d >> 1), d);
}
