// warp_emu.h — a single-warp CUDA emulator for HOST-SIDE TESTS of the kernel source.
//
// TEST INFRASTRUCTURE ONLY.  It lets tests/test_kernel_emulated.py compile
// hyperfluorescence_b200/csrc/hf_frm.cuh with g++ and run the *same* kernel code, lane by lane,
// against the golden vectors on a machine without a GPU.  It is not a CPU backend of the
// product: it is not shipped in the package, not reachable from the Python façade, runs one
// warp at a time, and is ~10^5 times slower than the device.
//
// Model: the 32 lanes of one warp are ucontext coroutines scheduled round-robin.  Every
// warp-collective (__ballot_sync, __shfl_*_sync, __syncwarp) deposits the lane's operand,
// yields, and reads the peers' operands after the whole warp has arrived.  The emulator checks
// that all lanes execute the same sequence of collectives (same call kind, full mask) and aborts
// with a message otherwise, so divergent use of a collective is a test failure rather than a
// silent hang on the GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__
#define __align__(x)
#define __restrict__

struct emu_dim3 { unsigned x, y, z; };
struct int2 { int x, y; };
static inline int2 make_int2(int x, int y) { int2 v = {x, y}; return v; }
static emu_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};
alignas(16) unsigned char hf_smem[1 << 20];

// ---- scheduler ---------------------------------------------------------------------------------
enum { EMU_BALLOT = 1, EMU_SHFL = 2, EMU_SHFL_XOR = 3, EMU_SYNCWARP = 4, EMU_SHFL_UP = 5 };

struct emu_warp {
    ucontext_t main_ctx, lane_ctx[32];
    unsigned char *stacks[32];
    int done[32];
    int cur;
    uint64_t slot[2][32];
    int op[2][32];
    uint64_t seq[32];
    void (*body)(void *);
    void *arg;
    const char *error;
};
static emu_warp g_emu;

static void emu_fail(const char *msg) {
    g_emu.error = msg;
    fprintf(stderr, "warp_emu: %s (lane %d)\n", msg, g_emu.cur);
    abort();
}

static void emu_trampoline() {
    g_emu.body(g_emu.arg);
    g_emu.done[g_emu.cur] = 1;
    swapcontext(&g_emu.lane_ctx[g_emu.cur], &g_emu.main_ctx);
}

// Run body(arg) on 32 lanes until every lane has returned.
static void emu_run_warp(void (*body)(void *), void *arg) {
    const size_t stack_size = 512 * 1024;
    g_emu.body = body; g_emu.arg = arg; g_emu.error = nullptr;
    for (int l = 0; l < 32; ++l) {
        g_emu.done[l] = 0; g_emu.seq[l] = 0;
        if (!g_emu.stacks[l]) g_emu.stacks[l] = (unsigned char *)malloc(stack_size);
        getcontext(&g_emu.lane_ctx[l]);
        g_emu.lane_ctx[l].uc_stack.ss_sp = g_emu.stacks[l];
        g_emu.lane_ctx[l].uc_stack.ss_size = stack_size;
        g_emu.lane_ctx[l].uc_link = &g_emu.main_ctx;
        makecontext(&g_emu.lane_ctx[l], emu_trampoline, 0);
    }
    for (;;) {
        int live = 0;
        for (int l = 0; l < 32; ++l) {
            if (g_emu.done[l]) continue;
            ++live;
            g_emu.cur = l;
            threadIdx.x = (unsigned)l;
            swapcontext(&g_emu.main_ctx, &g_emu.lane_ctx[l]);
        }
        if (!live) break;
        // all live lanes must be waiting at the same collective; a lane that returned while its
        // peers wait is divergent control flow around a full-mask collective
        int ndone = 0;
        uint64_t s0 = 0; int have = 0;
        for (int l = 0; l < 32; ++l) {
            if (g_emu.done[l]) { ++ndone; continue; }
            if (!have) { s0 = g_emu.seq[l]; have = 1; }
            else if (g_emu.seq[l] != s0) emu_fail("lanes are at different collectives");
        }
        if (have && ndone) emu_fail("some lanes exited while others wait at a collective");
    }
}

static uint64_t *emu_collective(int op, uint64_t v, unsigned mask) {
    if (mask != 0xffffffffu) emu_fail("only full-mask collectives are emulated");
    const int l = g_emu.cur;
    const int par = (int)(g_emu.seq[l] & 1);
    g_emu.seq[l] += 1;
    g_emu.slot[par][l] = v;
    g_emu.op[par][l] = op;
    swapcontext(&g_emu.lane_ctx[l], &g_emu.main_ctx);
    threadIdx.x = (unsigned)l;
    for (int i = 0; i < 32; ++i)
        if (g_emu.op[par][i] != op) emu_fail("lanes called different collectives");
    return g_emu.slot[par];
}

// ---- warp intrinsics -----------------------------------------------------------------------------
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint64_t *s = emu_collective(EMU_BALLOT, pred ? 1 : 0, mask);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(s[i] & 1) << i;
    return r;
}
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu_collective(EMU_SYNCWARP, 0, mask); }
// one emulated warp per block: a block barrier is a warp barrier
static inline void __syncthreads() { emu_collective(EMU_SYNCWARP, 0, 0xffffffffu); }

template <typename T>
static inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T>
static inline T emu_unbits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

template <typename T>
static inline T __shfl_sync(unsigned mask, T v, int src) {
    uint64_t *s = emu_collective(EMU_SHFL, emu_bits(v), mask);
    return emu_unbits<T>(s[src & 31]);
}
template <typename T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask) {
    const int me = g_emu.cur;
    uint64_t *s = emu_collective(EMU_SHFL_XOR, emu_bits(v), mask);
    return emu_unbits<T>(s[(me ^ lane_mask) & 31]);
}

template <typename T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    const int me = g_emu.cur;
    uint64_t *s = emu_collective(EMU_SHFL_UP, emu_bits(v), mask);
    return me >= (int)delta ? emu_unbits<T>(s[me - (int)delta]) : v;
}

// ---- scalar intrinsics -----------------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
template <typename T>
static inline T __ldg(const T *p) { return *p; }
// compiled with -ffp-contract=off: each of these is one correctly rounded IEEE operation
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline uint32_t __float_as_uint(float a) { uint32_t b; memcpy(&b, &a, 4); return b; }
static inline float __uint_as_float(uint32_t a) { float b; memcpy(&b, &a, 4); return b; }
static inline int __double2hiint(double a) { long long b; memcpy(&b, &a, 8); return (int)(b >> 32); }
static inline long long __double_as_longlong(double a) { long long b; memcpy(&b, &a, 8); return b; }
static inline double __longlong_as_double(long long a) { double b; memcpy(&b, &a, 8); return b; }
// fast-math intrinsics used by the screening pass (glibc's math.h reserves the __logf/__expf names)
#define __logf(x) logf(x)
#define __expf(x) expf(x)
struct double2 { double x, y; };
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    unsigned long long old = *p; *p += v; return old;
}
// lanes are coroutines of one host thread that only switch at collectives: plain read-modify-write is atomic enough
static inline unsigned atomicOr(unsigned *p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
static inline unsigned atomicCAS(unsigned *p, unsigned cmp, unsigned v) { const unsigned o = *p; if (o == cmp) *p = v; return o; }
