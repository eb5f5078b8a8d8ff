// A small CUDA thread-block emulator for CPU tests (test infrastructure, not product code).
//
// A kernel's source is compiled for the host and one thread block at a time is run by `blockDim.x` cooperative fibers
// (own stacks, a register-only context switch on x86-64, ucontext elsewhere) on ONE OS thread:
//   * threadIdx is restored whenever a fiber resumes; blockIdx / blockDim / gridDim are globals;
//   * __syncthreads() is a generation barrier over the block (a waiting fiber yields to the next one), __syncwarp() and
//     the warp collectives (__ballot_sync, __shfl_*_sync, __reduce_*_sync, __any/__all_sync) a barrier + exchange slots
//     over the 32 fibers of a warp (full-mask use only: every lane of the warp has to make the call, as CUDA requires for
//     mask 0xffffffff).  Kernels whose threads spin on each other WITHOUT a barrier would not make progress here;
//   * `__shared__` variables become function-local statics (one block runs at a time), dynamic shared memory is the
//     global array the including file defines under the kernel's `extern __shared__` name (the test's build step
//     rewrites `extern __shared__` to `extern`);
//   * atomics map to the GCC __atomic built-ins; the _rn arithmetic intrinsics to plain IEEE operations (compile with
//     -ffp-contract=off).
// Blocks run one after the other, which is a legal CUDA schedule for kernels whose blocks do not wait for each other.
#pragma once
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif
#include <algorithm>
#include <cuda_runtime.h>

struct EmuDim { unsigned x = 1, y = 1, z = 1; };
static EmuDim emu_threadIdx, emu_blockIdx;
static EmuDim emu_blockDim, emu_gridDim;
#define threadIdx emu_threadIdx
#define blockIdx emu_blockIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __shared__
#define __shared__ static
#undef __constant__
#define __constant__

// ---- fibers ------------------------------------------------------------------------------------------------------------
struct EmuBar { unsigned count = 0, expected = 0; unsigned long long gen = 0; };
struct EmuWarp {
    EmuBar bar;
    unsigned long long slot[32];
};
struct EmuFiber {
    void *sp = nullptr;  // saved stack pointer (x86-64) while the fiber is not running
#if !defined(__x86_64__)
    ucontext_t uc;
#endif
    unsigned tid = 0;
    EmuWarp *warp = nullptr;
    bool done = false;
    char *stack = nullptr;
};
static const size_t EMU_STACK_BYTES = 512 * 1024;
static EmuFiber *emu_current = nullptr;
static EmuBar emu_block_bar;
static EmuWarp *emu_warp = nullptr;
static int emu_lane = 0;
static std::function<void()> *emu_body = nullptr;

#if defined(__x86_64__)
// void emu_ctx_switch(void **save_sp, void *load_sp): callee-saved registers on the own stack, swap the stack pointers
extern "C" void emu_ctx_switch(void **save_sp, void *load_sp);
asm(".text\n"
    ".p2align 4\n"
    ".globl emu_ctx_switch\n"
    ".type emu_ctx_switch,@function\n"
    "emu_ctx_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size emu_ctx_switch,.-emu_ctx_switch\n");
static void *emu_main_sp = nullptr;
#else
static ucontext_t emu_main_uc;
#endif

static inline void emu_yield() {  // back to the scheduler; returns when this fiber is resumed
    EmuFiber *self = emu_current;
#if defined(__x86_64__)
    emu_ctx_switch(&self->sp, emu_main_sp);
#else
    swapcontext(&self->uc, &emu_main_uc);
#endif
    emu_threadIdx.x = self->tid;
    emu_warp = self->warp;
    emu_lane = (int)(self->tid & 31);
}
static inline void emu_bar_wait(EmuBar &b) {
    const unsigned long long gen = b.gen;
    if (++b.count == b.expected) {
        b.count = 0;
        b.gen++;
        return;
    }
    while (b.gen == gen) emu_yield();
}
static void emu_fiber_entry() {
    EmuFiber *self = emu_current;
    emu_threadIdx.x = self->tid;
    emu_warp = self->warp;
    emu_lane = (int)(self->tid & 31);
    (*emu_body)();
    self->done = true;
    for (;;) emu_yield();  // never resumed again
}

static inline void __syncthreads() { emu_bar_wait(emu_block_bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu_bar_wait(emu_warp->bar); }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline long long clock64() { return 0; }

// every lane publishes a 64-bit value; f(slots) is evaluated by every lane on the complete set
template <typename F>
static inline auto emu_warp_collective(unsigned long long mine, F f) {
    emu_warp->slot[emu_lane] = mine;
    emu_bar_wait(emu_warp->bar);
    auto r = f(emu_warp->slot);
    emu_bar_wait(emu_warp->bar);
    return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    return emu_warp_collective(pred ? 1ull : 0ull, [](const unsigned long long *s) {
        unsigned m = 0;
        for (int l = 0; l < 32; l++) m |= (unsigned)(s[l] & 1ull) << l;
        return m;
    });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src) {
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    const unsigned long long r = emu_warp_collective(b, [src](const unsigned long long *s) { return s[src & 31]; });
    T o;
    memcpy(&o, &r, sizeof(T));
    return o;
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    const int lane = emu_lane;
    const unsigned long long r = emu_warp_collective(b, [lane, delta](const unsigned long long *s) { return lane >= (int)delta ? s[lane - delta] : s[lane]; });
    T o;
    memcpy(&o, &r, sizeof(T));
    return o;
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    const int lane = emu_lane;
    const unsigned long long r = emu_warp_collective(b, [lane, delta](const unsigned long long *s) { return lane + (int)delta < 32 ? s[lane + delta] : s[lane]; });
    T o;
    memcpy(&o, &r, sizeof(T));
    return o;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) {
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    const int lane = emu_lane;
    const unsigned long long r = emu_warp_collective(b, [lane, mask](const unsigned long long *s) { return s[(lane ^ mask) & 31]; });
    T o;
    memcpy(&o, &r, sizeof(T));
    return o;
}
static inline int __reduce_max_sync(unsigned, int v) {
    return emu_warp_collective((unsigned long long)(long long)v, [](const unsigned long long *s) {
        int m = (int)(long long)s[0];
        for (int l = 1; l < 32; l++) m = std::max(m, (int)(long long)s[l]);
        return m;
    });
}
static inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    return emu_warp_collective((unsigned long long)v, [](const unsigned long long *s) {
        unsigned a = 0;
        for (int l = 0; l < 32; l++) a += (unsigned)s[l];
        return a;
    });
}
static inline unsigned __activemask() { return 0xffffffffu; }

#undef __grid_constant__
#define __grid_constant__
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __ldcs(const T *p) { return *p; }
static inline unsigned __vmaxu2(unsigned a, unsigned b) {
    const unsigned lo = std::max(a & 0xffffu, b & 0xffffu), hi = std::max(a >> 16, b >> 16);
    return (hi << 16) | lo;
}
static inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline unsigned __vmaxu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; k++) r |= std::max((a >> (8 * k)) & 0xffu, (b >> (8 * k)) & 0xffu) << (8 * k);
    return r;
}
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; k++) {
        const int x = (a >> (8 * k)) & 0xff, y = (b >> (8 * k)) & 0xff;
        r |= (unsigned)(x > y ? x - y : y - x) << (8 * k);
    }
    return r;
}
static inline unsigned __vcmpgtu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; k++)
        if (((a >> (8 * k)) & 0xff) > ((b >> (8 * k)) & 0xff)) r |= 0xffu << (8 * k);
    return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {  // unsigned bytes
    for (int k = 0; k < 4; k++) c += ((a >> (8 * k)) & 0xffu) * ((b >> (8 * k)) & 0xffu);
    return c;
}
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsqrt_rn(double a) { volatile double r = std::sqrt(a); return r; }
template <typename T> static inline T __ldcg(const T *p) { return *const_cast<const volatile T *>(p); }
template <typename T> static inline void __stcg(T *p, T v) { *p = v; }
static inline int4 __ldcg(const int4 *p) { int4 v; memcpy(&v, p, sizeof v); return v; }
static inline uint4 __ldcg(const uint4 *p) { uint4 v; memcpy(&v, p, sizeof v); return v; }
template <typename T, typename V> static inline T atomicAdd(T *p, V v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
template <typename T, typename V> static inline T atomicOr(T *p, V v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_SEQ_CST); }
template <typename T, typename V> static inline T atomicAnd(T *p, V v) { return __atomic_fetch_and(p, (T)v, __ATOMIC_SEQ_CST); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline int __float2int_rn(float v) { return (int)lrintf(v); }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
using std::max;
using std::min;

// run grid_x * grid_y * grid_z blocks of `block` threads (a multiple of 32), one block after the other: the `block` fibers are
// created once per launch and walk through the blocks together (a block barrier between two blocks keeps the function-local
// "shared memory" statics of one block from being touched by the next)
static long &emu_schedule_mode() {  // initial value from the environment; emu_set_schedule() changes it at run time
    static long mode = getenv("EMU_SCHEDULE") ? atol(getenv("EMU_SCHEDULE")) : 0;
    return mode;
}
extern "C" void emu_set_schedule(long mode) { emu_schedule_mode() = mode; }
static std::vector<char *> &emu_stacks() {
    static std::vector<char *> v;
    return v;
}
template <typename F>
static void emu_launch3(unsigned grid_x, unsigned grid_y, unsigned grid_z, unsigned block, F kernel_call) {
    emu_gridDim.x = grid_x;
    emu_gridDim.y = grid_y;
    emu_gridDim.z = grid_z;
    emu_blockDim.x = block;
    emu_block_bar = EmuBar();
    emu_block_bar.expected = block;
    std::vector<std::unique_ptr<EmuWarp>> warps;
    for (unsigned w = 0; w < (block + 31) / 32; w++) {
        warps.emplace_back(new EmuWarp());
        warps.back()->bar.expected = std::min(32u, block - 32 * w);
    }
    std::function<void()> body = [&] {
        for (unsigned bz = 0; bz < grid_z; bz++)
            for (unsigned by = 0; by < grid_y; by++)
                for (unsigned b = 0; b < grid_x; b++) {
                    // every fiber writes the same values; they all sit between the same two block barriers
                    emu_blockIdx.x = b;
                    emu_blockIdx.y = by;
                    emu_blockIdx.z = bz;
                    kernel_call();
                    emu_bar_wait(emu_block_bar);
                }
    };
    // blockIdx is a global: fibers leave the end-of-block barrier one after the other and the first one out moves blockIdx
    // on while the others still sit in the barrier loop -- they are past kernel_call, and each fiber rewrites the indices
    // before its own next kernel_call; between two end-of-block barriers all fibers are in the same block.
    emu_body = &body;
    std::vector<EmuFiber> fibers(block);
    auto &stacks = emu_stacks();
    while (stacks.size() < block) stacks.push_back((char *)aligned_alloc(64, EMU_STACK_BYTES));
    for (unsigned t = 0; t < block; t++) {
        EmuFiber &f = fibers[t];
        f.tid = t;
        f.warp = warps[t / 32].get();
        f.stack = stacks[t];
#if defined(__x86_64__)
        // initial frame: six callee-saved registers, the entry point as return address, a null return address above it
        uintptr_t top = ((uintptr_t)f.stack + EMU_STACK_BYTES) & ~(uintptr_t)15;
        void **sp = (void **)(top - 64);
        for (int k = 0; k < 6; k++) sp[k] = nullptr;
        sp[6] = (void *)&emu_fiber_entry;
        sp[7] = nullptr;
        f.sp = sp;
#else
        getcontext(&f.uc);
        f.uc.uc_stack.ss_sp = f.stack;
        f.uc.uc_stack.ss_size = EMU_STACK_BYTES;
        f.uc.uc_link = nullptr;
        makecontext(&f.uc, emu_fiber_entry, 0);
#endif
    }
    // EMU_SCHEDULE / emu_set_schedule(): 0 / unset = fibers resumed in ascending thread order, 1 = descending, >= 2 = a fresh pseudo-random
    // permutation every round (seed = the value).  Any order is a legal CUDA schedule: a kernel whose result depends on it
    // is missing a barrier (or an atomic).
    const long emu_schedule = emu_schedule_mode();
    static unsigned long long emu_sched_rng = 0x9E3779B97F4A7C15ull;
    emu_sched_rng ^= (unsigned long long)emu_schedule;
    std::vector<unsigned> order(block);
    for (unsigned t = 0; t < block; t++) order[t] = emu_schedule == 1 ? block - 1 - t : t;
    unsigned alive = block;
    while (alive) {
        if (emu_schedule >= 2)
            for (unsigned t = block - 1; t > 0; t--) {
                emu_sched_rng = emu_sched_rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[t], order[(unsigned)((emu_sched_rng >> 33) % (t + 1))]);
            }
        for (unsigned k = 0; k < block; k++) {
            const unsigned t = order[k];
            EmuFiber &f = fibers[t];
            if (f.done) continue;
            emu_current = &f;
#if defined(__x86_64__)
            emu_ctx_switch(&emu_main_sp, f.sp);
#else
            swapcontext(&emu_main_uc, &f.uc);
#endif
            if (f.done) alive--;
        }
    }
    emu_current = nullptr;
    emu_body = nullptr;
}
template <typename F>
static void emu_launch2(unsigned grid_x, unsigned grid_y, unsigned block, F kernel_call) { emu_launch3(grid_x, grid_y, 1, block, kernel_call); }
template <typename F>
static void emu_launch(unsigned grid, unsigned block, F kernel_call) { emu_launch3(grid, 1, 1, block, kernel_call); }
