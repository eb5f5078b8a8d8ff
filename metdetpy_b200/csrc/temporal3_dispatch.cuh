// Shape table and launcher of temporal3_kernel (temporal3_kernel.cuh).
// A window of n frames runs as (U, BL, P, K): U = frames per unrolled body (ring registers), P = n / U (1: whole ring in
// registers, 2: younger half in a shared-memory page), BL = van Herk block length (divides U), K = frames in flight per
// thread (divides U).  Windows without an entry (odd n > 32, n > 64, primes above 8 ...) keep the second-generation kernel.
#pragma once
#include "temporal3_kernel.cuh"

#define T3_SLACK_PLANES 32  // bit planes behind a batch that a thread's last, partial block may write (BL - 1 at most)

struct T3Variant {
    int variant = 0;  // test / tuning hook: alternative shape for the same window (0 = default)
};

struct T3Table {  // per-frame table of the bulk-copy kernels, one per predicate-bit buffer
    uint2 *d_tab[2] = {nullptr, nullptr};
    int cap = 0;
};

template <int U, int BL, int P, int K, int MINB, int FEED>
static inline int t3_launch_shape(bool masked, const FrameSrc &src, long long t0, int T, int HWG, const int *thr,
                                  uint8_t *bits, const T3Table &tabs, int parity, cudaStream_t st) {
    typedef t3::Layout<U, BL, P, FEED, K> LY;
    static_assert(BL - 1 <= T3_SLACK_PLANES, "slack planes");
    const size_t smem = LY::smem_bytes(T);
    const int grid = (HWG + T3_NT - 1) / T3_NT;
    if (smem > 200 * 1024) return -2;
    uint2 *gtab = nullptr;
    if (FEED == 2 && (((uintptr_t)src.cur & 15) || (src.HW & 15) || (HWG & 1))) return -2;
    if (FEED == 1) {
        // bulk copies move 16-byte units: frame base, frame stride and a CTA's span must be multiples of 16
        const int Tp = (T + BL - 1) / BL * BL;
        if (((uintptr_t)src.cur & 15) || (src.HW & 15) || (HWG & 1) || Tp > tabs.cap || !tabs.d_tab[parity]) return -2;
        gtab = tabs.d_tab[parity];
        t3_table_kernel<<<(Tp + 255) / 256, 256, 0, st>>>(thr, t0, T, Tp, LY::N, gtab);
    }
    if (masked) {
        auto kfn = temporal3_kernel<U, BL, P, K, true, MINB, FEED>;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        kfn<<<grid, T3_NT, smem, st>>>(src, t0, T, HWG, thr, gtab, bits);
    } else {
        auto kfn = temporal3_kernel<U, BL, P, K, false, MINB, FEED>;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        kfn<<<grid, T3_NT, smem, st>>>(src, t0, T, HWG, thr, gtab, bits);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// 0: launched; -1: CUDA error; -2: no shape for this window (caller falls back to temporal2_kernel)
static inline int temporal3_launch(int n, int variant, const FrameSrc &src, long long t0, int T, int HWG, const int *thr,
                                   uint8_t *bits, const T3Table &tabs, int parity, cudaStream_t st) {
    if (!src.cur || src.t0 != t0) return -2;  // batch frames must be contiguous in memory
    if ((unsigned long long)src.HW * 64ull >= (1ull << 32)) return -2;  // 32-bit frame offsets inside a body
    const bool m = src.mask != nullptr;
#define T3_CASE(N_, U, BL, P, K, MINB) \
    case N_: return t3_launch_shape<U, BL, P, K, MINB, 0>(m, src, t0, T, HWG, thr, bits, tabs, parity, st);
#define T3_BULK(N_, U, BL, P, K, MINB) T3_FEED(N_, U, BL, P, K, MINB, 1)
#define T3_PF(N_, U, BL, P, K, MINB) T3_FEED(N_, U, BL, P, K, MINB, 2)
#define T3_FEED(N_, U, BL, P, K, MINB, F)                                                                     \
    case N_: {                                                                                                \
        const int rc = t3_launch_shape<U, BL, P, K, MINB, F>(m, src, t0, T, HWG, thr, bits, tabs, parity, st); \
        if (rc != -2) return rc;                                                                              \
        break;                                                                                                \
    }
    if (n == 30 && variant) {
        switch (variant) {
            T3_CASE(1, 30, 15, 1, 5, 3)
            T3_CASE(2, 30, 10, 1, 6, 3)
            T3_CASE(3, 30, 10, 1, 5, 3)
            T3_CASE(4, 30, 15, 1, 6, 3)
            T3_CASE(5, 30, 10, 1, 10, 3)
            T3_CASE(6, 30, 6, 1, 6, 3)
            T3_BULK(7, 30, 10, 1, 5, 4)
            T3_BULK(8, 30, 10, 1, 5, 3)
            T3_BULK(9, 30, 15, 1, 5, 3)
            T3_BULK(10, 30, 10, 1, 10, 4)
            T3_BULK(11, 30, 6, 1, 6, 4)
            T3_BULK(12, 30, 10, 1, 3, 4)
            T3_PF(13, 30, 10, 1, 5, 4)
            T3_PF(14, 30, 10, 1, 3, 4)
            T3_PF(15, 30, 10, 1, 6, 4)
            T3_PF(16, 30, 10, 1, 10, 3)
            T3_PF(17, 30, 10, 1, 2, 4)
            T3_CASE(18, 15, 15, 2, 5, 4)
            T3_CASE(19, 30, 10, 1, 6, 4)
            T3_CASE(20, 15, 5, 2, 5, 5)
            T3_CASE(21, 15, 5, 2, 15, 5)
            T3_CASE(22, 15, 15, 2, 15, 5)
            T3_CASE(23, 30, 10, 1, 15, 3)
            default: break;
        }
    }
    if (n == 60 && variant) {
        switch (variant) {
            T3_CASE(1, 30, 10, 2, 10, 3)
            T3_CASE(2, 30, 10, 2, 15, 3)
            T3_CASE(3, 30, 15, 2, 15, 3)
            T3_CASE(4, 30, 15, 2, 10, 3)
            T3_CASE(5, 20, 10, 3, 10, 3)
            T3_CASE(6, 15, 15, 4, 15, 2)
            T3_CASE(7, 20, 20, 3, 10, 2)
            T3_CASE(8, 30, 10, 2, 6, 3)
            default: break;
        }
    }
    switch (n) {
        T3_CASE(2, 2, 2, 1, 2, 6)
        T3_CASE(3, 3, 3, 1, 3, 6)
        T3_CASE(4, 4, 4, 1, 4, 6)
        T3_CASE(5, 5, 5, 1, 5, 6)
        T3_CASE(6, 6, 6, 1, 6, 6)
        T3_CASE(7, 7, 7, 1, 7, 6)
        T3_CASE(8, 8, 8, 1, 8, 6)
        T3_CASE(9, 9, 9, 1, 3, 6)
        T3_CASE(10, 10, 10, 1, 5, 6)
        T3_CASE(12, 12, 12, 1, 6, 5)
        T3_CASE(14, 14, 14, 1, 7, 5)
        T3_CASE(15, 15, 15, 1, 5, 5)
        T3_CASE(16, 16, 16, 1, 8, 5)
        // measured at 4K, n = 30 (scripts/t3_tune.py): half of the ring in the shared-memory page and all of a body's
        // frames in flight (no spills at 128 registers, 16 warps per SM) beats the whole ring in registers
        T3_CASE(18, 9, 9, 2, 9, 4)
        T3_CASE(20, 10, 10, 2, 10, 4)
        T3_CASE(21, 21, 7, 1, 7, 3)
        T3_CASE(24, 12, 12, 2, 12, 4)
        T3_CASE(25, 25, 5, 1, 5, 3)
        T3_CASE(28, 14, 14, 2, 14, 4)
        T3_CASE(30, 15, 15, 2, 15, 4)
        T3_CASE(32, 16, 16, 2, 16, 4)
        // long windows: 3 CTAs per SM (the 60-frame ring is half registers, half page); half a body of frames in flight
        // (4K n = 60: 2.45 -> 1.59 ms against the second generation)
        T3_CASE(36, 18, 9, 2, 9, 3)
        T3_CASE(40, 20, 10, 2, 10, 3)
        T3_CASE(48, 24, 12, 2, 12, 3)
        T3_CASE(50, 25, 5, 2, 5, 3)
        T3_CASE(60, 30, 15, 2, 15, 3)
        T3_CASE(64, 32, 16, 2, 16, 3)
        default: return -2;
    }
#undef T3_CASE
#undef T3_BULK
#undef T3_PF
#undef T3_FEED
}
