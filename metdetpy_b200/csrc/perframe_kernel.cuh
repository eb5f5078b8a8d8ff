// Per-frame path with O(1) work in the window length (the reference's own call pattern, MetDetPy.py:197-198:
// `detector.update(frame); detector.detect()`).  SlidingWindow.update (utils.py:269-286) costs O(n) per frame in the
// reference because np.max rescans the ring; the batched kernels amortise the window over a batch.  For single frames
// the van Herk / Gil-Werman state itself stays resident in HBM instead:
//   S   u16 [HW]      running sum of the window           (utils.py:274-281)
//   P   u8  [HW]      prefix max of the current block of n frames (blocks anchored at global frame indices k*n)
//   SUF u8  [n][HW]   suffix maxima of the previous block: SUF[j] = max(x[b-n+j .. b-1])
// so that update(t) reads the new frame, the frame that leaves the window, S, P and one SUF plane (9 bytes per pixel
// instead of n + 2), writes the frame into its ring slot and emits the predicate bit max*L - sum > thr*L of
// Detector.py:325-328 (same bit plane the batched temporal kernels write; the spatial kernels take it from there).
// Once per n frames the suffix planes of the finished block are rebuilt (2(n-1) bytes per pixel, off the critical path).
// Frames are handled in groups of 16 pixels (one 16-byte load per plane); requires W % 32 == 0 like the streaming path.
#pragma once
#include "common.cuh"

#define PF_THREADS 256

__device__ __forceinline__ uint4 pf_ld(const uint8_t *p, size_t g) { return __ldg(reinterpret_cast<const uint4 *>(p) + g); }
__device__ __forceinline__ uint4 pf_max4(uint4 a, uint4 b) {
    return make_uint4(__vmaxu4(a.x, b.x), __vmaxu4(a.y, b.y), __vmaxu4(a.z, b.z), __vmaxu4(a.w, b.w));
}
// {0,1} mask bytes -> 0x00 / 0xff per byte, applied to the frame bytes (frame * mask, imgproc.py:96-101)
__device__ __forceinline__ uint4 pf_mask4(uint4 x, uint4 m) {
    return make_uint4(x.x & (m.x * 0xffu), x.y & (m.y * 0xffu), x.z & (m.z * 0xffu), x.w & (m.w * 0xffu));
}

// update(t): one thread = 16 pixels.
//  cur      the new frame (staging buffer), slot = its ring slot (written here), old = ring slot of frame t-n (may be
//           the same memory as slot; nullptr while t < n), suf = SUF[pos+1] (nullptr: window = current block only)
template <bool MASKED>
__global__ void __launch_bounds__(PF_THREADS)
pf_update_kernel(const uint8_t *__restrict__ cur, uint8_t *slot, const uint8_t *old, const uint8_t *__restrict__ mask,
                 uint16_t *S, uint8_t *P, const uint8_t *__restrict__ suf, int first_of_block, int L,
                 const int *__restrict__ thr_ptr, size_t g_begin, size_t groups, uint16_t *bits) {
    const size_t g = g_begin + blockIdx.x * (size_t)PF_THREADS + threadIdx.x;  // groups [g_begin, groups) of the frame
    if (g >= groups) return;
    uint4 x = pf_ld(cur, g);
    uint4 o = make_uint4(0, 0, 0, 0), sf = o, p = o;
    if (old) o = *(reinterpret_cast<const uint4 *>(old) + g);  // plain load: may alias `slot`
    if (suf) sf = pf_ld(suf, g);
    if (!first_of_block) p = *(reinterpret_cast<const uint4 *>(P) + g);
    uint4 s0 = *(reinterpret_cast<const uint4 *>(S) + 2 * g), s1 = *(reinterpret_cast<const uint4 *>(S) + 2 * g + 1);
    *(reinterpret_cast<uint4 *>(slot) + g) = x;  // the ring holds the frames as they arrive (masking is applied on load)
    if (MASKED) {
        const uint4 m = pf_ld(mask, g);
        x = pf_mask4(x, m);
        o = pf_mask4(o, m);
    }
    p = pf_max4(p, x);
    *(reinterpret_cast<uint4 *>(P) + g) = p;
    const uint4 mx = pf_max4(p, sf);
    const int thr = min(max(*thr_ptr, 0), 255);
    const int tl = thr * L;
    const unsigned xs[4] = {x.x, x.y, x.z, x.w}, os[4] = {o.x, o.y, o.z, o.w}, ms[4] = {mx.x, mx.y, mx.z, mx.w};
    unsigned sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};  // 16 u16 sums, pixel order
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int px = 4 * k + b;
            const int xv = (xs[k] >> (8 * b)) & 0xff, ov = (os[k] >> (8 * b)) & 0xff, mv = (ms[k] >> (8 * b)) & 0xff;
            int sv = (sw[px >> 1] >> (16 * (px & 1))) & 0xffff;
            sv = sv + xv - ov;
            sw[px >> 1] = (sw[px >> 1] & ~(0xffffu << (16 * (px & 1)))) | ((unsigned)sv << (16 * (px & 1)));
            out |= (unsigned)(mv * L - sv > tl) << px;
        }
    }
    *(reinterpret_cast<uint4 *>(S) + 2 * g) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
    *(reinterpret_cast<uint4 *>(S) + 2 * g + 1) = make_uint4(sw[4], sw[5], sw[6], sw[7]);
    bits[g] = (uint16_t)out;
}

// Suffix maxima over ring frames hi, hi-1, ..., lo (global frame indices, all >= 0): SUF[j0 - k] = max(x[hi-k .. hi]).
// After update(t) with t % n == n-1: hi = t, lo = t-n+2, j0 = n-1 (the finished block).
template <bool MASKED>
__global__ void __launch_bounds__(PF_THREADS)
pf_suffix_kernel(FrameSrc src, long long hi, long long lo, int j0, uint8_t *SUF, size_t groups) {
    const size_t g = blockIdx.x * (size_t)PF_THREADS + threadIdx.x;
    if (g >= groups) return;
    uint4 a = make_uint4(0, 0, 0, 0), m = a;
    if (MASKED) m = pf_ld(src.mask, g);
    int j = j0;
    for (long long t = hi; t >= lo; t--, j--) {
        uint4 x = pf_ld(src.frame(t), g);
        if (MASKED) x = pf_mask4(x, m);
        a = pf_max4(a, x);
        *(reinterpret_cast<uint4 *>(SUF + (size_t)j * src.HW) + g) = a;
    }
}

// State for "timer frames seen" rebuilt from the ring (after batched calls, reset or seek): S over the last
// min(n, timer) frames, P over the frames of the current block, SUF[j] (j > pos) of the previous block.
template <bool MASKED>
__global__ void __launch_bounds__(PF_THREADS)
pf_rebuild_kernel(FrameSrc src, long long timer, int n, uint16_t *S, uint8_t *P, uint8_t *SUF, size_t groups) {
    const size_t g = blockIdx.x * (size_t)PF_THREADS + threadIdx.x;
    if (g >= groups) return;
    uint4 m = make_uint4(0, 0, 0, 0);
    if (MASKED) m = pf_ld(src.mask, g);
    const int pos = (int)(timer % n);
    const long long b = timer - pos;  // first frame of the current block
    unsigned sw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint4 p = make_uint4(0, 0, 0, 0), a = p;
    const long long first = timer - n + 1 > 0 ? timer - n + 1 : 0;
    // previous block's tail, newest first: frames b-1 .. max(0, timer-n+1) -> SUF[n-1], SUF[n-2], ...
    int j = n - 1;
    for (long long t = b - 1; t >= first && j > pos; t--, j--) {
        uint4 x = pf_ld(src.frame(t), g);
        if (MASKED) x = pf_mask4(x, m);
        a = pf_max4(a, x);
        *(reinterpret_cast<uint4 *>(SUF + (size_t)j * src.HW) + g) = a;
        const unsigned xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            sw[2 * k] += (xs[k] & 0xff) | ((xs[k] << 8) & 0xff0000);
            sw[2 * k + 1] += ((xs[k] >> 16) & 0xff) | ((xs[k] >> 8) & 0xff0000);
        }
    }
    for (; j > pos; j--) *(reinterpret_cast<uint4 *>(SUF + (size_t)j * src.HW) + g) = a;  // frames before 0 are zeros
    if (timer - n >= 0) {  // the oldest frame of the window: in the sum, in no suffix that will still be used
        uint4 x = pf_ld(src.frame(timer - n), g);
        if (MASKED) x = pf_mask4(x, m);
        const unsigned xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            sw[2 * k] += (xs[k] & 0xff) | ((xs[k] << 8) & 0xff0000);
            sw[2 * k + 1] += ((xs[k] >> 16) & 0xff) | ((xs[k] >> 8) & 0xff0000);
        }
    }
    for (long long t = b > first ? b : first; t < timer; t++) {
        uint4 x = pf_ld(src.frame(t), g);
        if (MASKED) x = pf_mask4(x, m);
        p = pf_max4(p, x);
        const unsigned xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            sw[2 * k] += (xs[k] & 0xff) | ((xs[k] << 8) & 0xff0000);
            sw[2 * k + 1] += ((xs[k] >> 16) & 0xff) | ((xs[k] >> 8) & 0xff0000);
        }
    }
    *(reinterpret_cast<uint4 *>(P) + g) = p;
    *(reinterpret_cast<uint4 *>(S) + 2 * g) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
    *(reinterpret_cast<uint4 *>(S) + 2 * g + 1) = make_uint4(sw[4], sw[5], sw[6], sw[7]);
}
