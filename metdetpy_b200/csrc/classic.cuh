// ClassicDetector on the device (MetLib/Detector.py:245-299): 4-frame window,
//   diff23 = 255 - dilate3(|f(t-1) - f(t)| > thr)                       Detector.py:268-271
//   dst    = dilate3(|diff23 & f(t-3) - diff23 & f(t-2)| > thr)         Detector.py:274-281
// then cv2.HoughLinesP with the configured (not adaptive) maxLineGap (:284-289).  diff23 is 0 or 255, so
// the second difference is (~dilate(a)) & b with a = |f(t-1)-f(t)| > thr and b = |f(t-3)-f(t-2)| > thr:
//   classic_bits_kernel     one pass over the frames (each read once; a thread keeps its 32 pixels of the
//                           last three frames in registers), a and b as 1 bit per pixel
//   classic_spatial_kernel  dilate -> and-not -> dilate on 32-pixel words (warp strip, shuffles)
//   classic_expand_kernel   bits -> u8 mask + on-pixel count / list (input of the PPHT kernels)
// Any W, H (rows are not assumed to be word aligned).  Frames before the fourth give an empty mask
// (Detector.py:264-265 returns no lines there).
#pragma once
#include "common.cuh"
#include "spatial_kernel.cuh"

// the thread's (up to) 32 pixels of one frame as 8 packed u8x4 words; pixels beyond the row end read as 0
template <bool ALIGNED>
__device__ __forceinline__ void classic_load(unsigned (&w)[8], const uint8_t *p, int npx, const uint8_t *mask) {
    if (ALIGNED && npx == 32) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (4 * k + j < npx) v |= (unsigned)__ldg(p + 4 * k + j) << (8 * j);
            w[k] = v;
        }
    }
    if (mask) {  // Transform.mask_with on the fly ({0,1} bytes)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            unsigned m = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (4 * k + j < npx) m |= (unsigned)(mask[4 * k + j] ? 0xffu : 0u) << (8 * j);
            w[k] &= m;
        }
    }
}

// 32 bits: |x - y| > thr per pixel
__device__ __forceinline__ unsigned classic_gt(const unsigned (&x)[8], const unsigned (&y)[8], unsigned thr4) {
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const unsigned c = __vcmpgtu4(__vabsdiffu4(x[k], y[k]), thr4);  // 0xff per byte where greater
        r |= (((c & 0x08040201u) * 0x01010101u) >> 24) << (4 * k);
    }
    return r;
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256)
classic_bits_kernel(FrameSrc src, int W, int H, int Wb, long long t0, int T, const int *__restrict__ thr,
                    uint32_t *__restrict__ abits, uint32_t *__restrict__ bbits) {
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= H * Wb) return;
    const int wx = idx % Wb, y = idx / Wb;
    const int npx = min(32, W - wx * 32);
    const size_t off = (size_t)y * W + (size_t)wx * 32;
    const uint8_t *mk = src.mask ? src.mask + off : nullptr;
    unsigned f1[8], f2[8], f3[8], cur[8];
#pragma unroll
    for (int k = 0; k < 8; k++) f1[k] = f2[k] = f3[k] = 0;
    if (t0 - 1 >= 0) classic_load<ALIGNED>(f1, src.frame(t0 - 1) + off, npx, mk);
    if (t0 - 2 >= 0) classic_load<ALIGNED>(f2, src.frame(t0 - 2) + off, npx, mk);
    if (t0 - 3 >= 0) classic_load<ALIGNED>(f3, src.frame(t0 - 3) + off, npx, mk);
    // frames of the batch: contiguous in the caller's buffer (zero-copy) or consecutive ring slots
    int slot = src.cur ? (int)(t0 - src.t0) : (int)(t0 % src.R);
    const uint8_t *base = src.cur ? src.cur : src.ring;
    const int Rw = src.cur ? 0x7fffffff : src.R;
    for (int i = 0; i < T; i++) {
        classic_load<ALIGNED>(cur, base + (size_t)slot * src.HW + off, npx, mk);
        if (++slot == Rw) slot = 0;
        unsigned a = 0, b = 0;
        if (t0 + i >= 3) {
            const unsigned th = (unsigned)min(max(thr[i], 0), 255) * 0x01010101u;
            a = classic_gt(f1, cur, th);
            b = classic_gt(f3, f2, th);
        }
        const size_t o = ((size_t)i * H + y) * Wb + wx;
        abits[o] = a;
        bbits[o] = b;
#pragma unroll
        for (int k = 0; k < 8; k++) { f3[k] = f2[k]; f2[k] = f1[k]; f1[k] = cur[k]; }
    }
}

// dst = dilate3((~dilate3(a)) & b); out-of-image pixels are ignored by cv2.dilate (contribute 0).
// One warp = strip of SP_USE words (lanes 0 and 31 are halo words), walked top to bottom.
__global__ void __launch_bounds__(SP_WARPS * 32)
classic_spatial_kernel(const uint32_t *__restrict__ abits, const uint32_t *__restrict__ bbits, int H, int Wb, int rows,
                       int strips, int bands, uint32_t *__restrict__ dbits) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * SP_WARPS + warp;
    const int t = blockIdx.y;
    if (tile >= strips * bands) return;
    const int strip = tile % strips, band = tile / strips;
    const int wx = strip * SP_USE - 1 + lane;
    const bool lane_in = wx >= 0 && wx < Wb;
    const bool lane_out = lane_in && lane >= 1 && lane <= SP_USE;
    const int y0 = band * rows;
    const unsigned FULL = 0xffffffffu;
    const size_t fo = (size_t)t * H * Wb + (lane_in ? wx : 0);
    unsigned hdA0 = 0, hdA1 = 0, b_prev = 0, hc0 = 0, hc1 = 0;
    for (int yy = y0 - 2; yy < y0 + rows + 2; yy++) {
        const bool row_in = lane_in && yy >= 0 && yy < H;
        const unsigned a = row_in ? __ldg(abits + fo + (size_t)yy * Wb) : 0u;
        const unsigned b = row_in ? __ldg(bbits + fo + (size_t)yy * Wb) : 0u;
        const unsigned La = __shfl_up_sync(FULL, a, 1), Ra = __shfl_down_sync(FULL, a, 1);
        const unsigned hdA2 = a | (a << 1) | (La >> 31) | (a >> 1) | (Ra << 31);
        // row yy-1: c = ~dilate(a) & b   (b = 0 outside the image)
        const unsigned c = ~(hdA0 | hdA1 | hdA2) & b_prev;
        hdA0 = hdA1; hdA1 = hdA2; b_prev = b;
        const unsigned Lc = __shfl_up_sync(FULL, c, 1), Rc = __shfl_down_sync(FULL, c, 1);
        const unsigned hc2 = c | (c << 1) | (Lc >> 31) | (c >> 1) | (Rc << 31);
        // row yy-2: dst
        const int y = yy - 2;
        if (lane_out && y >= y0 && y < y0 + rows && y < H) {
            dbits[(size_t)t * H * Wb + (size_t)y * Wb + wx] = hc0 | hc1 | hc2;
        }
        hc0 = hc1; hc1 = hc2;
    }
}

// bits -> u8 mask (0 / 255) + on-pixel count and list; one thread per 32-pixel word; bits beyond the row
// end are masked off here
template <bool ALIGNED>
__global__ void __launch_bounds__(256)
classic_expand_kernel(const uint32_t *__restrict__ dbits, int W, int H, int Wb, uint8_t *__restrict__ dst,
                      unsigned *__restrict__ npoints, uint32_t *__restrict__ points, int cap) {
    const int idx = blockIdx.x * 256 + threadIdx.x, t = blockIdx.y;
    if (idx >= H * Wb) return;
    const int wx = idx % Wb, y = idx / Wb;
    const int npx = min(32, W - wx * 32);
    unsigned bits = dbits[(size_t)t * H * Wb + idx];
    if (npx < 32) bits &= (1u << npx) - 1u;
    uint8_t *o = dst + ((size_t)t * H + y) * W + (size_t)wx * 32;
    if (ALIGNED && npx == 32) {
        uint4 *o4 = reinterpret_cast<uint4 *>(o);
        o4[0] = make_uint4(nib_to_bytes(bits & 15u), nib_to_bytes((bits >> 4) & 15u), nib_to_bytes((bits >> 8) & 15u),
                           nib_to_bytes((bits >> 12) & 15u));
        o4[1] = make_uint4(nib_to_bytes((bits >> 16) & 15u), nib_to_bytes((bits >> 20) & 15u),
                           nib_to_bytes((bits >> 24) & 15u), nib_to_bytes(bits >> 28));
    } else {
        for (int j = 0; j < npx; j++) o[j] = ((bits >> j) & 1u) ? 255 : 0;
    }
    if (bits) {
        unsigned slot = atomicAdd(npoints + t, __popc(bits));
        unsigned ob = bits;
        while (ob) {
            const int bpos = __ffs(ob) - 1;
            ob &= ob - 1;
            if (slot < (unsigned)cap) points[(size_t)t * cap + slot] = ((unsigned)y << 16) | (unsigned)(wx * 32 + bpos);
            slot++;
        }
    }
}
