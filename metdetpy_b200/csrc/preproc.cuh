// Loader preprocessing on the device: what MetLib's video loader does to
// every decoded frame before the detector sees it,
//     cv2.resize(INTER_LINEAR) -> cv2.cvtColor(BGR2GRAY) -> img * mask        (Transform, imgproc.py:82-101,
//                                                                             chain built at videoloader.py:300-308)
//     -> np.max over exp_frame consecutive frames                              (MergeFunction.max, utils.py:203-204)
// fused into one pass: a thread produces one output pixel of one merged frame.  The arithmetic is
// OpenCV's 8-bit fixed point (11-bit weights, (((b*(h>>4))>>16)+...+2)>>2 vertical pass, 15-bit gray
// weights), restated and pinned by the CPU checker of the test-suite; results are bit-exact.
#pragma once
#include "common.cuh"

struct PreTap {  // one destination coordinate: two source offsets (already multiplied by the pixel / row
    int s0, s1;  // stride) and their 11-bit weights
    int w0, w1;
};

struct PreParams {
    int src_w, src_h, channels;  // 1 or 3 interleaved channels
    int dst_w, dst_h;
    int resize;                  // 0: same size, taps unused
    int rgb;                     // channel order of a 3-channel source: 0 = BGR, 1 = RGB
    int exp_frame;
    const PreTap *xt, *yt;       // [dst_w], [dst_h]
    const uint8_t *mask;         // [dst_h][dst_w] {0,1} or nullptr
};

__device__ __forceinline__ unsigned pre_vert(int h0, int h1, int b0, int b1) {
    return (unsigned)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

template <int C>
__global__ void __launch_bounds__(256)
preproc_kernel(PreParams P, const uint8_t *__restrict__ frames, int T, uint8_t *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    const int g = blockIdx.z;  // merged output frame
    if (x >= P.dst_w) return;
    const size_t src_row = (size_t)P.src_w * C, src_frame = src_row * P.src_h;
    PreTap tx, ty;
    if (P.resize) {
        tx = P.xt[x];
        ty = P.yt[y];
    } else {
        tx.s0 = tx.s1 = x * C; tx.w0 = 2048; tx.w1 = 0;
        ty.s0 = ty.s1 = y; ty.w0 = 2048; ty.w1 = 0;
    }
    const int f0 = g * P.exp_frame, f1 = min(f0 + P.exp_frame, T);
    unsigned best = 0;
    for (int f = f0; f < f1; f++) {
        const uint8_t *r0 = frames + (size_t)f * src_frame + (size_t)ty.s0 * src_row;
        const uint8_t *r1 = frames + (size_t)f * src_frame + (size_t)ty.s1 * src_row;
        unsigned v[C];
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (P.resize) {
                const int h0 = (int)__ldg(r0 + tx.s0 + c) * tx.w0 + (int)__ldg(r0 + tx.s1 + c) * tx.w1;
                const int h1 = (int)__ldg(r1 + tx.s0 + c) * tx.w0 + (int)__ldg(r1 + tx.s1 + c) * tx.w1;
                v[c] = pre_vert(h0, h1, ty.w0, ty.w1);
            } else {
                v[c] = __ldg(r0 + tx.s0 + c);
            }
        }
        unsigned gray;
        if (C == 3) {
            const unsigned b = P.rgb ? v[2 % C] : v[0], r = P.rgb ? v[0] : v[2 % C];
            gray = (b * 3735u + v[1 % C] * 19235u + r * 9798u + 16384u) >> 15;
        } else {
            gray = v[0];
        }
        best = max(best, gray);
    }
    if (P.mask) best *= P.mask[(size_t)y * P.dst_w + x];
    out[((size_t)g * P.dst_h + y) * P.dst_w + x] = (uint8_t)best;
}
