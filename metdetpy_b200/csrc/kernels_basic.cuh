// Noise-sample, threshold-recurrence, per-frame fused mask (generic path), stack readback,
// ordered compaction and max-stack kernels.  sm_100a.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// SNR_SW.update noise sample (MetLib/Detector.py:81-91), as exact integer sums:
//   per ROI pixel  sx = sum_t x, sxx = sum_t x^2 over the L window frames, m = sx // L,
//   acc[0] += sx - L*m (= sum of d),  acc[1] += sxx - 2*m*sx + L*m^2 (= sum of d^2).
// grid = (blocks over ROI pixels, T frames of the batch); frames that are no sample exit at once.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_noise_sample(long long tau, int n, long long std_interval) {
    return (tau > 1 && tau <= n) || (tau > n && std_interval > 0 && tau % std_interval == 0);
}

struct SampleList {  // frames of the batch that are noise samples (usually a handful)
    int count;       // < 0: not listed, every frame of the batch gets a grid row and tests itself
    int idx[63];
};

__global__ void __launch_bounds__(256)
noise_sample_kernel(FrameSrc src, int W, int n, long long timer0, long long std_interval, int r0,
                    int c0, int rh, int rw, unsigned long long *acc, long long min_tau,
                    const __grid_constant__ SampleList sl) {
    const int i = sl.count < 0 ? blockIdx.y : sl.idx[blockIdx.y];
    const long long tau = timer0 + i + 1;
    if (tau < min_tau || !is_noise_sample(tau, n, std_interval)) return;
    const int L = (int)(tau < n ? tau : n);
    const long long t = tau - 1;
    __shared__ const uint8_t *fp[256];  // window frame pointers, resolved once per block (longer windows resolve inline)
    const bool big = L > 256;
    for (int k = threadIdx.x; k < min(L, 256); k += blockDim.x) fp[k] = src.frame(t - k);
    __syncthreads();
    unsigned long long d1 = 0, d2 = 0;
    const int total = rh * rw;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int y = r0 + q / rw, x = c0 + q % rw;
        const size_t p = (size_t)y * W + x;
        const unsigned mk = src.mask ? src.mask[p] : 1u;
        unsigned sx = 0, sxx = 0;
#pragma unroll 4
        for (int k = 0; k < L; k++) {
            const unsigned v = (big ? src.frame(t - k) : fp[k])[p] * mk;
            sx += v;
            sxx += v * v;
        }
        const unsigned m = sx / (unsigned)L;
        d1 += sx - (unsigned)L * m;
        d2 += (unsigned long long)((long long)sxx - 2ll * m * sx + (long long)L * m * m);
    }
    // block reduction (integers: order-independent, deterministic)
    for (int o = 16; o; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    __shared__ unsigned long long s1[8], s2[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s1[w] = d1; s2[w] = d2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) { d1 += s1[k]; d2 += s2[k]; }
        atomicAdd(&acc[2 * i], d1);
        atomicAdd(&acc[2 * i + 1], d2);
    }
}

// Same sums, sixteen pixels per 16-byte load (W % 16 == 0, n <= 128, 16-byte aligned frames): the ROI rows are widened
// to 16-pixel alignment and the pixels outside [c0, c0 + rw) are masked out.  Per 32-bit word and frame: sum of squares
// of the four bytes by one IDP.4A (it only enters d2 summed over pixels), per-pixel sums as u16x2 pairs of the even and
// the odd bytes.  The loads of a window's frames are independent, so a thread keeps several in flight.
__global__ void __launch_bounds__(256)
noise_sample16_kernel(FrameSrc src, int W, int n, long long timer0, long long std_interval, int r0, int c0, int rh,
                      int rw, unsigned long long *acc, long long min_tau, const __grid_constant__ SampleList sl) {
    const int i = sl.count < 0 ? blockIdx.y : sl.idx[blockIdx.y];
    const long long tau = timer0 + i + 1;
    if (tau < min_tau || !is_noise_sample(tau, n, std_interval)) return;
    const int L = (int)(tau < n ? tau : n);
    const long long t = tau - 1;
    __shared__ const uint8_t *fp[128];
    for (int k = threadIdx.x; k < L; k += blockDim.x) fp[k] = src.frame(t - k);
    __syncthreads();
    const int c0a = c0 & ~15, c1a = (c0 + rw + 15) & ~15;
    const int gpr = (c1a - c0a) >> 4;  // 16-pixel groups per ROI row
    const int total = rh * gpr;
    unsigned long long d1 = 0, d2 = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int y = r0 + q / gpr, x = c0a + ((q % gpr) << 4);
        const size_t p = (size_t)y * W + x;
        unsigned keep[4];  // 0xff for the bytes that are ROI pixels (and unmasked)
#pragma unroll
        for (int w = 0; w < 4; w++) {
            keep[w] = 0u;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int xx = x + 4 * w + b;
                if (xx >= c0 && xx < c0 + rw) keep[w] |= 0xffu << (8 * b);
            }
        }
        if (src.mask) {
            const uint4 m = *reinterpret_cast<const uint4 *>(src.mask + p);
            keep[0] &= m.x * 0xffu; keep[1] &= m.y * 0xffu; keep[2] &= m.z * 0xffu; keep[3] &= m.w * 0xffu;
        }
        unsigned se[4] = {0, 0, 0, 0}, so[4] = {0, 0, 0, 0}, sq = 0;
#pragma unroll 10
        for (int k = 0; k < L; k++) {
            const uint4 v4 = __ldg(reinterpret_cast<const uint4 *>(fp[k] + p));
            const unsigned v[4] = {v4.x & keep[0], v4.y & keep[1], v4.z & keep[2], v4.w & keep[3]};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                se[w] += v[w] & 0x00ff00ffu;
                so[w] += (v[w] >> 8) & 0x00ff00ffu;
                sq = __dp4a(v[w], v[w], sq);  // 16 * 128 * 255^2 < 2^32
            }
        }
        d2 += sq;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const unsigned sx4[4] = {se[w] & 0xffffu, so[w] & 0xffffu, se[w] >> 16, so[w] >> 16};
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const unsigned sx = sx4[b], m = sx / (unsigned)L;
                d1 += sx - (unsigned)L * m;
                // sum of d^2 = sxx - 2 m sx + L m^2; the sxx part is already in; a masked-out pixel has sx = m = 0
                d2 += (unsigned long long)((long long)L * m * m - 2ll * m * sx);
            }
        }
    }
    for (int o = 16; o; o >>= 1) {
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
        d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    __shared__ unsigned long long s1[8], s2[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s1[w] = d1; s2[w] = d2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); k++) { d1 += s1[k]; d2 += s2[k]; }
        atomicAdd(&acc[2 * i], d1);
        atomicAdd(&acc[2 * i + 1], d2);
    }
}

static inline void launch_noise_samples(const FrameSrc &src, int W, int n, long long timer0, long long std_interval,
                                        const int *roi, unsigned long long *acc, long long min_tau, const SampleList &sl,
                                        int T, cudaStream_t st) {
    const int rh = roi[2] - roi[0], rw = roi[3] - roi[1];
    const int rows = sl.count < 0 ? T : sl.count;
    const bool mask_ok = !src.mask || ((uintptr_t)src.mask & 15) == 0;
    const bool base_ok = ((uintptr_t)src.ring & 15) == 0 && ((uintptr_t)src.cur & 15) == 0 && (src.HW & 15) == 0;
    if (W % 16 == 0 && n <= 128 && mask_ok && base_ok) {
        const int groups = rh * ((((roi[1] + rw + 15) & ~15) - (roi[1] & ~15)) >> 4);
        const int gx = std::max(1, std::min((groups + 255) / 256, 1184));
        noise_sample16_kernel<<<dim3(gx, rows), 256, 0, st>>>(src, W, n, timer0, std_interval, roi[0], roi[1], rh, rw, acc,
                                                            min_tau, sl);
    } else {
        const int gx = std::max(1, std::min((rh * rw + 255) / 256, 592));
        noise_sample_kernel<<<dim3(gx, rows), 256, 0, st>>>(src, W, n, timer0, std_interval, roi[0], roi[1], rh, rw, acc,
                                                          min_tau, sl);
    }
}

// ------------------------------------------------------------------------------------------
// Scalar recurrence of one batch, one thread: EMA.update (MetLib/utils.py:334-368) on every noise
// sample, then LineDetector.update's threshold rule (MetLib/Detector.py:225-229, :177-183).
// Explicit _rn intrinsics keep nvcc from contracting a*b+c into an FMA: the reference is plain
// IEEE double arithmetic in CPython.
// ------------------------------------------------------------------------------------------
__global__ void threshold_kernel(DevState *st, const unsigned long long *acc, int T,
                                 long long timer0, int n, long long std_interval,
                                 long long roi_pixels, int adaptive, int sens, int *thr_out,
                                 double *thrf_out, double *snr_out) {
    if (threadIdx.x || blockIdx.x) return;
    DevState s = *st;
    for (int i = 0; i < T; i++) {
        const long long tau = timer0 + i + 1;
        if (is_noise_sample(tau, n, std_interval)) {
            const int L = (int)(tau < n ? tau : n);
            const double N = (double)(L * roi_pixels);
            const double s1 = (double)acc[2 * i], s2 = (double)acc[2 * i + 1];
            const double mean = __ddiv_rn(s1, N);
            double var = __dsub_rn(__ddiv_rn(s2, N), __dmul_rn(mean, mean));
            if (var < 0) var = 0;
            const double sigma = __dsqrt_rn(var);
            if (s.ema_warm != 0.0) {
                const double k =
                    __dmul_rn(__dmul_rn((double)s.ema_t, __dsub_rn(1.0, s.ema_init_m)), s.ema_warm);
                if (k < 1.0) {
                    const double u = __dsub_rn(1.0, k);
                    s.ema_cur_m = __dmul_rn(s.ema_init_m, __dsub_rn(1.0, __dmul_rn(u, u)));
                } else {
                    s.ema_warm = 0.0;
                    s.ema_cur_m = s.ema_init_m;
                }
            }
            s.ema_value = __dadd_rn(__dmul_rn(s.ema_cur_m, s.ema_value),
                                    __dmul_rn(__dsub_rn(1.0, s.ema_cur_m), sigma));
            s.ema_t++;
        }
        if (adaptive && s.ema_value != 0.0) {
            const double x2 = __dmul_rn(s.ema_value, s.ema_value);
            const double a = sens == MDB_SENS_LOW ? 2.0 : (sens == MDB_SENS_NORMAL ? 1.2 : 0.9);
            const double b = sens == MDB_SENS_LOW ? 4.4 : (sens == MDB_SENS_NORMAL ? 3.6 : 3.0);
            s.thr_float = __dadd_rn(__dmul_rn(a, x2), b);
            s.bi_threshold = (int)rint(s.thr_float);  // Python round(): half to even
        }
        thr_out[i] = s.bi_threshold;
        thrf_out[i] = s.thr_float;
        snr_out[i] = s.ema_value;
    }
    *st = s;
}

// ------------------------------------------------------------------------------------------
// Generic per-frame fused mask kernel ("v1"): one launch per frame, the n-frame ring is re-read.
// It serves the per-frame drop-in API (update(); detect()), arbitrary W/H, and is the on-device
// cross-check of the time-tiled streaming kernel.
//   stack max / floor-mean  (utils.py:288-307)  -> diff (Detector.py:327-328)
//   -> medianBlur 3 (:329) -> threshold (:332) -> MORPH_CLOSE 3x3 (:335)
//   -> dynamic mask: not on in all of the last L act frames, erode 3x3, multiply (:234-242)
//   -> dst, on-pixel list.
// Tile 64x16 outputs, 4-pixel halo (median 1 + dilate 1 + erode 1 + dy-erode 1).
// ------------------------------------------------------------------------------------------
#define V1_TW 64
#define V1_TH 16
#define V1_RW (V1_TW + 8)
#define V1_RH (V1_TH + 8)

__device__ __forceinline__ void cswap(int &a, int &b) {
    int lo = min(a, b), hi = max(a, b);
    a = lo; b = hi;
}
__device__ __forceinline__ int median9(int p0, int p1, int p2, int p3, int p4, int p5, int p6,
                                       int p7, int p8) {
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p1); cswap(p3, p4); cswap(p6, p7);
    cswap(p1, p2); cswap(p4, p5); cswap(p7, p8); cswap(p0, p3); cswap(p5, p8); cswap(p4, p7);
    cswap(p3, p6); cswap(p1, p4); cswap(p2, p5); cswap(p4, p7); cswap(p4, p2); cswap(p6, p4);
    cswap(p4, p2);
    return p4;
}

__global__ void __launch_bounds__(256)
fused_frame_kernel(FrameSrc src, int W, int H, int n, long long t, int L, long long d, int Ldy,
                   int dy_on, const int *thr_ptr, ActRing ring, uint8_t *dst, unsigned *npoints,
                   uint32_t *points, int cap) {
    __shared__ uint8_t sA[V1_RH][V1_RW], sB[V1_RH][V1_RW];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * V1_TW - 4, y0 = blockIdx.y * V1_TH - 4;
    const int thr = *thr_ptr;
    const int nwin = (int)((t + 1) < n ? (t + 1) : n);  // frames that exist in the window
    __shared__ const uint8_t *fp[256];  // window frame pointers, resolved once per block (longer windows resolve inline)
    const bool big = nwin > 256;
    for (int k = tid; k < min(nwin, 256); k += 256) fp[k] = src.frame(t - k);
    __syncthreads();
    // stage 0: diff on the whole region, replicated border (medianBlur's border mode)
    for (int q = tid; q < V1_RW * V1_RH; q += 256) {
        const int rx = q % V1_RW, ry = q / V1_RW;
        const int gx = min(max(x0 + rx, 0), W - 1), gy = min(max(y0 + ry, 0), H - 1);
        const size_t p = (size_t)gy * W + gx;
        const unsigned mk = src.mask ? src.mask[p] : 1u;
        unsigned mx = 0, sm = 0;
#pragma unroll 4
        for (int k = 0; k < nwin; k++) {
            const unsigned v = (big ? src.frame(t - k) : fp[k])[p] * mk;
            mx = max(mx, v);
            sm += v;
        }
        sA[ry][rx] = (uint8_t)(mx - sm / (unsigned)L);
    }
    __syncthreads();
    // stage 1: bin = median3x3(diff) > thr on region shrunk by 1; outside the image -> 0
    for (int q = tid; q < (V1_RW - 2) * (V1_RH - 2); q += 256) {
        const int rx = 1 + q % (V1_RW - 2), ry = 1 + q / (V1_RW - 2);
        const int gx = x0 + rx, gy = y0 + ry;
        int v = 0;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            int m = median9(sA[ry - 1][rx - 1], sA[ry - 1][rx], sA[ry - 1][rx + 1], sA[ry][rx - 1],
                            sA[ry][rx], sA[ry][rx + 1], sA[ry + 1][rx - 1], sA[ry + 1][rx],
                            sA[ry + 1][rx + 1]);
            v = m > thr;
        }
        sB[ry][rx] = (uint8_t)v;
    }
    __syncthreads();
    // stage 2: dilate on region shrunk by 2; outside the image -> 1 (ignored by the erosion)
    for (int q = tid; q < (V1_RW - 4) * (V1_RH - 4); q += 256) {
        const int rx = 2 + q % (V1_RW - 4), ry = 2 + q / (V1_RW - 4);
        const int gx = x0 + rx, gy = y0 + ry;
        int v = 1;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            v = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) v |= sB[ry + dy][rx + dx];
        }
        sA[ry][rx] = (uint8_t)v;
    }
    __syncthreads();
    // stage 3: erode -> act on region shrunk by 3 (+ run counters and dy candidate mask m)
    for (int q = tid; q < (V1_RW - 6) * (V1_RH - 6); q += 256) {
        const int rx = 3 + q % (V1_RW - 6), ry = 3 + q / (V1_RW - 6);
        const int gx = x0 + rx, gy = y0 + ry;
        int act = 0, m = 1;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            act = 1;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) act &= sA[ry + dy][rx + dx];
            if (dy_on && act) {  // m = 0 only if the pixel was on in all of the last Ldy act frames
                m = 0;
                for (int k = 1; k < Ldy; k++)
                    if (!((ring.frame(d - k)[(size_t)gy * ring.Wb + (gx >> 5)] >> (gx & 31)) & 1u)) { m = 1; break; }
            }
        }
        sB[ry][rx] = (uint8_t)(act | (m << 1));
    }
    __syncthreads();
    // stage 4: dst = act & erode(m) on the tile; emit mask bytes and the on-pixel list
    for (int q = tid; q < V1_TW * V1_TH; q += 256) {
        const int rx = 4 + q % V1_TW, ry = 4 + q / V1_TW;
        const int gx = x0 + rx, gy = y0 + ry;
        int on = 0;
        const bool inside = gx < W && gy < H;
        {   // publish this frame's act bits (32 consecutive pixels of one row per warp)
            const unsigned ab = __ballot_sync(0xffffffffu, inside && (sB[ry][rx] & 1));
            if ((tid & 31) == 0 && inside) ring.frame(d)[(size_t)gy * ring.Wb + (gx >> 5)] = ab;
        }
        if (inside) {
            int mm = 2;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++)
#pragma unroll
                for (int dx = -1; dx <= 1; dx++) mm &= sB[ry + dy][rx + dx];
            on = (sB[ry][rx] & 1) & (mm >> 1);
            dst[(size_t)gy * W + gx] = on ? 255 : 0;
        }
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (b) {
            const int lane = tid & 31;
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(npoints, __popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (on) {
                const unsigned slot = base + __popc(b & ((1u << lane) - 1));
                if (slot < (unsigned)cap) points[slot] = ((unsigned)gy << 16) | (unsigned)gx;
            }
        }
    }
}

// SlidingWindow.max / .mean / .sum of the main window (utils.py:288-300), for API read-back.
__global__ void stack_readback_kernel(FrameSrc src, size_t HW, int n, long long t, int L,
                                      uint8_t *mx_out, uint8_t *mean_out, uint32_t *sum_out) {
    const int nwin = (int)((t + 1) < n ? (t + 1) : n);
    __shared__ const uint8_t *fp[256];
    const bool big = nwin > 256;
    for (int k = threadIdx.x; k < min(nwin, 256); k += blockDim.x) fp[k] = src.frame(t - k);
    __syncthreads();
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < HW;
         p += (size_t)gridDim.x * blockDim.x) {
        const unsigned mk = src.mask ? src.mask[p] : 1u;
        unsigned mx = 0, sm = 0;
        for (int k = 0; k < nwin; k++) {
            const unsigned v = (big ? src.frame(t - k) : fp[k])[p] * mk;
            mx = max(mx, v);
            sm += v;
        }
        if (mx_out) mx_out[p] = (uint8_t)mx;
        if (mean_out) mean_out[p] = (uint8_t)(sm / (unsigned)L);
        if (sum_out) sum_out[p] = sm;
    }
}

// one ring frame as the reference's window holds it (masked when the library does the loader's mask_with)
__global__ void window_frame_kernel(const uint8_t *frame, const uint8_t *mask, size_t HW, uint8_t *out) {
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < HW; p += (size_t)gridDim.x * blockDim.x)
        out[p] = (uint8_t)(frame[p] * mask[p]);
}

// SlidingWindow.std with calc_std=True, force_int=True, dtype=uint8 (MetLib/utils.py:309-321):
// sqrt(mean((square_sum - square(sum) // L) // L)) in numpy's uint32 arithmetic; this kernel leaves the exact integer total
// of the per-pixel terms in *total (the mean and the square root are two double operations on the host side of the ABI).
__global__ void stack_std_kernel(FrameSrc src, size_t HW, int n, long long t, int L, unsigned long long *total) {
    const int nwin = (int)((t + 1) < n ? (t + 1) : n);
    __shared__ const uint8_t *fp[256];
    const bool big = nwin > 256;
    for (int k = threadIdx.x; k < min(nwin, 256); k += blockDim.x) fp[k] = src.frame(t - k);
    __syncthreads();
    unsigned long long acc = 0;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < HW; p += (size_t)gridDim.x * blockDim.x) {
        const unsigned mk = src.mask ? src.mask[p] : 1u;
        unsigned sm = 0, sq = 0;
        for (int k = 0; k < nwin; k++) {
            const unsigned v = (big ? src.frame(t - k) : fp[k])[p] * mk;
            sm += v;
            sq += v * v;
        }
        acc += (unsigned)(sq - (unsigned)(sm * sm) / (unsigned)L) / (unsigned)L;  // uint32 wrap-around as in numpy
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(total, acc);
}

// stacker.MaxImgContainer / MergeFunction.max: out = max(out?, frames[0..T)) element-wise.
// 16-byte lanes; per-byte max via the masked even/odd u16x2 trick (VIMNMX.U16x2 is native on
// sm_100a, the u8x4 SIMD max is emulated).
__device__ __forceinline__ unsigned bytemax4(unsigned a, unsigned b) {
    const unsigned ae = a & 0x00ff00ffu, be = b & 0x00ff00ffu;
    const unsigned ao = a & 0xff00ff00u, bo = b & 0xff00ff00u;
    return __vmaxu2(ae, be) | __vmaxu2(ao, bo);
}

__global__ void __launch_bounds__(256)
max_stack_kernel(const uint8_t *frames, int T, size_t frame_bytes, uint8_t *out, int accumulate) {
    const size_t nvec = frame_bytes / 16;
    const bool aligned = (((uintptr_t)frames | (uintptr_t)out | frame_bytes) & 15) == 0;
    if (aligned) {
        for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec;
             v += (size_t)gridDim.x * blockDim.x) {
            uint4 acc = accumulate ? reinterpret_cast<const uint4 *>(out)[v]
                                   : make_uint4(0, 0, 0, 0);
            for (int t = 0; t < T; t++) {
                const uint4 x = __ldcs(reinterpret_cast<const uint4 *>(frames + t * frame_bytes) + v);
                acc.x = bytemax4(acc.x, x.x); acc.y = bytemax4(acc.y, x.y);
                acc.z = bytemax4(acc.z, x.z); acc.w = bytemax4(acc.w, x.w);
            }
            reinterpret_cast<uint4 *>(out)[v] = acc;
        }
    } else {
        for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < frame_bytes;
             p += (size_t)gridDim.x * blockDim.x) {
            unsigned acc = accumulate ? out[p] : 0;
            for (int t = 0; t < T; t++) acc = max(acc, (unsigned)frames[t * frame_bytes + p]);
            out[p] = (uint8_t)acc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// FastGaussianContainer (MetLib/stacker.py:52-59) / FastGaussianParam.__add__ (MetLib/utils.py:485-493):
// per element the sum of the frames as uint16 and the sum of their squares as uint32 -- both WRAP exactly
// as numpy's fixed-width adds do (the reference documents the overflow as a limit of the class,
// utils.py:427).  One pass over the T frames; a thread owns 4 consecutive bytes.
__global__ void __launch_bounds__(256)
gauss_stack_kernel(const uint8_t *__restrict__ frames, int T, size_t frame_bytes, uint16_t *__restrict__ sum,
                   uint32_t *__restrict__ sq, int accumulate) {
    const bool aligned = (((uintptr_t)frames | (uintptr_t)sum | (uintptr_t)sq | frame_bytes) & 15) == 0;
    if (aligned) {  // 16 bytes per thread, four frames' loads in flight
        const size_t nvec = frame_bytes / 16;
        for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += (size_t)gridDim.x * blockDim.x) {
            unsigned s[16], q[16];
#pragma unroll
            for (int j = 0; j < 16; j++) s[j] = q[j] = 0;
            if (accumulate) {
#pragma unroll
                for (int j = 0; j < 16; j++) { s[j] = sum[v * 16 + j]; q[j] = sq[v * 16 + j]; }
            }
            const uint4 *fp = reinterpret_cast<const uint4 *>(frames) + v;
            const size_t fstride = frame_bytes / 16;
            int t = 0;
            for (; t + 4 <= T; t += 4) {
                uint4 x[4];
#pragma unroll
                for (int u = 0; u < 4; u++) x[u] = __ldcs(fp + (size_t)(t + u) * fstride);
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const unsigned w[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        const unsigned b = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                        s[j] += b;
                        q[j] += b * b;
                    }
                }
            }
            for (; t < T; t++) {
                const uint4 xv = __ldcs(fp + (size_t)t * fstride);
                const unsigned w[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const unsigned b = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
                    s[j] += b;
                    q[j] += b * b;
                }
            }
            uint4 *so = reinterpret_cast<uint4 *>(sum + v * 16);
            so[0] = make_uint4((s[0] & 0xffffu) | (s[1] << 16), (s[2] & 0xffffu) | (s[3] << 16),
                               (s[4] & 0xffffu) | (s[5] << 16), (s[6] & 0xffffu) | (s[7] << 16));
            so[1] = make_uint4((s[8] & 0xffffu) | (s[9] << 16), (s[10] & 0xffffu) | (s[11] << 16),
                               (s[12] & 0xffffu) | (s[13] << 16), (s[14] & 0xffffu) | (s[15] << 16));
            uint4 *qo = reinterpret_cast<uint4 *>(sq + v * 16);
#pragma unroll
            for (int j = 0; j < 4; j++) qo[j] = make_uint4(q[4 * j], q[4 * j + 1], q[4 * j + 2], q[4 * j + 3]);
        }
        return;
    }
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < frame_bytes; p += (size_t)gridDim.x * blockDim.x) {
        unsigned s = accumulate ? sum[p] : 0u, q = accumulate ? sq[p] : 0u;
        for (int t = 0; t < T; t++) {
            const unsigned b = frames[(size_t)t * frame_bytes + p];
            s += b;
            q += b * b;
        }
        sum[p] = (uint16_t)s;
        sq[p] = q;
    }
}
