// Exact-order progressive probabilistic Hough transform, one CTA per frame, many frames in flight.
// Replaces cv2.HoughLinesP(dst, 1, pi/180, threshold, minLineLength, maxLineGap) at
// MetLib/Detector.py:347-352 and reproduces its visiting order (row-major point list + OpenCV's
// MWC RNG seeded per call), its float32 rho rounding and its in-loop un-voting, so the emitted
// segments are identical.
//
// Layout: every CTA ("slot") owns one int32 accumulator [180][numrho] and one H*W-bit mask in
// global memory (L2-resident in practice: only cells of actual points are ever touched, and they
// are reset by replaying the point list -- the arrays are never memset per frame).
// Thread n < 180 owns accumulator row n, so votes need no atomics; the arg-max over angles is a
// warp REDUX + one shared-memory hop.
#pragma once
#include "common.cuh"

__constant__ float c_trig[2 * MDB_HOUGH_ANGLES];  // (float)cos(n*theta), (float)sin(n*theta); host-computed

#define HOUGH_THREADS 256

__device__ __forceinline__ int rho_of(int x, int y, int n) {
    // plain float32 multiply/add, no FMA contraction (matches the compiled OpenCV loop)
    const float r = __fadd_rn(__fmul_rn((float)x, c_trig[2 * n]), __fmul_rn((float)y, c_trig[2 * n + 1]));
    return __float2int_rn(r);  // cvRound: round half to even
}

__device__ __forceinline__ void bitonic_sort_u32(uint32_t *a, int npow2, int tid, int nthreads) {
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += nthreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

// Process one frame's point list. keys: N sorted (row-major) point keys (y<<16|x); idx: N u32
// scratch. Both may live in shared or global memory.
__device__ void ppht_frame(const HoughParams &P, uint32_t *keys, uint32_t *idx, int N, int line_gap,
                           int32_t *accum, uint32_t *bitmap, uint32_t *walk, int32_t *lines_out,
                           int *nlines_out) {
    __shared__ int s_red[HOUGH_THREADS / 32];
    __shared__ int s_ctl[8];  // [0]=good, [1]=#walk pixels, [2]=found lines
    const int tid = threadIdx.x;
    const int W = P.W, H = P.H, numrho = P.numrho, half = (numrho - 1) / 2;
    volatile uint32_t *vbitmap = bitmap;

    // visiting order: OpenCV draws idx = rng % count and swap-removes; an in-place Fisher-Yates
    // over an index array leaves the visit sequence in idx[N-1], idx[N-2], ..., idx[0].
    for (int i = tid; i < N; i += HOUGH_THREADS) {
        idx[i] = i;
        const uint32_t k = keys[i];
        const size_t p = (size_t)(k >> 16) * W + (k & 0xffffu);
        atomicOr(&bitmap[p >> 5], 1u << (p & 31));
    }
    if (tid == 0) s_ctl[2] = 0;
    __syncthreads();
    if (tid == 0) {
        unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
        for (int count = N; count > 0; count--) {
            state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
            const unsigned r = (unsigned)state % (unsigned)count;
            const uint32_t a = idx[r], b = idx[count - 1];
            idx[r] = b;
            idx[count - 1] = a;
        }
    }
    __syncthreads();

    int32_t *myrow = accum + (size_t)(tid < MDB_HOUGH_ANGLES ? tid : 0) * numrho + half;
    for (int s = N - 1; s >= 0; s--) {
        const uint32_t key = keys[idx[s]];
        const int x = key & 0xffffu, y = key >> 16;
        const size_t p = (size_t)y * W + x;
        if (!((vbitmap[p >> 5] >> (p & 31)) & 1u)) continue;  // removed by an earlier line (uniform)
        int best = INT_MIN;
        if (tid < MDB_HOUGH_ANGLES) {
            const int r = rho_of(x, y, tid);
            const int v = myrow[r] + 1;
            myrow[r] = v;
            best = v * 256 + (255 - tid);  // max value first, lowest angle on ties
        }
        best = __reduce_max_sync(0xffffffffu, best);
        if ((tid & 31) == 0) s_red[tid >> 5] = best;
        __syncthreads();
        best = s_red[0];
#pragma unroll
        for (int k = 1; k < HOUGH_THREADS / 32; k++) best = max(best, s_red[k]);
        const int max_val = best >> 8;  // arithmetic shift: floor for negatives
        if (max_val < P.threshold) { __syncthreads(); continue; }
        const int max_n = 255 - (best & 255);

        if (tid == 0) {
            const float a = -c_trig[2 * max_n + 1], b = c_trig[2 * max_n];
            int x0 = x, y0 = y, dx0, dy0;
            bool xflag;
            if (fabsf(a) > fabsf(b)) {
                xflag = true;
                dx0 = a > 0 ? 1 : -1;
                dy0 = __float2int_rn(__fdiv_rn(__fmul_rn(b, 65536.0f), fabsf(a)));
                y0 = (y0 << 16) + 32768;
            } else {
                xflag = false;
                dy0 = b > 0 ? 1 : -1;
                dx0 = __float2int_rn(__fdiv_rn(__fmul_rn(a, 65536.0f), fabsf(b)));
                x0 = (x0 << 16) + 32768;
            }
            int ex[2] = {0, 0}, ey[2] = {0, 0};
            for (int k = 0; k < 2; k++) {
                int gap = 0, xx = x0, yy = y0;
                const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
                for (;; xx += dx, yy += dy) {
                    const int j1 = xflag ? xx : xx >> 16, i1 = xflag ? yy >> 16 : yy;
                    if (j1 < 0 || j1 >= W || i1 < 0 || i1 >= H) break;
                    const size_t q = (size_t)i1 * W + j1;
                    if ((vbitmap[q >> 5] >> (q & 31)) & 1u) { gap = 0; ex[k] = j1; ey[k] = i1; }
                    else if (++gap > line_gap) break;
                }
            }
            const int good = abs(ex[1] - ex[0]) >= P.min_len || abs(ey[1] - ey[0]) >= P.min_len;
            int nw = 0;
            for (int k = 0; k < 2; k++) {
                int xx = x0, yy = y0;
                const int dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
                for (;; xx += dx, yy += dy) {
                    const int j1 = xflag ? xx : xx >> 16, i1 = xflag ? yy >> 16 : yy;
                    const size_t q = (size_t)i1 * W + j1;
                    const uint32_t wv = vbitmap[q >> 5];
                    if ((wv >> (q & 31)) & 1u) {
                        if (good && nw < P.walk_cap) walk[nw++] = ((unsigned)i1 << 16) | (unsigned)j1;
                        vbitmap[q >> 5] = wv & ~(1u << (q & 31));
                    }
                    if (i1 == ey[k] && j1 == ex[k]) break;
                }
            }
            if (good) {
                const int li = s_ctl[2];
                if (li < P.max_lines) {
                    lines_out[4 * li] = ex[0]; lines_out[4 * li + 1] = ey[0];
                    lines_out[4 * li + 2] = ex[1]; lines_out[4 * li + 3] = ey[1];
                }
                s_ctl[2] = li + 1;
            }
            s_ctl[0] = good;
            s_ctl[1] = nw;
            __threadfence_block();
        }
        __syncthreads();
        if (s_ctl[0] && tid < MDB_HOUGH_ANGLES) {
            const int nw = s_ctl[1];
            volatile uint32_t *vwalk = walk;
            for (int k = 0; k < nw; k++) {
                const uint32_t wk = vwalk[k];
                myrow[rho_of(wk & 0xffffu, wk >> 16, tid)]--;
            }
        }
        __syncthreads();
    }
    // reset: zero every accumulator cell and mask word a point of this frame can have touched
    for (long long q = tid; q < (long long)N * MDB_HOUGH_ANGLES; q += HOUGH_THREADS) {
        const int i = (int)(q / MDB_HOUGH_ANGLES), n = (int)(q % MDB_HOUGH_ANGLES);
        const uint32_t k = keys[i];
        accum[(size_t)n * numrho + half + rho_of(k & 0xffffu, k >> 16, n)] = 0;
    }
    for (int i = tid; i < N; i += HOUGH_THREADS) {
        const uint32_t k = keys[i];
        const size_t p = (size_t)(k >> 16) * W + (k & 0xffffu);
        bitmap[p >> 5] = 0;
    }
    __syncthreads();
    if (tid == 0) *nlines_out = s_ctl[2];
    __syncthreads();
}

__device__ __forceinline__ int line_gap_of(const HoughParams &P, unsigned n_on) {
    // Detector.py:342-344: dst_sum = count / mask_area * 100; gap = max(0, 1 - dst_sum/0.05) * max_gap
    const double dst_sum = __dmul_rn(__ddiv_rn((double)n_on, P.mask_area), 100.0);
    double g = __dsub_rn(1.0, __ddiv_rn(dst_sum, 0.05));
    if (!(g > 0.0)) g = 0.0;
    g = __dmul_rn(g, (double)P.max_gap);
    return (int)rint(g);  // cvRound(maxLineGap)
}

// Shared-memory path: persistent CTAs loop over the frames of the batch.
__global__ void __launch_bounds__(HOUGH_THREADS)
hough_batch_kernel(HoughParams P, int T, const unsigned *npoints, const uint32_t *points,
                   int32_t *accum_slots, uint32_t *bitmap_slots, uint32_t *walk_slots,
                   int32_t *lines_out, int *nlines_out) {
    extern __shared__ uint32_t sm[];
    uint32_t *keys = sm, *idx = sm + P.cap;
    const size_t bm_words = ((size_t)P.W * P.H + 31) / 32;
    int32_t *accum = accum_slots + (size_t)blockIdx.x * MDB_HOUGH_ANGLES * P.numrho;
    uint32_t *bitmap = bitmap_slots + (size_t)blockIdx.x * bm_words;
    uint32_t *walk = walk_slots + (size_t)blockIdx.x * P.walk_cap;
    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const unsigned N = npoints[t];
        if (N == 0) { if (threadIdx.x == 0) nlines_out[t] = 0; continue; }
        if (N > (unsigned)P.cap) { if (threadIdx.x == 0) nlines_out[t] = -1; continue; }  // overflow path
        int np2 = 1;
        while (np2 < (int)N) np2 <<= 1;
        for (int i = threadIdx.x; i < np2; i += HOUGH_THREADS)
            keys[i] = i < (int)N ? points[(size_t)t * P.cap + i] : 0xFFFFFFFFu;
        __syncthreads();
        bitonic_sort_u32(keys, np2, threadIdx.x, HOUGH_THREADS);
        ppht_frame(P, keys, idx, (int)N, line_gap_of(P, N), accum, bitmap, walk,
                   lines_out + (size_t)t * P.max_lines * 4, nlines_out + t);
    }
}

// Overflow path: one frame, keys already sorted in global memory (compact_ordered_kernel).
__global__ void __launch_bounds__(HOUGH_THREADS)
hough_global_kernel(HoughParams P, const unsigned *n_ptr, uint32_t *keys, uint32_t *idx,
                    int32_t *accum, uint32_t *bitmap, uint32_t *walk, int32_t *lines_out,
                    int *nlines_out) {
    const unsigned N = *n_ptr;
    if (N == 0) { if (threadIdx.x == 0) *nlines_out = 0; return; }
    ppht_frame(P, keys, idx, (int)N, line_gap_of(P, N), accum, bitmap, walk, lines_out, nlines_out);
}
