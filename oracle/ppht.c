/*
 * oracle/ppht.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the progressive probabilistic Hough transform that the reference calls
 * at MetLib/Detector.py:347-352 (`cv2.HoughLinesP(dst, rho=1, theta=PI, threshold, minLineLength,
 * maxLineGap)`).  The arithmetic lives in the un-vendored third-party wheel opencv-python
 * (requirements.txt:1 `opencv-python>=4.9.0`; 4.13.0 in this image); this file restates the
 * published algorithm (Matas, Galambos, Kittler 2000, as implemented by OpenCV's
 * HoughLinesProbabilistic) and is pinned against cv2 itself in tests/test_oracle_hough.py and the
 * fixtures in tests/golden/hough.npz.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 *
 * Semantics restated:
 *  - 180 angles (theta = (float)(pi/180)), numrho = 2*(W+H)+1 for rho = 1, int32 accumulator
 *  - non-zero pixels collected in row-major order; RNG = OpenCV's MWC generator seeded with
 *    0xFFFFFFFFFFFFFFFF per call; idx = next() % count; swap-remove
 *  - vote over all angles, first strict maximum >= threshold wins
 *  - 16.16 fixed-point walk in both directions with a `lineGap` tolerance, Chebyshev length test,
 *    second walk clears the mask and (for accepted lines) un-votes the pixels
 *  - float32 rho evaluation; `vote_fma` / `dec_fma` select whether x*cos + y*sin is contracted
 *    into fmaf(x, cos, y*sin) (what the compiled wheel does in the voting loop) or not.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int round_half_even_f(float v) { return (int)lrintf(v); }

static inline int rho_index(int x, int y, const float *tr, int n, int use_fma) {
    float c = tr[2 * n], s = tr[2 * n + 1];
    float r;
    if (use_fma) {
        r = fmaf((float)x, c, (float)y * s);
    } else {
        volatile float t0 = (float)x * c; /* volatile: forbid contraction by the C compiler */
        volatile float t1 = (float)y * s;
        r = t0 + t1;
    }
    return round_half_even_f(r);
}

/* returns number of segments written (<= max_lines); total_found gets the uncapped count */
int oracle_ppht(const uint8_t *img, int W, int H, int threshold, int line_length, int line_gap,
                int max_lines, int32_t *out, int vote_fma, int dec_fma, int *total_found) {
    const int numangle = 180;
    const int numrho = 2 * (W + H) + 1;
    const float theta = (float)(3.14159265358979323846 / 180.0);
    float tr[360];
    for (int n = 0; n < numangle; n++) {
        tr[2 * n] = (float)cos((double)n * (double)theta);
        tr[2 * n + 1] = (float)sin((double)n * (double)theta);
    }
    int32_t *accum = (int32_t *)calloc((size_t)numangle * numrho, sizeof(int32_t));
    uint8_t *mask = (uint8_t *)malloc((size_t)W * H);
    int32_t *nz = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)W * H);
    int count = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint8_t on = img[(size_t)y * W + x] != 0;
            mask[(size_t)y * W + x] = on;
            if (on) { nz[2 * count] = x; nz[2 * count + 1] = y; count++; }
        }
    uint64_t state = 0xFFFFFFFFFFFFFFFFull;
    int nlines = 0, found = 0;
    const int half = (numrho - 1) / 2;
    for (; count > 0; count--) {
        state = (uint64_t)(uint32_t)state * 4164903690u + (uint32_t)(state >> 32);
        int idx = (int)((uint32_t)state % (uint32_t)count);
        int j = nz[2 * idx], i = nz[2 * idx + 1];
        nz[2 * idx] = nz[2 * (count - 1)];
        nz[2 * idx + 1] = nz[2 * (count - 1) + 1];
        if (!mask[(size_t)i * W + j]) continue;
        int max_val = threshold - 1, max_n = 0;
        for (int n = 0; n < numangle; n++) {
            int r = rho_index(j, i, tr, n, vote_fma) + half;
            int v = ++accum[(size_t)n * numrho + r];
            if (max_val < v) { max_val = v; max_n = n; }
        }
        if (max_val < threshold) continue;
        float a = -tr[2 * max_n + 1], b = tr[2 * max_n];
        int x0 = j, y0 = i, dx0, dy0, xflag;
        if (fabsf(a) > fabsf(b)) {
            xflag = 1;
            dx0 = a > 0 ? 1 : -1;
            dy0 = round_half_even_f(b * 65536.0f / fabsf(a));
            y0 = (y0 << 16) + 32768;
        } else {
            xflag = 0;
            dy0 = b > 0 ? 1 : -1;
            dx0 = round_half_even_f(a * 65536.0f / fabsf(b));
            x0 = (x0 << 16) + 32768;
        }
        int ex[2] = {0, 0}, ey[2] = {0, 0};
        for (int k = 0; k < 2; k++) {
            int gap = 0, x = x0, y = y0, dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
            for (;; x += dx, y += dy) {
                int j1 = xflag ? x : x >> 16, i1 = xflag ? y >> 16 : y;
                if (j1 < 0 || j1 >= W || i1 < 0 || i1 >= H) break;
                if (mask[(size_t)i1 * W + j1]) { gap = 0; ex[k] = j1; ey[k] = i1; }
                else if (++gap > line_gap) break;
            }
        }
        int good = abs(ex[1] - ex[0]) >= line_length || abs(ey[1] - ey[0]) >= line_length;
        for (int k = 0; k < 2; k++) {
            int x = x0, y = y0, dx = k ? -dx0 : dx0, dy = k ? -dy0 : dy0;
            for (;; x += dx, y += dy) {
                int j1 = xflag ? x : x >> 16, i1 = xflag ? y >> 16 : y;
                uint8_t *m = &mask[(size_t)i1 * W + j1];
                if (*m) {
                    if (good)
                        for (int n = 0; n < numangle; n++)
                            accum[(size_t)n * numrho + rho_index(j1, i1, tr, n, dec_fma) + half]--;
                    *m = 0;
                }
                if (i1 == ey[k] && j1 == ex[k]) break;
            }
        }
        if (good) {
            if (nlines < max_lines) {
                out[4 * nlines] = ex[0]; out[4 * nlines + 1] = ey[0];
                out[4 * nlines + 2] = ex[1]; out[4 * nlines + 3] = ey[1];
                nlines++;
            }
            found++;
        }
    }
    free(accum); free(mask); free(nz);
    if (total_found) *total_found = found;
    return nlines;
}
