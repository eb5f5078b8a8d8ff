// The exact-order PPHT kernels (csrc/hough.cuh: ppht_order_kernel, hough_smem_kernel tiers 1a / 1b, hough_tier2_kernel)
// run on the CPU by the block emulator (cuda_block_emu.h: one fiber per CUDA thread, barriers and warp collectives
// emulated) against oracle/ppht.c -- the restatement of cv2.HoughLinesP that is pinned on cv2 itself.  The kernel source
// is the product's, with two mechanical edits made by the test's build step (tests/test_hough_emu_cpu.py):
// `extern __shared__` -> `extern` and the PTX prefetch hints removed.  Test infrastructure.
#include "cuda_block_emu.h"

#include <cstdio>
#include <cstdlib>

uint32_t h_sm[96 * 1024];   // dynamic shared memory of the PPHT kernels (384 KB: more than any tier asks for)
uint16_t o_sm[8192];        // ... of ppht_order_kernel
#include "hough_emu.cuh"

extern "C" int oracle_ppht(const uint8_t *img, int W, int H, int threshold, int line_length, int line_gap, int max_lines,
                           int32_t *out, int vote_fma, int dec_fma, int *total_found);

static unsigned rng_state = 99u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static void draw_line(std::vector<uint8_t> &img, int W, int H, int x0, int y0, int x1, int y1, int thick) {
    const int steps = std::max(abs(x1 - x0), abs(y1 - y0)) + 1;
    for (int s = 0; s < steps; s++) {
        const int x = x0 + (int)lrint((double)(x1 - x0) * s / std::max(steps - 1, 1));
        const int y = y0 + (int)lrint((double)(y1 - y0) * s / std::max(steps - 1, 1));
        for (int dy = 0; dy < thick; dy++)
            for (int dx = 0; dx < thick; dx++)
                if (x + dx >= 0 && x + dx < W && y + dy >= 0 && y + dy < H) img[(size_t)(y + dy) * W + x + dx] = 255;
    }
}

int main() {
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int k = 0; k < MDB_HOUGH_ANGLES; k++) {  // as metdet.cu fills the constant table
        c_trig[2 * k] = (float)cos((double)k * (double)theta);
        c_trig[2 * k + 1] = (float)sin((double)k * (double)theta);
    }
    struct Case { int W, H, nlines, thick, noise, thr, minlen, gap, far; };
    const Case cases[] = {
        {160, 120, 1, 1, 0, 10, 10, 3, 0},   {160, 120, 2, 2, 20, 10, 10, 5, 0}, {320, 200, 3, 3, 60, 10, 10, 10, 0},
        {97, 61, 2, 1, 10, 6, 6, 2, 0},      {640, 360, 1, 2, 0, 10, 10, 10, 0}, {640, 360, 2, 2, 30, 10, 10, 0, 1},
        {1920, 1080, 2, 3, 0, 10, 10, 10, 1}, {256, 160, 6, 2, 200, 8, 8, 4, 0},  {128, 96, 0, 1, 150, 5, 5, 1, 0},
        {3840, 2160, 1, 3, 0, 10, 10, 10, 0}, {320, 200, 2, 2, 4600, 12, 10, 2, 0},  // > 4096 points: tier 2
        {192, 128, 3, 3, 40000, 40, 12, 1, 0},                                         // > 16384 points: tier 3 (dense)
    };
    const int T = (int)(sizeof cases / sizeof cases[0]);
    int bad = 0;
    for (int ci = 0; ci < T; ci++) {
        const Case &c = cases[ci];
        const int W = c.W, H = c.H;
        std::vector<uint8_t> img((size_t)W * H, 0);
        for (int l = 0; l < c.nlines; l++) {
            int x0 = rnd() % W, y0 = rnd() % H;
            int len = 20 + rnd() % std::min(W, H) / 2;
            if (c.far) { x0 = l ? W - 200 : 30; y0 = l ? H - 150 : 20; len = 100; }  // two far-apart objects: two rho intervals
            const double a = (rnd() % 360) * 3.14159265 / 180.0;
            draw_line(img, W, H, x0, y0, x0 + (int)(len * cos(a)), y0 + (int)(len * sin(a)), c.thick);
        }
        for (int k = 0; k < c.noise; k++) img[(size_t)(rnd() % H) * W + rnd() % W] = 255;
        // on-pixel list in scrambled order, as the dst kernels emit it
        std::vector<uint32_t> pts;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++)
                if (img[(size_t)y * W + x]) pts.push_back(((uint32_t)y << 16) | (uint32_t)x);
        for (size_t i = pts.size(); i > 1; i--) std::swap(pts[i - 1], pts[rnd() % i]);
        const unsigned N = (unsigned)pts.size();
        HoughParams P;
        P.W = W; P.H = H; P.numrho = 2 * (W + H) + 1;
        P.threshold = c.thr; P.min_len = c.minlen; P.max_gap = c.gap; P.mask_area = (double)W * H;
        P.cap = MDB_POINT_CAP; P.max_lines = 512; P.walk_cap = W + H + 2; P.fixed_gap = c.gap;
        std::vector<uint32_t> points(MDB_POINT_CAP, 0);
        std::copy(pts.begin(), pts.begin() + std::min<size_t>(pts.size(), MDB_POINT_CAP), points.begin());
        std::vector<uint16_t> order(HOUGH_ORDER_CAP, 0);
        std::vector<int32_t> lines(512 * 4, -1);
        std::vector<int32_t> accum((size_t)MDB_HOUGH_ANGLES * P.numrho, 0);
        int nlines = -99;
        unsigned queue[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned npoints = N;
        emu_launch(1, 32, [&] { ppht_order_kernel(1, HOUGH_ORDER_CAP, &npoints, order.data()); });
        emu_launch(1, HOUGH_THREADS, [&] {
            hough_smem_kernel(P, 1, &npoints, points.data(), order.data(), lines.data(), &nlines, queue, nullptr, HOUGH_CAP_SMALL,
                              HOUGH_TABLE_BYTES_SMALL, 0);
        });
        const char *tier = "1a";
        if (nlines == -2) {
            tier = "1b";
            emu_launch(1, HOUGH_THREADS, [&] {
                hough_smem_kernel(P, 1, &npoints, points.data(), order.data(), lines.data(), &nlines, queue + 1, nullptr,
                                  HOUGH_CAP_LARGE, HOUGH_TABLE_BYTES, 1);
            });
        }
        if (nlines == -3) {
            tier = "2";
            emu_launch(1, HOUGH_THREADS, [&] {
                hough_tier2_kernel(P, 1, &npoints, points.data(), accum.data(), lines.data(), &nlines, nullptr, queue + 7);
            });
            for (int32_t v : accum)
                if (v != 0) { fprintf(stderr, "case %d: tier 2 left a dirty accumulator\n", ci); bad++; break; }
        }
        if (nlines == -1) {  // dense: ordered compaction of the mask on the device side, global point list
            tier = "3";
            const size_t HWs = (size_t)W * H;
            std::vector<uint32_t> okeys(HWs), oidx(HWs), bitmap((HWs + 31) / 32, 0), walk(P.walk_cap, 0);
            emu_launch(1, HOUGH_THREADS, [&] {
                hough_tier3_kernel(P, 1, img.data(), okeys.data(), oidx.data(), accum.data(), bitmap.data(), walk.data(), lines.data(),
                                   &nlines, queue + 2, nullptr);
            });
            for (int32_t v : accum)
                if (v != 0) { fprintf(stderr, "case %d: tier 3 left a dirty accumulator\n", ci); bad++; break; }
            for (uint32_t v : bitmap)
                if (v != 0) { fprintf(stderr, "case %d: tier 3 left a dirty bitmap\n", ci); bad++; break; }
        }
        std::vector<int32_t> ref(4 * std::max(1u, N));
        int total = 0;
        const int nref = oracle_ppht(img.data(), W, H, c.thr, c.minlen, c.gap, (int)std::max(1u, N), ref.data(), 0, 0, &total);
        bool ok = nlines == nref && nref == total;
        for (int k = 0; ok && k < 4 * nref && k < 4 * 512; k++) ok = lines[k] == ref[k];
        printf("case %d: %dx%d, %u points, tier %s: %d segments, oracle %d: %s\n", ci, W, H, N, tier, nlines, nref, ok ? "ok" : "DIFFERENT");
        if (!ok) bad++;
    }
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
