// The spatial half of the mask chain (csrc/spatial_kernel.cuh: act4_kernel / act_kernel, dst_sparse_kernel,
// dst_dense_kernel, the overflow hand-over between them, the persistent mask buffer with its shadow bits and word lists) run on
// the CPU by the thread-block emulator against a per-pixel statement of MetLib/Detector.py:329-335 (median, threshold,
// close) and :234-242 (dynamic mask: not on in ALL of the last L act frames, eroded; dst = act * m).  Batches reuse the
// same buffers, as the product's batch contexts do.  Launch geometry as in stream_kernel_launch (csrc/stream_kernel.cuh).
// Test infrastructure: built and run by tests/test_spatial_emu_cpu.py.
#include "cuda_block_emu.h"

#include <cstdio>
#include <cstdlib>
#include <set>

#include "spatial_kernel_emu.cuh"

static unsigned rng_state = 31337u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static int run_case(int W, int H, int n, int T, int batches, int dy_on, int mode /*0 sparse, 1 forced dense*/, int rows_act) {
    const int Wb = W / 32, RA = n - 1 + 2 * T;
    const size_t HW = (size_t)W * H, FW = (size_t)H * Wb;
    std::vector<uint32_t> ringbuf((size_t)RA * FW, 0), dstbits((size_t)T * FW, 0), alist((size_t)T * SPX_ACAP), wlist((size_t)T * SPX_WCAP),
        points((size_t)T * 4096), bits((size_t)T * FW);
    std::vector<unsigned> acount(T, 0), wcount(T, 0), dense(T + 1, 0), npoints(T, 0);
    std::vector<uint8_t> dst((size_t)T * HW, 0);
    ActRing ring; ring.base = ringbuf.data(); ring.RA = RA; ring.Wb = Wb; ring.frame_words = FW;
    SparseLists sl; sl.alist = alist.data(); sl.acount = acount.data(); sl.wlist = wlist.data(); sl.wcount = wcount.data(); sl.dense = dense.data();
    std::vector<std::vector<uint8_t>> act_hist;  // reference act frames by dy index
    // persistent blobs (they stay on for a while: dynamic-mask food) + flicker
    struct Blob { int x, y, w, h, t0, t1; };
    std::vector<Blob> blobs;
    for (int k = 0; k < 7; k++) blobs.push_back({(int)(rnd() % (W - 12)), (int)(rnd() % std::max(1, H - 6)), 4 + (int)(rnd() % 8), 3 + (int)(rnd() % 3),
                                                 (int)(rnd() % (T * batches)), 0});
    for (auto &b : blobs) b.t1 = b.t0 + 1 + (int)(rnd() % (2 * n));
    int bad = 0;
    long long dy0 = 0;
    for (int bt = 0; bt < batches && bad < 5; bt++, dy0 += T) {
        std::vector<std::vector<uint8_t>> pred(T, std::vector<uint8_t>(HW, 0));
        for (int i = 0; i < T; i++) {
            const int t = bt * T + i;
            for (const auto &b : blobs)
                if (t >= b.t0 && t < b.t1)
                    for (int y = b.y; y < std::min(H, b.y + b.h); y++)
                        for (int x = b.x; x < std::min(W, b.x + b.w); x++) pred[i][(size_t)y * W + x] = 1;
            for (int k = 0; k < (int)(HW / 40); k++) pred[i][rnd() % HW] ^= 1;
            if (t % 5 == 0) for (int x = 3; x < std::min(W, 70); x++) pred[i][(size_t)(H / 2) * W + x] = pred[i][(size_t)(H / 2 + 1 < H ? H / 2 + 1 : H / 2) * W + x] = 1;
        }
        std::fill(bits.begin(), bits.end(), 0u);
        for (int i = 0; i < T; i++)
            for (size_t p = 0; p < HW; p++)
                if (pred[i][p]) bits[(size_t)i * FW + (p / W) * Wb + (p % W) / 32] |= 1u << ((p % W) % 32);
        // ---- the product's launch sequence (stream_kernel_launch) ------------------------------------------------------
        std::fill(npoints.begin(), npoints.end(), 0u);
        std::fill(acount.begin(), acount.end(), 0u);
        const int strips = (Wb + SP_USE - 1) / SP_USE;
        if (Wb % 4 == 0) {
            const int chunks = Wb / 4, bands = (H + rows_act - 1) / rows_act;
            emu_launch2((chunks * bands + A4_THREADS - 1) / A4_THREADS, T, A4_THREADS,
                        [&] { act4_kernel(bits.data(), H, Wb, rows_act, chunks, bands, ring, dy0, sl); });
        } else {
            const int bands = (H + rows_act - 1) / rows_act;
            emu_launch2((strips * bands + SP_WARPS - 1) / SP_WARPS, T, SP_WARPS * 32,
                        [&] { act_kernel(bits.data(), W, H, T, rows_act, strips, bands, ring, dy0, sl); });
        }
        dense[0] = 0;
        if (mode == 1) emu_launch((T + 127) / 128, 128, [&] { dst_force_dense_kernel(T, sl); });
        else emu_launch(T, 256, [&] { dst_sparse_kernel(ring, W, H, n, dy0, dy_on, dst.data(), dstbits.data(), npoints.data(), points.data(), 4096, sl); });
        const int dst_rows = 32, dbands = (H + dst_rows - 1) / dst_rows;
        emu_launch2((strips * dbands + SP_WARPS - 1) / SP_WARPS, std::min(T, DENSE_GY), SP_WARPS * 32, [&] {
            dst_dense_kernel(ring, W, H, n, dy0, dy_on, dst_rows, strips, dbands, dst.data(), dstbits.data(), npoints.data(), points.data(), 4096, sl);
        });
        // ---- reference, per pixel -----------------------------------------------------------------------------------------
        for (int i = 0; i < T && bad < 5; i++) {
            const long long d = dy0 + i;
            auto clampi = [](int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); };
            std::vector<uint8_t> bin(HW), dil(HW), act(HW), m(HW), out(HW);
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int c = 0;
                    for (int dy = -1; dy <= 1; dy++)
                        for (int dx = -1; dx <= 1; dx++) c += pred[i][(size_t)clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)];
                    bin[(size_t)y * W + x] = c >= 5;
                }
            auto at = [&](const std::vector<uint8_t> &im, int y, int x, int outside) { return (y < 0 || y >= H || x < 0 || x >= W) ? outside : (int)im[(size_t)y * W + x]; };
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int v = 0;
                    for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) v |= at(bin, y + dy, x + dx, 0);
                    dil[(size_t)y * W + x] = (uint8_t)v;
                }
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int v = 1;
                    for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) v &= at(dil, y + dy, x + dx, 1);
                    act[(size_t)y * W + x] = (uint8_t)v;
                }
            act_hist.push_back(act);
            const int L = (int)std::min<long long>(n, d + 1);
            for (size_t p = 0; p < HW; p++) {
                int all = 1;
                for (int k = 0; k < L; k++) all &= act_hist[(size_t)(d - k)][p];
                m[p] = !all;
            }
            unsigned cnt = 0;
            std::set<unsigned> want_pts;
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int e = 1;
                    if (dy_on) for (int dy = -1; dy <= 1; dy++) for (int dx = -1; dx <= 1; dx++) e &= at(m, y + dy, x + dx, 1);
                    const int v = act[(size_t)y * W + x] & e;
                    out[(size_t)y * W + x] = v ? 255 : 0;
                    if (v) { cnt++; want_pts.insert(((unsigned)y << 16) | (unsigned)x); }
                }
            const uint8_t *got = dst.data() + (size_t)i * HW;
            for (size_t p = 0; p < HW && bad < 5; p++)
                if (got[p] != out[p]) { fprintf(stderr, "W=%d n=%d dy=%d mode=%d batch %d frame %d pixel (%zu,%zu): want %d got %d\n", W, n, dy_on, mode, bt, i, p % W, p / W, out[p], got[p]); bad++; }
            if (npoints[i] != cnt) { fprintf(stderr, "frame %d: on-pixel count %u, want %u\n", i, npoints[i], cnt); bad++; }
            std::set<unsigned> got_pts(points.begin() + (size_t)i * 4096, points.begin() + (size_t)i * 4096 + std::min(npoints[i], 4096u));
            if (cnt <= 4096 && got_pts != want_pts) { fprintf(stderr, "frame %d: on-pixel list differs\n", i); bad++; }
            for (size_t w = 0; w < FW && bad < 5; w++) {  // shadow bits describe the buffer
                unsigned wb = 0;
                for (int b = 0; b < 32; b++) wb |= (unsigned)(got[(w / Wb) * W + (w % Wb) * 32 + b] != 0) << b;
                if (dstbits[(size_t)i * FW + w] != wb) { fprintf(stderr, "frame %d: shadow word %zu differs\n", i, w); bad++; }
            }
        }
    }
    return bad;
}

int main() {
    int bad = 0, cases = 0;
    for (int W : {128, 160})             // Wb = 4: act4_kernel; Wb = 5: the warp-strip act_kernel
        for (int dy_on = 0; dy_on < 2; dy_on++)
            for (int mode = 0; mode < 2; mode++) {
                bad += run_case(W, 37, 4, 5, 3, dy_on, mode, mode ? 64 : 8);
                cases++;
            }
    bad += run_case(256, 70, 6, 4, 2, 1, 0, 64); cases++;
    printf("%d cases: %s\n", cases, bad ? "FAILED" : "ALL OK");
    return bad ? 1 : 0;
}
