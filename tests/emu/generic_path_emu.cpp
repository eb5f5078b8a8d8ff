// The generic per-frame path of the product -- noise sample, EMA / threshold recurrence, fused_frame_kernel
// (stack -> diff -> median -> threshold -> close -> dynamic mask, csrc/kernels_basic.cuh) and the PPHT kernels
// (csrc/hough.cuh) -- run frame by frame on the CPU by the thread-block emulator, driven like mdb_update / mdb_detect drive
// them (csrc/metdet.cu: launch_noise_thr, launch_fused's generic branch, launch_hough_kernels).  Built as a shared
// library; tests/test_generic_emu_cpu.py feeds it the golden trajectories of the live reference.  Test infrastructure.
#include "cuda_block_emu.h"

#include <cstdio>
#include <cstdlib>

uint32_t h_sm[96 * 1024];
uint16_t o_sm[8192];
#include "kernels_basic_emu.cuh"
#include "hough_emu.cuh"

extern "C" int emu_generic_path(const uint8_t *frames, int T, int W, int H, int n, const uint8_t *mask, int apply_mask,
                                int adaptive, int init_value, int sensitivity, int nz_interval, const int *roi,
                                int hough_thr, int hough_min_len, int hough_max_gap, int dy_on, double mask_area,
                                int *thr_out, double *thrf_out, double *snr_out, uint8_t *dst_out, int *n_on_out,
                                int *lines_num_out, int32_t *raw_out /*[T][512][4]*/) {
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int k = 0; k < MDB_HOUGH_ANGLES; k++) {
        c_trig[2 * k] = (float)cos((double)k * (double)theta);
        c_trig[2 * k + 1] = (float)sin((double)k * (double)theta);
    }
    const size_t HW = (size_t)W * H;
    const int R = n, Wb = (W + 31) / 32, RA = n;  // max_batch = 1: R = n - 1 + 1, RA = n - 1 + 1
    std::vector<uint8_t> ringbuf((size_t)R * HW, 0), dst(HW, 0);
    std::vector<uint32_t> actbuf((size_t)RA * H * Wb, 0), points(MDB_POINT_CAP), okeys(HW), oidx(HW), bitmap((HW + 31) / 32, 0);
    std::vector<uint16_t> order(HOUGH_ORDER_CAP);
    std::vector<int32_t> lines(512 * 4), accum((size_t)MDB_HOUGH_ANGLES * (2 * (W + H) + 1), 0);
    std::vector<uint32_t> walk(W + H + 2, 0);
    FrameSrc src; src.ring = ringbuf.data(); src.cur = nullptr; src.mask = apply_mask ? mask : nullptr; src.t0 = 0; src.R = R; src.HW = HW;
    ActRing ring; ring.base = actbuf.data(); ring.RA = RA; ring.Wb = Wb; ring.frame_words = (size_t)H * Wb;
    DevState st;  // initial_state() of csrc/metdet.cu
    memset(&st, 0, sizeof st);
    st.ema_init_m = 1.0 - (double)nz_interval / 60.0;
    st.ema_cur_m = st.ema_init_m;
    st.ema_warm = (double)n;
    static const int abs_sens[3] = {7, 5, 3};
    st.bi_threshold = adaptive ? abs_sens[sensitivity] : init_value;
    st.thr_float = (double)st.bi_threshold;
    HoughParams P;
    P.W = W; P.H = H; P.numrho = 2 * (W + H) + 1; P.threshold = hough_thr; P.min_len = hough_min_len; P.max_gap = hough_max_gap;
    P.mask_area = mask_area; P.cap = MDB_POINT_CAP; P.max_lines = 512; P.walk_cap = W + H + 2; P.fixed_gap = -1;
    const int rh = roi[2] - roi[0], rw = roi[3] - roi[1];
    const long long std_interval = (long long)nz_interval * n;
    for (int t = 0; t < T; t++) {
        memcpy(&ringbuf[(size_t)(t % R) * HW], frames + (size_t)t * HW, HW);  // copy_to_ring
        // ---- launch_noise_thr --------------------------------------------------------------------------------------
        unsigned long long noise[2] = {0, 0};
        const long long tau = t + 1;
        SampleList sl; sl.count = 0;
        if ((tau > 1 && tau <= n) || (tau > n && std_interval > 0 && tau % std_interval == 0)) sl.idx[sl.count++] = 0;
        if (sl.count) {
            const int gx = std::max(1, std::min((rh * rw + 255) / 256, 8));
            emu_launch2(gx, 1, 256, [&] { noise_sample_kernel(src, W, n, (long long)t, std_interval, roi[0], roi[1], rh, rw, noise, 0, sl); });
        }
        int thr = 0; double thrf = 0, snr = 0;
        emu_launch(1, 32, [&] { threshold_kernel(&st, noise, 1, (long long)t, n, std_interval, (long long)rh * rw, adaptive, sensitivity, &thr, &thrf, &snr); });
        thr_out[t] = thr; thrf_out[t] = thrf; snr_out[t] = snr;
        // ---- launch_fused, generic branch ----------------------------------------------------------------------------
        unsigned npoints = 0;
        const int L = (int)std::min<long long>(n, t + 1), Ldy = L;
        emu_launch2((W + V1_TW - 1) / V1_TW, (H + V1_TH - 1) / V1_TH, 256, [&] {
            fused_frame_kernel(src, W, H, n, (long long)t, L, (long long)t, Ldy, dy_on, &thr, ring, dst.data(), &npoints, points.data(), MDB_POINT_CAP);
        });
        memcpy(dst_out + (size_t)t * HW, dst.data(), HW);
        n_on_out[t] = (int)npoints;
        // ---- launch_hough_kernels ---------------------------------------------------------------------------------------
        unsigned queue[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int nlines = -99;
        if (npoints == 0) nlines = 0;  // (what tier 1a writes for an empty mask)
        else {
            emu_launch(1, 32, [&] { ppht_order_kernel(1, HOUGH_ORDER_CAP, &npoints, order.data()); });
            emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, 1, &npoints, points.data(), order.data(), lines.data(), &nlines, queue, nullptr, HOUGH_CAP_SMALL, HOUGH_TABLE_BYTES_SMALL, 0); });
        }
        // the product launches every tier and lets the device decide; here a tier whose flag is not set is skipped (its
        // kernel would exit at once) to spare the emulator 256 thread start-ups per launch
        if (nlines == -2) emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, 1, &npoints, points.data(), order.data(), lines.data(), &nlines, queue + 1, nullptr, HOUGH_CAP_LARGE, HOUGH_TABLE_BYTES, 1); });
        if (nlines == -3) emu_launch(1, HOUGH_THREADS, [&] { hough_tier2_kernel(P, 1, &npoints, points.data(), accum.data(), lines.data(), &nlines, nullptr, queue + 7); });
        if (nlines == -1) emu_launch(1, HOUGH_THREADS, [&] { hough_tier3_kernel(P, 1, dst.data(), okeys.data(), oidx.data(), accum.data(), bitmap.data(), walk.data(), lines.data(), &nlines, queue + 2, nullptr); });
        if (nlines < 0) return -(t + 1);
        lines_num_out[t] = nlines;
        memcpy(raw_out + (size_t)t * 512 * 4, lines.data(), (size_t)std::min(nlines, 512) * 16);
    }
    return 0;
}

// Clip stackers of kernels_basic.cuh (mdb_max_stack / mdb_gauss_stack in csrc/metdet.cu feed them chunk by chunk):
// MaxImgContainer (MetLib/stacker.py:43-49) and FastGaussianContainer (:52-59, uint16 / uint32 wrap-around)
extern "C" int emu_max_stack(const uint8_t *frames, int T, size_t frame_bytes, int chunk, unsigned grid, uint8_t *out) {
    for (int t0 = 0; t0 < T; t0 += chunk) {
        const int c = std::min(chunk, T - t0);
        const uint8_t *src = frames + (size_t)t0 * frame_bytes;
        emu_launch(grid, 256, [&] { max_stack_kernel(src, c, frame_bytes, out, t0 > 0); });
    }
    return 0;
}
extern "C" int emu_gauss_stack(const uint8_t *frames, int T, size_t frame_bytes, int chunk, unsigned grid, int accumulate, uint16_t *sum,
                               uint32_t *sq) {
    for (int t0 = 0; t0 < T; t0 += chunk) {
        const int c = std::min(chunk, T - t0);
        const uint8_t *src = frames + (size_t)t0 * frame_bytes;
        emu_launch(grid, 256, [&] { gauss_stack_kernel(src, c, frame_bytes, sum, sq, accumulate || t0 > 0); });
    }
    return 0;
}

// API read-backs (mdb_get_stack / mdb_get_window / mdb_get_std in csrc/metdet.cu): SlidingWindow.max / .mean / .sum / .std and
// the window's frames (MetLib/utils.py:269-321).  frames[0 .. t] are all frames fed so far, written into a ring of R slots
// as the library keeps them; the kernels read the ring only.
extern "C" int emu_stack_readback(const uint8_t *frames, long long t, int W, int H, int n, int R, const uint8_t *mask, unsigned grid,
                                  uint8_t *mx, uint8_t *mean, uint32_t *sum, unsigned long long *std_total, uint8_t *newest_masked) {
    const size_t HW = (size_t)W * H;
    if (R < n) return -1;
    std::vector<uint8_t> ring((size_t)R * HW, 0);
    for (long long k = std::max<long long>(0, t - R + 1); k <= t; k++) memcpy(&ring[(size_t)(k % R) * HW], frames + (size_t)k * HW, HW);
    FrameSrc src; src.ring = ring.data(); src.cur = nullptr; src.mask = mask; src.t0 = 0; src.R = R; src.HW = HW;
    const int L = (int)std::min<long long>(n, t + 1);
    emu_launch(grid, 256, [&] { stack_readback_kernel(src, HW, n, t, L, mx, mean, sum); });
    *std_total = 0;
    emu_launch(grid, 256, [&] { stack_std_kernel(src, HW, n, t, L, std_total); });
    if (mask && newest_masked) emu_launch(grid, 256, [&] { window_frame_kernel(src.frame(t), mask, HW, newest_masked); });
    return 0;
}
