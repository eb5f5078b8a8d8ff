// ClassicDetector's device path on the CPU: batch-wise noise samples and threshold recurrence over the
// 4-frame window, classic_bits_kernel / classic_spatial_kernel / classic_expand_kernel (csrc/classic.cuh) and the PPHT
// kernels with the configured maxLineGap, under the thread-block emulator, issued like launch_fused's ClassicDetector
// branch issues them (csrc/metdet.cu).  Also the loader's preprocessing kernel (csrc/preproc.cuh).  Built as a shared
// library; tests/test_classic_emu_cpu.py feeds it golden vectors of the live reference.  Test infrastructure.
#include "cuda_block_emu.h"

#include <cstdio>
#include <cstdlib>

uint32_t h_sm[96 * 1024];
uint16_t o_sm[8192];
#include "kernels_basic_emu.cuh"
#include "classic_emu.cuh"
#include "preproc_emu.cuh"
#include "hough_emu.cuh"

extern "C" int emu_classic_path(const uint8_t *frames, int Ttot, int W, int H, int B, int adaptive, int init_value, int sensitivity,
                                int nz_interval, const int *roi, int hough_thr, int hough_min_len, int hough_max_gap, double mask_area,
                                int *thr_out, double *snr_out, uint8_t *dst_out, int *lines_num_out, int32_t *raw_out, int raw_cap) {
    const int n = 4;  // ClassicDetector.classic_max_size
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int k = 0; k < MDB_HOUGH_ANGLES; k++) {
        c_trig[2 * k] = (float)cos((double)k * (double)theta);
        c_trig[2 * k + 1] = (float)sin((double)k * (double)theta);
    }
    const size_t HW = (size_t)W * H;
    const int Wb = (W + 31) / 32, R = n - 1 + 2 * B;
    const size_t plane = (size_t)B * H * Wb;
    std::vector<uint8_t> ringbuf((size_t)R * HW, 0), dst((size_t)B * HW, 0);
    std::vector<uint32_t> cbits(3 * plane, 0), points((size_t)B * MDB_POINT_CAP), okeys(HW), oidx(HW), bitmap((HW + 31) / 32, 0), walk(W + H + 2, 0);
    std::vector<unsigned> npoints(B, 0);
    std::vector<uint16_t> order((size_t)B * HOUGH_ORDER_CAP);
    std::vector<int32_t> lines((size_t)B * raw_cap * 4), accum((size_t)MDB_HOUGH_ANGLES * (2 * (W + H) + 1), 0);
    std::vector<int> nlines(B), thr(B);
    std::vector<double> thrf(B), snr(B);
    std::vector<unsigned long long> noise((size_t)B * 2);
    DevState st;
    memset(&st, 0, sizeof st);
    st.ema_init_m = 1.0 - (double)nz_interval / 60.0;
    st.ema_cur_m = st.ema_init_m;
    st.ema_warm = (double)n;
    static const int abs_sens[3] = {7, 5, 3};
    st.bi_threshold = adaptive ? abs_sens[sensitivity] : init_value;
    st.thr_float = (double)st.bi_threshold;
    HoughParams P;
    P.W = W; P.H = H; P.numrho = 2 * (W + H) + 1; P.threshold = hough_thr; P.min_len = hough_min_len; P.max_gap = hough_max_gap;
    P.mask_area = mask_area; P.cap = MDB_POINT_CAP; P.max_lines = raw_cap; P.walk_cap = W + H + 2; P.fixed_gap = hough_max_gap;
    const int rh = roi[2] - roi[0], rw = roi[3] - roi[1];
    const long long std_interval = (long long)nz_interval * n;
    uint32_t *ab = cbits.data(), *bb = cbits.data() + plane, *db = cbits.data() + 2 * plane;
    for (long long t0 = 0; t0 < Ttot; t0 += B) {
        const int T = (int)std::min<long long>(B, Ttot - t0);
        FrameSrc src; src.ring = ringbuf.data(); src.cur = frames + (size_t)t0 * HW; src.mask = nullptr; src.t0 = t0; src.R = R; src.HW = HW;
        std::fill(noise.begin(), noise.end(), 0ull);
        SampleList sml; sml.count = 0;
        for (int i = 0; i < T && sml.count >= 0; i++) {
            const long long tau = t0 + i + 1;
            if ((tau > 1 && tau <= n) || (tau > n && std_interval > 0 && tau % std_interval == 0)) {
                if (sml.count < 63) sml.idx[sml.count++] = i; else sml.count = -1;
            }
        }
        if (sml.count != 0) {
            const int rows = sml.count < 0 ? T : sml.count, gx = std::max(1, std::min((rh * rw + 255) / 256, 8));
            emu_launch2(gx, rows, 256, [&] { noise_sample_kernel(src, W, n, t0, std_interval, roi[0], roi[1], rh, rw, noise.data(), 0, sml); });
        }
        emu_launch(1, 32, [&] { threshold_kernel(&st, noise.data(), T, t0, n, std_interval, (long long)rh * rw, adaptive, sensitivity, thr.data(), thrf.data(), snr.data()); });
        std::fill(npoints.begin(), npoints.end(), 0u);
        const int nthreads = H * Wb;
        if (W % 16 == 0) emu_launch((nthreads + 255) / 256, 256, [&] { classic_bits_kernel<true>(src, W, H, Wb, t0, T, thr.data(), ab, bb); });
        else emu_launch((nthreads + 255) / 256, 256, [&] { classic_bits_kernel<false>(src, W, H, Wb, t0, T, thr.data(), ab, bb); });
        const int rows = 64, strips = (Wb + SP_USE - 1) / SP_USE, bands = (H + rows - 1) / rows;
        emu_launch2((strips * bands + SP_WARPS - 1) / SP_WARPS, T, SP_WARPS * 32, [&] { classic_spatial_kernel(ab, bb, H, Wb, rows, strips, bands, db); });
        if (W % 16 == 0) emu_launch2((nthreads + 255) / 256, T, 256, [&] { classic_expand_kernel<true>(db, W, H, Wb, dst.data(), npoints.data(), points.data(), MDB_POINT_CAP); });
        else emu_launch2((nthreads + 255) / 256, T, 256, [&] { classic_expand_kernel<false>(db, W, H, Wb, dst.data(), npoints.data(), points.data(), MDB_POINT_CAP); });
        for (long long t = t0 + T - std::min(T, n); t < t0 + T; t++) memcpy(&ringbuf[(size_t)(t % R) * HW], frames + (size_t)t * HW, HW);
        unsigned queue[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        emu_launch(T, 32, [&] { ppht_order_kernel(T, HOUGH_ORDER_CAP, npoints.data(), order.data()); });
        emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, T, npoints.data(), points.data(), order.data(), lines.data(), nlines.data(), queue, nullptr, HOUGH_CAP_SMALL, HOUGH_TABLE_BYTES_SMALL, 0); });
        bool f1 = false, f2 = false, f3 = false;
        for (int i = 0; i < T; i++) f1 |= nlines[i] == -2;
        if (f1) emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, T, npoints.data(), points.data(), order.data(), lines.data(), nlines.data(), queue + 1, nullptr, HOUGH_CAP_LARGE, HOUGH_TABLE_BYTES, 1); });
        for (int i = 0; i < T; i++) f2 |= nlines[i] == -3;
        if (f2) emu_launch(1, HOUGH_THREADS, [&] { hough_tier2_kernel(P, T, npoints.data(), points.data(), accum.data(), lines.data(), nlines.data(), nullptr, queue + 7); });
        for (int i = 0; i < T; i++) f3 |= nlines[i] == -1;
        if (f3) emu_launch(1, HOUGH_THREADS, [&] { hough_tier3_kernel(P, T, dst.data(), okeys.data(), oidx.data(), accum.data(), bitmap.data(), walk.data(), lines.data(), nlines.data(), queue + 2, nullptr); });
        for (int i = 0; i < T; i++) {
            if (nlines[i] < 0) return -(int)(t0 + i + 1);
            thr_out[t0 + i] = thr[i]; snr_out[t0 + i] = snr[i];
            lines_num_out[t0 + i] = nlines[i];
            memcpy(dst_out + (size_t)(t0 + i) * HW, dst.data() + (size_t)i * HW, HW);
            memcpy(raw_out + (size_t)(t0 + i) * raw_cap * 4, lines.data() + (size_t)i * raw_cap * 4, (size_t)std::min(nlines[i], raw_cap) * 16);
        }
    }
    return 0;
}

// the loader's Transform chain: xt / yt = tap tables as mdb_preproc_axis_taps builds them (s0, s1, w0, w1 per coordinate)
extern "C" int emu_preproc(const uint8_t *frames, int T, int src_w, int src_h, int channels, int rgb, int dst_w, int dst_h, int resize,
                           int exp_frame, const int *xt, const int *yt, const uint8_t *mask, uint8_t *out) {
    PreParams P;
    P.src_w = src_w; P.src_h = src_h; P.channels = channels; P.dst_w = dst_w; P.dst_h = dst_h; P.resize = resize; P.rgb = rgb;
    P.exp_frame = exp_frame; P.xt = reinterpret_cast<const PreTap *>(xt); P.yt = reinterpret_cast<const PreTap *>(yt); P.mask = mask;
    const int G = (T + exp_frame - 1) / exp_frame;
    if (channels == 3) emu_launch3((dst_w + 255) / 256, dst_h, G, 256, [&] { preproc_kernel<3>(P, frames, T, out); });
    else emu_launch3((dst_w + 255) / 256, dst_h, G, 256, [&] { preproc_kernel<1>(P, frames, T, out); });
    return 0;
}
