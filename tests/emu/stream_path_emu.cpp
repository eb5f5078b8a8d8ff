// The product's batched (streaming) path on the CPU: noise samples + threshold recurrence for a whole batch, temporal3's
// per-thread code (csrc/temporal3_kernel.cuh, host build of the same source), act4 / act, dst_sparse / dst_dense
// (csrc/spatial_kernel.cuh) and the PPHT kernels (csrc/hough.cuh) under the thread-block emulator, batch after batch with
// the frame ring, act ring and mask buffers carried over -- the sequence submit_impl / stream_kernel_launch /
// launch_hough_kernels issue (csrc/metdet.cu, csrc/stream_kernel.cuh).  Built as a shared library;
// tests/test_stream_emu_cpu.py feeds it golden trajectories of the live reference.  Test infrastructure.
#include "cuda_block_emu.h"

#include <cstdio>
#include <cstdlib>

uint32_t h_sm[96 * 1024];
uint16_t o_sm[8192];
#define T3_HOST_EMU 1
#include "temporal3_kernel_emu.cuh"
#include "kernels_basic_emu.cuh"
#include "spatial_kernel_emu.cuh"
#include "hough_emu.cuh"
#include "perframe_kernel_emu.cuh"
// temporal2_kernel (csrc/temporal_kernel.cuh): the second-generation temporal pass the product launches for windows without a
// temporal3 shape (n = 11, 13, 17, 19, 22, 23, 26, 27, 29, 31 ...) and when temporal_version = 2 is forced
uint4 t_smem[16 * 1024];  // 256 KB of emulated dynamic shared memory
#define __cvta_generic_to_shared(p) ((size_t)0)
#include "temporal_kernel_emu.cuh"

static int g_temporal_version = 3;
extern "C" void emu_set_temporal_version(int v) { g_temporal_version = v; }  // the library's "temporal_version" option
static int g_t2_launches = 0;
extern "C" int emu_temporal2_launches() { return g_t2_launches; }

// stream_choose_kdiv / stream_state_config / stream_state_init (csrc/stream_kernel.cuh:65-111), restated: sub-blocks per
// window, words per thread and CTA size of temporal2 for a window of n frames
struct T2Config { int kdiv, wpt, nt; };
static bool t2_config(int n, int max_batch, T2Config &c) {
    c.kdiv = 1;
    if (n >= 48)
        for (int k = T2_KMAX; k >= 2; k--)
            if (n % k == 0 && n / k >= 15) { c.kdiv = k; break; }
    auto config = [&](int wpt) {
        const size_t sm_bytes = 228 * 1024, cta_max = 220 * 1024, reserved = 1024;
        const size_t per_thread = (size_t)(n + T2_K + n / c.kdiv) * 4 * wpt, table = ((size_t)2 * max_batch + 15) & ~(size_t)15;
        int best_nt = 0;
        size_t best_warps = 0;
        for (int nt = 128; nt >= 32; nt >>= 1) {
            const size_t cta = per_thread * nt + table;
            if (cta > cta_max) continue;
            const size_t warps = sm_bytes / (cta + reserved) * (nt / 32);
            if (warps > best_warps) { best_warps = warps; best_nt = nt; }
        }
        c.wpt = wpt; c.nt = best_nt;
        return best_nt != 0;
    };
    const bool wide = (size_t)(2 * n + T2_K) * 16 * 32 * 8 <= (size_t)220 * 1024;
    return config(wide ? 4 : 2) || config(2);
}
template <bool M, int WPT, int NT>
static void temporal2_launch_nt(const FrameSrc &src, long long t0, int T, int n, int kdiv, int HWG, const int *thr, uint8_t *bits) {
    const unsigned grid = (unsigned)((HWG + NT - 1) / NT);
    if (kdiv > 1) emu_launch(grid, NT, [&] { temporal2_kernel<M, WPT, NT, true>(src, t0, T, n, kdiv, HWG, thr, bits); });
    else emu_launch(grid, NT, [&] { temporal2_kernel<M, WPT, NT, false>(src, t0, T, n, 1, HWG, thr, bits); });
}
template <bool M, int WPT>
static void temporal2_launch_wpt(int nt, const FrameSrc &src, long long t0, int T, int n, int kdiv, int HWG, const int *thr, uint8_t *bits) {
    if (nt == 32) temporal2_launch_nt<M, WPT, 32>(src, t0, T, n, kdiv, HWG, thr, bits);
    else if (nt == 64) temporal2_launch_nt<M, WPT, 64>(src, t0, T, n, kdiv, HWG, thr, bits);
    else temporal2_launch_nt<M, WPT, 128>(src, t0, T, n, kdiv, HWG, thr, bits);
}
// the temporal2 branch of stream_kernel_launch (csrc/stream_kernel.cuh:160-176)
static int temporal2_batch(const FrameSrc &src, long long t0, int T, int n, int max_batch, size_t HW, const int *thr, uint8_t *bits) {
    T2Config c;
    if (!t2_config(n, max_batch, c)) return -1;
    const size_t smem = (size_t)(n + T2_K + n / c.kdiv) * 4 * c.wpt * c.nt + (((size_t)2 * T + 15) & ~(size_t)15);
    if (smem > sizeof t_smem) return -1;
    const int HWG = (int)(HW / (4 * c.wpt));
    g_t2_launches++;
    if (c.wpt == 2) { if (src.mask) temporal2_launch_wpt<true, 2>(c.nt, src, t0, T, n, c.kdiv, HWG, thr, bits); else temporal2_launch_wpt<false, 2>(c.nt, src, t0, T, n, c.kdiv, HWG, thr, bits); }
    else { if (src.mask) temporal2_launch_wpt<true, 4>(c.nt, src, t0, T, n, c.kdiv, HWG, thr, bits); else temporal2_launch_wpt<false, 4>(c.nt, src, t0, T, n, c.kdiv, HWG, thr, bits); }
    return 0;
}

// apply_mask = 1 of the product: frames arrive unmasked, every kernel that loads a frame applies this mask (FrameSrc::mask);
// set by emu_set_device_mask() for the following calls, nullptr = frames are already masked (or there is no mask)
static const uint8_t *g_dev_mask = nullptr;
extern "C" void emu_set_device_mask(const uint8_t *mask) { g_dev_mask = mask; }

static int g_noise16_launches = 0;
extern "C" int emu_noise16_launches() { return g_noise16_launches; }  // how often the 16-byte-load variant was the one selected
// launch_noise_samples (csrc/kernels_basic.cuh) with the launches replaced by emulated ones: same kernel selection, same grids
static void emu_noise_samples(const FrameSrc &src, int W, int n, long long timer0, long long std_interval, const int *roi,
                              unsigned long long *acc, long long min_tau, const SampleList &sl, int T) {
    const int rh = roi[2] - roi[0], rw = roi[3] - roi[1];
    const int rows = sl.count < 0 ? T : sl.count;
    const bool mask_ok = !src.mask || ((uintptr_t)src.mask & 15) == 0;
    const bool base_ok = ((uintptr_t)src.ring & 15) == 0 && ((uintptr_t)src.cur & 15) == 0 && (src.HW & 15) == 0;
    if (W % 16 == 0 && n <= 128 && mask_ok && base_ok) {
        const int groups = rh * ((((roi[1] + rw + 15) & ~15) - (roi[1] & ~15)) >> 4);
        const int gx = std::max(1, std::min((groups + 255) / 256, 1184));
        g_noise16_launches++;
        emu_launch2(gx, rows, 256, [&] { noise_sample16_kernel(src, W, n, timer0, std_interval, roi[0], roi[1], rh, rw, acc, min_tau, sl); });
    } else {
        const int gx = std::max(1, std::min((rh * rw + 255) / 256, 592));
        emu_launch2(gx, rows, 256, [&] { noise_sample_kernel(src, W, n, timer0, std_interval, roi[0], roi[1], rh, rw, acc, min_tau, sl); });
    }
}

template <int U, int BL, int P, int K>
static void temporal3_batch(const FrameSrc &src, long long t0, int T, int HWG, const int *thr, uint8_t *bits) {
    typedef t3::Layout<U, BL, P, 0, K> LY;
    std::vector<unsigned char> smem(LY::smem_bytes(T) + 64);
    const int Tp = (T + BL - 1) / BL * BL;
    for (int c = 0; c * T3_NT < HWG; c++) {  // one CTA after the other, its threads one after the other
        uint2 *tab = reinterpret_cast<uint2 *>(smem.data() + LY::tab_off);
        for (int i = 0; i < Tp; i++) tab[i] = t3::table_entry(thr[i < T ? i : T - 1], t0 + i, LY::N);
        for (int tid = 0; tid < T3_NT; tid++) {
            const int g = c * T3_NT + tid;
            if (g >= HWG) break;
            if (src.mask) t3::thread_main<U, BL, P, K, true, 0>(src, t0, T, g, tid, smem.data(), bits, (size_t)HWG, nullptr, T3_NT);
            else t3::thread_main<U, BL, P, K, false, 0>(src, t0, T, g, tid, smem.data(), bits, (size_t)HWG, nullptr, T3_NT);
        }
    }
}

// frames[0] is global frame t_first (mdb_seek); the first `halo` frames are look-back frames whose results are not wanted
// (MDB_SUBMIT_HALO: window and act history only); thr_in != nullptr: thresholds supplied by the caller (mdb_submit_batch_thr)
static int run_range(const uint8_t *frames, int Ttot, long long t_first, int halo, const int *thr_in, int W, int H, int n, int B,
                     int adaptive, int init_value, int sensitivity, int nz_interval, const int *roi, int hough_thr,
                     int hough_min_len, int hough_max_gap, int dy_on, double mask_area, int *thr_out, double *snr_out,
                     uint8_t *dst_out, int *n_on_out, int *lines_num_out, int32_t *raw_out /*[Ttot][512][4]*/) {
    if (W % 32 || n < 2 || n > 128) return -1000;  // stream_state_init: the generic kernels serve these
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int k = 0; k < MDB_HOUGH_ANGLES; k++) {
        c_trig[2 * k] = (float)cos((double)k * (double)theta);
        c_trig[2 * k + 1] = (float)sin((double)k * (double)theta);
    }
    const size_t HW = (size_t)W * H;
    const int Wb = W / 32, R = n - 1 + 2 * B, RA = R, HWG = (int)(HW / 8);
    const size_t FW = (size_t)H * Wb;
    std::vector<uint8_t> ringbuf((size_t)R * HW, 0), dst((size_t)B * HW, 0);
    std::vector<uint32_t> actbuf((size_t)RA * FW, 0), bits((size_t)(B + 32) * FW, 0), dstbits((size_t)B * FW, 0), alist((size_t)B * SPX_ACAP),
        wlist((size_t)B * SPX_WCAP), points((size_t)B * MDB_POINT_CAP), okeys(HW), oidx(HW), bitmap((HW + 31) / 32, 0), walk(W + H + 2, 0);
    std::vector<unsigned> acount(B, 0), wcount(B, 0), dense(B + 1, 0), npoints(B, 0);
    std::vector<uint16_t> order((size_t)B * HOUGH_ORDER_CAP);
    std::vector<int32_t> lines((size_t)B * 512 * 4), accum((size_t)MDB_HOUGH_ANGLES * (2 * (W + H) + 1), 0);
    std::vector<int> nlines(B), thr(B);
    std::vector<double> thrf(B), snr(B);
    std::vector<unsigned long long> noise((size_t)B * 2);
    ActRing ring; ring.base = actbuf.data(); ring.RA = RA; ring.Wb = Wb; ring.frame_words = FW;
    SparseLists sl; sl.alist = alist.data(); sl.acount = acount.data(); sl.wlist = wlist.data(); sl.wcount = wcount.data(); sl.dense = dense.data();
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
    P.mask_area = mask_area; P.cap = MDB_POINT_CAP; P.max_lines = 512; P.walk_cap = W + H + 2; P.fixed_gap = -1;
    const int rh = roi[2] - roi[0], rw = roi[3] - roi[1];
    const long long std_interval = (long long)nz_interval * n;
    const long long t_end = t_first + Ttot;
    for (long long t0 = t_first; t0 < t_end;) {
        // halo frames go first, in batches of their own (results dropped); then the frames whose results are wanted
        const bool is_halo = t0 < t_first + halo;
        const int T = (int)std::min<long long>(B, (is_halo ? t_first + halo : t_end) - t0);
        const uint8_t *bframes = frames + (size_t)(t0 - t_first) * HW;
        FrameSrc src; src.ring = ringbuf.data(); src.cur = bframes; src.mask = g_dev_mask; src.t0 = t0; src.R = R; src.HW = HW;
        // ---- launch_noise_thr (or the caller's thresholds) --------------------------------------------------------------
        if (thr_in) {
            for (int i = 0; i < T; i++) { thr[i] = thr_in[t0 - t_first + i]; snr[i] = 0.0; }
        } else {
        std::fill(noise.begin(), noise.end(), 0ull);
        SampleList sml; sml.count = 0;
        for (int i = 0; i < T && sml.count >= 0; i++) {
            const long long tau = t0 + i + 1;
            if ((tau > 1 && tau <= n) || (tau > n && std_interval > 0 && tau % std_interval == 0)) {
                if (sml.count < 63) sml.idx[sml.count++] = i; else sml.count = -1;
            }
        }
        if (sml.count != 0) emu_noise_samples(src, W, n, t0, std_interval, roi, noise.data(), 0, sml, T);
        emu_launch(1, 32, [&] { threshold_kernel(&st, noise.data(), T, t0, n, std_interval, (long long)rh * rw, adaptive, sensitivity, thr.data(), thrf.data(), snr.data()); });
        }
        // ---- temporal pass (shape table of temporal3_dispatch.cuh) -------------------------------------------------------
        uint8_t *bits8 = reinterpret_cast<uint8_t *>(bits.data());
        bool t3_done = g_temporal_version == 3;
        if (t3_done) switch (n) {  // generated by tests/emu_build.py from the product's table
#define T3_SHAPE_ARGS src, t0, T, HWG, thr.data(), bits8
#include "t3_shapes_emu.inc"
#undef T3_SHAPE_ARGS
            default: t3_done = false;  // no shape: the product runs temporal2_kernel for this window
        }
        if (!t3_done && temporal2_batch(src, t0, T, n, B, HW, thr.data(), bits8) != 0) return -1001;
        // history for the next batch: the last min(T, n) frames go into the ring (copy_to_ring)
        for (long long t = t0 + T - std::min(T, n); t < t0 + T; t++) memcpy(&ringbuf[(size_t)(t % R) * HW], frames + (size_t)(t - t_first) * HW, HW);
        // ---- stream_kernel_launch: act, dst -------------------------------------------------------------------------------
        std::fill(npoints.begin(), npoints.end(), 0u);
        std::fill(acount.begin(), acount.end(), 0u);
        const int rows_act = 64, strips = (Wb + SP_USE - 1) / SP_USE, bands = (H + rows_act - 1) / rows_act;
        if (Wb % 4 == 0) {
            const int chunks = Wb / 4;
            emu_launch2((chunks * bands + A4_THREADS - 1) / A4_THREADS, T, A4_THREADS, [&] { act4_kernel(bits.data(), H, Wb, rows_act, chunks, bands, ring, t0, sl); });
        } else {
            emu_launch2((strips * bands + SP_WARPS - 1) / SP_WARPS, T, SP_WARPS * 32, [&] { act_kernel(bits.data(), W, H, T, rows_act, strips, bands, ring, t0, sl); });
        }
        if (is_halo) { t0 += T; continue; }  // act_only: no dst, no PPHT for look-back frames
        dense[0] = 0;
        emu_launch(T, 256, [&] { dst_sparse_kernel(ring, W, H, n, t0, dy_on, dst.data(), dstbits.data(), npoints.data(), points.data(), MDB_POINT_CAP, sl); });
        const int dst_rows = 32, dbands = (H + dst_rows - 1) / dst_rows;
        if (dense[0])
            emu_launch2((strips * dbands + SP_WARPS - 1) / SP_WARPS, std::min(T, DENSE_GY), SP_WARPS * 32, [&] {
                dst_dense_kernel(ring, W, H, n, t0, dy_on, dst_rows, strips, dbands, dst.data(), dstbits.data(), npoints.data(), points.data(), MDB_POINT_CAP, sl);
            });
        // ---- launch_hough_kernels (one emulated CTA takes every frame from the queue) ------------------------------------------
        unsigned queue[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        emu_launch(T, 32, [&] { ppht_order_kernel(T, HOUGH_ORDER_CAP, npoints.data(), order.data()); });
        emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, T, npoints.data(), points.data(), order.data(), lines.data(), nlines.data(), queue, nullptr, HOUGH_CAP_SMALL, HOUGH_TABLE_BYTES_SMALL, 0); });
        bool f2 = false, f3 = false, f1 = false;
        for (int i = 0; i < T; i++) f1 |= nlines[i] == -2;
        if (f1) emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, T, npoints.data(), points.data(), order.data(), lines.data(), nlines.data(), queue + 1, nullptr, HOUGH_CAP_LARGE, HOUGH_TABLE_BYTES, 1); });
        for (int i = 0; i < T; i++) f2 |= nlines[i] == -3;
        if (f2) emu_launch(1, HOUGH_THREADS, [&] { hough_tier2_kernel(P, T, npoints.data(), points.data(), accum.data(), lines.data(), nlines.data(), nullptr, queue + 7); });
        for (int i = 0; i < T; i++) f3 |= nlines[i] == -1;
        if (f3) emu_launch(1, HOUGH_THREADS, [&] { hough_tier3_kernel(P, T, dst.data(), okeys.data(), oidx.data(), accum.data(), bitmap.data(), walk.data(), lines.data(), nlines.data(), queue + 2, nullptr); });
        for (int i = 0; i < T; i++) {
            if (nlines[i] < 0) return -(int)(t0 + i + 1);
            const size_t k = (size_t)(t0 - t_first + i);
            thr_out[k] = thr[i]; snr_out[k] = snr[i];
            n_on_out[k] = (int)npoints[i]; lines_num_out[k] = nlines[i];
            memcpy(dst_out + k * HW, dst.data() + (size_t)i * HW, HW);
            memcpy(raw_out + k * 512 * 4, lines.data() + (size_t)i * 512 * 4, (size_t)std::min(nlines[i], 512) * 16);
        }
        t0 += T;
    }
    return 0;
}

extern "C" int emu_stream_path(const uint8_t *frames, int Ttot, int W, int H, int n, int B, int adaptive, int init_value,
                               int sensitivity, int nz_interval, const int *roi, int hough_thr, int hough_min_len,
                               int hough_max_gap, int dy_on, double mask_area, int *thr_out, double *snr_out, uint8_t *dst_out,
                               int *n_on_out, int *lines_num_out, int32_t *raw_out) {
    return run_range(frames, Ttot, 0, 0, nullptr, W, H, n, B, adaptive, init_value, sensitivity, nz_interval, roi, hough_thr,
                     hough_min_len, hough_max_gap, dy_on, mask_area, thr_out, snr_out, dst_out, n_on_out, lines_num_out, raw_out);
}

// one rank's share of a time-sharded run: seek to t_first, `halo` look-back frames, then the chunk with replayed thresholds
extern "C" int emu_stream_chunk(const uint8_t *frames, int Ttot, long long t_first, int halo, const int *thr_in, int W, int H, int n,
                                int B, const int *roi, int hough_thr, int hough_min_len, int hough_max_gap, int dy_on, double mask_area,
                                int *thr_out, double *snr_out, uint8_t *dst_out, int *n_on_out, int *lines_num_out, int32_t *raw_out) {
    return run_range(frames, Ttot, t_first, halo, thr_in, W, H, n, B, 0, 0, 1, 1, roi, hough_thr, hough_min_len, hough_max_gap, dy_on,
                     mask_area, thr_out, snr_out, dst_out, n_on_out, lines_num_out, raw_out);
}

// integer noise sums (sum d, sum d^2) of the sample timers taus[] (what mdb_noise_sums_dev computes for a rank's chunk);
// frames[0] is global frame t_first and must reach back far enough for every sample's window
extern "C" int emu_noise_sums(const uint8_t *frames, int Ttot, long long t_first, int W, int H, int n, int nz_interval, const int *roi,
                              const long long *taus, int ntaus, unsigned long long *sums) {
    const size_t HW = (size_t)W * H;
    std::vector<uint8_t> ringbuf(HW, 0);
    std::vector<unsigned long long> acc((size_t)Ttot * 2, 0ull);
    const int rh = roi[2] - roi[0], rw = roi[3] - roi[1];
    const long long std_interval = (long long)nz_interval * n;
    FrameSrc src; src.ring = ringbuf.data(); src.cur = frames; src.mask = g_dev_mask; src.t0 = t_first; src.R = 1; src.HW = HW;
    for (int k = 0; k < ntaus; k++) {
        const long long tau = taus[k], L = tau < n ? tau : n;
        const long long i = tau - 1 - t_first;
        if (i < 0 || i >= Ttot || tau - L < t_first) return -1;
        SampleList sml; sml.count = 1; sml.idx[0] = (int)i;
        emu_noise_samples(src, W, n, t_first, std_interval, roi, acc.data(), 0, sml, Ttot);
        sums[2 * k] = acc[2 * i]; sums[2 * k + 1] = acc[2 * i + 1];
    }
    return 0;
}


// The per-frame API's resident-state path (pf_update() / mdb_detect() with bits_ready in csrc/metdet.cu): per frame the
// staging copy, noise sample + threshold on (staging, ring), pf_update_kernel in two halves, the suffix rebuild at block ends,
// then act (one-frame batch: 8-row bands), dst_sparse and the PPHT on the predicate bits it left behind.
extern "C" int emu_perframe_path(const uint8_t *frames, int Ttot, int W, int H, int n, int adaptive, int init_value, int sensitivity,
                                 int nz_interval, const int *roi, int hough_thr, int hough_min_len, int hough_max_gap, int dy_on,
                                 double mask_area, int *thr_out, double *snr_out, uint8_t *dst_out, int *n_on_out, int *lines_num_out,
                                 int32_t *raw_out) {
    if (W % 32 || n < 2 || n > 128) return -1000;  // stream_state_init: the generic kernels serve these
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int k = 0; k < MDB_HOUGH_ANGLES; k++) {
        c_trig[2 * k] = (float)cos((double)k * (double)theta);
        c_trig[2 * k + 1] = (float)sin((double)k * (double)theta);
    }
    const size_t HW = (size_t)W * H, groups = HW / 16;
    const int Wb = W / 32, R = n, RA = n;  // max_batch = 1
    const size_t FW = (size_t)H * Wb;
    std::vector<uint8_t> ringbuf((size_t)R * HW, 0), dst(HW, 0), stage(HW, 0), Pm(HW, 0xEE), SUF((size_t)n * HW, 0xEE);
    std::vector<uint16_t> S(HW, 0xEEEE);
    std::vector<uint32_t> actbuf((size_t)RA * FW, 0), bits(33 * FW, 0), dstbits(FW, 0), alist(SPX_ACAP), wlist(SPX_WCAP), points(MDB_POINT_CAP),
        okeys(HW), oidx(HW), bitmap((HW + 31) / 32, 0), walk(W + H + 2, 0);
    std::vector<uint16_t> order(HOUGH_ORDER_CAP);
    std::vector<int32_t> lines(512 * 4), accum((size_t)MDB_HOUGH_ANGLES * (2 * (W + H) + 1), 0);
    unsigned acount = 0, wcount = 0, dense[2] = {0, 0}, npoints = 0;
    ActRing ring; ring.base = actbuf.data(); ring.RA = RA; ring.Wb = Wb; ring.frame_words = FW;
    SparseLists sl; sl.alist = alist.data(); sl.acount = &acount; sl.wlist = wlist.data(); sl.wcount = &wcount; sl.dense = dense;
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
    P.mask_area = mask_area; P.cap = MDB_POINT_CAP; P.max_lines = 512; P.walk_cap = W + H + 2; P.fixed_gap = -1;
    const int rh = roi[2] - roi[0], rw = roi[3] - roi[1];
    const long long std_interval = (long long)nz_interval * n;
    const unsigned pgrid = (unsigned)((groups + PF_THREADS - 1) / PF_THREADS);
    FrameSrc ringsrc; ringsrc.ring = ringbuf.data(); ringsrc.cur = nullptr; ringsrc.mask = g_dev_mask; ringsrc.t0 = 0; ringsrc.R = R; ringsrc.HW = HW;
    bool suffix_pending = false;
    const uint8_t *dm = g_dev_mask;
    emu_launch(pgrid, PF_THREADS, [&] {  // pf_timer = -1
        if (dm) pf_rebuild_kernel<true>(ringsrc, 0, n, S.data(), Pm.data(), SUF.data(), groups);
        else pf_rebuild_kernel<false>(ringsrc, 0, n, S.data(), Pm.data(), SUF.data(), groups);
    });
    for (long long t = 0; t < Ttot; t++) {
        if (suffix_pending) {  // pf_launch_suffix: the block that ended with frame t-1
            emu_launch(pgrid, PF_THREADS, [&] {
                if (dm) pf_suffix_kernel<true>(ringsrc, t - 1, t - n + 1, n - 1, SUF.data(), groups);
                else pf_suffix_kernel<false>(ringsrc, t - 1, t - n + 1, n - 1, SUF.data(), groups);
            });
            suffix_pending = false;
        }
        memcpy(stage.data(), frames + (size_t)t * HW, HW);
        FrameSrc src = ringsrc; src.cur = stage.data(); src.t0 = t;
        unsigned long long noise[2] = {0, 0};
        const long long tau = t + 1;
        SampleList sml; sml.count = 0;
        if ((tau > 1 && tau <= n) || (tau > n && std_interval > 0 && tau % std_interval == 0)) sml.idx[sml.count++] = 0;
        if (sml.count) emu_noise_samples(src, W, n, t, std_interval, roi, noise, 0, sml, 1);
        int thr = 0; double thrf = 0, snr = 0;
        emu_launch(1, 32, [&] { threshold_kernel(&st, noise, 1, t, n, std_interval, (long long)rh * rw, adaptive, sensitivity, &thr, &thrf, &snr); });
        const int pos = (int)(t % n), L = (int)std::min<long long>(n, t + 1);
        uint8_t *slot = ringbuf.data() + (size_t)(t % R) * HW;
        const uint8_t *old = t >= n ? ringbuf.data() + (size_t)((t - n) % R) * HW : nullptr;
        const uint8_t *suf = pos < n - 1 ? SUF.data() + (size_t)(pos + 1) * HW : nullptr;
        const size_t gA = groups / 2;
        for (int half = 0; half < 2; half++) {
            const size_t g0 = half ? gA : 0, g1 = half ? groups : gA;
            if (g1 == g0) continue;
            emu_launch((unsigned)((g1 - g0 + PF_THREADS - 1) / PF_THREADS), PF_THREADS, [&] {
                if (dm) pf_update_kernel<true>(stage.data(), slot, old, dm, S.data(), Pm.data(), suf, pos == 0, L, &thr, g0, g1,
                                               reinterpret_cast<uint16_t *>(bits.data()));
                else pf_update_kernel<false>(stage.data(), slot, old, nullptr, S.data(), Pm.data(), suf, pos == 0, L, &thr, g0, g1,
                                             reinterpret_cast<uint16_t *>(bits.data()));
            });
        }
        suffix_pending = pos == n - 1;
        // ---- mdb_detect with bits_ready: spatial passes on the bits, PPHT ------------------------------------------------
        npoints = 0; acount = 0;
        const int rows_act = 8, strips = (Wb + SP_USE - 1) / SP_USE, bands = (H + rows_act - 1) / rows_act;
        if (Wb % 4 == 0) {
            const int chunks = Wb / 4;
            emu_launch2((chunks * bands + A4_THREADS - 1) / A4_THREADS, 1, A4_THREADS, [&] { act4_kernel(bits.data(), H, Wb, rows_act, chunks, bands, ring, t, sl); });
        } else {
            const int bands64 = (H + 63) / 64;
            emu_launch2((strips * bands64 + SP_WARPS - 1) / SP_WARPS, 1, SP_WARPS * 32, [&] { act_kernel(bits.data(), W, H, 1, 64, strips, bands64, ring, t, sl); });
        }
        dense[0] = 0;
        emu_launch(1, 256, [&] { dst_sparse_kernel(ring, W, H, n, t, dy_on, dst.data(), dstbits.data(), &npoints, points.data(), MDB_POINT_CAP, sl); });
        const int dst_rows = 32, dbands = (H + dst_rows - 1) / dst_rows;
        if (dense[0])
            emu_launch2((strips * dbands + SP_WARPS - 1) / SP_WARPS, 1, SP_WARPS * 32, [&] {
                dst_dense_kernel(ring, W, H, n, t, dy_on, dst_rows, strips, dbands, dst.data(), dstbits.data(), &npoints, points.data(), MDB_POINT_CAP, sl);
            });
        unsigned queue[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int nlines = -99;
        if (npoints == 0) nlines = 0;
        else {
            emu_launch(1, 32, [&] { ppht_order_kernel(1, HOUGH_ORDER_CAP, &npoints, order.data()); });
            emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, 1, &npoints, points.data(), order.data(), lines.data(), &nlines, queue, nullptr, HOUGH_CAP_SMALL, HOUGH_TABLE_BYTES_SMALL, 0); });
        }
        if (nlines == -2) emu_launch(1, HOUGH_THREADS, [&] { hough_smem_kernel(P, 1, &npoints, points.data(), order.data(), lines.data(), &nlines, queue + 1, nullptr, HOUGH_CAP_LARGE, HOUGH_TABLE_BYTES, 1); });
        if (nlines == -3) emu_launch(1, HOUGH_THREADS, [&] { hough_tier2_kernel(P, 1, &npoints, points.data(), accum.data(), lines.data(), &nlines, nullptr, queue + 7); });
        if (nlines == -1) emu_launch(1, HOUGH_THREADS, [&] { hough_tier3_kernel(P, 1, dst.data(), okeys.data(), oidx.data(), accum.data(), bitmap.data(), walk.data(), lines.data(), &nlines, queue + 2, nullptr); });
        if (nlines < 0) return -(int)(t + 1);
        thr_out[t] = thr; snr_out[t] = snr; n_on_out[t] = (int)npoints; lines_num_out[t] = nlines;
        memcpy(dst_out + (size_t)t * HW, dst.data(), HW);
        memcpy(raw_out + (size_t)t * 512 * 4, lines.data(), (size_t)std::min(nlines, 512) * 16);
    }
    return 0;
}
