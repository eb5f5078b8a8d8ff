// The whole MFNR mix stacker of the product on the CPU: the kernels of csrc/mfnr.cuh run by the thread-block emulator
// (tests/emu/cuda_block_emu.h) in the order mdb_mfnr_append / mdb_mfnr_finish (csrc/metdet.cu) launch them, host arithmetic
// between the launches restated from there (Gumbel mean handed in by the caller as metdetpy_b200/stacker.py does, the
// cv2.getGaussianKernel taps, the left-to-right constants).  Compared by tests/test_mfnr_emu_cpu.py with golden images of the
// live mfnr_mix_stacker (MetLib/stacker.py:296-403).  Test infrastructure, not product code.
#include "cuda_block_emu.h"
#include "mfnr_emu.cuh"

extern "C" int emu_mfnr_mix(const uint8_t *frames, int N, int H, int W, int C, int chunk, int bg_algorithm, double highlight_preserve,
                            int ks, double blur_sigma, double sigma_high, double sigma_low, double bg_fix_factor, double gumbel_mean,
                            int med_block_size, uint8_t *out, double *stats) {
    if (N < 2 || ks < 1 || ks % 2 == 0 || bg_algorithm < 0 || bg_algorithm > 3 || chunk < 1) return -1;
    const size_t P = (size_t)H * W, E = P * C;
    const bool clip = bg_algorithm == 1, med = bg_algorithm >= 2;
    int med_block = 0;
    if (bg_algorithm == 3 && N > 16) {
        med_block = med_block_size > 0 ? med_block_size : (int)std::sqrt((double)N);
        if (med_block < 1 || (N - 1) / med_block + 1 > MF_MAX_BLOCKS) return -2;
    }
    // ---- mdb_mfnr_append, chunk by chunk ---------------------------------------------------------------------------
    std::vector<uint8_t> mx(E, 0xEE);
    std::vector<uint16_t> sum(E, 0xEEEE);
    std::vector<uint32_t> sq(E, 0xEEEEEEEEu);
    std::vector<const uint8_t *> cptr, fptr;
    std::vector<int> ccnt;
    for (int t0 = 0; t0 < N; t0 += chunk) {
        const int T = std::min(chunk, N - t0);
        const uint8_t *buf = frames + (size_t)t0 * E;
        const int first = t0 == 0;
        if (E % 4 == 0)
            emu_launch((unsigned)((E / 4 + MF_THREADS - 1) / MF_THREADS), MF_THREADS, [&] { mfnr_accum_kernel<4>(buf, T, E, mx.data(), sum.data(), sq.data(), first); });
        else
            emu_launch((unsigned)((E + MF_THREADS - 1) / MF_THREADS), MF_THREADS, [&] { mfnr_accum_kernel<1>(buf, T, E, mx.data(), sum.data(), sq.data(), first); });
        cptr.push_back(buf); ccnt.push_back(T);
        for (int t = 0; t < T; t++) fptr.push_back(buf + (size_t)t * E);
    }
    // ---- mdb_mfnr_finish -------------------------------------------------------------------------------------------
    const unsigned gE = (unsigned)((E + MF_THREADS - 1) / MF_THREADS), gP = (unsigned)((P + MF_THREADS - 1) / MF_THREADS);
    std::vector<double> pv(MF_PARTS), row(P), blur(P);
    std::vector<unsigned long long> pc(MF_PARTS);
    std::vector<uint8_t> fg(P);
    std::vector<uint16_t> sum2(clip ? E : 0);
    std::vector<uint32_t> sq2(clip ? E : 0);
    std::vector<int32_t> n2(clip ? E : 0);
    std::vector<float> mu(med ? E : 0);
    const uint16_t *s = sum.data();
    const uint32_t *q = sq.data();
    const int32_t *n_arr = nullptr;
    const float *mu_arr = med ? mu.data() : nullptr;
    if (clip) {
        MfnrChunks ch; ch.ptr = cptr.data(); ch.count = ccnt.data(); ch.n = (int)cptr.size();
        if (E % 4 == 0)
            emu_launch((unsigned)((E / 4 + MF_THREADS - 1) / MF_THREADS), MF_THREADS,
                       [&] { mfnr_sigma_kernel<4>(ch, E, N, sigma_high, sigma_low, sum.data(), sq.data(), sum2.data(), sq2.data(), n2.data()); });
        else
            emu_launch(gE, MF_THREADS, [&] { mfnr_sigma_kernel<1>(ch, E, N, sigma_high, sigma_low, sum.data(), sq.data(), sum2.data(), sq2.data(), n2.data()); });
        s = sum2.data(); q = sq2.data(); n_arr = n2.data();
    }
    if (med) {
        if (E % 4 == 0)
            emu_launch((unsigned)((E / 4 + MF_THREADS - 1) / MF_THREADS), MF_THREADS, [&] { mfnr_median_kernel<4>(fptr.data(), E, N, med_block, mu.data()); });
        else
            emu_launch(gE, MF_THREADS, [&] { mfnr_median_kernel<1>(fptr.data(), E, N, med_block, mu.data()); });
    }
    double tot = 0.0;
    unsigned long long cnt = 0;
    emu_launch(MF_PARTS, MF_THREADS, [&] { mfnr_sqrtvar_kernel(E, N, s, q, n_arr, pv.data(), pc.data()); });
    emu_launch(1, 32, [&] { mfnr_final_reduce_kernel(MF_PARTS, pv.data(), pc.data(), &tot, &cnt); });
    const double est_bg_var = tot / (double)E;
    double g = gumbel_mean;
    if (!(g > 0.0)) {
        const double s2 = std::sqrt(2.0 * std::log((double)N));
        g = s2 - (std::log(std::log((double)N)) + std::log(4.0 * 3.141592653589793)) / (2.0 * s2) + 0.5772 / s2;
    }
    volatile double c2v = est_bg_var * g;
    volatile double c1v = c2v * bg_fix_factor;
    const double c1 = c1v, c2 = c2v;
    emu_launch(MF_PARTS, MF_THREADS, [&] { mfnr_diffpos_kernel(E, N, c1, mx.data(), s, n_arr, mu_arr, pv.data(), pc.data()); });
    emu_launch(1, 32, [&] { mfnr_final_reduce_kernel(MF_PARTS, pv.data(), pc.data(), &tot, &cnt); });
    const double avg = tot / (double)cnt;
    std::vector<double> taps(ks);
    {
        const double sigma = blur_sigma > 0 ? blur_sigma : 3.0;
        double ssum = 0.0;
        for (int i = 0; i < ks; i++) {
            const double x = i - (ks - 1) * 0.5;
            taps[i] = std::exp(-(x * x) / (2.0 * sigma * sigma));
            ssum += taps[i];
        }
        for (int i = 0; i < ks; i++) taps[i] = taps[i] / ssum;
    }
    volatile double hl = 255.0 * highlight_preserve, omh = 1.0 - highlight_preserve;
    const double hl_ = hl, omh_ = omh;
    emu_launch(gP, MF_THREADS, [&] { mfnr_mask_kernel(P, C, N, c1, avg, hl_, mx.data(), s, n_arr, mu_arr, fg.data()); });
    emu_launch(gP, MF_THREADS, [&] { mfnr_blur_row_kernel(H, W, ks, taps.data(), fg.data(), row.data()); });
    emu_launch(gP, MF_THREADS, [&] { mfnr_blur_col_kernel(H, W, ks, taps.data(), row.data(), blur.data()); });
    emu_launch(gE, MF_THREADS, [&] { mfnr_mix_kernel(E, C, N, c2, highlight_preserve, omh_, mx.data(), s, n_arr, mu_arr, blur.data(), out); });
    if (stats) { stats[0] = est_bg_var; stats[1] = g; stats[2] = avg; stats[3] = (double)cnt; }
    return 0;
}
