// Host emulation of the frame-pass kernels of the MFNR stacker (csrc/mfnr.cuh: mfnr_accum_kernel, mfnr_sigma_kernel,
// mfnr_median_kernel; same source, CUDA built-ins emulated) against brute-force statements of MaxImgContainer /
// FastGaussianContainer (MetLib/stacker.py:43-59, uint16 / uint32 wrap-around), single_sigma_clipping (:94-115) and
// np.median / median_of_medians (:62-78).  Test infrastructure: built and run by tests/test_mfnr_emu_cpu.py with g++.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

struct EmuIdx { unsigned x, y, z; };
static EmuIdx emu_blockIdx, emu_threadIdx, emu_gridDim;
#define blockIdx emu_blockIdx
#define threadIdx emu_threadIdx
#define gridDim emu_gridDim
#undef __launch_bounds__
#define __launch_bounds__(...)
#define __shared__ static
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }  // reductions are not exercised here
static inline void __syncthreads() {}
using std::max;
using std::min;
#include "../../metdetpy_b200/csrc/mfnr.cuh"

static unsigned rng_state = 777u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

template <typename F>
static void launch(size_t threads, F f) {
    const unsigned grid = (unsigned)((threads + MF_THREADS - 1) / MF_THREADS);
    emu_gridDim.x = grid;
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < MF_THREADS; t++) { emu_blockIdx.x = b; emu_threadIdx.x = t; f(); }
}

static double np_median(std::vector<double> v) {
    std::sort(v.begin(), v.end());
    const size_t m = v.size();
    return m & 1 ? v[m / 2] : (v[m / 2 - 1] + v[m / 2]) * 0.5;
}

template <int V>
static int run_case(size_t E, int N, int style, int chunk) {
    std::vector<uint8_t> frames((size_t)N * E);
    for (auto &x : frames) {
        const unsigned v = rnd();
        x = style == 0 ? (uint8_t)(40 + v % 9) : style == 1 ? (uint8_t)(v & 0xff) : (uint8_t)((v % 11) == 0 ? 255 : 200 + v % 56);
    }
    int bad = 0;
    // ---- accumulation in chunks ---------------------------------------------------------------------------------
    std::vector<uint8_t> mx(E, 0xEE);
    std::vector<uint16_t> sum(E, 0xEEEE);
    std::vector<uint32_t> sq(E, 0xEEEEEEEEu);
    std::vector<const uint8_t *> cptr;
    std::vector<int> ccnt;
    for (int t0 = 0; t0 < N; t0 += chunk) {
        const int c = std::min(chunk, N - t0);
        const uint8_t *p = &frames[(size_t)t0 * E];
        launch(E / V, [&] { mfnr_accum_kernel<V>(p, c, E, mx.data(), sum.data(), sq.data(), t0 == 0); });
        cptr.push_back(p); ccnt.push_back(c);
    }
    for (size_t i = 0; i < E && bad < 5; i++) {
        unsigned m = 0; uint16_t s = 0; uint32_t q = 0;
        for (int t = 0; t < N; t++) { const unsigned x = frames[(size_t)t * E + i]; m = std::max(m, x); s = (uint16_t)(s + x); q += x * x; }
        if (mx[i] != m || sum[i] != s || sq[i] != q) { fprintf(stderr, "accum E=%zu N=%d i=%zu\n", E, N, i); bad++; }
    }
    // ---- sigma clipping --------------------------------------------------------------------------------------------
    std::vector<uint16_t> sum2(E); std::vector<uint32_t> sq2(E); std::vector<int32_t> n2(E);
    MfnrChunks ch; ch.ptr = cptr.data(); ch.count = ccnt.data(); ch.n = (int)cptr.size();
    launch(E / V, [&] { mfnr_sigma_kernel<V>(ch, E, N, 3.0, 3.0, sum.data(), sq.data(), sum2.data(), sq2.data(), n2.data()); });
    for (size_t i = 0; i < E && bad < 5; i++) {
        const int n16 = (int)(int16_t)N;
        const double mu = std::nearbyint((double)sum[i] / (double)n16);
        const unsigned s2 = (unsigned)sum[i] * (unsigned)sum[i];
        const double var = ((double)sq[i] - (double)s2 / (double)n16) / (double)(n16 - 1), sd = std::sqrt(var);
        auto u8 = [](double v) { return (unsigned)std::min(std::max(std::nearbyint(v), 0.0), 255.0); };
        const unsigned hi = u8(mu + 3.0 * sd), lo = u8(mu - 3.0 * sd);
        uint16_t cs = 0; uint32_t cq = 0; uint16_t cn = 0;
        for (int t = 0; t < N; t++) { const unsigned x = frames[(size_t)t * E + i]; if (x > hi || x < lo) { cs = (uint16_t)(cs + x); cq += x * x; cn++; } }
        if (sum2[i] != (uint16_t)(sum[i] - cs) || sq2[i] != sq[i] - cq || n2[i] != n16 - (int)cn) {
            fprintf(stderr, "sigma E=%zu N=%d i=%zu\n", E, N, i); bad++;
        }
    }
    // ---- medians: plain and block-wise --------------------------------------------------------------------------------
    std::vector<const uint8_t *> fptr(N);
    for (int t = 0; t < N; t++) fptr[t] = &frames[(size_t)t * E];
    for (int bs : {0, std::max(1, (int)std::sqrt((double)N))}) {
        std::vector<float> mu(E, -1.f);
        launch(E / V, [&] { mfnr_median_kernel<V>(fptr.data(), E, N, bs, mu.data()); });
        for (size_t i = 0; i < E && bad < 5; i++) {
            double want;
            if (bs == 0) {
                std::vector<double> v(N);
                for (int t = 0; t < N; t++) v[t] = frames[(size_t)t * E + i];
                want = np_median(v);
            } else {
                std::vector<double> meds;
                for (int b0 = 0; b0 < N; b0 += bs) {
                    std::vector<double> v;
                    for (int t = b0; t < std::min(b0 + bs, N); t++) v.push_back(frames[(size_t)t * E + i]);
                    meds.push_back(np_median(v));
                }
                want = np_median(meds);
            }
            if ((double)mu[i] != want) { fprintf(stderr, "median E=%zu N=%d bs=%d i=%zu want %.3f got %.3f\n", E, N, bs, i, want, (double)mu[i]); bad++; }
        }
    }
    return bad;
}

int main() {
    int bad = 0, cases = 0;
    for (int N : {2, 3, 4, 7, 12, 16, 17, 24, 50, 97, 300})
        for (int style = 0; style < 3; style++) {
            bad += run_case<4>(4 * 37, N, style, 5);
            bad += run_case<1>(29, N, style, N);
            cases += 2;
        }
    printf("%d cases: %s\n", cases, bad ? "FAILED" : "ALL OK");
    return bad ? 1 : 0;
}
