// Host emulation of temporal3_kernel's per-thread code (same source, intrinsics emulated) against a
// brute-force statement of the predicate  max(window)*L - sum(window) > thr*L  (MetLib/utils.py:269-307,
// MetLib/Detector.py:327-332).  Test infrastructure: built and run by tests/test_t3_emu_cpu.py with g++.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#define T3_HOST_EMU 1
#include "../../metdetpy_b200/csrc/temporal3_kernel.cuh"

static unsigned rng_state = 12345u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

template <int U, int BL, int P, int K, bool MASKED, int FEED>
static int run_case(int HWG, int T, long long t0, int maxbatch, int style) {
    typedef t3::Layout<U, BL, P, FEED, K> LY;
    const int N = LY::N;
    const size_t HW = (size_t)HWG * 8;
    const int R = N - 1 + 2 * maxbatch;
    // the whole stream 0 .. t0+T-1 (frames before 0 are zeros)
    const long long total = t0 + T;
    std::vector<uint8_t> stream((size_t)total * HW);
    for (size_t k = 0; k < stream.size(); k++) {
        unsigned v = rnd();
        if (style == 0) stream[k] = (uint8_t)(40 + (v % 9));                 // noise around a level
        else if (style == 1) stream[k] = (uint8_t)(v & 0xff);                // full range
        else stream[k] = (uint8_t)((v % 97) == 0 ? 255 : (v % 5 == 0 ? 0 : 30 + v % 4));  // spikes and zeros
    }
    std::vector<uint8_t> mask(HW, 1);
    if (MASKED) for (size_t k = 0; k < HW; k++) mask[k] = (rnd() % 4) != 0;
    std::vector<uint8_t> ring((size_t)R * HW, 0xAB);  // garbage where nothing was stored
    for (long long t = (t0 - N + 1 > 0 ? t0 - N + 1 : 0); t < t0; t++)
        memcpy(&ring[(size_t)(t % R) * HW], &stream[(size_t)t * HW], HW);
    std::vector<int> thr(T);
    for (int i = 0; i < T; i++) thr[i] = style == 1 ? (int)(rnd() % 60) : (int)(1 + rnd() % 6);
    FrameSrc src;
    src.ring = ring.data(); src.cur = &stream[(size_t)t0 * HW]; src.mask = MASKED ? mask.data() : nullptr;
    src.t0 = t0; src.R = R; src.HW = HW;
    std::vector<uint8_t> bits((size_t)(T + BL) * HWG, 0xCD);  // BL-1 slack planes (see thread_main)
    const int ctas = (HWG + T3_NT - 1) / T3_NT;
    std::vector<unsigned char> smem(LY::smem_bytes(T) + 64);
    std::vector<uint2> gtab((T + BL - 1) / BL * BL);
    for (size_t i = 0; i < gtab.size(); i++) gtab[i] = t3::table_entry(thr[(int)i < T ? i : T - 1], t0 + (long long)i, N);
    for (int c = 0; c < ctas; c++) {
        if (!FEED) {
            uint2 *tab = reinterpret_cast<uint2 *>(smem.data() + LY::tab_off);
            for (int i = 0; i < (T + BL - 1) / BL * BL; i++) tab[i] = t3::table_entry(thr[i < T ? i : T - 1], t0 + i, N);
        }
        for (int tid = 0; tid < T3_NT; tid++) {
            const int g = c * T3_NT + tid;
            if (g >= HWG) break;
            t3::thread_main<U, BL, P, K, MASKED, FEED>(src, t0, T, g, tid, smem.data(), bits.data(), (size_t)HWG, gtab.data(), T3_NT);
        }
    }
    // brute force
    int bad = 0;
    for (int i = 0; i < T && bad < 5; i++) {
        const long long t = t0 + i;
        const long long L = t + 1 < N ? t + 1 : N;
        for (size_t p = 0; p < HW; p++) {
            int mx = 0, sum = 0;
            for (long long s = t - N + 1; s <= t; s++) {
                if (s < 0) continue;
                const int v = stream[(size_t)s * HW + p] * (MASKED ? mask[p] : 1);
                mx = v > mx ? v : mx; sum += v;
            }
            const int want = mx * (int)L - sum > thr[i] * (int)L;
            const int got = (bits[(size_t)i * HWG + p / 8] >> (p % 8)) & 1;
            if (want != got) {
                if (bad < 5) fprintf(stderr, "U=%d BL=%d P=%d K=%d M=%d F=%d t0=%lld T=%d: frame %d pixel %zu want %d got %d\n",
                                     U, BL, P, K, (int)MASKED, FEED, t0, T, i, p, want, got);
                bad++;
            }
        }
    }
    return bad;
}

template <int U, int BL, int P, int K, int FEED = 0>
static int run_shape() {
    int bad = 0;
    const int N = U * P;
    const long long t0s[] = {0, 1, N - 2 > 0 ? N - 2 : 0, N, 3 * N + 1, 1000};
    const int Ts[] = {1, 2, BL, BL + 1, U - 1, U, U + 1, 2 * U + 3, 5 * U - 1, 64};
    for (long long t0 : t0s)
        for (int T : Ts) {
            if (T < 1) continue;
            for (int style = 0; style < 3; style++) {
                bad += run_case<U, BL, P, K, false, FEED>(13, T, t0, 64, style);
                if (style == 0) bad += run_case<U, BL, P, K, true, FEED>(130, T, t0, 64, style);
            }
        }
    printf("shape U=%d BL=%d P=%d K=%d FEED=%d: %s\n", U, BL, P, K, FEED, bad ? "FAIL" : "ok");
    return bad;
}

int main() {
    int bad = 0;
    bad += run_shape<15, 15, 2, 15>();  // the default shape for n = 30
    bad += run_shape<30, 15, 2, 15>();  // ... for n = 60
    bad += run_shape<5, 5, 1, 5>();     // ... for n = 5
    bad += run_shape<30, 10, 1, 6>();
    bad += run_shape<30, 15, 1, 5>();
    bad += run_shape<30, 10, 2, 6>();
    bad += run_shape<5, 5, 1, 5>();
    bad += run_shape<6, 6, 1, 6>();
    bad += run_shape<6, 3, 1, 3>();
    bad += run_shape<25, 5, 1, 5>();
    bad += run_shape<12, 4, 2, 4>();
    bad += run_shape<2, 2, 1, 2>();
    bad += run_shape<15, 15, 4, 15>();
    bad += run_shape<20, 10, 3, 10>();
    bad += run_shape<6, 3, 4, 3>();
    bad += run_shape<30, 10, 1, 5, 1>();
    bad += run_shape<30, 10, 2, 5, 1>();
    bad += run_shape<24, 12, 1, 6, 1>();
    printf(bad ? "FAILED\n" : "ALL OK\n");
    return bad ? 1 : 0;
}
