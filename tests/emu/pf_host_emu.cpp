// Host emulation of the per-frame resident-state kernels (csrc/perframe_kernel.cuh: pf_update_kernel, pf_suffix_kernel,
// pf_rebuild_kernel; same source, CUDA built-ins emulated) against a brute-force statement of
// SlidingWindow.update + the predicate  max(window)*L - sum(window) > thr*L  (MetLib/utils.py:269-307,
// MetLib/Detector.py:327-332).  The driver below mirrors pf_update() / pf_launch_suffix() of csrc/metdet.cu: staging
// buffer, ring slot written by the update kernel, suffix planes rebuilt at block ends, state rebuilt from the ring
// after a jump (what batched calls, reset and seek cause).  Test infrastructure: built and run by
// tests/test_pf_emu_cpu.py with g++.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

struct EmuIdx { unsigned x, y, z; };
static EmuIdx emu_blockIdx, emu_threadIdx;
#define blockIdx emu_blockIdx
#define threadIdx emu_threadIdx
#undef __launch_bounds__
#define __launch_bounds__(...)
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __vmaxu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int k = 0; k < 4; k++) {
        const unsigned x = (a >> (8 * k)) & 0xff, y = (b >> (8 * k)) & 0xff;
        r |= (x > y ? x : y) << (8 * k);
    }
    return r;
}
using std::max;
using std::min;
#include "../../metdetpy_b200/csrc/perframe_kernel.cuh"

static unsigned rng_state = 2024u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

template <typename F>
static void launch(size_t groups, F f) {  // every thread of the grid, serially
    const unsigned grid = (unsigned)((groups + PF_THREADS - 1) / PF_THREADS);
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < PF_THREADS; t++) {
            emu_blockIdx.x = b; emu_threadIdx.x = t;
            f();
        }
}

struct Emu {
    int n, R;
    size_t HW, groups;
    bool masked;
    std::vector<uint8_t> ring, mask, stage, P, SUF;
    std::vector<uint16_t> S, bits;
    long long timer = 0, pf_timer = -1;
    bool suffix_pending = false;
    FrameSrc src() const {
        FrameSrc s;
        s.ring = ring.data(); s.cur = nullptr; s.mask = masked ? mask.data() : nullptr; s.t0 = 0; s.R = R; s.HW = HW;
        return s;
    }
    void suffix() {
        const FrameSrc s = src();
        const long long hi = timer - 1, lo = timer - n + 1;
        if (masked) launch(groups, [&] { pf_suffix_kernel<true>(s, hi, lo, n - 1, SUF.data(), groups); });
        else launch(groups, [&] { pf_suffix_kernel<false>(s, hi, lo, n - 1, SUF.data(), groups); });
        suffix_pending = false;
    }
    void update(const uint8_t *frame, int thr) {
        if (suffix_pending && pf_timer == timer) suffix();
        if (pf_timer != timer) {
            const FrameSrc s = src();
            if (masked) launch(groups, [&] { pf_rebuild_kernel<true>(s, timer, n, S.data(), P.data(), SUF.data(), groups); });
            else launch(groups, [&] { pf_rebuild_kernel<false>(s, timer, n, S.data(), P.data(), SUF.data(), groups); });
            suffix_pending = false;
        }
        const long long t = timer;
        memcpy(stage.data(), frame, HW);
        const int pos = (int)(t % n), L = (int)std::min<long long>(n, t + 1);
        uint8_t *slot = ring.data() + (size_t)(t % R) * HW;
        const uint8_t *old = t >= n ? ring.data() + (size_t)((t - n) % R) * HW : nullptr;
        const uint8_t *suf = pos < n - 1 ? SUF.data() + (size_t)(pos + 1) * HW : nullptr;
        const size_t gA = groups / 2;
        for (int half = 0; half < 2; half++) {
            const size_t g0 = half ? gA : 0, g1 = half ? groups : gA;
            if (g1 == g0) continue;
            if (masked) launch(g1 - g0, [&] { pf_update_kernel<true>(stage.data(), slot, old, mask.data(), S.data(), P.data(), suf, pos == 0, L, &thr, g0, g1, bits.data()); });
            else launch(g1 - g0, [&] { pf_update_kernel<false>(stage.data(), slot, old, nullptr, S.data(), P.data(), suf, pos == 0, L, &thr, g0, g1, bits.data()); });
        }
        timer += 1;
        pf_timer = timer;
        suffix_pending = pos == n - 1;
    }
};

// frames t of the stream (frames before 0 are zeros); jump_at: the emulated handle "loses" its resident state at these timers
// (as after a batched call) while the ring keeps the last n frames -- written here the way the batched path leaves them
static int run_case(int n, int Rextra, size_t HW, bool masked, int T, int style, const std::vector<int> &jumps) {
    Emu e;
    e.n = n; e.R = n + Rextra; e.HW = HW; e.groups = HW / 16; e.masked = masked;
    e.ring.assign((size_t)e.R * HW, 0); e.mask.assign(HW, 1); e.stage.assign(HW, 0); e.P.assign(HW, 0xEE);
    e.SUF.assign((size_t)n * HW, 0xEE); e.S.assign(HW, 0xEEEE); e.bits.assign(HW / 16, 0xEEEE);
    if (masked) for (size_t k = 0; k < HW; k++) e.mask[k] = (rnd() % 4) != 0;
    std::vector<uint8_t> stream((size_t)T * HW);
    for (size_t k = 0; k < stream.size(); k++) {
        const unsigned v = rnd();
        stream[k] = style == 0 ? (uint8_t)(40 + v % 9) : style == 1 ? (uint8_t)(v & 0xff) : (uint8_t)((v % 53) == 0 ? 255 : (v % 7 == 0 ? 0 : 30 + v % 4));
    }
    int bad = 0;
    for (int t = 0; t < T && bad < 5; t++) {
        if (std::find(jumps.begin(), jumps.end(), t) != jumps.end()) e.pf_timer = -1;  // forces the rebuild from the ring
        const int thr = style == 1 ? (int)(rnd() % 60) : (int)(1 + rnd() % 6);
        e.update(&stream[(size_t)t * HW], thr);
        const long long L = t + 1 < n ? t + 1 : n;
        for (size_t p = 0; p < HW && bad < 5; p++) {
            int mx = 0, sum = 0;
            for (long long s = t - n + 1; s <= t; s++) {
                if (s < 0) continue;
                const int v = stream[(size_t)s * HW + p] * (masked ? e.mask[p] : 1);
                mx = v > mx ? v : mx; sum += v;
            }
            const int want = mx * (int)L - sum > thr * (int)L;
            const int got = (e.bits[p / 16] >> (p % 16)) & 1;
            const bool ring_ok = e.ring[(size_t)(t % e.R) * HW + p] == stream[(size_t)t * HW + p];
            if (want != got || e.S[p] != sum || !ring_ok) {
                fprintf(stderr, "n=%d R=%d masked=%d style=%d frame %d pixel %zu: predicate want %d got %d, sum want %d got %d, ring %s\n",
                        n, e.R, (int)masked, style, t, p, want, got, sum, (int)e.S[p], ring_ok ? "ok" : "WRONG");
                bad++;
            }
        }
    }
    return bad;
}

int main() {
    int bad = 0, cases = 0;
    const int ns[] = {2, 3, 5, 8, 12, 30};
    for (int n : ns)
        for (int Rextra : {0, 1, 7})           // R == n (max_batch = 1) and larger rings (max_batch > 1)
            for (int masked = 0; masked < 2; masked++)
                for (int style = 0; style < 3; style++) {
                    const int T = 3 * n + 5;
                    bad += run_case(n, Rextra, 16 * 11, masked != 0, T, style, {});
                    bad += run_case(n, Rextra, 16 * 6, masked != 0, T, style, {1, n - 1, n, n + 1, 2 * n, 2 * n + 3});
                    cases += 2;
                }
    printf("%d cases: %s\n", cases, bad ? "FAILED" : "ALL OK");
    return bad ? 1 : 0;
}
