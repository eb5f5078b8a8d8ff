// Host emulation of act4_kernel (csrc/spatial_kernel.cuh: 3x3 median of the thresholded difference as a majority of 9
// predicate bits, then MORPH_CLOSE = dilate + erode, on 32-pixel words; same source, CUDA built-ins emulated) against a
// per-pixel statement of cv2.medianBlur(.,3) (border replicated) -> threshold -> cv2.morphologyEx(CLOSE, 3x3) (pixels
// outside the image ignored), MetLib/Detector.py:329-335.  Also checks the list of non-zero act words the kernel emits.
// Test infrastructure: built and run by tests/test_act4_emu_cpu.py with g++.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

struct EmuIdx { unsigned x, y, z; };
static EmuIdx emu_blockIdx, emu_threadIdx, emu_blockDim, emu_gridDim;
#define blockIdx emu_blockIdx
#define threadIdx emu_threadIdx
#define blockDim emu_blockDim
#define gridDim emu_gridDim
#undef __launch_bounds__
#define __launch_bounds__(...)
#define __shared__ static
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v) { const unsigned o = *p; *p = o | v; return o; }
// warp-level built-ins only appear in kernels this harness does not run
template <typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int) { return v; }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline unsigned __activemask() { return 1u; }
static inline int __any_sync(unsigned, int p) { return p; }
static inline int __all_sync(unsigned, int p) { return p; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
using std::max;
using std::min;
#include "../../metdetpy_b200/csrc/spatial_kernel.cuh"

static unsigned rng_state = 4242u;
static unsigned rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

static int run_case(int W, int H, int rows, int density) {
    const int Wb = W / 32;
    std::vector<uint8_t> pred((size_t)H * W);
    for (auto &p : pred) p = (rnd() % 100) < (unsigned)density;
    if (density > 5)  // a few solid blobs and lines so that the close has something to close
        for (int k = 0; k < 6; k++) {
            const int y0 = rnd() % H, x0 = rnd() % W;
            for (int y = y0; y < std::min(H, y0 + 3 + (int)(rnd() % 4)); y++)
                for (int x = x0; x < std::min(W, x0 + 5 + (int)(rnd() % 40)); x++) pred[(size_t)y * W + x] = 1;
        }
    std::vector<uint32_t> bits((size_t)H * Wb, 0);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            if (pred[(size_t)y * W + x]) bits[(size_t)y * Wb + x / 32] |= 1u << (x % 32);
    // reference, per pixel
    auto at = [&](const std::vector<uint8_t> &im, int y, int x, int outside) {
        return (y < 0 || y >= H || x < 0 || x >= W) ? outside : (int)im[(size_t)y * W + x];
    };
    std::vector<uint8_t> bin((size_t)H * W), dil((size_t)H * W), act((size_t)H * W);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int c = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++)
                    c += pred[(size_t)std::min(std::max(y + dy, 0), H - 1) * W + std::min(std::max(x + dx, 0), W - 1)];
            bin[(size_t)y * W + x] = c >= 5;  // median of nine {0,1} values
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int v = 0;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) v |= at(bin, y + dy, x + dx, 0);
            dil[(size_t)y * W + x] = (uint8_t)v;
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int v = 1;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) v &= at(dil, y + dy, x + dx, 1);
            act[(size_t)y * W + x] = (uint8_t)v;
        }
    // kernel
    const int RA = 3;
    std::vector<uint32_t> ringbuf((size_t)RA * H * Wb, 0xDEADBEEFu), alist(SPX_ACAP), wlist(1), dense(2);
    std::vector<unsigned> acount(1, 0), wcount(1, 0);
    ActRing ring; ring.base = ringbuf.data(); ring.RA = RA; ring.Wb = Wb; ring.frame_words = (size_t)H * Wb;
    SparseLists sl; sl.alist = alist.data(); sl.acount = acount.data(); sl.wlist = wlist.data(); sl.wcount = wcount.data(); sl.dense = dense.data();
    const long long dy0 = 7;
    const int chunks = Wb / 4, bands = (H + rows - 1) / rows;
    const unsigned grid = (unsigned)((chunks * bands + A4_THREADS - 1) / A4_THREADS);
    emu_blockIdx.y = 0;
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < A4_THREADS; t++) {
            emu_blockIdx.x = b; emu_threadIdx.x = t;
            act4_kernel(bits.data(), H, Wb, rows, chunks, bands, ring, dy0, sl);
        }
    const uint32_t *out = ring.frame(dy0);
    int bad = 0;
    std::set<unsigned> nz;
    for (int y = 0; y < H; y++)
        for (int wx = 0; wx < Wb; wx++) {
            unsigned want = 0;
            for (int b = 0; b < 32; b++) want |= (unsigned)act[(size_t)y * W + wx * 32 + b] << b;
            if (want) nz.insert(((unsigned)y << 12) | (unsigned)wx);
            if (out[(size_t)y * Wb + wx] != want && bad < 5) {
                fprintf(stderr, "W=%d H=%d rows=%d density=%d: row %d word %d want %08x got %08x\n", W, H, rows, density, y, wx, want,
                        out[(size_t)y * Wb + wx]);
                bad++;
            }
        }
    if (acount[0] <= SPX_ACAP) {
        std::set<unsigned> got(alist.begin(), alist.begin() + acount[0]);
        if (got != nz || got.size() != acount[0]) { fprintf(stderr, "W=%d H=%d: act word list differs (%zu vs %zu)\n", W, H, got.size(), nz.size()); bad++; }
    }
    return bad;
}

int main() {
    int bad = 0, cases = 0;
    for (int W : {128, 256, 384})
        for (int H : {1, 2, 3, 5, 17, 64, 67})
            for (int rows : {8, 64})
                for (int density : {0, 2, 30, 60, 100}) { bad += run_case(W, H, rows, density); cases++; }
    printf("%d cases: %s\n", cases, bad ? "FAILED" : "ALL OK");
    return bad ? 1 : 0;
}
