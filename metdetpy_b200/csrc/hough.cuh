// Exact-order progressive probabilistic Hough transform, one CTA per frame, many frames in flight.
// Replaces cv2.HoughLinesP(dst, 1, pi/180, threshold, minLineLength, maxLineGap) at
// MetLib/Detector.py:347-352 and reproduces its visiting order (row-major point list + OpenCV's
// MWC RNG seeded per call), its float32 rho rounding and its in-loop un-voting, so the emitted
// segments are identical.
//
// Tiers (a frame takes the first one it fits; the choice never changes the result):
//   1a / 1b  hough_smem_kernel: the frame's whole PPHT on chip.  On-pixels sorted in shared memory (row-major == cv2's
//            nzloc order); "is pixel q still on" = binary search in the sorted keys + a removed-bit per point; the
//            accumulator is a shared-memory table of per-angle rho INTERVALS (int16 cells; two intervals per angle when
//            one does not fit: two far-apart objects); thread n < 180 owns angle n, so votes need no atomics; votes of
//            four points per barrier (exact: a thread reads its cell right after its own increment, votes behind a
//            line-yielding point are rolled back); arg-max = warp REDUX + one shared-memory hop; line walks 256 steps
//            at a time by the whole CTA (ballots).  1a: <= 2048 points, 90 KB table, two CTAs per SM; 1b: <= 4096
//            points, 184 KB table.  CTAs take frames from a queue.
//   2        hough_tier2_kernel: frames of up to MDB_POINT_CAP points whose intervals do not fit on chip (8K streaks):
//            same scheme with the accumulator rows in global memory ([180][numrho] int32 per slot), cells of the next
//            visits prefetched into L2, accumulator cleared by per-angle rho intervals.
//   3        hough_tier3_kernel: dense masks beyond MDB_POINT_CAP.  Ordered row-wise compaction of the mask into a
//            global point list, visiting order shuffled in shared memory (up to H3_ORDER_CAP points), visits staged
//            256 at a time, four votes per barrier, ballot walks on a pixel bitmap, isolated pixels skipped by a
//            neighbour flag; up to one CTA per SM, each with its own scratch slot.
#pragma once
#include <limits.h>

#include "common.cuh"

__constant__ float c_trig[2 * MDB_HOUGH_ANGLES];  // (float)cos(n*theta), (float)sin(n*theta); host-computed

#define HOUGH_THREADS 256
#define HOUGH_PREFETCH 6  // visits of look-ahead for the accumulator-cell L2 prefetch

__device__ __forceinline__ int rho_of(int x, int y, int n) {
    // plain float32 multiply/add, no FMA contraction (matches the compiled OpenCV loop)
    const float r = __fadd_rn(__fmul_rn((float)x, c_trig[2 * n]), __fmul_rn((float)y, c_trig[2 * n + 1]));
    return __float2int_rn(r);  // cvRound: round half to even
}

// same, with the angle's cos/sin already in registers (thread n owns angle n: reading c_trig[2*n]
// from every lane would serialise on the constant cache).  The int <-> float conversions are done with
// magic-number adds instead of I2F / F2I: those run on the quarter-rate XU pipe, and 180 threads x
// (votes + un-votes + range scan) made it the busiest unit of the kernel.  Exact: 0x4B000000 | v is the
// float 2^23 + v for 0 <= v < 2^23, and adding 1.5 * 2^23 to |r| < 2^22 rounds to the nearest integer,
// ties to even, exactly like cvRound / F2I.RN.
__device__ __forceinline__ float u16_to_float(unsigned v) {
    return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f);
}
__device__ __forceinline__ int rho_cs(int x, int y, float c, float s) {
    const float r = __fadd_rn(__fmul_rn(u16_to_float((unsigned)x), c), __fmul_rn(u16_to_float((unsigned)y), s));
    return __float_as_int(__fadd_rn(r, 12582912.0f)) - 0x4B400000;
}

__device__ __forceinline__ void bitonic_sort_u32(uint32_t *a, int npow2, int tid, int nthreads) {
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npow2; i += nthreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t x = a[i], y = a[ixj];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

__device__ __forceinline__ int line_gap_of(const HoughParams &P, unsigned n_on) {
    if (P.fixed_gap >= 0) return P.fixed_gap;  // ClassicDetector: maxLineGap = hough_cfg.max_gap (Detector.py:284-289)
    // Detector.py:342-344: dst_sum = count / mask_area * 100; gap = max(0, 1 - dst_sum/0.05) * max_gap
    const double dst_sum = __dmul_rn(__ddiv_rn((double)n_on, P.mask_area), 100.0);
    double g = __dsub_rn(1.0, __ddiv_rn(dst_sum, 0.05));
    if (!(g > 0.0)) g = 0.0;
    g = __dmul_rn(g, (double)P.max_gap);
    return (int)rint(g);  // cvRound(maxLineGap)
}

// walk geometry of OpenCV's 16.16 fixed-point line tracer
struct Walk {
    int x0, y0, dx0, dy0;
    bool xflag;
    __device__ __forceinline__ void init(int x, int y, int max_n) {
        const float a = -c_trig[2 * max_n + 1], b = c_trig[2 * max_n];
        x0 = x; y0 = y;
        if (fabsf(a) > fabsf(b)) {
            xflag = true;
            dx0 = a > 0 ? 1 : -1;
            dy0 = __float2int_rn(__fdiv_rn(__fmul_rn(b, 65536.0f), fabsf(a)));
            y0 = (y0 << 16) + 32768;
        } else {
            xflag = false;
            dy0 = b > 0 ? 1 : -1;
            dx0 = __float2int_rn(__fdiv_rn(__fmul_rn(a, 65536.0f), fabsf(b)));
            x0 = (x0 << 16) + 32768;
        }
    }
    // pixel of step i in direction k
    __device__ __forceinline__ void at(int k, int i, int &j1, int &i1) const {
        const int xx = x0 + (k ? -dx0 : dx0) * i, yy = y0 + (k ? -dy0 : dy0) * i;
        j1 = xflag ? xx : xx >> 16;
        i1 = xflag ? yy >> 16 : yy;
    }
};

// index of key in sorted keys[0..N) or -1
__device__ __forceinline__ int find_key(const uint32_t *keys, int N, uint32_t key) {
    int lo = 0, hi = N;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    return (lo < N && keys[lo] == key) ? lo : -1;
}

// ------------------------------------------------------------------------------------------
// Visiting order. OpenCV draws idx = rng % count and swap-removes; an in-place Fisher-Yates over an
// index array leaves the visit sequence in idx[N-1], idx[N-2], ..., idx[0].  It depends only on
// the NUMBER of on-pixels, so it is produced for all frames of the batch up front (one CTA per
// frame, one lane runs the sequential generator in shared memory, the warp writes it out).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
ppht_order_kernel(int T, int cap, const unsigned *__restrict__ npoints, uint16_t *__restrict__ order) {
    extern __shared__ uint16_t o_sm[];
    const int t = blockIdx.x, lane = threadIdx.x;
    const unsigned Nu = npoints[t];
    if (Nu == 0 || Nu > (unsigned)cap) return;
    const int N = (int)Nu;
    for (int i = lane; i < N; i += 32) o_sm[i] = (uint16_t)i;
    __syncwarp();
    if (lane == 0) {
        unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
        for (int count = N; count > 0; count--) {
            state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
            const unsigned r = (unsigned)state % (unsigned)count;
            const uint16_t a = o_sm[r], b = o_sm[count - 1];
            o_sm[r] = b;
            o_sm[count - 1] = a;
        }
    }
    __syncwarp();
    for (int i = lane; i < N; i += 32) order[(size_t)t * cap + i] = o_sm[i];
}

// ------------------------------------------------------------------------------------------
// Tier 1: accumulator in SHARED memory.  Row n only needs the rho interval [min_n, max_n] that the
// frame's points project onto, so the table has sum_n (max_n - min_n + 1) int16 cells (un-voting a
// long line drives its cell to minus the line length, so int8 is not enough); when that fits (a
// streak up to ~800 px long does) the whole PPHT of the frame runs on-chip.  Frames whose table
// would not fit are flagged (-2) for tier 2.
// ------------------------------------------------------------------------------------------
#define HOUGH_TABLE_BYTES (184 * 1024)       // tier 1b: one CTA per SM
#define HOUGH_TABLE_BYTES_SMALL (90 * 1024)  // tier 1a: two CTAs per SM
// per-point shared memory of the tier-1 kernel: key u32, visit order u16, inverse order u16, line pixels u16,
// removed bits by sorted position and by visit position
#define HOUGH1_POINT_BYTES(cap) ((cap) * 10 + (cap) / 4)
#define HOUGH_CAP_SMALL 2048
#define HOUGH_SPEC 4  // points whose votes are taken per barrier in the shared-memory tiers

__global__ void __launch_bounds__(HOUGH_THREADS)
hough_smem_kernel(HoughParams P, int T, const unsigned *__restrict__ npoints,
                  const uint32_t *__restrict__ points, const uint16_t *__restrict__ order,
                  int32_t *lines_out, int *nlines_out, unsigned *queue, long long *prof, int lcap,
                  int table_bytes, int stage) {
    // stage 0 (tier 1a): every frame; up to lcap = 2048 points and a 90 KB table, two CTAs per SM.
    // stage 1 (tier 1b): frames stage 0 flagged -2; up to 4096 points and a 184 KB table (more points: tier 2).
    extern __shared__ uint32_t h_sm[];
    uint32_t *keys = h_sm;                                            // [lcap]
    uint16_t *idx = reinterpret_cast<uint16_t *>(keys + lcap);        // [lcap] visiting order
    uint16_t *wl = idx + lcap;                                        // [lcap] pixels of the current line
    uint16_t *inv = wl + lcap;                                        // [lcap] sorted position -> visit position
    uint32_t *rm = reinterpret_cast<uint32_t *>(inv + lcap);          // [lcap/32] removed bits by sorted position
    uint32_t *rmv = rm + lcap / 32;                                   // [lcap/32] removed bits by visit position
    int16_t *table = reinterpret_cast<int16_t *>(rmv + lcap / 32);    // [table_bytes / 2]
    const int fail_flag = stage == 0 ? -2 : -3;
    __shared__ int s_red[2][HOUGH_THREADS / 32][HOUGH_SPEC];
    __shared__ unsigned s_on[HOUGH_THREADS / 32], s_inb[HOUGH_THREADS / 32];
    __shared__ int s_ctl[8];
    __shared__ int s_base[MDB_HOUGH_ANGLES + 1];
    __shared__ int s_next;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = P.W, H = P.H;
    const float my_c = c_trig[2 * (tid < MDB_HOUGH_ANGLES ? tid : 0)], my_s = c_trig[2 * (tid < MDB_HOUGH_ANGLES ? tid : 0) + 1];

    for (;;) {
        // frames are handed out dynamically: their cost varies by orders of magnitude
        __syncthreads();
        if (tid == 0) s_next = (int)atomicAdd(queue, 1u);
        __syncthreads();
        const int t = s_next;
        if (t >= T) break;
        const unsigned Nu = npoints[t];
        if (stage == 0) {
            if (Nu == 0) { if (tid == 0) nlines_out[t] = 0; continue; }
            if (Nu > (unsigned)P.cap) { if (tid == 0) nlines_out[t] = -1; continue; }  // tier 3 (dense)
            if (Nu > (unsigned)lcap) { if (tid == 0) nlines_out[t] = -2; continue; }    // tier 1b
        } else {
            if (nlines_out[t] != -2) continue;
            if (Nu > (unsigned)lcap) { if (tid == 0) nlines_out[t] = -3; continue; }  // too many points for tier 1b: tier 2
        }
        const int N = (int)Nu;
        const long long pc0 = clock64();
        long long p_setup = 0, p_vote = 0, p_walk = 0, p_unvote = 0, n_vote = 0, n_line = 0;
        const int line_gap = line_gap_of(P, Nu);
        int32_t *lines = lines_out + (size_t)t * P.max_lines * 4;
        int np2 = 1;
        while (np2 < N) np2 <<= 1;
        for (int i = tid; i < np2; i += HOUGH_THREADS)
            keys[i] = i < N ? points[(size_t)t * P.cap + i] : 0xFFFFFFFFu;
        for (int i = tid; i < (N + 31) / 32; i += HOUGH_THREADS) rm[i] = rmv[i] = 0;
        for (int i = tid; i < N; i += HOUGH_THREADS) {
            const uint16_t o = order[(size_t)t * HOUGH_ORDER_CAP + i];
            idx[i] = o;
            inv[o] = (uint16_t)i;
        }
        if (tid == 0) { s_ctl[2] = 0; s_ctl[3] = 0; }
        __syncthreads();
        bitonic_sort_u32(keys, np2, tid, HOUGH_THREADS);
        const long long p_sort = clock64() - pc0;  // loads + sort (profile slot 5 of the shared-memory tiers)
        // per-angle rho interval of this frame's points
        int mn = INT_MAX, mx = INT_MIN;
        if (tid < MDB_HOUGH_ANGLES) {
#pragma unroll 4
            for (int i = 0; i < N; i++) {
                const uint32_t k = keys[i];
                const int r = rho_cs(k & 0xffffu, k >> 16, my_c, my_s);
                mn = min(mn, r); mx = max(mx, r);
            }
            // row length in cells, padded to an ODD number of 32-bit words: thread n's cell address is roughly affine
            // in n (row start + projection), and an even word stride between neighbouring rows would put a whole
            // warp's 32 votes into a handful of shared-memory banks
            s_base[tid + 1] = ((((mx - mn + 1) + 1) >> 1) | 1) << 1;
        }
        if (tid == 0) s_base[0] = 0;
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int k = 1; k <= MDB_HOUGH_ANGLES; k++) { acc += s_base[k]; s_base[k] = acc; }
        }
        __syncthreads();
        int total = s_base[MDB_HOUGH_ANGLES];
        // rho >= gap_b lies gap_skip cells further down the row: rows are ONE interval [mn, mx] by default
        int gap_b = INT_MAX, gap_skip = 0;
        if (total > table_bytes / 2) {
            // Two far-apart objects in one window: per angle the points project onto two clusters with a long empty
            // stretch in between.  Each thread looks for the longest empty run of its row on a 64-bin occupancy mask
            // (thread-private: thread n owns angle n) and leaves it out of the table: the row becomes two intervals.
            __syncthreads();
            if (tid < MDB_HOUGH_ANGLES) {
                const int range = mx - mn + 1;
                unsigned long long occ = 0ull;
#pragma unroll 4
                for (int i = 0; i < N; i++) {
                    const uint32_t k = keys[i];
                    const int r = rho_cs(k & 0xffffu, k >> 16, my_c, my_s);
                    occ |= 1ull << (unsigned)(((long long)(r - mn) * 64) / range);
                }
                // longest run of empty bins between occupied ones (bins 0 and 63 are occupied: they hold mn and mx)
                int best_len = 0, best_start = 0, run = 0;
                for (int bnum = 0; bnum < 64; bnum++) {
                    if ((occ >> bnum) & 1ull) run = 0;
                    else if (++run > best_len) { best_len = run; best_start = bnum - run + 1; }
                }
                if (best_len >= 2) {
                    // a = largest rho whose bin is below the run, b = smallest rho whose bin is at or behind its end
                    const long long gs = best_start, ge = best_start + best_len;
                    const int a = mn + (int)((gs * range + 63) / 64) - 1;
                    const int b = mn + (int)((ge * range + 63) / 64);
                    gap_b = b;
                    gap_skip = b - a - 1;
                }
                s_base[tid + 1] = ((((range - gap_skip) + 1) >> 1) | 1) << 1;
            }
            if (tid == 0) s_base[0] = 0;
            __syncthreads();
            if (tid == 0) {
                int acc = 0;
                for (int k = 1; k <= MDB_HOUGH_ANGLES; k++) { acc += s_base[k]; s_base[k] = acc; }
            }
            __syncthreads();
            total = s_base[MDB_HOUGH_ANGLES];
        }
        if (total > table_bytes / 2) {  // does not fit in this tier's table: next tier
            if (tid == 0) nlines_out[t] = fail_flag;
            __syncthreads();
            continue;
        }
        for (int i = tid; i < (total + 1) / 2; i += HOUGH_THREADS) reinterpret_cast<uint32_t *>(table)[i] = 0;
        int16_t *myrow0 = table + (tid < MDB_HOUGH_ANGLES ? s_base[tid] - mn : 0);
        // cell of rho r in this thread's row
#define H1_CELL(r) (myrow0[(r) - ((r) >= gap_b ? gap_skip : 0)])
        __syncthreads();

        int par = 0;
        bool sat = false;
        p_setup = clock64() - pc0;
        // Votes are taken HOUGH_SPEC points at a time: thread n adds the votes of the next (not yet removed)
        // points to its own row one after the other, reading each cell right after its own increment, so every
        // point's arg-max is the one the sequential algorithm sees; the four arg-max reductions share one
        // barrier.  Lines are rare (a few per frame): when point j of a group yields one, the votes of the points
        // after it are taken back and the scan resumes right behind j, after the line has been removed.
        int s = N - 1;
        for (;;) {
            long long c0 = clock64();
            int cs[HOUGH_SPEC], cnt = 0;
            uint32_t ck[HOUGH_SPEC];
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) { cs[j] = -1; ck[j] = 0; }
            while (cnt < HOUGH_SPEC && s >= 0) {  // uniform: every thread walks the same bit list
                // not yet removed visit positions <= s of this 32-position word, highest first; one load serves
                // up to HOUGH_SPEC candidates
                const int wbase = s & ~31;
                unsigned m = ~rmv[s >> 5] & (0xffffffffu >> (31 - (s & 31)));
                while (m && cnt < HOUGH_SPEC) {
                    const int b = 31 - __clz(m);
                    m ^= 1u << b;
#pragma unroll
                    for (int j = 0; j < HOUGH_SPEC; j++)
                        if (j == cnt) cs[j] = wbase + b;
                    cnt++;
                    s = wbase + b - 1;
                }
                if (!m) s = min(s, wbase - 1);
            }
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++)
                if (j < cnt) ck[j] = keys[idx[cs[j]]];
            if (cnt == 0) break;
            int bj[HOUGH_SPEC], rj[HOUGH_SPEC];
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) { bj[j] = INT_MIN; rj[j] = 0; }
            if (tid < MDB_HOUGH_ANGLES) {
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) {
                        const int r = rho_cs(ck[j] & 0xffffu, ck[j] >> 16, my_c, my_s);
                        const int v = (int)H1_CELL(r) + 1;
                        H1_CELL(r) = (int16_t)v;
                        sat |= v >= 32767;
                        bj[j] = v * 256 + (255 - tid);  // max value first, lowest angle on ties
                        rj[j] = r;
                    }
            }
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) {
                bj[j] = __reduce_max_sync(0xffffffffu, bj[j]);
                if (lane == 0) s_red[par][warp][j] = bj[j];
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++)
#pragma unroll
                for (int k = 0; k < HOUGH_THREADS / 32; k++) bj[j] = max(bj[j], s_red[par][k][j]);
            par ^= 1;
            int trig = -1;
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++)
                if (trig < 0 && j < cnt && (bj[j] >> 8) >= P.threshold) trig = j;
            p_vote += clock64() - c0; n_vote += trig < 0 ? cnt : trig + 1;
            if (trig < 0) continue;
            c0 = clock64(); n_line++;
            uint32_t key = 0;
            int best = 0;
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) {
                if (j == trig) { key = ck[j]; best = bj[j]; s = cs[j] - 1; }
                if (j > trig && j < cnt && tid < MDB_HOUGH_ANGLES) H1_CELL(rj[j]) = (int16_t)((int)H1_CELL(rj[j]) - 1);
            }
            const int x = key & 0xffffu, y = key >> 16;
            const int max_n = 255 - (best & 255);

            // ---- line: find both ends (mask unchanged meanwhile), 256 steps per round ----------
            Walk wk;
            wk.init(x, y, max_n);
            int ends[2];
            for (int k = 0; k < 2; k++) {
                int last_on = 0;
                bool done = false;
                for (int base = 0; !done; base += HOUGH_THREADS) {
                    const int i = base + tid;
                    int j1, i1;
                    wk.at(k, i, j1, i1);
                    const bool inb = j1 >= 0 && j1 < W && i1 >= 0 && i1 < H;
                    bool on = false;
                    if (inb) {
                        const int f = find_key(keys, N, ((unsigned)i1 << 16) | (unsigned)j1);
                        on = f >= 0 && !((rm[f >> 5] >> (f & 31)) & 1u);
                    }
                    const unsigned bo = __ballot_sync(0xffffffffu, on), bi = __ballot_sync(0xffffffffu, inb);
                    if (lane == 0) { s_on[warp] = bo; s_inb[warp] = bi; }
                    __syncthreads();
                    for (int w = 0; w < HOUGH_THREADS / 32 && !done; w++) {
                        unsigned onw = s_on[w];
                        const unsigned inw = s_inb[w];
                        const int wbase = base + w * 32;
                        const int ob = inw == 0xffffffffu ? INT_MAX : wbase + __ffs(~inw) - 1;
                        if (onw == 0xffffffffu && wbase - last_on - 1 <= line_gap) {  // solid run: no per-bit work
                            last_on = wbase + 31;
                            continue;
                        }
                        while (onw) {
                            const int io = wbase + __ffs(onw) - 1;
                            onw &= onw - 1;
                            if (io - last_on - 1 > line_gap) { done = true; break; }
                            last_on = io;
                        }
                        if (!done && (ob != INT_MAX || wbase + 31 - last_on > line_gap)) done = true;
                    }
                    __syncthreads();
                }
                ends[k] = last_on;
            }
            int ex0, ey0, ex1, ey1;
            wk.at(0, ends[0], ex0, ey0);
            wk.at(1, ends[1], ex1, ey1);
            const bool good = abs(ex1 - ex0) >= P.min_len || abs(ey1 - ey0) >= P.min_len;
            if (tid == 0) s_ctl[1] = 0;
            __syncthreads();
            p_walk += clock64() - c0; c0 = clock64();
            for (int k = 0; k < 2; k++)
                for (int i = tid + k; i <= ends[k]; i += HOUGH_THREADS) {  // k=1 skips the shared start pixel
                    int j1, i1;
                    wk.at(k, i, j1, i1);
                    const int f = find_key(keys, N, ((unsigned)i1 << 16) | (unsigned)j1);
                    if (f >= 0 && !((rm[f >> 5] >> (f & 31)) & 1u)) {
                        atomicOr(&rm[f >> 5], 1u << (f & 31));
                        const int sv = inv[f];
                        atomicOr(&rmv[sv >> 5], 1u << (sv & 31));
                        if (good) wl[atomicAdd(&s_ctl[1], 1)] = (uint16_t)f;
                    }
                }
            __syncthreads();
            if (good) {
                const int nw = s_ctl[1];
                if (tid < MDB_HOUGH_ANGLES) {
                    // four pixels per round: the four cell reads are issued together; equal cells inside
                    // a round are forwarded through registers so the result equals the sequential order
                    int q = 0;
                    for (; q + 4 <= nw; q += 4) {
                        int r[4], v[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t k2 = keys[wl[q + u]];
                            r[u] = rho_cs(k2 & 0xffffu, k2 >> 16, my_c, my_s);
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) v[u] = (int)H1_CELL(r[u]);
                        v[0] -= 1;
                        v[1] = (r[1] == r[0] ? v[0] : v[1]) - 1;
                        v[2] = (r[2] == r[1] ? v[1] : (r[2] == r[0] ? v[0] : v[2])) - 1;
                        v[3] = (r[3] == r[2] ? v[2] : (r[3] == r[1] ? v[1] : (r[3] == r[0] ? v[0] : v[3]))) - 1;
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            H1_CELL(r[u]) = (int16_t)v[u];
                            sat |= v[u] <= -32768;
                        }
                    }
                    for (; q < nw; q++) {
                        const uint32_t k2 = keys[wl[q]];
                        const int rc = rho_cs(k2 & 0xffffu, k2 >> 16, my_c, my_s);
                        const int v1 = (int)H1_CELL(rc) - 1;
                        H1_CELL(rc) = (int16_t)v1;
                        sat |= v1 <= -32768;
                    }
                }
                if (tid == 0) {
                    const int li = s_ctl[2];
                    if (li < P.max_lines) {
                        lines[4 * li] = ex0; lines[4 * li + 1] = ey0;
                        lines[4 * li + 2] = ex1; lines[4 * li + 3] = ey1;
                    }
                    s_ctl[2] = li + 1;
                }
            }
            __syncthreads();
            p_unvote += clock64() - c0;
        }
        if (prof && tid == 0) {
            long long *o = prof + (size_t)t * 10;
            o[0] = N; o[1] = p_setup; o[2] = p_vote; o[3] = p_walk; o[4] = p_unvote; o[5] = p_sort;
            o[6] = n_vote; o[7] = n_line; o[8] = clock64() - pc0; o[9] = s_ctl[2];
        }
        if (sat) s_ctl[3] = 1;
        __syncthreads();
        if (tid == 0) {
            nlines_out[t] = s_ctl[3] ? -3 : s_ctl[2];  // a saturated cell (impossible for N <= cap): tier 2
            if (!s_ctl[3]) atomicAdd(queue + 4, 1u);  // frames this tier resolved (statistics)
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
#undef H1_CELL
// Tier 2: same algorithm with the accumulator in global memory ([180][numrho] int32 per slot)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HOUGH_THREADS)
hough_tier2_kernel(HoughParams P, int T, const unsigned *__restrict__ npoints,
                   const uint32_t *__restrict__ points, int32_t *accum_slots, int32_t *lines_out,
                   int *nlines_out, long long *prof, unsigned *tier_count) {
    extern __shared__ uint32_t h_sm[];
    uint32_t *keys = h_sm;                                            // [cap]
    uint16_t *idx = reinterpret_cast<uint16_t *>(keys + P.cap);       // [cap] visiting order
    uint16_t *wl = idx + P.cap;                                       // [cap] pixels of the current line
    uint32_t *rm = reinterpret_cast<uint32_t *>(wl + P.cap);          // [cap/32] removed bits
    __shared__ int s_red[2][HOUGH_THREADS / 32][HOUGH_SPEC];
    __shared__ unsigned s_on[HOUGH_THREADS / 32], s_inb[HOUGH_THREADS / 32];
    __shared__ int s_ctl[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = P.W, H = P.H, numrho = P.numrho, half = (numrho - 1) / 2;
    int32_t *accum = accum_slots + (size_t)blockIdx.x * MDB_HOUGH_ANGLES * numrho;
    int32_t *myrow = accum + (size_t)(tid < MDB_HOUGH_ANGLES ? tid : 0) * numrho + half;
    const float my_c = c_trig[2 * (tid < MDB_HOUGH_ANGLES ? tid : 0)], my_s = c_trig[2 * (tid < MDB_HOUGH_ANGLES ? tid : 0) + 1];

    for (int t = blockIdx.x; t < T; t += gridDim.x) {
        const unsigned Nu = npoints[t];
        if (nlines_out[t] != -3) continue;  // only frames the shared-memory tiers handed over
        __syncthreads();
        const int N = (int)Nu;
        long long pc0 = clock64(), p_setup = 0, p_vote = 0, p_walk = 0, p_unvote = 0, p_reset = 0, n_vote = 0, n_line = 0;
        const int line_gap = line_gap_of(P, Nu);
        int32_t *lines = lines_out + (size_t)t * P.max_lines * 4;
        int np2 = 1;
        while (np2 < N) np2 <<= 1;
        for (int i = tid; i < np2; i += HOUGH_THREADS)
            keys[i] = i < N ? points[(size_t)t * P.cap + i] : 0xFFFFFFFFu;
        for (int i = tid; i < (N + 31) / 32; i += HOUGH_THREADS) rm[i] = 0;
        for (int i = tid; i < N; i += HOUGH_THREADS) idx[i] = (uint16_t)i;
        __syncthreads();
        bitonic_sort_u32(keys, np2, tid, HOUGH_THREADS);
        // visiting order: OpenCV draws idx = rng % count and swap-removes; an in-place Fisher-Yates
        // over an index array leaves the visit sequence in idx[N-1], idx[N-2], ..., idx[0].
        if (tid == 0) {
            unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
            for (int count = N; count > 0; count--) {
                state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
                const unsigned r = (unsigned)state % (unsigned)count;
                const uint16_t a = idx[r], b = idx[count - 1];
                idx[r] = b;
                idx[count - 1] = a;
            }
            s_ctl[2] = 0;  // lines found
        }
        __syncthreads();
        // warm the first visits' cells
        if (tid < MDB_HOUGH_ANGLES)
            for (int s = N - 1; s >= 0 && s >= N - HOUGH_PREFETCH; s--) {
                const uint32_t k = keys[idx[s]];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(myrow + rho_cs(k & 0xffffu, k >> 16, my_c, my_s)));
            }
        int par = 0;
        p_setup = clock64() - pc0;
        // votes are taken HOUGH_SPEC points per barrier as in the shared-memory tiers: here it is the L2 round trips
        // of the four cell reads that overlap (the cells were prefetched a few visits ahead)
        int s = N - 1;
        for (;;) {
            long long c0 = clock64();
            int cs[HOUGH_SPEC], cnt = 0;
            uint32_t ck[HOUGH_SPEC];
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) { cs[j] = -1; ck[j] = 0; }
            while (cnt < HOUGH_SPEC && s >= 0) {  // uniform: every thread walks the same list
                const int pi = idx[s];
                if (tid < MDB_HOUGH_ANGLES && s >= HOUGH_PREFETCH) {
                    const uint32_t k = keys[idx[s - HOUGH_PREFETCH]];
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(myrow + rho_cs(k & 0xffffu, k >> 16, my_c, my_s)));
                }
                if (!((rm[pi >> 5] >> (pi & 31)) & 1u)) {
                    const uint32_t kk = keys[pi];
#pragma unroll
                    for (int j = 0; j < HOUGH_SPEC; j++)
                        if (j == cnt) { cs[j] = s; ck[j] = kk; }
                    cnt++;
                }
                s--;
            }
            if (cnt == 0) break;
            int bj[HOUGH_SPEC], rj[HOUGH_SPEC], vj[HOUGH_SPEC];
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) { bj[j] = INT_MIN; rj[j] = 0; vj[j] = 0; }
            if (tid < MDB_HOUGH_ANGLES) {
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) rj[j] = rho_cs(ck[j] & 0xffffu, ck[j] >> 16, my_c, my_s);
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) vj[j] = __ldcg(myrow + rj[j]);  // L2 only: cells are also updated by REDs
                // the four reads are issued together; equal cells inside the group are forwarded through
                // registers so that every point sees the count the sequential order gives it
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) {
                        int base = vj[j];
#pragma unroll
                        for (int i = 0; i < j; i++)
                            if (rj[i] == rj[j]) base = vj[i];  // the latest earlier vote for the same cell wins
                        vj[j] = base + 1;
                        bj[j] = vj[j] * 256 + (255 - tid);  // max value first, lowest angle on ties
                    }
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) __stcg(myrow + rj[j], vj[j]);
            }
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) {
                bj[j] = __reduce_max_sync(0xffffffffu, bj[j]);
                if (lane == 0) s_red[par][warp][j] = bj[j];
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++)
#pragma unroll
                for (int k = 0; k < HOUGH_THREADS / 32; k++) bj[j] = max(bj[j], s_red[par][k][j]);
            par ^= 1;
            int trig = -1;
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++)
                if (trig < 0 && j < cnt && (bj[j] >> 8) >= P.threshold) trig = j;
            p_vote += clock64() - c0; n_vote += trig < 0 ? cnt : trig + 1;
            if (trig < 0) continue;
            c0 = clock64(); n_line++;
            uint32_t key = 0;
            int best = 0;
#pragma unroll
            for (int j = 0; j < HOUGH_SPEC; j++) {
                if (j == trig) { key = ck[j]; best = bj[j]; s = cs[j] - 1; }
                // take back the votes of the points behind the one that yielded a line
                if (j > trig && j < cnt && tid < MDB_HOUGH_ANGLES) __stcg(myrow + rj[j], __ldcg(myrow + rj[j]) - 1);
            }
            const int x = key & 0xffffu, y = key >> 16;
            const int max_n = 255 - (best & 255);

            // ---- line: find both ends (mask unchanged meanwhile), 256 steps per round ----------
            Walk wk;
            wk.init(x, y, max_n);
            int ends[2];  // step index of the last on-pixel per direction
            for (int k = 0; k < 2; k++) {
                int last_on = 0;  // step 0 is the start pixel, which is on
                bool done = false;
                for (int base = 0; !done; base += HOUGH_THREADS) {
                    const int i = base + tid;
                    int j1, i1;
                    wk.at(k, i, j1, i1);
                    const bool inb = j1 >= 0 && j1 < W && i1 >= 0 && i1 < H;
                    bool on = false;
                    if (inb) {
                        const int f = find_key(keys, N, ((unsigned)i1 << 16) | (unsigned)j1);
                        on = f >= 0 && !((rm[f >> 5] >> (f & 31)) & 1u);
                    }
                    const unsigned bo = __ballot_sync(0xffffffffu, on), bi = __ballot_sync(0xffffffffu, inb);
                    if (lane == 0) { s_on[warp] = bo; s_inb[warp] = bi; }
                    __syncthreads();
                    // every thread scans the 8 ballot words identically (uniform, no divergence)
                    for (int w = 0; w < HOUGH_THREADS / 32 && !done; w++) {
                        unsigned onw = s_on[w];
                        const unsigned inw = s_inb[w];
                        const int wbase = base + w * 32;
                        // first out-of-bounds step in this word (bounds are monotone along the walk)
                        const int ob = inw == 0xffffffffu ? INT_MAX : wbase + __ffs(~inw) - 1;
                        while (onw) {
                            const int io = wbase + __ffs(onw) - 1;
                            onw &= onw - 1;
                            if (io - last_on - 1 > line_gap) { done = true; break; }
                            last_on = io;
                        }
                        if (!done) {
                            const int wend = min(wbase + 31, ob == INT_MAX ? INT_MAX : ob - 1);  // last in-bounds step seen
                            if (ob != INT_MAX || wend - last_on > line_gap) done = true;
                        }
                    }
                    __syncthreads();
                }
                ends[k] = last_on;
            }
            int ex0, ey0, ex1, ey1;
            wk.at(0, ends[0], ex0, ey0);
            wk.at(1, ends[1], ex1, ey1);
            const bool good = abs(ex1 - ex0) >= P.min_len || abs(ey1 - ey0) >= P.min_len;
            // ---- second pass: clear the on-pixels up to both ends; un-vote them if the line counts
            if (tid == 0) s_ctl[1] = 0;
            __syncthreads();
            p_walk += clock64() - c0; c0 = clock64();
            for (int k = 0; k < 2; k++)
                for (int i = tid + k; i <= ends[k]; i += HOUGH_THREADS) {  // k=1 skips the shared start pixel
                    int j1, i1;
                    wk.at(k, i, j1, i1);
                    const int f = find_key(keys, N, ((unsigned)i1 << 16) | (unsigned)j1);
                    if (f >= 0 && !((rm[f >> 5] >> (f & 31)) & 1u)) {
                        atomicOr(&rm[f >> 5], 1u << (f & 31));
                        if (good) wl[atomicAdd(&s_ctl[1], 1)] = (uint16_t)f;
                    }
                }
            __syncthreads();
            if (good) {
                const int nw = s_ctl[1];
                if (tid < MDB_HOUGH_ANGLES)
                    for (int q = 0; q < nw; q++) {
                        const uint32_t k2 = keys[wl[q]];
                        atomicAdd(myrow + rho_cs(k2 & 0xffffu, k2 >> 16, my_c, my_s), -1);  // RED, no return
                    }
                if (tid == 0) {
                    const int li = s_ctl[2];
                    if (li < P.max_lines) {
                        lines[4 * li] = ex0; lines[4 * li + 1] = ey0;
                        lines[4 * li + 2] = ex1; lines[4 * li + 3] = ey1;
                    }
                    s_ctl[2] = li + 1;
                }
            }
            __syncthreads();
            p_unvote += clock64() - c0;
        }
        long long c1 = clock64();
        // reset: zero every accumulator cell a point of this frame can have touched
        // Row n was only touched inside the rho interval its points project onto: find the intervals (thread n
        // owns row n, cos/sin in registers) and clear them with coalesced stores -- N * 180 scattered 4-byte
        // stores from one SM are bound by its load/store transaction rate (1.3 M cycles for a 6.5 k-point frame).
        {
            float vmin = 3.0e38f, vmax = -3.0e38f;
            if (tid < MDB_HOUGH_ANGLES) {
#pragma unroll 4
                for (int i = 0; i < N; i++) {
                    const uint32_t k = keys[i];
                    const float v = __fadd_rn(__fmul_rn(u16_to_float(k & 0xffffu), my_c), __fmul_rn(u16_to_float(k >> 16), my_s));
                    vmin = fminf(vmin, v);
                    vmax = fmaxf(vmax, v);
                }
            }
            // rounding is monotone: the rounded extremes bound every rounded projection
            int *s_lo = reinterpret_cast<int *>(wl), *s_hi = s_lo + MDB_HOUGH_ANGLES;  // wl is free now (>= 2 * 180 ints)
            __syncthreads();
            if (tid < MDB_HOUGH_ANGLES) {
                s_lo[tid] = __float_as_int(__fadd_rn(vmin, 12582912.0f)) - 0x4B400000;
                s_hi[tid] = __float_as_int(__fadd_rn(vmax, 12582912.0f)) - 0x4B400000;
            }
            __syncthreads();
            for (int n = warp; n < MDB_HOUGH_ANGLES; n += HOUGH_THREADS / 32) {
                int32_t *row = accum + (size_t)n * numrho + half;
                for (int r = s_lo[n] + lane; r <= s_hi[n]; r += 32) __stcg(row + r, 0);
            }
        }
        __syncthreads();
        p_reset = clock64() - c1;
        if (tid == 0) { nlines_out[t] = s_ctl[2]; atomicAdd(tier_count, 1u); }
        if (prof && tid == 0) {
            long long *o = prof + (size_t)t * 10;
            o[0] = N; o[1] = p_setup; o[2] = p_vote; o[3] = p_walk; o[4] = p_unvote; o[5] = p_reset;
            o[6] = n_vote; o[7] = n_line; o[8] = clock64() - pc0; o[9] = s_ctl[2];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Tier 3: dense masks beyond the point capacity of tier 2 (sensor noise above the threshold, clouds, dawn: tens of
// thousands of on-pixels).  Same exact-order algorithm with tier 2's parallel machinery -- votes of four points per
// barrier with the cell reads in flight together, 256-step ballot walks, fire-and-forget un-votes -- but nothing
// frame-sized lives in shared memory: the row-major point list and the visiting order are in global memory and are
// staged 256 visits at a time, "is this pixel still on" is a per-slot pixel bitmap in global memory (L2).  One CTA per
// scratch slot; the CTAs claim the dense frames of the batch from a queue, so up to `slots3` frames are worked on side
// by side.  Entirely on the device (no host round trip: batches stay pipelined).
// ------------------------------------------------------------------------------------------
#define H3_STAGE 256          // visits staged per round (= HOUGH_THREADS)
#define H3_ORDER_CAP 98304    // most points whose visiting order is shuffled in shared memory (u16 + one high bit each)
#define H3_ORDER_SMEM (H3_ORDER_CAP * 2 + H3_ORDER_CAP / 8)

// Row-major list of the on-pixels of one mask: per-row counts, a scan over the rows, then every warp writes its rows in
// order.  row_off: global scratch of at least H + 1 words.  Returns the number of points.
__device__ int compact_rows(const uint8_t *dst, int W, int H, uint32_t *keys, uint32_t *row_off) {
    __shared__ unsigned c_part[HOUGH_THREADS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool vec = (W & 15) == 0 && ((uintptr_t)dst & 15) == 0;
    // pass 1: on-pixels per row
    for (int y = warp; y < H; y += HOUGH_THREADS / 32) {
        const uint8_t *row = dst + (size_t)y * W;
        unsigned c = 0;
        if (vec) {
            for (int x = lane * 16; x < W; x += 512) {
                const uint4 v = *reinterpret_cast<const uint4 *>(row + x);
                // mask bytes are 0 or 255: the top bit of every byte
                c += __popc(v.x & 0x80808080u) + __popc(v.y & 0x80808080u) + __popc(v.z & 0x80808080u) + __popc(v.w & 0x80808080u);
            }
        } else {
            for (int x = lane; x < W; x += 32) c += row[x] != 0;
        }
        c = __reduce_add_sync(0xffffffffu, c);
        if (lane == 0) row_off[y] = c;
    }
    __syncthreads();
    // exclusive scan over the rows: every thread owns a contiguous range of rows
    const int per = (H + HOUGH_THREADS - 1) / HOUGH_THREADS;
    const int y0 = min(tid * per, H), y1 = min(y0 + per, H);
    unsigned sum = 0;
    for (int y = y0; y < y1; y++) sum += row_off[y];
    c_part[tid] = sum;
    __syncthreads();
    unsigned base = 0, total = 0;
    for (int k = 0; k < HOUGH_THREADS; k++) {
        if (k < tid) base += c_part[k];
        total += c_part[k];
    }
    for (int y = y0; y < y1; y++) {
        const unsigned c = row_off[y];
        row_off[y] = base;
        base += c;
    }
    __syncthreads();
    // pass 2: write the keys of every row in x order
    for (int y = warp; y < H; y += HOUGH_THREADS / 32) {
        const uint8_t *row = dst + (size_t)y * W;
        unsigned off = row_off[y];
        const int step = vec ? 512 : 32;
        for (int xb = 0; xb < W; xb += step) {
            unsigned bits = 0;  // this lane's on-pixels of the chunk (16 consecutive pixels, or one)
            if (vec) {
                const int x = xb + lane * 16;
                if (x < W) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(row + x);
                    const unsigned w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; q++)
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (w4[q] >> (8 * b + 7) & 1u) bits |= 1u << (4 * q + b);
                }
            } else {
                const int x = xb + lane;
                if (x < W && row[x]) bits = 1u;
            }
            const unsigned c = __popc(bits);
            unsigned inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            unsigned w = off + inc - c;
            const int x0 = vec ? xb + lane * 16 : xb + lane;
            while (bits) {
                const int b = __ffs(bits) - 1;
                bits &= bits - 1;
                keys[w++] = ((unsigned)y << 16) | (unsigned)(x0 + b);
            }
            off += __shfl_sync(0xffffffffu, inc, 31);
        }
    }
    __syncthreads();
    return (int)total;
}

__global__ void __launch_bounds__(HOUGH_THREADS)
hough_tier3_kernel(HoughParams P, int T, const uint8_t *dst, uint32_t *keys, uint32_t *idx, int32_t *accum,
                   uint32_t *bitmap, uint32_t *walk, int32_t *lines_all, int *nlines_all, unsigned *queue, long long *prof) {
    extern __shared__ uint32_t h_sm[];  // order shuffle: u16 [H3_ORDER_CAP] + high bits
    __shared__ int s_red[2][HOUGH_THREADS / 32][HOUGH_SPEC];
    __shared__ unsigned s_on[HOUGH_THREADS / 32], s_inb[HOUGH_THREADS / 32];
    __shared__ int s_ctl[8];
    __shared__ int s_next, s_any;
    __shared__ uint32_t st_key[H3_STAGE];
    __shared__ unsigned st_alive[H3_STAGE / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = P.W, H = P.H, numrho = P.numrho, half = (numrho - 1) / 2;
    {
        const size_t HWs = (size_t)W * H;
        keys += (size_t)blockIdx.x * HWs;
        idx += (size_t)blockIdx.x * HWs;
        accum += (size_t)blockIdx.x * MDB_HOUGH_ANGLES * numrho;
        bitmap += (size_t)blockIdx.x * ((HWs + 31) / 32);
        walk += (size_t)blockIdx.x * P.walk_cap;
    }
    const float my_c = c_trig[2 * (tid < MDB_HOUGH_ANGLES ? tid : 0)], my_s = c_trig[2 * (tid < MDB_HOUGH_ANGLES ? tid : 0) + 1];
    int32_t *myrow = accum + (size_t)(tid < MDB_HOUGH_ANGLES ? tid : 0) * numrho + half;
    if (tid == 0) s_any = 0;
    __syncthreads();
    for (int t = tid; t < T; t += HOUGH_THREADS)
        if (nlines_all[t] == -1) s_any = 1;
    __syncthreads();
    if (!s_any) return;  // the usual case: no dense frame in this batch
    for (;;) {
        __syncthreads();
        if (tid == 0) s_next = (int)atomicAdd(queue, 1u);
        __syncthreads();
        const int t = s_next;
        if (t >= T) break;
        if (nlines_all[t] != -1) continue;  // (only the CTA that claimed t ever writes nlines_all[t])
        int32_t *lines = lines_all + (size_t)t * P.max_lines * 4;
        const long long pc0 = clock64();
        long long p_setup = 0, p_vote = 0, p_walk = 0, p_unvote = 0, p_stage = 0, n_vote = 0, n_line = 0, n_iso = 0;
        const int N = compact_rows(dst + (size_t)t * W * H, W, H, keys, walk);
        if (N == 0) { if (tid == 0) nlines_all[t] = 0; continue; }
        const int line_gap = line_gap_of(P, (unsigned)N);
        for (int i = tid; i < N; i += HOUGH_THREADS) {
            const uint32_t k = keys[i];
            const size_t p = (size_t)(k >> 16) * W + (k & 0xffffu);
            atomicOr(&bitmap[p >> 5], 1u << (p & 31));
        }
        // ---- visiting order: Fisher-Yates with OpenCV's RNG; idx[N-1], idx[N-2], ... is the visit sequence ----------
        if (N <= H3_ORDER_CAP) {
            uint16_t *lo = reinterpret_cast<uint16_t *>(h_sm);
            uint32_t *hi = h_sm + H3_ORDER_CAP / 2;  // bit i = bit 16 of entry i
            for (int i = tid; i < N; i += HOUGH_THREADS) lo[i] = (uint16_t)i;
            for (int i = tid; i < (N + 31) / 32; i += HOUGH_THREADS) {
                // entries 65536 .. N-1 start with their high bit set
                unsigned m = 0;
                if (i * 32 >= 65536) m = 0xffffffffu;
                hi[i] = m;
            }
            __syncthreads();
            if (tid == 0) {
                unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
                for (int count = N; count > 0; count--) {
                    state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
                    const unsigned r = (unsigned)state % (unsigned)count, c = (unsigned)count - 1;
                    const uint16_t a = lo[r], b = lo[c];
                    lo[r] = b; lo[c] = a;
                    if (N > 65536) {
                        const unsigned ha = (hi[r >> 5] >> (r & 31)) & 1u, hb = (hi[c >> 5] >> (c & 31)) & 1u;
                        if (ha != hb) { hi[r >> 5] ^= 1u << (r & 31); hi[c >> 5] ^= 1u << (c & 31); }
                    }
                }
            }
            __syncthreads();
            for (int i = tid; i < N; i += HOUGH_THREADS) idx[i] = (uint32_t)lo[i] | (((hi[i >> 5] >> (i & 31)) & 1u) << 16);
        } else {
            for (int i = tid; i < N; i += HOUGH_THREADS) idx[i] = i;
            __syncthreads();
            if (tid == 0) {
                unsigned long long state = 0xFFFFFFFFFFFFFFFFull;
                for (int count = N; count > 0; count--) {
                    state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
                    const unsigned r = (unsigned)state % (unsigned)count;
                    const uint32_t a = idx[r], b = idx[count - 1];
                    idx[r] = b;
                    idx[count - 1] = a;
                }
            }
        }
        if (tid == 0) s_ctl[2] = 0;
        __syncthreads();
        __threadfence_block();

        int par = 0;
        bool hot = false;  // the last point looked at yielded a line
        p_setup = clock64() - pc0;
        // ---- visits, H3_STAGE at a time -----------------------------------------------------------------------
        for (int top = N - 1; top >= 0; top -= H3_STAGE) {
            const int cnt_stage = min(H3_STAGE, top + 1);
            long long c0 = clock64();
            // stage: keys of visits top, top-1, ... and whether their pixels are still on
            auto refresh = [&](int from) {  // alive bits of staged positions >= from (others cleared)
                bool on = false;
                if (tid >= from && tid < cnt_stage) {
                    const uint32_t k = st_key[tid];
                    const size_t p = (size_t)(k >> 16) * W + (k & 0xffffu);
                    on = (__ldcg(&bitmap[p >> 5]) >> (p & 31)) & 1u;
                }
                const unsigned b = __ballot_sync(0xffffffffu, on);
                if (lane == 0) st_alive[warp] = b;
            };
            if (tid < cnt_stage) st_key[tid] = keys[idx[top - tid]];
            __syncthreads();
            refresh(0);
            // warm the cells of the staged points (L2)
            if (tid < MDB_HOUGH_ANGLES)
                for (int j = 0; j < cnt_stage; j++) {
                    const uint32_t k = st_key[j];
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(myrow + rho_cs(k & 0xffffu, k >> 16, my_c, my_s)));
                }
            __syncthreads();
            int j0 = 0;  // next staged position to look at
            p_stage += clock64() - c0;
            for (;;) {
                c0 = clock64();
                // the next (up to HOUGH_SPEC) staged points whose pixels are still on: uniform scan of the alive bits.
                // Right after a point that yielded a line only one is taken: in saturated accumulators (dense noise)
                // nearly every point yields one, and votes taken ahead would only have to be taken back.
                const int spec = hot ? 1 : HOUGH_SPEC;
                hot = false;
                int cs[HOUGH_SPEC], cnt = 0;
                uint32_t ck[HOUGH_SPEC];
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++) { cs[j] = -1; ck[j] = 0; }
                while (cnt < spec && j0 < cnt_stage) {
                    unsigned m = st_alive[j0 >> 5] & (0xffffffffu << (j0 & 31));
                    const int wbase = j0 & ~31;
                    while (m && cnt < spec) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1;
#pragma unroll
                        for (int j = 0; j < HOUGH_SPEC; j++)
                            if (j == cnt) cs[j] = wbase + b;
                        cnt++;
                        j0 = wbase + b + 1;
                    }
                    if (!m && cnt < spec) j0 = max(j0, wbase + 32);
                }
                if (cnt == 0) break;
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) ck[j] = st_key[cs[j]];
                int bj[HOUGH_SPEC], rj[HOUGH_SPEC], vj[HOUGH_SPEC];
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++) { bj[j] = INT_MIN; rj[j] = 0; vj[j] = 0; }
                if (tid < MDB_HOUGH_ANGLES) {
#pragma unroll
                    for (int j = 0; j < HOUGH_SPEC; j++)
                        if (j < cnt) rj[j] = rho_cs(ck[j] & 0xffffu, ck[j] >> 16, my_c, my_s);
#pragma unroll
                    for (int j = 0; j < HOUGH_SPEC; j++)
                        if (j < cnt) vj[j] = __ldcg(myrow + rj[j]);  // L2 only: cells are also updated by REDs
                    // maxLineGap 0 (every mask this dense): a walk ends at the first off pixel, so a point whose two
                    // neighbours along the winning angle are off is an isolated pixel.  Thread n looks at the neighbours
                    // along ITS angle while its cell read is in flight; the flag rides in the low bit of the arg-max key
                    // and the usual outcome in noise -- isolated -- costs no further round trip.
                    unsigned nbj = 0xfu;
                    if (line_gap == 0) {
                        nbj = 0u;
#pragma unroll
                        for (int j = 0; j < HOUGH_SPEC; j++)
                            if (j < cnt) {
                                Walk wn;
                                wn.init(ck[j] & 0xffffu, ck[j] >> 16, tid);
#pragma unroll
                                for (int k = 0; k < 2; k++) {
                                    int j1, i1;
                                    wn.at(k, 1, j1, i1);
                                    if (j1 >= 0 && j1 < W && i1 >= 0 && i1 < H) {
                                        const size_t q = (size_t)i1 * W + j1;
                                        nbj |= ((__ldcg(&bitmap[q >> 5]) >> (q & 31)) & 1u) << j;
                                    }
                                }
                            }
                    }
#pragma unroll
                    for (int j = 0; j < HOUGH_SPEC; j++)
                        if (j < cnt) {
                            int base = vj[j];
#pragma unroll
                            for (int i = 0; i < j; i++)
                                if (rj[i] == rj[j]) base = vj[i];  // the latest earlier vote for the same cell wins
                            vj[j] = base + 1;
                            // max value first, lowest angle on ties; bit 0: this angle's walk would find a neighbour
                            bj[j] = (vj[j] * 256 + (255 - tid)) * 2 + (int)((nbj >> j) & 1u);
                        }
#pragma unroll
                    for (int j = 0; j < HOUGH_SPEC; j++)
                        if (j < cnt) __stcg(myrow + rj[j], vj[j]);
                }
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) {  // uniform
                        bj[j] = __reduce_max_sync(0xffffffffu, bj[j]);
                        if (lane == 0) s_red[par][warp][j] = bj[j];
                    }
                __syncthreads();
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (j < cnt) {
#pragma unroll
                        for (int k = 0; k < HOUGH_THREADS / 32; k++) bj[j] = max(bj[j], s_red[par][k][j]);
                    }
                par ^= 1;
                int trig = -1;
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++)
                    if (trig < 0 && j < cnt && (bj[j] >> 9) >= P.threshold) trig = j;
                p_vote += clock64() - c0; n_vote += trig < 0 ? cnt : trig + 1;
                if (trig < 0) continue;
                c0 = clock64(); n_line++;
                uint32_t key = 0;
                int best = 0;
                bool lonely = false;
#pragma unroll
                for (int j = 0; j < HOUGH_SPEC; j++) {
                    if (j == trig) { key = ck[j]; best = bj[j] >> 1; lonely = !(bj[j] & 1); j0 = cs[j] + 1; }
                    // take back the votes of the points behind the one that yielded a line
                    if (j > trig && j < cnt && tid < MDB_HOUGH_ANGLES) __stcg(myrow + rj[j], __ldcg(myrow + rj[j]) - 1);
                }
                const int x = key & 0xffffu, y = key >> 16;
                const int max_n = 255 - (best & 255);
                if (lonely) {
                    // an isolated pixel (known from the neighbour flag): only the start pixel leaves the mask
                    if (tid == 0) {
                        const size_t q = (size_t)y * W + x;
                        atomicAnd(&bitmap[q >> 5], ~(1u << (q & 31)));
                    }
                    hot = true;
                    p_walk += clock64() - c0; n_iso++;
                    continue;  // the next vote round's barrier orders the bitmap update before any later probe
                }
                // ---- line: find both ends (mask unchanged meanwhile) -----------------------------------------
                Walk wk;
                wk.init(x, y, max_n);
                int ends[2];
                bool slow[2];
                {
                    // first round: warp k probes steps 1 .. 32 of direction k (dense masks have maxLineGap 0 or close to
                    // it: nearly every walk ends here, after one barrier)
                    if (warp < 2) {
                        int j1, i1;
                        wk.at(warp, lane + 1, j1, i1);
                        const bool inb = j1 >= 0 && j1 < W && i1 >= 0 && i1 < H;
                        bool on = false;
                        if (inb) {
                            const size_t q = (size_t)i1 * W + j1;
                            on = (__ldcg(&bitmap[q >> 5]) >> (q & 31)) & 1u;
                        }
                        const unsigned bo = __ballot_sync(0xffffffffu, on), bi = __ballot_sync(0xffffffffu, inb);
                        if (lane == 0) { s_on[warp] = bo; s_inb[warp] = bi; }
                    }
                    __syncthreads();
#pragma unroll
                    for (int k = 0; k < 2; k++) {
                        unsigned onw = s_on[k];
                        const unsigned inw = s_inb[k];
                        int last_on = 0;
                        bool done = false;
                        const int ob = inw == 0xffffffffu ? INT_MAX : __ffs(~inw);  // first out-of-bounds step (bit b = step b+1)
                        while (onw) {
                            const int io = __ffs(onw);
                            onw &= onw - 1;
                            if (io - last_on - 1 > line_gap) { done = true; break; }
                            last_on = io;
                        }
                        if (!done) {
                            const int wend = min(32, ob == INT_MAX ? INT_MAX : ob - 1);  // last in-bounds step seen
                            if (ob != INT_MAX || wend - last_on > line_gap) done = true;
                        }
                        ends[k] = last_on;
                        slow[k] = !done;
                    }
                    __syncthreads();
                }
                for (int k = 0; k < 2; k++) {
                    if (!slow[k]) continue;  // uniform
                    int last_on = 0;  // step 0 is the start pixel, which is on
                    bool done = false;
                    for (int base = 0; !done; base += HOUGH_THREADS) {
                        const int i = base + tid;
                        int j1, i1;
                        wk.at(k, i, j1, i1);
                        const bool inb = j1 >= 0 && j1 < W && i1 >= 0 && i1 < H;
                        bool on = false;
                        if (inb) {
                            const size_t q = (size_t)i1 * W + j1;
                            on = (__ldcg(&bitmap[q >> 5]) >> (q & 31)) & 1u;
                        }
                        const unsigned bo = __ballot_sync(0xffffffffu, on), bi = __ballot_sync(0xffffffffu, inb);
                        if (lane == 0) { s_on[warp] = bo; s_inb[warp] = bi; }
                        __syncthreads();
                        for (int w = 0; w < HOUGH_THREADS / 32 && !done; w++) {
                            unsigned onw = s_on[w];
                            const unsigned inw = s_inb[w];
                            const int wbase = base + w * 32;
                            const int ob = inw == 0xffffffffu ? INT_MAX : wbase + __ffs(~inw) - 1;
                            while (onw) {
                                const int io = wbase + __ffs(onw) - 1;
                                onw &= onw - 1;
                                if (io - last_on - 1 > line_gap) { done = true; break; }
                                last_on = io;
                            }
                            if (!done) {
                                const int wend = min(wbase + 31, ob == INT_MAX ? INT_MAX : ob - 1);
                                if (ob != INT_MAX || wend - last_on > line_gap) done = true;
                            }
                        }
                        __syncthreads();
                    }
                    ends[k] = last_on;
                }
                if (ends[0] == 0 && ends[1] == 0) {
                    // an isolated pixel (the usual outcome in noise): only the start pixel leaves the mask, nothing to
                    // un-vote (a zero-length segment never counts), no other staged point is affected
                    if (tid == 0) {
                        const size_t q = (size_t)y * W + x;
                        atomicAnd(&bitmap[q >> 5], ~(1u << (q & 31)));
                    }
                    hot = true;
                    p_walk += clock64() - c0; n_iso++;
                    continue;  // the next vote round's barrier orders the bitmap update before any later probe
                }
                int ex0, ey0, ex1, ey1;
                wk.at(0, ends[0], ex0, ey0);
                wk.at(1, ends[1], ex1, ey1);
                const bool good = abs(ex1 - ex0) >= P.min_len || abs(ey1 - ey0) >= P.min_len;
                p_walk += clock64() - c0; c0 = clock64();
                // ---- second pass: clear the on-pixels up to both ends; un-vote them if the line counts
                if (tid == 0) s_ctl[1] = 0;
                __syncthreads();
                for (int k = 0; k < 2; k++)
                    for (int i = tid + k; i <= ends[k]; i += HOUGH_THREADS) {  // k=1 skips the shared start pixel
                        int j1, i1;
                        wk.at(k, i, j1, i1);
                        const size_t q = (size_t)i1 * W + j1;
                        const unsigned bit = 1u << (q & 31);
                        if (__ldcg(&bitmap[q >> 5]) & bit) {
                            atomicAnd(&bitmap[q >> 5], ~bit);
                            if (good) {
                                const int wi = atomicAdd(&s_ctl[1], 1);
                                if (wi < P.walk_cap) walk[wi] = ((unsigned)i1 << 16) | (unsigned)j1;
                            }
                        }
                    }
                __syncthreads();
                __threadfence_block();
                if (good) {
                    const int nw = min(s_ctl[1], P.walk_cap);
                    if (tid < MDB_HOUGH_ANGLES)
                        for (int q = 0; q < nw; q++) {
                            const uint32_t k2 = __ldcg(&walk[q]);
                            atomicAdd(myrow + rho_cs(k2 & 0xffffu, k2 >> 16, my_c, my_s), -1);  // RED, no return
                        }
                    if (tid == 0) {
                        const int li = s_ctl[2];
                        if (li < P.max_lines) {
                            lines[4 * li] = ex0; lines[4 * li + 1] = ey0;
                            lines[4 * li + 2] = ex1; lines[4 * li + 3] = ey1;
                        }
                        s_ctl[2] = li + 1;
                    }
                }
                hot = true;
                __syncthreads();
                refresh(j0);  // the line may have removed pixels of points staged behind this one
                __syncthreads();
                p_unvote += clock64() - c0;
            }
            __syncthreads();
        }
        // dense frame: its points project onto (nearly) the whole accumulator -- clear the slot with coalesced stores
        {
            int4 *a4 = reinterpret_cast<int4 *>(accum);
            const size_t cells = (size_t)MDB_HOUGH_ANGLES * numrho, n4 = cells / 4;
            for (size_t q = tid; q < n4; q += HOUGH_THREADS) __stcg(a4 + q, make_int4(0, 0, 0, 0));
            for (size_t q = n4 * 4 + tid; q < cells; q += HOUGH_THREADS) accum[q] = 0;
        }
        for (int i = tid; i < N; i += HOUGH_THREADS) {
            const uint32_t k = keys[i];
            const size_t p = (size_t)(k >> 16) * W + (k & 0xffffu);
            bitmap[p >> 5] = 0;
        }
        __syncthreads();
        if (tid == 0) { nlines_all[t] = s_ctl[2]; atomicAdd(queue + 4, 1u); }
        if (prof && tid == 0) {
            long long *o = prof + (size_t)t * 10;
            o[0] = N; o[1] = p_setup; o[2] = p_vote; o[3] = p_walk; o[4] = p_unvote; o[5] = p_stage;
            o[6] = n_vote; o[7] = n_line; o[8] = clock64() - pc0; o[9] = n_iso;
        }
        __syncthreads();
    }
}
