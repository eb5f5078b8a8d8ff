// Spatial half of the streaming path, sm_100a: everything after the per-pixel predicate bit.
//
//   act4_kernel       bits -> majority-of-9 (== medianBlur 3 + threshold, Detector.py:329-332)
//                     -> dilate -> erode (MORPH_CLOSE, :335) = `act` bit-frame.  Dense pass over the
//                     predicate bits (1/8 byte per pixel); each thread owns 4 consecutive 32-pixel
//                     words of a row (one 16-byte load) and walks down a band of rows.  The 3 pixels
//                     it needs from either side travel in ONE extra register word (`X`), so the three
//                     3x3 operators are funnel shifts + LOP3 on a circular 5-word register row: no
//                     warp shuffles, no shared memory.  Non-zero act words are also appended to a
//                     per-frame list.
//   dst_sparse_kernel dynamic mask (:234-242) + mask bytes + on-pixel list, driven by that list: the
//                     mask after the close is almost empty, so work is proportional to the number of
//                     non-zero words (~1e3 per 4K frame), not to H*W.  The u8 mask buffer is
//                     persistent; a second list remembers which of its 32-pixel words are non-zero so
//                     that they can be cleared when the slot is reused.
//   dst_dense_kernel  same result by a full scan; only runs for frames whose lists overflowed
//                     (dense masks: clouds, dawn, camera shake).
#pragma once
#include "common.cuh"

#define SPX_ACAP 8192  // per-frame capacity of the non-zero act word list
#define SPX_WCAP 8192  // per-slot capacity of the "non-zero words of the mask buffer" list
#define A4_THREADS 128
#define A4_MLP 3       // rows loaded per thread before they are processed
#define SP_WARPS 4     // warps per CTA in the strip kernels (act_kernel, dst_dense_kernel)
#define SP_MLP 8       // rows loaded per warp before they are processed (memory-level parallelism)
#define SP_USE 30      // useful 32-px words per warp strip (lanes 1..30; lanes 0 and 31 are halo)
#define DENSE_GY 16    // grid.y of dst_dense_kernel (frames are taken from the overflow list)

struct SparseLists {
    uint32_t *alist;   // [T][SPX_ACAP]  (y << 12) | wx  of every non-zero act word of frame t
    unsigned *acount;  // [T]
    uint32_t *wlist;   // [T][SPX_WCAP]  non-zero 32-px words of mask slot t
    unsigned *wcount;  // [T]
    unsigned *dense;   // [0] = number of frames whose lists overflowed, [1 + k] = their indices
};

__device__ __forceinline__ unsigned maj3(unsigned a, unsigned b, unsigned c) { return (a & b) | (c & (a | b)); }

__device__ __forceinline__ unsigned nib_to_bytes(unsigned nib) {
    return ((nib & 1u) ? 0xffu : 0u) | ((nib & 2u) ? 0xff00u : 0u) | ((nib & 4u) ? 0xff0000u : 0u) |
           ((nib & 8u) ? 0xff000000u : 0u);
}

__device__ __forceinline__ void act_list_append(const SparseLists &sl, int t, int y, int wx) {
    const unsigned k = atomicAdd(sl.acount + t, 1u);
    if (k < SPX_ACAP) sl.alist[(size_t)t * SPX_ACAP + k] = ((unsigned)y << 12) | (unsigned)wx;
}

// ------------------------------------------------------------------------------------------
// act4: requires Wb % 4 == 0.  chunks = Wb / 4; thread = (chunk, band) of frame blockIdx.y.
// Register row w[0..3] = the thread's words, w[4] = X: bits 0..2 are the 3 pixels right of w[3],
// bits 29..31 the 3 pixels left of w[0] (circular neighbours: left(i) = w[(i+4)%5], right(i) =
// w[(i+1)%5]).  Three 3x3 stages consume one pixel of halo each, so X's middle bits never reach an
// output bit.  Borders as cv2: replicate for the median; outside pixels ignored by dilate (0) and
// erode (1).
__global__ void __launch_bounds__(A4_THREADS)
act4_kernel(const uint32_t *__restrict__ bits, int H, int Wb, int rows, int chunks, int bands, ActRing ring,
            long long dy0, SparseLists sl) {
    const int idx = blockIdx.x * A4_THREADS + threadIdx.x;
    if (idx >= chunks * bands) return;
    const int t = blockIdx.y;
    const int chunk = idx % chunks, band = idx / chunks;
    const int y0 = band * rows;
    const int yend = min(y0 + rows, H);
    const unsigned FULL = 0xffffffffu;
    const bool ledge = chunk == 0, redge = chunk == chunks - 1;
    unsigned xv = FULL;  // X bits that are inside the image
    if (ledge) xv &= 0x1fffffffu;
    if (redge) xv &= ~7u;
    const uint32_t *fb = bits + (size_t)t * H * Wb + 4 * chunk;
    uint32_t *ob = ring.frame(dy0 + t) + 4 * chunk;
    const int lo = ledge ? 0 : -1, ro = redge ? 3 : 4;  // neighbour words (dummies at the image edge)
    unsigned hs0[5], hc0[5], hs1[5], hc1[5], hd0[5], hd1[5], he0[4], he1[4];
#pragma unroll
    for (int i = 0; i < 5; i++) hs0[i] = hc0[i] = hs1[i] = hc1[i] = hd0[i] = hd1[i] = 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) he0[i] = he1[i] = FULL;
    const int y_stop = yend + 3;  // rows y0-3 .. yend+2 of the predicate bits are consumed
    // rows are loaded A4_MLP at a time, one group ahead of the arithmetic (the pass is otherwise bound
    // by DRAM latency: a warp alternates between waiting for its rows and ~100 ALU ops per row)
    struct Rows {
        uint4 rb[A4_MLP];
        unsigned rl[A4_MLP], rr[A4_MLP];
    };
    auto load_rows = [&](Rows &r, int yb) {
        // rows outside the image repeat the edge row (medianBlur replicates)
        const uint32_t *pr = fb + (size_t)min(max(yb, 0), H - 1) * Wb;
#pragma unroll
        for (int u = 0; u < A4_MLP; u++) {
            r.rb[u] = __ldg(reinterpret_cast<const uint4 *>(pr));
            r.rl[u] = __ldg(pr + lo);
            r.rr[u] = __ldg(pr + ro);
            if ((unsigned)(yb + u) < (unsigned)(H - 1)) pr += Wb;
        }
    };
    auto process_rows = [&](const Rows &r, int yb) {
#pragma unroll
        for (int u = 0; u < A4_MLP; u++) {
            const int yy = yb + u;  // rows past y_stop are computed too (loads are clamped, stores guarded)
            unsigned w[5];
            w[0] = r.rb[u].x; w[1] = r.rb[u].y; w[2] = r.rb[u].z; w[3] = r.rb[u].w;
            const unsigned Lw = ledge ? 0u - (w[0] & 1u) : r.rl[u];
            const unsigned Rw = redge ? 0u - (w[3] >> 31) : r.rr[u];
            w[4] = (Lw & 0xe0000000u) | (Rw & 0x1fffffffu);
            // ---- horizontal 3-sums of predicate row yy: l + w + r = s + 2c ---------------------
            unsigned hs2[5], hc2[5];
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const unsigned l = __funnelshift_l(w[(i + 4) % 5], w[i], 1);
                const unsigned r = __funnelshift_r(w[i], w[(i + 1) % 5], 1);
                hs2[i] = l ^ w[i] ^ r;
                hc2[i] = maj3(l, w[i], r);
            }
            // ---- bin row yy-1: at least 5 of the 9 bits ------------------------------------------
            unsigned bin[5];
            const bool bin_in = (unsigned)(yy - 1) < (unsigned)H;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const unsigned ones = hs0[i] ^ hs1[i] ^ hs2[i], c1 = maj3(hs0[i], hs1[i], hs2[i]);
                const unsigned twos = hc0[i] ^ hc1[i] ^ hc2[i], c2 = maj3(hc0[i], hc1[i], hc2[i]);
                // count = ones + 2 (c1 + twos) + 4 c2 >= 5  <=>  c2 ? (ones | c1 | twos) : (ones & c1 & twos)
                // and  c ? (a | b) : (a & b) == maj3(a, b, c): two LOP3
                const unsigned m5 = maj3(maj3(ones, c1, c2), twos, c2);
                bin[i] = bin_in ? m5 : 0u;
                hs0[i] = hs1[i]; hc0[i] = hc1[i]; hs1[i] = hs2[i]; hc1[i] = hc2[i];
            }
            bin[4] &= xv;
            // ---- dil row yy-2 (outside the image: ones, which the erosion ignores) -------------
            unsigned dil[5];
            const bool dil_in = (unsigned)(yy - 2) < (unsigned)H;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const unsigned hd2 = bin[i] | __funnelshift_l(bin[(i + 4) % 5], bin[i], 1) |
                                     __funnelshift_r(bin[i], bin[(i + 1) % 5], 1);
                dil[i] = dil_in ? (hd0[i] | hd1[i] | hd2) : FULL;
                hd0[i] = hd1[i]; hd1[i] = hd2;
            }
            dil[4] |= ~xv;
            // ---- act row yy-3 = erosion of dil ---------------------------------------------------
            unsigned a[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const unsigned he2 = dil[i] & __funnelshift_l(dil[(i + 4) % 5], dil[i], 1) &
                                     __funnelshift_r(dil[i], dil[i + 1], 1);
                a[i] = he0[i] & he1[i] & he2;
                he0[i] = he1[i]; he1[i] = he2;
            }
            const int y = yy - 3;
            if (y >= y0 && y < yend) {
                *reinterpret_cast<uint4 *>(ob + (size_t)y * Wb) = make_uint4(a[0], a[1], a[2], a[3]);
                if (a[0] | a[1] | a[2] | a[3]) {
#pragma unroll
                    for (int i = 0; i < 4; i++)
                        if (a[i]) act_list_append(sl, t, y, 4 * chunk + i);
                }
            }
        }
    };
    Rows ra, rb2;
    int yb = y0 - 3;
    load_rows(ra, yb);
    for (;;) {  // two groups per iteration: the register sets ping-pong, no moves
        const bool more1 = yb + A4_MLP < y_stop;
        if (more1) load_rows(rb2, yb + A4_MLP);
        process_rows(ra, yb);
        if (!more1) break;
        yb += A4_MLP;
        const bool more2 = yb + A4_MLP < y_stop;
        if (more2) load_rows(ra, yb + A4_MLP);
        process_rows(rb2, yb);
        if (!more2) break;
        yb += A4_MLP;
    }
}

// ------------------------------------------------------------------------------------------
// Strip variant for widths whose word count is not a multiple of 4: one warp = one strip of SP_USE
// words (960 px; lanes 0 and 31 are halo columns) walked top to bottom, neighbours by warp shuffle.
struct RowH {  // horizontal 3-sums of one bit row: s = parity, c = carry (l + w + r = s + 2c)
    unsigned s, c;
};

__global__ void __launch_bounds__(SP_WARPS * 32)
act_kernel(const uint32_t *__restrict__ bits, int W, int H, int T, int rows, int strips, int bands,
           ActRing ring, long long dy0, SparseLists sl) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * SP_WARPS + warp;
    const int t = blockIdx.y;
    if (tile >= strips * bands) return;
    const int strip = tile % strips, band = tile / strips;
    const int Wb = W >> 5;
    const int wx = strip * SP_USE - 1 + lane;
    const bool lane_in = wx >= 0 && wx < Wb;
    const bool lane_out = lane_in && lane >= 1 && lane <= SP_USE;
    const int y0 = band * rows;
    const unsigned FULL = 0xffffffffu;
    const uint32_t *fb = bits + (size_t)t * H * Wb + (lane_in ? wx : 0);
    uint32_t *ob = ring.frame(dy0 + t) + (lane_in ? wx : 0);
    RowH h0 = {0, 0}, h1 = {0, 0};
    unsigned hd0 = 0, hd1 = 0, he0 = FULL, he1 = FULL;
    // rows y0-3 .. y0+rows+2 of b are needed; SP_MLP rows are loaded at a time so that every warp keeps
    // several independent 128-byte requests in flight (the pass is DRAM-latency-bound otherwise)
    const int y_first = y0 - 3, y_last = y0 + rows + 3;
    for (int yb = y_first; yb < y_last; yb += SP_MLP) {
    unsigned rowbuf[SP_MLP];
    {   // the row pointer advances only inside the image: rows outside repeat the edge row (medianBlur)
        const uint32_t *pr = fb + (unsigned)min(max(yb, 0), H - 1) * (unsigned)Wb;
#pragma unroll
        for (int u = 0; u < SP_MLP; u++) {
            rowbuf[u] = __ldg(pr);
            if ((unsigned)(yb + u) < (unsigned)(H - 1)) pr += Wb;
        }
    }
#pragma unroll
    for (int u = 0; u < SP_MLP; u++) {
        const int yy = yb + u;
        if (yy >= y_last) break;
        const unsigned bw = lane_in ? rowbuf[u] : 0u;
        // ---- horizontal sums of b row yy (rows/cols replicated outside the image: medianBlur) --
        unsigned Lw = __shfl_up_sync(FULL, bw, 1), Rw = __shfl_down_sync(FULL, bw, 1);
        if (wx == 0) Lw = (bw & 1u) << 31;
        if (wx == Wb - 1) Rw = bw >> 31;
        const unsigned l = (bw << 1) | (Lw >> 31), r = (bw >> 1) | (Rw << 31);
        RowH h2;
        h2.s = l ^ bw ^ r;
        h2.c = maj3(l, bw, r);
        // ---- bin row yy-1 = at least 5 of the 9 bits ----------------------------------------
        unsigned bin = 0;
        {
            const int y = yy - 1;
            if (lane_in && y >= 0 && y < H) {
                const unsigned ones = h0.s ^ h1.s ^ h2.s, c1 = maj3(h0.s, h1.s, h2.s);
                const unsigned twos = h0.c ^ h1.c ^ h2.c, c2 = maj3(h0.c, h1.c, h2.c);
                const unsigned t0b = c1 ^ twos, t1b = c1 & twos;  // weights 2 and 4
                bin = (t1b & c2) | ((t1b ^ c2) & (t0b | ones));
            }
        }
        h0 = h1; h1 = h2;
        // ---- dil row yy-2 (outside the image: ones, ignored by the erosion) -------------------
        unsigned dil;
        {
            const unsigned Lb = __shfl_up_sync(FULL, bin, 1), Rb = __shfl_down_sync(FULL, bin, 1);
            const unsigned hd2 = bin | (bin << 1) | (Lb >> 31) | (bin >> 1) | (Rb << 31);
            const int y = yy - 2;
            dil = (lane_in && y >= 0 && y < H) ? (hd0 | hd1 | hd2) : FULL;
            hd0 = hd1; hd1 = hd2;
        }
        // ---- act row yy-3 = erosion of dil ---------------------------------------------------
        {
            const unsigned Ld = __shfl_up_sync(FULL, dil, 1), Rd = __shfl_down_sync(FULL, dil, 1);
            const unsigned he2 = dil & ((dil << 1) | (Ld >> 31)) & ((dil >> 1) | (Rd << 31));
            const int y = yy - 3;
            if (lane_out && y >= y0 && y < y0 + rows && y < H) {
                const unsigned a = he0 & he1 & he2;
                ob[(size_t)y * Wb] = a;
                if (a) act_list_append(sl, t, y, wx);
            }
            he0 = he1; he1 = he2;
        }
    }
    }
}

// ------------------------------------------------------------------------------------------
// Dynamic mask (Detector.py:234-242): m = not(on in ALL of the last L act frames); dst = act &
// erode3x3(m).  Outside the image m = ones (ignored by the erosion).  `a0` = act(d) of the word;
// history is read four frames at a time (independent loads) and only while the AND is non-empty.
__device__ __forceinline__ unsigned dy_m(const ActRing &ring, long long d, int L, size_t off, unsigned a0) {
    unsigned acc = a0;
    int k = 1;
    while (k < L && acc) {
        unsigned v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (k + j < L) ? __ldg(ring.frame(d - k - j) + off) : 0xffffffffu;
        acc &= v[0] & v[1] & v[2] & v[3];
        k += 4;
    }
    return ~acc;
}

__device__ __forceinline__ unsigned dst_word(const ActRing &ring, long long d, int L, int dy_on, int y, int wx,
                                             int H, int Wb, unsigned act) {
    if (!dy_on) return act;
    const uint32_t *cur = ring.frame(d);
    unsigned a[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const int yy = y + r - 1, xx = wx + c - 1;
            const bool in = (unsigned)yy < (unsigned)H && (unsigned)xx < (unsigned)Wb;
            a[r][c] = (r == 1 && c == 1) ? act : (in ? __ldg(cur + (size_t)yy * Wb + xx) : 0u);
        }
    unsigned out = act;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        unsigned m[3];
#pragma unroll
        for (int c = 0; c < 3; c++)  // a == 0 (also: outside the image) -> m = ones without touching the history
            m[c] = a[r][c] ? dy_m(ring, d, L, (size_t)(y + r - 1) * Wb + (wx + c - 1), a[r][c]) : 0xffffffffu;
        out &= m[1] & ((m[1] << 1) | (m[0] >> 31)) & ((m[1] >> 1) | (m[2] << 31));
    }
    return out;
}

// One 32-pixel word of mask slot t: bytes, shadow bits, on-pixel count + list, non-zero word list.
// `prev` = what the buffer holds now (shadow).  Nothing is written when buffer and result are both 0.
__device__ __forceinline__ void dst_emit(uint8_t *__restrict__ dst_t, uint32_t *__restrict__ dstbits_t, int W, int Wb,
                                         int y, int wx, unsigned prev, unsigned out_bits, unsigned *npoints_t,
                                         uint32_t *points_t, int cap, const SparseLists &sl, int t) {
    if (!(prev | out_bits)) return;
    uint4 a = make_uint4(nib_to_bytes(out_bits & 15u), nib_to_bytes((out_bits >> 4) & 15u),
                         nib_to_bytes((out_bits >> 8) & 15u), nib_to_bytes((out_bits >> 12) & 15u));
    uint4 b = make_uint4(nib_to_bytes((out_bits >> 16) & 15u), nib_to_bytes((out_bits >> 20) & 15u),
                         nib_to_bytes((out_bits >> 24) & 15u), nib_to_bytes(out_bits >> 28));
    uint4 *o = reinterpret_cast<uint4 *>(dst_t + (size_t)y * W + (size_t)wx * 32);
    o[0] = a;
    o[1] = b;
    if (prev != out_bits) dstbits_t[(size_t)y * Wb + wx] = out_bits;
    if (out_bits) {
        const unsigned c = __popc(out_bits);
        unsigned slot = atomicAdd(npoints_t, c);
        unsigned ob = out_bits;
        while (ob) {
            const int bpos = __ffs(ob) - 1;
            ob &= ob - 1;
            if (slot < (unsigned)cap) points_t[slot] = ((unsigned)y << 16) | (unsigned)(wx * 32 + bpos);
            slot++;
        }
        const unsigned k = atomicAdd(sl.wcount + t, 1u);
        if (k < SPX_WCAP) sl.wlist[(size_t)t * SPX_WCAP + k] = ((unsigned)y << 12) | (unsigned)wx;
    }
}

// grid = T blocks; block t handles frame t of the batch (dy index dy0 + t, mask slot t)
__global__ void __launch_bounds__(256)
dst_sparse_kernel(ActRing ring, int W, int H, int n, long long dy0, int dy_on, uint8_t *__restrict__ dst,
                  uint32_t *__restrict__ dstbits, unsigned *__restrict__ npoints, uint32_t *__restrict__ points,
                  int cap, SparseLists sl) {
    const int t = blockIdx.x, tid = threadIdx.x;
    const int Wb = W >> 5;
    const unsigned na = sl.acount[t], nw = sl.wcount[t];
    if (na > SPX_ACAP || nw > SPX_WCAP) {  // a list overflowed: the dense kernel rescans this frame
        if (tid == 0) {
            sl.dense[1 + atomicAdd(sl.dense, 1u)] = (unsigned)t;
            sl.wcount[t] = 0;
        }
        return;
    }
    const long long d = dy0 + t;
    const int L = (int)((d + 1) < n ? (d + 1) : n);  // SlidingWindow.length of the dy window
    const uint32_t *cur = ring.frame(d);
    uint8_t *dst_t = dst + (size_t)t * W * H;
    uint32_t *dstbits_t = dstbits + (size_t)t * H * Wb;
    // words that are non-zero in the buffer and get no new content: clear them
    for (unsigned e = tid; e < nw; e += blockDim.x) {
        const unsigned v = sl.wlist[(size_t)t * SPX_WCAP + e];
        const int y = (int)(v >> 12), wx = (int)(v & 4095u);
        if (__ldg(cur + (size_t)y * Wb + wx) == 0u) {
            uint4 *o = reinterpret_cast<uint4 *>(dst_t + (size_t)y * W + (size_t)wx * 32);
            o[0] = make_uint4(0u, 0u, 0u, 0u);
            o[1] = make_uint4(0u, 0u, 0u, 0u);
            dstbits_t[(size_t)y * Wb + wx] = 0u;
        }
    }
    __syncthreads();
    if (tid == 0) sl.wcount[t] = 0;
    __syncthreads();
    for (unsigned e = tid; e < na; e += blockDim.x) {
        const unsigned v = sl.alist[(size_t)t * SPX_ACAP + e];
        const int y = (int)(v >> 12), wx = (int)(v & 4095u);
        const unsigned act = __ldg(cur + (size_t)y * Wb + wx);
        const unsigned out_bits = dst_word(ring, d, L, dy_on, y, wx, H, Wb, act);
        const unsigned prev = dstbits_t[(size_t)y * Wb + wx];
        dst_emit(dst_t, dstbits_t, W, Wb, y, wx, prev, out_bits, npoints + t, points + (size_t)t * cap, cap, sl, t);
    }
}

// test hook: put every frame on the overflow list
__global__ void dst_force_dense_kernel(int T, SparseLists sl) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    sl.dense[1 + atomicAdd(sl.dense, 1u)] = (unsigned)t;
    sl.wcount[t] = 0;
}

// Full-scan variant for the frames on the overflow list; warp = strip of SP_USE words walked top-down.
__global__ void __launch_bounds__(SP_WARPS * 32)
dst_dense_kernel(ActRing ring, int W, int H, int n, long long dy0, int dy_on, int rows, int strips, int bands,
                 uint8_t *__restrict__ dst, uint32_t *__restrict__ dstbits, unsigned *__restrict__ npoints,
                 uint32_t *__restrict__ points, int cap, SparseLists sl) {
    const unsigned ndense = sl.dense[0];
    if (ndense == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * SP_WARPS + warp;
    if (tile >= strips * bands) return;
    const int strip = tile % strips, band = tile / strips;
    const int Wb = W >> 5;
    const int wx = strip * SP_USE - 1 + lane;
    const bool lane_in = wx >= 0 && wx < Wb;
    const bool lane_out = lane_in && lane >= 1 && lane <= SP_USE;
    const int y0 = band * rows;
    const unsigned FULL = 0xffffffffu;
    const int ylast = dy_on ? y0 + rows + 1 : y0 + rows;
    const int yfirst = dy_on ? y0 - 1 : y0;
    const int ylag = dy_on ? 1 : 0;  // the output row trails the input row by this much
    const unsigned lmask = lane_in ? FULL : 0u, omask = lane_out ? FULL : 0u;
    for (unsigned kd = blockIdx.y; kd < ndense; kd += gridDim.y) {
    const int t = (int)sl.dense[1 + kd];
    const long long d = dy0 + t;
    const int L = (int)((d + 1) < n ? (d + 1) : n);  // SlidingWindow.length of the dy window
    const uint32_t *cur = ring.frame(d) + (lane_in ? wx : 0);
    unsigned hm0 = FULL, hm1 = FULL, act_prev = 0;
    uint32_t *dstbits_t = dstbits + (size_t)t * H * Wb;
    const uint32_t *dbase = dstbits_t + (lane_in ? wx : 0);
    for (int yb = yfirst; yb < ylast; yb += SP_MLP) {
    unsigned actbuf[SP_MLP], prevbuf[SP_MLP];
    {   // row pointers advance by one row only inside the image (rows outside repeat the edge row: always a
        // valid address, no predicated loads, no per-load index arithmetic)
        const uint32_t *pa = cur + (unsigned)min(max(yb, 0), H - 1) * (unsigned)Wb;
        const uint32_t *pd = dbase + (unsigned)min(max(yb - ylag, 0), H - 1) * (unsigned)Wb;
#pragma unroll
        for (int u = 0; u < SP_MLP; u++) {
            const int yy = yb + u;
            actbuf[u] = __ldg(pa);
            prevbuf[u] = *pd;
            if ((unsigned)yy < (unsigned)(H - 1)) pa += Wb;
            if ((unsigned)(yy - ylag) < (unsigned)(H - 1)) pd += Wb;
        }
    }
#pragma unroll
    for (int u = 0; u < SP_MLP; u++) {
        const int yy = yb + u;
        if (yy >= ylast) break;
        const int yo_u = yy - ylag;
        const unsigned act = ((unsigned)yy < (unsigned)H) ? (actbuf[u] & lmask) : 0u;
        const unsigned prev_u = ((unsigned)(yo_u - y0) < (unsigned)rows && yo_u < H) ? (prevbuf[u] & omask) : 0u;
        // row segment empty now, empty one row ago, zeros in the mask buffer: nothing to compute or write
        if (!__any_sync(FULL, (act | act_prev | prev_u) != 0u)) {
            hm0 = hm1;
            hm1 = FULL;
            act_prev = 0;
            continue;
        }
        unsigned out_bits;
        int yo;
        if (dy_on) {
            // m = not(on in all of the last L frames); the loop ends as soon as the AND is empty
            unsigned acc = act;
            for (int k = 1; k < L && acc; k++) acc &= __ldg(ring.frame(d - k) + wx + (size_t)yy * Wb);
            const unsigned m = ~acc;  // rows / columns outside the image: act = 0 -> m = ones
            const unsigned Lm = __shfl_up_sync(FULL, m, 1), Rm = __shfl_down_sync(FULL, m, 1);
            const unsigned hm2 = m & ((m << 1) | (Lm >> 31)) & ((m >> 1) | (Rm << 31));
            out_bits = act_prev & hm0 & hm1 & hm2;  // row yy-1
            hm0 = hm1; hm1 = hm2;
            act_prev = act;
            yo = yy - 1;
        } else {
            out_bits = act;
            yo = yy;
        }
        if (yo >= y0 && yo < y0 + rows && yo < H && lane_out)
            dst_emit(dst + (size_t)t * W * H, dstbits_t, W, Wb, yo, wx, prev_u, out_bits, npoints + t,
                     points + (size_t)t * cap, cap, sl, t);
    }
    }
    }
}
