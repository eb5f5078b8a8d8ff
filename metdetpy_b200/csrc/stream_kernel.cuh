// Time-tiled streaming path for whole batches (the throughput path), sm_100a.
//
// The reference recomputes max/mean over the n-frame ring for every frame (utils.py:269-307), i.e.
// (n+2) bytes of traffic per pixel per frame.  Here every frame is read from HBM ONCE:
//
//   temporal_kernel  (stack -> diff -> threshold, fused)     reads  u8 frame, writes 1 bit / pixel
//       each thread owns 16 consecutive pixels for the whole batch.  Its last n frames live in a
//       private shared-memory ring fed by cp.async (LDGSTS, K frames in flight, no CTA barrier in
//       the steady state); the sliding max is van Herk / Gil-Werman (prefix max in registers +
//       suffix max ring in shared memory, one backward scan every n frames), the sliding sum is a
//       running u16x2 register.  The decision  median(max - floor(sum/L)) > thr  is rewritten as
//       "at least 5 of the 9 neighbours have  max*L - sum > thr*L",  so this kernel only emits the
//       per-pixel predicate (exact integer arithmetic, no division) as a bit.
//   act4_kernel / dst_sparse_kernel (spatial_kernel.cuh: 3x3 median == majority of 9 bits, 3x3 close,
//       dynamic mask, dst): read the bits, write the u8 mask.  All 3x3 operators are bitwise on
//       32-pixel words.
//
// Per frame HBM traffic: H*W (frame) + 4 * H*W/8 (predicate bits and act bits, out and in)
// + the mask bytes that change (the u8 mask buffer is persistent) = 1.44 H*W measured, against the algorithmic
// 2 H*W of SURVEY.md section 8(d).
#pragma once
#include "common.cuh"
#include "spatial_kernel.cuh"
#include "temporal_kernel.cuh"
#include "temporal3_dispatch.cuh"

#define ST_K 8          // frames in flight per thread (cp.async groups)

struct StreamState {
    int ok = 0;
    int W = 0, H = 0, n = 0, max_batch = 0, device = 0;
    int t_threads = 32;     // CTA size of the temporal kernel
    int t_wpt = 2;          // 32-bit words (4 px) per thread in the temporal kernel
    int t_version = 3;      // 3: temporal3_kernel (register ring; falls back to 2 for windows without a shape or frames
                            // that are not contiguous); 2: temporal2_kernel (temporal_kernel.cuh); 1: the first-generation kernel below
    int t3_variant = 0;     // tuning hook (temporal3_dispatch.cuh)
    int t_last = 0;         // which generation the last launch used
    int t_kdiv = 1;         // temporal2: sub-blocks per window (divides n)
    int t_kdiv_req = 0;     // test hook: force this many sub-blocks (0 = choose)
    int dst_rows = 32;      // output rows per warp strip in dst_dense_kernel
    int sp_rows_single = 8; // act4 band height when the batch is one frame
    int single_dense = 0;   // one-frame batches take the full-scan dst kernel (measured slower than the list walk: off)
    int force_dense = 0;    // test hook: dst of every frame by the full-scan kernel
    int force_strip = 0;    // test hook: act by the warp-strip kernel even when W % 128 == 0
    size_t t_smem_per_thread = 0;
    int sp_rows = 8;        // output rows per warp strip in the spatial kernel
    // predicate bits [max_batch][H][W/32], two buffers: batch k+1's temporal pass (front stream) writes one
    // while batch k's act pass (back stream) still reads the other
    uint32_t *d_bits = nullptr, *d_bits2 = nullptr;
    T3Table t3tab;
};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// even bytes (0,2) / odd bytes (1,3) of a packed u8x4 word as u16x2
__device__ __forceinline__ unsigned ev(unsigned x) { return x & 0x00ff00ffu; }
__device__ __forceinline__ unsigned od(unsigned x) { return prmt(x, 0u, 0x4341u); }
__device__ __forceinline__ unsigned pack_eo(unsigned e, unsigned o) { return e | (o << 8); }

// WPT = 32-bit words (4 pixels each) per thread: 4 (16 px, 16-byte accesses) or 2 (8 px, 8-byte
// accesses -- twice the resident warps for the same shared-memory footprint per pixel).
template <int WPT> struct VecT;
template <> struct VecT<4> { typedef uint4 type; };
template <> struct VecT<2> { typedef uint2 type; };

template <int WPT>
__device__ __forceinline__ void cp_async_vec(uint32_t saddr, const void *g) {
    if (WPT == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
template <int WPT>
__device__ __forceinline__ void lds_vec(unsigned (&w)[WPT], uint32_t saddr) {
    if (WPT == 4) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(saddr));
    else asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(saddr));
}
template <int WPT>
__device__ __forceinline__ void sts_vec(uint32_t saddr, const unsigned (&w)[WPT]) {
    if (WPT == 4) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
    else asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(saddr), "r"(w[0]), "r"(w[1]) : "memory");
}

template <bool MASKED, int WPT>
__global__ void __launch_bounds__(128)
temporal_kernel(FrameSrc src, long long t0, int T, int n, int HWG, const int *__restrict__ thr,
                uint8_t *__restrict__ bits) {
    constexpr int VB = WPT * 4;  // bytes (= pixels) per thread per frame
    extern __shared__ uint4 t_smem[];
    const int nt = blockDim.x, tid = threadIdx.x;
    const int RS = n + ST_K;
    // shared memory: ring [RS][nt] raw frames (slot RS-1 doubles as "frame t0-n"), smx [n][nt] suffix
    // max of the previous block by position, thr_s [T]
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(t_smem);
    const uint32_t slot_stride = nt * VB;
    const uint32_t ring_s = smem0 + tid * VB;
    const uint32_t smx_s = ring_s + RS * slot_stride;
    uint8_t *thr_s = reinterpret_cast<uint8_t *>(t_smem) + (size_t)(RS + n) * slot_stride;
    for (int i = tid; i < T; i += nt) thr_s[i] = (uint8_t)min(max(thr[i], 0), 255);
    __syncthreads();
    const int g = blockIdx.x * nt + tid;
    if (g >= HWG) return;
    // frames of this batch: contiguous in the caller's buffer (zero-copy) or slots of the ring
    const uint8_t *gbase = (src.cur ? src.cur : src.ring) + (size_t)g * VB;
    const int Rw = src.cur ? 0x7fffffff : src.R;  // no wrap in the caller's buffer

    unsigned mk[WPT];
#pragma unroll
    for (int k = 0; k < WPT; k++) mk[k] = ~0u;
    if (MASKED) {
        const unsigned *m = reinterpret_cast<const unsigned *>(src.mask + (size_t)g * VB);
#pragma unroll
        for (int k = 0; k < WPT; k++) mk[k] = m[k] * 0xffu;  // {0,1} -> {0x00,0xff}
    }
    unsigned zero[WPT];
#pragma unroll
    for (int k = 0; k < WPT; k++) zero[k] = 0;

    // ---- history: frames t0-n+1 .. t0-1 -> slots 0 .. n-2 ; slot RS-1 = zeros -----------------
    sts_vec<WPT>(ring_s + (RS - 1) * slot_stride, zero);
    for (int p = 1; p < n; p++) {
        const long long th = t0 - n + p;
        if (th >= 0) cp_async_vec<WPT>(ring_s + (p - 1) * slot_stride, src.frame(th) + (size_t)g * VB);
        else sts_vec<WPT>(ring_s + (p - 1) * slot_stride, zero);
    }
    cp_async_commit();
    // ---- prime the pipeline: frames t0 .. t0+K-1 -> slots n-1 .. n+K-2 -------------------------
    int pf_slot = src.cur ? (int)(t0 - src.t0) : (int)(t0 % src.R);  // slot of the next frame to prefetch
    for (int i = 0; i < ST_K; i++) {
        if (i < T) cp_async_vec<WPT>(ring_s + (n - 1 + i) * slot_stride, gbase + (size_t)pf_slot * src.HW);
        cp_async_commit();
        if (++pf_slot == Rw) pf_slot = 0;
    }
    cp_async_wait<ST_K>();  // history landed

    unsigned sE[WPT], sO[WPT];  // window sums, u16x2 (even / odd pixels)
    {
        unsigned aE[WPT], aO[WPT];
#pragma unroll
        for (int k = 0; k < WPT; k++) sE[k] = sO[k] = aE[k] = aO[k] = 0;
        for (int p = n - 1; p >= 1; p--) {
            unsigned w[WPT], o[WPT];
            lds_vec<WPT>(w, ring_s + (p - 1) * slot_stride);
            if (MASKED) {
#pragma unroll
                for (int k = 0; k < WPT; k++) w[k] &= mk[k];
                sts_vec<WPT>(ring_s + (p - 1) * slot_stride, w);
            }
#pragma unroll
            for (int k = 0; k < WPT; k++) {
                const unsigned e = ev(w[k]), d = od(w[k]);
                sE[k] += e; sO[k] += d;
                aE[k] = __vmaxu2(aE[k], e); aO[k] = __vmaxu2(aO[k], d);
                o[k] = pack_eo(aE[k], aO[k]);
            }
            sts_vec<WPT>(smx_s + (p - 1) * slot_stride, o);
        }
    }
    // suffix-max slots: position p (1..n-1) lives in slot p-1; slot n-1 stays zero ("position n")
    sts_vec<WPT>(smx_s + (n - 1) * slot_stride, zero);

    uint32_t a_cur = ring_s + (n - 1) * slot_stride;        // smem address of the current frame's slot
    uint32_t a_old = ring_s + (RS - 1) * slot_stride;       // frame t-n; destination of the next prefetch
    const uint32_t a_end = ring_s + RS * slot_stride;
    uint8_t *bout = bits + (size_t)g * WPT / 2;             // WPT*4 bits per thread and frame
    const size_t bstride = (size_t)HWG * WPT / 2;
    // next frame to prefetch = slot pf_slot: contiguous in the caller's buffer, or ring slots that wrap
    int pf_left = T - ST_K;                                  // frames still to be prefetched
    const uint32_t thr_sa = smem0 + (RS + n) * slot_stride;  // shared address of thr_s
    int L = (int)(t0 + 1 < n ? t0 + 1 : n);                  // SlidingWindow.length of the current frame
    const unsigned one = (unsigned)(n > 0), neg1 = 0u - one; // opaque to the compiler on purpose
    int i = 0;
    while (i < T) {
        const int nb = min(n, T - i);  // frames of this block
        unsigned pE[WPT], pO[WPT];     // prefix max of the current block (0 = identity: first frame sets it)
#pragma unroll
        for (int k = 0; k < WPT; k++) pE[k] = pO[k] = 0;
        uint32_t a_smx = smx_s;        // suffix max at position j+1
        for (int j = 0; j < nb; j++, i++) {
            cp_async_wait<ST_K - 1>();  // this thread's copy of frame i has landed
            unsigned xw[WPT], ow[WPT], mw[WPT];
            lds_vec<WPT>(xw, a_cur);
            lds_vec<WPT>(ow, a_old);
            lds_vec<WPT>(mw, a_smx);
            if (MASKED) {
#pragma unroll
                for (int k = 0; k < WPT; k++) xw[k] &= mk[k];
                sts_vec<WPT>(a_cur, xw);
            }
            // slot a_old is free now: fetch frame i+K into it
            if (pf_left > 0) cp_async_vec<WPT>(a_old, gbase + (size_t)pf_slot * src.HW);
            cp_async_commit();
            pf_left--;
            if (++pf_slot == Rw) pf_slot = 0;

            unsigned thr_i;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(thr_i) : "r"(thr_sa + i));
            const unsigned Tq = thr_i * (unsigned)L;                          // <= 255*128
            const unsigned Cpk = (0x7fffu - Tq) * 0x00010001u;                // per-half bias
            unsigned M[WPT];
#pragma unroll
            for (int k = 0; k < WPT; k++) {
                const unsigned e = ev(xw[k]), d = od(xw[k]);
                // running sums on the FMA pipe (IMAD with a run-time 1 / -1): the ALU pipe, which carries
                // every LOP3 / PRMT / VIMNMX of this loop, is the bottleneck of the kernel
                sE[k] = ev(ow[k]) * neg1 + (e * one + sE[k]);
                sO[k] = od(ow[k]) * neg1 + (d * one + sO[k]);
                pE[k] = __vmaxu2(pE[k], e);
                pO[k] = __vmaxu2(pO[k], d);
                const unsigned wE = __vmaxu2(pE[k], ev(mw[k])), wO = __vmaxu2(pO[k], od(mw[k]));
                // per half: max*L - sum + 0x7fff - thr*L ; bit 15 set <=> max*L - sum > thr*L
                const unsigned vE = wE * (unsigned)L + Cpk - sE[k];
                const unsigned vO = wO * (unsigned)L + Cpk - sO[k];
                M[k] = prmt(vE, vO, 0xFBD9u);  // sign-replicate bytes 1,5,3,7 -> 0x00/0xff per pixel
            }
            const unsigned q01 = (M[0] & 0x08040201u) | (M[1] & 0x80402010u);
            const unsigned r01 = q01 * 0x01010101u;
            if (WPT == 4) {
                const unsigned q23 = (M[2 % WPT] & 0x08040201u) | (M[3 % WPT] & 0x80402010u);
                const unsigned r23 = q23 * 0x01010101u;
                *reinterpret_cast<uint16_t *>(bout) = (uint16_t)prmt(r01, r23, 0x4473u);
            } else {
                *bout = (uint8_t)(r01 >> 24);
            }
            bout += bstride;
            L += (L < n);
            a_smx += slot_stride;
            a_cur += slot_stride; if (a_cur == a_end) a_cur = ring_s;
            a_old += slot_stride; if (a_old == a_end) a_old = ring_s;
        }
        if (nb == n && i < T) {  // block complete and more frames follow: suffix max by position 1..n-1
            unsigned aE[WPT], aO[WPT];
#pragma unroll
            for (int k = 0; k < WPT; k++) aE[k] = aO[k] = 0;
            uint32_t a = (a_cur == ring_s) ? a_end - slot_stride : a_cur - slot_stride;  // last frame of the block
            uint32_t o_s = smx_s + (n - 2) * slot_stride;
            for (int p = n - 1; p >= 1; p--) {
                unsigned w[WPT], o[WPT];
                lds_vec<WPT>(w, a);
#pragma unroll
                for (int k = 0; k < WPT; k++) {
                    aE[k] = __vmaxu2(aE[k], ev(w[k]));
                    aO[k] = __vmaxu2(aO[k], od(w[k]));
                    o[k] = pack_eo(aE[k], aO[k]);
                }
                sts_vec<WPT>(o_s, o);
                o_s -= slot_stride;
                a = (a == ring_s) ? a_end - slot_stride : a - slot_stride;
            }
        }
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------
static inline void stream_state_free(StreamState &s) {
    if (s.d_bits) cudaFree(s.d_bits);
    if (s.d_bits2) cudaFree(s.d_bits2);
    s.d_bits = s.d_bits2 = nullptr;
    for (uint2 *&p : s.t3tab.d_tab) { if (p) cudaFree(p); p = nullptr; }
    s.ok = 0;
}

// sub-blocks per window for temporal2.  Splitting the window saves shared memory (more resident warps) but every
// block end costs a register FIFO update and a scan set-up, and runs get shorter: it pays for long windows only
// (measured at 4K: n = 30 is fastest with one block -- chain 2.25 ms vs 2.47 / 2.71 ms with 3 / 6 blocks; n = 60:
// 3.44 ms with one block, 3.12 / 3.01 / 2.98 / 3.12 ms with 2 / 3 / 4 / 6).
static inline int stream_choose_kdiv(int n, int req) {
    if (req >= 1 && req <= T2_KMAX && n % req == 0 && n / req >= 2) return req;
    if (n < 48) return 1;
    for (int k = T2_KMAX; k >= 2; k--)
        if (n % k == 0 && n / k >= 15) return k;
    return 1;
}

// shared memory of one temporal CTA: per-thread ring + suffix slots, u16 bias per frame, one "bias changes" bit per frame
static inline size_t stream_temporal_smem(const StreamState &s, int nt, int T, int version) {
    const int slots = version == 2 ? s.n + ST_K + s.n / s.t_kdiv : 2 * s.n + ST_K;
    return (size_t)slots * 4 * s.t_wpt * nt + (((size_t)2 * T + 15) & ~(size_t)15);
}

// choose words-per-thread of the temporal kernel and derive its CTA size / shared memory
// (nt_req > 0 forces the CTA size; otherwise the largest CTA that keeps the most warps per SM)
static inline int stream_state_config(StreamState &s, int wpt, int nt_req = 0) {
    if (wpt != 2 && wpt != 4) return -1;
    const size_t sm_bytes = 228 * 1024, cta_max = 220 * 1024, reserved = 1024;
    s.t_kdiv = stream_choose_kdiv(s.n, s.t_kdiv_req);
    const size_t per_thread_v1 = (size_t)(2 * s.n + ST_K) * 4 * wpt;  // the first-generation kernel must fit too
    const size_t per_thread = (size_t)(s.n + ST_K + s.n / s.t_kdiv) * 4 * wpt;
    const size_t table = ((size_t)2 * s.max_batch + 15) & ~(size_t)15;
    int best_nt = 0;
    size_t best_warps = 0;
    for (int nt = 128; nt >= 32; nt >>= 1) {
        if (nt_req && nt != nt_req) continue;
        if (per_thread_v1 * nt + table > cta_max) continue;
        const size_t cta = per_thread * nt + table;
        const size_t warps = sm_bytes / (cta + reserved) * (nt / 32);
        if (warps > best_warps) { best_warps = warps; best_nt = nt; }
    }
    if (!best_nt) return -1;
    s.t_wpt = wpt;
    s.t_smem_per_thread = per_thread;
    s.t_threads = best_nt;
    return 0;
}

static inline int stream_state_init(StreamState &s, int W, int H, int n, int device, int max_batch) {
    s.W = W; s.H = H; s.n = n; s.device = device; s.max_batch = max_batch; s.ok = 0;
    if (W % 32 != 0 || n < 2 || n > 128 || max_batch > 4096) return 0;  // generic kernel serves these
    const size_t budget = 220 * 1024;
    // 16 px per thread needs fewer instructions per pixel, 8 px per thread doubles the resident warps:
    // the wider variant only pays when its ring still leaves >= 8 warps per SM
    const bool wide = (size_t)(2 * n + ST_K) * 16 * 32 * 8 <= budget;
    if (stream_state_config(s, wide ? 4 : 2) != 0 && stream_state_config(s, 2) != 0) return 0;
    if (cudaMalloc((void **)&s.d_bits, (size_t)(max_batch + T3_SLACK_PLANES) * H * (W / 32) * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc((void **)&s.d_bits2, (size_t)(max_batch + T3_SLACK_PLANES) * H * (W / 32) * sizeof(uint32_t)) != cudaSuccess) {
        cudaGetLastError();
        stream_state_free(s);
        return 0;  // fall back to the generic per-frame kernel
    }
    s.sp_rows = 64;
    s.t3tab.cap = max_batch + T3_SLACK_PLANES;
    for (uint2 *&p : s.t3tab.d_tab)
        if (cudaMalloc((void **)&p, (size_t)s.t3tab.cap * sizeof(uint2)) != cudaSuccess) { cudaGetLastError(); p = nullptr; }
#define ST_SETATTR(K)                                                                                      \
    if (cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget) != cudaSuccess || \
        cudaFuncSetAttribute(K, cudaFuncAttributePreferredSharedMemoryCarveout, 100) != cudaSuccess)       \
        return -1;
    ST_SETATTR((temporal_kernel<false, 2>)) ST_SETATTR((temporal_kernel<true, 2>))
    ST_SETATTR((temporal_kernel<false, 4>)) ST_SETATTR((temporal_kernel<true, 4>))
#define ST_SETATTR2(NT)                                                                                          \
    ST_SETATTR((temporal2_kernel<false, 2, NT, false>)) ST_SETATTR((temporal2_kernel<true, 2, NT, false>))      \
    ST_SETATTR((temporal2_kernel<false, 4, NT, false>)) ST_SETATTR((temporal2_kernel<true, 4, NT, false>))      \
    ST_SETATTR((temporal2_kernel<false, 2, NT, true>)) ST_SETATTR((temporal2_kernel<true, 2, NT, true>))        \
    ST_SETATTR((temporal2_kernel<false, 4, NT, true>)) ST_SETATTR((temporal2_kernel<true, 4, NT, true>))
    ST_SETATTR2(32) ST_SETATTR2(64) ST_SETATTR2(128)
#undef ST_SETATTR2
#undef ST_SETATTR
    s.ok = 1;
    return 0;
}

static inline bool stream_kernel_supported(const StreamState &s, int T) { return s.ok && T >= 1 && T <= s.max_batch; }

// Launches the temporal pass (stream st1) and act + dst (stream st2, after `ev_f1`) for frames
// timer0 .. timer0+T-1 (dy indices dy0 ..).  The split lets the spatial passes of batch k (ALU-bound bit
// logic, list-driven dst) run beside the temporal pass of batch k+1, which leaves a third of the ALU pipe
// and most of the register file idle.  `parity` selects the predicate-bit buffer.
// Returns 0 / -1; *launches gets the number of kernel launches.
static inline int stream_kernel_launch(StreamState &s, FrameSrc src, long long timer0, long long dy0, int T,
                                       int dy_on, const int *d_thr, ActRing ring, uint8_t *dst, uint32_t *dstbits,
                                       unsigned *npoints, uint32_t *points, int cap, SparseLists sl, cudaStream_t st1,
                                       cudaStream_t st2, cudaEvent_t ev_f1, cudaEvent_t ev_d0, int parity,
                                       int *launches, bool act_only = false, bool skip_temporal = false) {
    uint32_t *const bits = parity ? s.d_bits2 : s.d_bits;
    const int HWG = (int)((size_t)s.W * s.H / (4 * s.t_wpt));  // pixel groups = threads
    const int nt = s.t_threads;
    const size_t smem = stream_temporal_smem(s, nt, T, s.t_version == 1 ? 1 : 2);
    const int grid = (HWG + nt - 1) / nt;
    uint8_t *bits8 = reinterpret_cast<uint8_t *>(bits);
    int t3rc = -2;
    if (skip_temporal) {  // the predicate bits are already in the buffer (per-frame O(1) path, perframe_kernel.cuh)
        t3rc = 0;
    } else if (s.t_version == 3) {
        t3rc = temporal3_launch(s.n, s.t3_variant, src, timer0, T, (int)((size_t)s.W * s.H / 8), d_thr, bits8, s.t3tab, parity, st1);
        if (t3rc == -1) return -1;
    }
    s.t_last = skip_temporal ? 4 : t3rc == 0 ? 3 : (s.t_version == 1 ? 1 : 2);
    if (t3rc == 0) {
    } else if (s.t_version != 1) {
#define T2_LAUNCH(M, WP, NT)                                                                                            \
    do {                                                                                                                \
        if (s.t_kdiv > 1) temporal2_kernel<M, WP, NT, true><<<grid, NT, smem, st1>>>(src, timer0, T, s.n, s.t_kdiv, HWG, d_thr, bits8); \
        else temporal2_kernel<M, WP, NT, false><<<grid, NT, smem, st1>>>(src, timer0, T, s.n, 1, HWG, d_thr, bits8);   \
    } while (0)
#define T2_NT(M, WP)                                              \
    do {                                                          \
        if (nt == 32) T2_LAUNCH(M, WP, 32);                       \
        else if (nt == 64) T2_LAUNCH(M, WP, 64);                  \
        else T2_LAUNCH(M, WP, 128);                               \
    } while (0)
        if (s.t_wpt == 2) { if (src.mask) T2_NT(true, 2); else T2_NT(false, 2); }
        else { if (src.mask) T2_NT(true, 4); else T2_NT(false, 4); }
#undef T2_NT
#undef T2_LAUNCH
    } else if (s.t_wpt == 2) {
        if (src.mask) temporal_kernel<true, 2><<<grid, nt, smem, st1>>>(src, timer0, T, s.n, HWG, d_thr, bits8);
        else temporal_kernel<false, 2><<<grid, nt, smem, st1>>>(src, timer0, T, s.n, HWG, d_thr, bits8);
    } else {
        if (src.mask) temporal_kernel<true, 4><<<grid, nt, smem, st1>>>(src, timer0, T, s.n, HWG, d_thr, bits8);
        else temporal_kernel<false, 4><<<grid, nt, smem, st1>>>(src, timer0, T, s.n, HWG, d_thr, bits8);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    if (cudaEventRecord(ev_f1, st1) != cudaSuccess) return -1;
    if (cudaStreamWaitEvent(st2, ev_f1, 0) != cudaSuccess) return -1;
    if (cudaEventRecord(ev_d0, st2) != cudaSuccess) return -1;
    if (cudaMemsetAsync(sl.acount, 0, (size_t)T * sizeof(unsigned), st2) != cudaSuccess) return -1;
    const int Wb = s.W / 32;
    const int strips = (Wb + SP_USE - 1) / SP_USE;
    if (Wb % 4 == 0 && !s.force_strip) {
        // a single frame (per-frame API) is latency-bound: short bands give the one frame enough CTAs to cover the SMs
        const int rows = T == 1 ? s.sp_rows_single : s.sp_rows;
        const int chunks = Wb / 4, bands = (s.H + rows - 1) / rows;
        dim3 g((chunks * bands + A4_THREADS - 1) / A4_THREADS, T);
        act4_kernel<<<g, A4_THREADS, 0, st2>>>(bits, s.H, Wb, rows, chunks, bands, ring, dy0, sl);
    } else {
        const int bands = (s.H + s.sp_rows - 1) / s.sp_rows;
        dim3 g((strips * bands + SP_WARPS - 1) / SP_WARPS, T);
        act_kernel<<<g, SP_WARPS * 32, 0, st2>>>(bits, s.W, s.H, T, s.sp_rows, strips, bands, ring, dy0, sl);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    if (act_only) {  // halo frames of a time-sharded run: only the window and the act history are wanted
        *launches = 2;
        return 0;
    }
    if (cudaMemsetAsync(sl.dense, 0, sizeof(unsigned), st2) != cudaSuccess) return -1;
    if (s.force_dense || (T == 1 && s.single_dense)) {  // test hooks: every frame takes the full-scan path
        dst_force_dense_kernel<<<(T + 127) / 128, 128, 0, st2>>>(T, sl);
    } else {
        dst_sparse_kernel<<<T, 256, 0, st2>>>(ring, s.W, s.H, s.n, dy0, dy_on, dst, dstbits, npoints, points, cap, sl);
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    const int dbands = (s.H + s.dst_rows - 1) / s.dst_rows;
    dim3 gd((strips * dbands + SP_WARPS - 1) / SP_WARPS, std::min(T, DENSE_GY));
    dst_dense_kernel<<<gd, SP_WARPS * 32, 0, st2>>>(ring, s.W, s.H, s.n, dy0, dy_on, s.dst_rows, strips, dbands, dst,
                                                    dstbits, npoints, points, cap, sl);
    if (cudaGetLastError() != cudaSuccess) return -1;
    *launches = 4;
    return 0;
}
