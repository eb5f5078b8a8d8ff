// Time-tiled streaming path for whole batches (the throughput path), sm_100a.
//
// The reference recomputes max/mean over the n-frame ring for every frame (utils.py:269-307), i.e.
// (n+2) bytes of traffic per pixel per frame.  Here every frame is read from HBM ONCE:
//
//   temporal3_kernel / temporal2_kernel (stack -> diff -> threshold, fused; temporal3_kernel.cuh, temporal_kernel.cuh)
//       read the u8 frame, write 1 bit / pixel.  A thread owns 8 (or 16) consecutive pixels for the whole batch and
//       keeps its last n frames in registers (+ a shared-memory page) or, second generation, in a private
//       shared-memory ring fed by cp.async; the sliding max is van Herk / Gil-Werman, the sliding sum a running
//       u16x2 register.  The decision  median(max - floor(sum/L)) > thr  is rewritten as
//       "at least 5 of the 9 neighbours have  max*L - sum > thr*L",  so these kernels only emit the
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
                            // that are not contiguous); 2: temporal2_kernel (temporal_kernel.cuh)
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
static inline size_t stream_temporal_smem(const StreamState &s, int nt, int T) {
    const int slots = s.n + ST_K + s.n / s.t_kdiv;
    return (size_t)slots * 4 * s.t_wpt * nt + (((size_t)2 * T + 15) & ~(size_t)15);
}

// choose words-per-thread of the temporal kernel and derive its CTA size / shared memory
// (nt_req > 0 forces the CTA size; otherwise the largest CTA that keeps the most warps per SM)
static inline int stream_state_config(StreamState &s, int wpt, int nt_req = 0) {
    if (wpt != 2 && wpt != 4) return -1;
    const size_t sm_bytes = 228 * 1024, cta_max = 220 * 1024, reserved = 1024;
    s.t_kdiv = stream_choose_kdiv(s.n, s.t_kdiv_req);
    const size_t per_thread = (size_t)(s.n + ST_K + s.n / s.t_kdiv) * 4 * wpt;
    const size_t table = ((size_t)2 * s.max_batch + 15) & ~(size_t)15;
    int best_nt = 0;
    size_t best_warps = 0;
    for (int nt = 128; nt >= 32; nt >>= 1) {
        if (nt_req && nt != nt_req) continue;
        const size_t cta = per_thread * nt + table;
        if (cta > cta_max) continue;
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
    const size_t smem = stream_temporal_smem(s, nt, T);
    const int grid = (HWG + nt - 1) / nt;
    uint8_t *bits8 = reinterpret_cast<uint8_t *>(bits);
    int t3rc = -2;
    if (skip_temporal) {  // the predicate bits are already in the buffer (per-frame O(1) path, perframe_kernel.cuh)
        t3rc = 0;
    } else if (s.t_version == 3) {
        t3rc = temporal3_launch(s.n, s.t3_variant, src, timer0, T, (int)((size_t)s.W * s.H / 8), d_thr, bits8, s.t3tab, parity, st1);
        if (t3rc == -1) return -1;
    }
    s.t_last = skip_temporal ? 4 : t3rc == 0 ? 3 : 2;
    if (t3rc == 0) {
    } else {
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
