// libmetdet_b200.so -- host side of the C ABI declared in include/metdet_b200.h.
// One handle = one detector (M3Detector, MetLib/Detector.py:302-392) with its device state:
// frame ring, dy-mask run counters, EMA/threshold scalars, per-batch outputs, Hough slots.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "hough.cuh"
#include "classic.cuh"
#include "perframe_kernel.cuh"
#include "mfnr.cuh"
#include "kernels_basic.cuh"
#include "preproc.cuh"
#include "stream_kernel.cuh"

#define HOUGH_SMEM_BYTES (MDB_POINT_CAP * 8 + MDB_POINT_CAP / 8)  // keys u32 + order u16 + line u16 + removed bits
#define HOUGH_SMEM_SMALL HOUGH1_POINT_BYTES(HOUGH_CAP_SMALL)   // tier 1a
#define HOUGH_SMEM_LARGE HOUGH1_POINT_BYTES(HOUGH_CAP_LARGE)   // tier 1b

// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(MDB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

#define NCTX 3  // batches that may be in flight (submit .. collect) per handle

// host-visible results of one batch (pinned) + its completion events
struct BatchCtx {
    int *h_thr = nullptr, *h_nlines = nullptr;
    double *h_thrf = nullptr, *h_snr = nullptr;
    unsigned *h_npoints = nullptr;
    unsigned *h_tiers = nullptr;     // [4] frames resolved by PPHT tiers 1a, 1b, 3, 2
    int32_t *h_lines = nullptr;
    int *d_thr = nullptr;            // per-frame thresholds of this batch (device)
    double *d_thrf = nullptr, *d_snr = nullptr;
    uint8_t *d_dst = nullptr;        // [T][H][W] masks of this batch (persistent: zero where dstbits is zero)
    uint32_t *d_dstbits = nullptr;   // [T][H][Wb] what d_dst currently holds, 1 bit per pixel (streaming path)
    bool dst_dirty = false;          // the generic kernel rewrote d_dst without maintaining d_dstbits
    uint32_t *d_alist = nullptr;     // [T][SPX_ACAP] non-zero act words per frame (streaming path)
    unsigned *d_acount = nullptr;    // [T]
    uint32_t *d_wlist = nullptr;     // [T][SPX_WCAP] non-zero words of each d_dst slot
    unsigned *d_wcount = nullptr;    // [T]
    unsigned *d_dense = nullptr;     // [1 + T] frames whose lists overflowed
    unsigned *d_npoints = nullptr;   // [T] on-pixel counts
    uint32_t *d_points = nullptr;    // [T][cap] on-pixel lists
    uint16_t *d_order = nullptr;     // [T][cap] PPHT visiting order per frame
    unsigned *d_queue = nullptr;     // work-queue head of the tier-1 Hough kernel
    int32_t *d_lines = nullptr;      // [T][MDB_MAX_LINES][4]
    int *d_nlines = nullptr;         // [T]
    // ev_f0..ev_f1: temporal pass on the front stream; ev_d0..ev_d1: act + dst on the back stream
    cudaEvent_t ev_f0 = nullptr, ev_f1 = nullptr, ev_d0 = nullptr, ev_d1 = nullptr, ev_done = nullptr;
    cudaEvent_t ev_thr = nullptr;    // thresholds of this batch are on the device (scalar stream)
    cudaEvent_t ev_src = nullptr;    // the scalar stream is done with this batch's frames (noise samples, history copy)
    cudaEvent_t tl[12] = {};         // optional timeline marks (debug)
    // thr_float | snr | thr | nlines | npoints | queue | lines live in ONE device block and one pinned mirror, so that a
    // batch's results come back in a single copy (the pointers above point into these)
    uint8_t *d_res = nullptr, *h_res = nullptr;
    size_t res_head = 0;  // bytes in front of the line rows
    int T = 0;  // 0: free
    int bits_parity = 0;  // which predicate-bit buffer this batch uses
    bool halo = false;    // results not wanted (look-back frames of a time-sharded run): no dst, no Hough
    long long timer0 = 0;
    long long seq = 0;
};

struct mdb_detector {
    mdb_config cfg;
    int W, H, n, R;
    size_t HW;
    int slots;
    int slots3 = 1;  // tier-3 scratch slots (dense frames worked on side by side)
    int sm_count = 148;
    // stream: noise, thresholds, temporal, act | stream2: dst | stream3: Hough, result copy-out |
    // cstream: host->device frames.  The stages of consecutive batches overlap (three batches in flight):
    // the write-only dst pass and the shared-memory-bound Hough pass hide under the ALU-bound temporal pass.
    // sstream ("scalar"): noise samples + threshold recurrence + history copy of batch k+1 run beside batch k's temporal pass
    cudaStream_t stream = nullptr, stream2 = nullptr, stream3 = nullptr, cstream = nullptr, sstream = nullptr;
    cudaEvent_t ev_copy = nullptr, ev_front = nullptr;
    bool front_dirty = false;        // the per-frame API touched the scalar state on the front stream
    // device
    uint8_t *d_ring = nullptr, *d_mask = nullptr;
    uint8_t *d_stage[2] = {nullptr, nullptr};  // host-fed batches land here (contiguous), alternating; allocated on first use
    uint32_t *d_act = nullptr;  // act bit-frame ring [RA][H][Wb]
    int RA = 0, Wb = 0;
    DevState *d_state = nullptr;
    unsigned long long *d_noise = nullptr;
    unsigned long long *d_noise2 = nullptr;  // mdb_noise_sums_dev: [max_batch + n][2]
    unsigned long long *h_noise2 = nullptr;  // pinned staging of the same size
    size_t noise2_cap = 0;
    int32_t *d_accum = nullptr;   // tier-2/3 accumulators [slots][180][numrho]
    uint32_t *d_bitmap = nullptr, *d_walk = nullptr, *d_okeys = nullptr, *d_oidx = nullptr;  // tier 3
    long long *d_prof = nullptr;  // optional per-frame PPHT phase cycle counters (debug)
    StreamState sk;
    uint32_t *d_cbits = nullptr;  // ClassicDetector: three bit planes [3][max_batch][H][Wb] (a, b, dst)
    BatchCtx ctx[NCTX];
    // host state
    long long timer = 0, dy_timer = 0, seek0 = 0;
    long long launches = 0;
    long long submitted = 0, collected = 0;  // batch sequence numbers
    int last_T = 0;                           // frames in the most recently finished batch
    int last_ctx = 0;                         // ... and which context holds its dst
    float fused_ms = 0.f, temporal_ms = 0.f, spatial_ms = 0.f;
    unsigned tiers_last[4] = {0, 0, 0, 0};        // PPHT tier statistics of the batch collected last (1a, 1b, 3, 2)
    unsigned long long tiers_total[4] = {0, 0, 0, 0};
    int fused_launches = 0, last_fused_launches = 0;
    HoughParams hp;
    int use_stream_kernel = 1;
    // per-frame O(1) path (perframe_kernel.cuh): resident window state, valid for `pf_timer` frames seen
    bool single_stream = false;      // set around the per-frame chain: everything on the front stream
    cudaEvent_t ev_suffix = nullptr; // the suffix planes of the block that just ended are rebuilt (back stream)
    cudaEvent_t ev_pf_a = nullptr, ev_pf_b = nullptr, ev_pf_done = nullptr;  // per-frame copy halves / staging buffer free
    unsigned long long last_digest = 0;  // of the batch finished last (finish_batch)
    int hough_ctas = 0;              // > 0: grid cap of the shared-memory PPHT tiers
    int per_frame_fast = 1;
    uint16_t *d_pf_sum = nullptr;
    uint8_t *d_pf_pmax = nullptr, *d_pf_suf = nullptr, *d_pf_stage = nullptr;
    long long pf_timer = -1;         // frames the state has absorbed (-1: rebuild from the ring before use)
    long long pf_bits_timer = -1;    // the predicate bits in sk.d_bits belong to frame pf_bits_timer - 1
    bool pf_suffix_pending = false;  // the block that just ended still needs its suffix planes
    bool pf_suffix_wait = false;     // ... they are being rebuilt: the next update kernel waits for ev_suffix
    int timeline = 0;                // debug: record per-kernel timeline events
    cudaEvent_t tl_base = nullptr;
};

#define TL(c, i, st)                                                  \
    do {                                                              \
        if (h->timeline) cudaEventRecord((c).tl[i], st);              \
    } while (0)

extern "C" const char *mdb_last_error(void) { return g_err; }
extern "C" int mdb_version(void) { return 102; }
extern "C" int mdb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

static FrameSrc frame_src(const mdb_detector *h, const uint8_t *cur, long long t0) {
    FrameSrc s;
    s.ring = h->d_ring;
    s.cur = cur;
    s.t0 = t0;
    s.mask = h->cfg.apply_mask ? h->d_mask : nullptr;
    s.R = h->R;
    s.HW = h->HW;
    return s;
}

static ActRing act_ring(const mdb_detector *h) {
    ActRing r;
    r.base = h->d_act;
    r.RA = h->RA;
    r.Wb = h->Wb;
    r.frame_words = (size_t)h->H * h->Wb;
    return r;
}

static void free_all(mdb_detector *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->stream2) cudaStreamSynchronize(h->stream2);
    if (h->stream3) cudaStreamSynchronize(h->stream3);
    if (h->cstream) cudaStreamSynchronize(h->cstream);
    if (h->sstream) cudaStreamSynchronize(h->sstream);
    if (h->h_noise2) cudaFreeHost(h->h_noise2);
    void *dev[] = {h->d_ring, h->d_mask, h->d_stage[0], h->d_stage[1], h->d_act, h->d_state, h->d_noise, h->d_noise2, h->d_accum,
                   h->d_bitmap, h->d_walk, h->d_okeys, h->d_oidx, h->d_prof, h->d_cbits, h->d_pf_sum, h->d_pf_pmax, h->d_pf_suf,
                   h->d_pf_stage};
    for (void *p : dev)
        if (p) cudaFree(p);
    for (BatchCtx &c : h->ctx) {
        void *cd[] = {c.d_res, c.d_dst, c.d_dstbits, c.d_points, c.d_order, c.d_alist, c.d_acount, c.d_wlist, c.d_wcount,
                      c.d_dense};
        for (void *p : cd)
            if (p) cudaFree(p);
    }
    stream_state_free(h->sk);
    for (BatchCtx &c : h->ctx) {
        void *pin[] = {c.h_res};
        for (void *p : pin)
            if (p) cudaFreeHost(p);
        if (c.ev_f0) cudaEventDestroy(c.ev_f0);
        if (c.ev_f1) cudaEventDestroy(c.ev_f1);
        if (c.ev_d0) cudaEventDestroy(c.ev_d0);
        if (c.ev_d1) cudaEventDestroy(c.ev_d1);
        if (c.ev_done) cudaEventDestroy(c.ev_done);
        if (c.ev_thr) cudaEventDestroy(c.ev_thr);
        if (c.ev_src) cudaEventDestroy(c.ev_src);
    }
    if (h->ev_copy) cudaEventDestroy(h->ev_copy);
    if (h->ev_front) cudaEventDestroy(h->ev_front);
    if (h->ev_suffix) cudaEventDestroy(h->ev_suffix);
    for (cudaEvent_t e : {h->ev_pf_a, h->ev_pf_b, h->ev_pf_done})
        if (e) cudaEventDestroy(e);
    if (h->sstream) cudaStreamDestroy(h->sstream);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    if (h->cstream) cudaStreamDestroy(h->cstream);
    delete h;
}

// initial scalar state: LineDetector.__init__ (Detector.py:204-209), SNR_SW.__init__ (:58-61)
static DevState initial_state(const mdb_detector *h) {
    DevState st;
    memset(&st, 0, sizeof st);
    st.ema_init_m = 1.0 - (double)h->cfg.nz_interval / 60.0;
    st.ema_cur_m = st.ema_init_m;
    st.ema_warm = (double)h->n;
    st.ema_value = 0.0;
    static const int abs_sens[3] = {7, 5, 3};  // low, normal, high
    st.bi_threshold = h->cfg.adaptive ? abs_sens[h->cfg.sensitivity] : h->cfg.init_value;
    st.thr_float = (double)st.bi_threshold;
    return st;
}

extern "C" int mdb_create(const mdb_config *cfg, const uint8_t *mask, mdb_handle *out) {
    if (!cfg || !mask || !out) return fail(MDB_ERR_INVALID, "mdb_create: null argument");
    *out = nullptr;
    if (cfg->width < 1 || cfg->height < 1 || cfg->width > 65535 || cfg->height > 65535)
        return fail(MDB_ERR_INVALID, "mdb_create: unsupported frame size %dx%d", cfg->width, cfg->height);
    if (cfg->window < 1 || cfg->window > MDB_MAX_WINDOW)
        return fail(MDB_ERR_INVALID, "mdb_create: window n=%d outside 1..%d", cfg->window, MDB_MAX_WINDOW);
    if (cfg->max_batch < 1) return fail(MDB_ERR_INVALID, "mdb_create: max_batch must be >= 1");
    if (cfg->sensitivity < 0 || cfg->sensitivity > 2)
        return fail(MDB_ERR_INVALID, "mdb_create: bad sensitivity %d", cfg->sensitivity);
    const int r0 = cfg->roi[0], c0 = cfg->roi[1], r1 = cfg->roi[2], c1 = cfg->roi[3];
    if (r0 < 0 || c0 < 0 || r1 > cfg->height || c1 > cfg->width || r1 <= r0 || c1 <= c0)
        return fail(MDB_ERR_INVALID, "mdb_create: bad std_roi (%d,%d,%d,%d)", r0, c0, r1, c1);
    if (cfg->nz_interval < 0) return fail(MDB_ERR_INVALID, "mdb_create: negative interval");
    int ndev = mdb_device_count();
    if (ndev == 0)
        return fail(MDB_ERR_CUDA, "mdb_create: no CUDA device -- this library has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(MDB_ERR_INVALID, "mdb_create: device %d of %d", cfg->device, ndev);
    CK(cudaSetDevice(cfg->device));

    mdb_detector *h = new (std::nothrow) mdb_detector();
    if (!h) return fail(MDB_ERR_NOMEM, "mdb_create: out of host memory");
    h->cfg = *cfg;
    if (cfg->detector != 0 && cfg->detector != 1) { delete h; return fail(MDB_ERR_INVALID, "mdb_create: unknown detector kind %d", cfg->detector); }
    if (cfg->detector == 1) { h->cfg.window = 4; h->cfg.dy_mask = 0; }  // ClassicDetector.classic_max_size, Detector.py:249-254
    h->W = cfg->width; h->H = cfg->height; h->n = h->cfg.window;
    h->HW = (size_t)h->W * h->H;
    const int T = cfg->max_batch;
    // ring: history (n-1) + two batches, so the copy of batch k+1 never lands on frames batch k reads
    h->R = h->n - 1 + (T > 1 ? 2 * T : 1);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, cfg->device);
    if (e != cudaSuccess) { delete h; return fail(MDB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    h->sm_count = prop.multiProcessorCount;

    HoughParams &hp = h->hp;
    hp.W = h->W; hp.H = h->H; hp.numrho = 2 * (h->W + h->H) + 1;
    hp.threshold = cfg->hough_threshold; hp.min_len = cfg->hough_min_len; hp.max_gap = cfg->hough_max_gap;
    unsigned long long area = 0;
    for (size_t i = 0; i < h->HW; i++) area += mask[i];
    hp.mask_area = (double)area;
    hp.cap = MDB_POINT_CAP; hp.max_lines = MDB_MAX_LINES; hp.walk_cap = h->W + h->H + 2;
    hp.fixed_gap = cfg->detector == 1 ? cfg->hough_max_gap : -1;
    // tier-2 Hough slots (global-memory accumulators): one per frame in flight, 4 GB budget
    {
        const size_t per_slot = (size_t)MDB_HOUGH_ANGLES * hp.numrho * sizeof(int32_t);
        size_t s = std::min<size_t>((size_t)T, (size_t)prop.multiProcessorCount * 2);
        s = std::min<size_t>(s, std::max<size_t>(1, (4ull << 30) / per_slot));
        h->slots = (int)s;
    }

#define ALLOC(ptr, bytes)                                                                   \
    do {                                                                                    \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));                              \
        if (e_ != cudaSuccess) {                                                            \
            int rc_ = fail(MDB_ERR_NOMEM, "cudaMalloc(%zu bytes) for %s: %s", (size_t)(bytes), #ptr, \
                           cudaGetErrorString(e_));                                         \
            free_all(h);                                                                    \
            return rc_;                                                                     \
        }                                                                                   \
    } while (0)
#define CKH(call)                                                                           \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            int rc_ = fail(MDB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));          \
            free_all(h);                                                                    \
            return rc_;                                                                     \
        }                                                                                   \
    } while (0)

    {   // the back half of batch k (dst, Hough: short, latency-bound) must not queue behind the thousands of
        // CTAs of batch k+1's temporal pass: it runs at a higher stream priority
        int prio_lo = 0, prio_hi = 0;
        CKH(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CKH(cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, prio_lo));
        CKH(cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_hi));
        CKH(cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, prio_hi));
        CKH(cudaStreamCreateWithPriority(&h->cstream, cudaStreamNonBlocking, prio_lo));
        CKH(cudaStreamCreateWithPriority(&h->sstream, cudaStreamNonBlocking, prio_hi));
    }
    CKH(cudaEventCreateWithFlags(&h->ev_front, cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_copy, cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_suffix, cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_pf_a, cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_pf_b, cudaEventDisableTiming));
    CKH(cudaEventCreateWithFlags(&h->ev_pf_done, cudaEventDisableTiming));
    const size_t bm_words = (h->HW + 31) / 32;
    h->Wb = (h->W + 31) / 32;
    h->RA = h->n - 1 + (T > 1 ? 2 * T : 1);  // two batches: act of batch k+1 is written while dst of batch k reads
    ALLOC(h->d_ring, (size_t)h->R * h->HW);
    ALLOC(h->d_mask, h->HW);
    ALLOC(h->d_act, (size_t)h->RA * h->H * h->Wb * sizeof(uint32_t));
    ALLOC(h->d_state, sizeof(DevState));
    ALLOC(h->d_noise, (size_t)T * 2 * sizeof(unsigned long long));
    for (int k = 0; k < NCTX; k++) {
        BatchCtx &c = h->ctx[k];
        ALLOC(c.d_dst, (size_t)T * h->HW);
        ALLOC(c.d_points, (size_t)T * MDB_POINT_CAP * sizeof(uint32_t));
        ALLOC(c.d_order, (size_t)T * HOUGH_ORDER_CAP * sizeof(uint16_t));
        {   // result block: [thr_float T][snr T][thr T][nlines T][npoints T][queue 8][pad][lines T*MDB_MAX_LINES*4]
            const size_t o_thrf = 0, o_snr = o_thrf + (size_t)T * 8, o_thr = o_snr + (size_t)T * 8, o_nl = o_thr + (size_t)T * 4,
                         o_np = o_nl + (size_t)T * 4, o_q = o_np + (size_t)T * 4;
            c.res_head = (o_q + 8 * sizeof(unsigned) + 15) / 16 * 16;
            const size_t total = c.res_head + (size_t)T * MDB_MAX_LINES * 4 * sizeof(int32_t);
            ALLOC(c.d_res, total);
            CKH(cudaMemsetAsync(c.d_res, 0, total, h->stream));
            CKH(cudaHostAlloc((void **)&c.h_res, total, cudaHostAllocDefault));
            memset(c.h_res, 0, total);
            c.d_thrf = (double *)(c.d_res + o_thrf); c.h_thrf = (double *)(c.h_res + o_thrf);
            c.d_snr = (double *)(c.d_res + o_snr);   c.h_snr = (double *)(c.h_res + o_snr);
            c.d_thr = (int *)(c.d_res + o_thr);      c.h_thr = (int *)(c.h_res + o_thr);
            c.d_nlines = (int *)(c.d_res + o_nl);    c.h_nlines = (int *)(c.h_res + o_nl);
            c.d_npoints = (unsigned *)(c.d_res + o_np); c.h_npoints = (unsigned *)(c.h_res + o_np);
            // queue: [0..2] work-queue heads of tiers 1a, 1b, 3; [4..7] frames resolved by tiers 1a, 1b, 3, 2
            c.d_queue = (unsigned *)(c.d_res + o_q);  c.h_tiers = (unsigned *)(c.h_res + o_q) + 4;
            c.d_lines = (int32_t *)(c.d_res + c.res_head); c.h_lines = (int32_t *)(c.h_res + c.res_head);
        }
        CKH(cudaMemsetAsync(c.d_dst, 0, (size_t)T * h->HW, h->stream));
        ALLOC(c.d_dstbits, (size_t)T * h->H * h->Wb * sizeof(uint32_t));
        CKH(cudaMemsetAsync(c.d_dstbits, 0, (size_t)T * h->H * h->Wb * sizeof(uint32_t), h->stream));
        ALLOC(c.d_alist, (size_t)T * SPX_ACAP * sizeof(uint32_t));
        ALLOC(c.d_acount, T * sizeof(unsigned));
        ALLOC(c.d_wlist, (size_t)T * SPX_WCAP * sizeof(uint32_t));
        ALLOC(c.d_wcount, T * sizeof(unsigned));
        ALLOC(c.d_dense, (size_t)(T + 1) * sizeof(unsigned));
        CKH(cudaMemsetAsync(c.d_acount, 0, T * sizeof(unsigned), h->stream));
        CKH(cudaMemsetAsync(c.d_wcount, 0, T * sizeof(unsigned), h->stream));
        CKH(cudaMemsetAsync(c.d_dense, 0, (size_t)(T + 1) * sizeof(unsigned), h->stream));
    }
    if (cfg->detector == 1) ALLOC(h->d_cbits, (size_t)3 * T * h->H * h->Wb * sizeof(uint32_t));
    ALLOC(h->d_accum, (size_t)h->slots * MDB_HOUGH_ANGLES * hp.numrho * sizeof(int32_t));
    // tier-3 scratch slots: point list + visiting order (2 x HW words each), 6 GB at most
    h->slots3 = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(std::min(h->slots, T), (size_t)h->sm_count), (12ull << 30) / (h->HW * 8)));
    ALLOC(h->d_bitmap, (size_t)h->slots3 * bm_words * sizeof(uint32_t));
    ALLOC(h->d_walk, (size_t)h->slots3 * hp.walk_cap * sizeof(uint32_t));
    CKH(cudaMemsetAsync(h->d_ring, 0, (size_t)h->R * h->HW, h->stream));
    CKH(cudaMemsetAsync(h->d_act, 0, (size_t)h->RA * h->H * h->Wb * sizeof(uint32_t), h->stream));
    CKH(cudaMemsetAsync(h->d_accum, 0, (size_t)h->slots * MDB_HOUGH_ANGLES * hp.numrho * sizeof(int32_t), h->stream));
    CKH(cudaMemsetAsync(h->d_bitmap, 0, (size_t)h->slots3 * bm_words * sizeof(uint32_t), h->stream));
    CKH(cudaMemcpyAsync(h->d_mask, mask, h->HW, cudaMemcpyHostToDevice, h->stream));
    for (BatchCtx &c : h->ctx) {
        CKH(cudaEventCreate(&c.ev_f0));
        CKH(cudaEventCreate(&c.ev_f1));
        CKH(cudaEventCreate(&c.ev_d0));
        CKH(cudaEventCreate(&c.ev_d1));
        CKH(cudaEventCreateWithFlags(&c.ev_done, cudaEventDisableTiming));
        CKH(cudaEventCreateWithFlags(&c.ev_thr, cudaEventDisableTiming));
        CKH(cudaEventCreateWithFlags(&c.ev_src, cudaEventDisableTiming));
    }

    const DevState st = initial_state(h);
    CKH(cudaMemcpyAsync(h->d_state, &st, sizeof st, cudaMemcpyHostToDevice, h->stream));

    // trig table exactly as OpenCV builds it: (float)cos((double)n * (double)(float)theta)
    float trig[2 * MDB_HOUGH_ANGLES];
    const float theta = (float)(3.14159265358979323846 / 180.0);
    for (int k = 0; k < MDB_HOUGH_ANGLES; k++) {
        trig[2 * k] = (float)cos((double)k * (double)theta);
        trig[2 * k + 1] = (float)sin((double)k * (double)theta);
    }
    CKH(cudaMemcpyToSymbolAsync(c_trig, trig, sizeof trig, 0, cudaMemcpyHostToDevice, h->stream));
    CKH(cudaFuncSetAttribute(hough_tier2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             HOUGH_SMEM_BYTES));
    CKH(cudaFuncSetAttribute(hough_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             HOUGH_SMEM_LARGE + HOUGH_TABLE_BYTES));
    CKH(cudaFuncSetAttribute(hough_tier3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, H3_ORDER_SMEM));
    CKH(cudaStreamSynchronize(h->stream));
    {
        int rc = cfg->detector == 1 ? 0 : stream_state_init(h->sk, h->W, h->H, h->n, cfg->device, cfg->max_batch);
        if (rc != 0) { free_all(h); return fail(MDB_ERR_CUDA, "stream kernel init failed: %s", cudaGetErrorString(cudaGetLastError())); }
    }
    *out = h;
    return MDB_OK;
}

extern "C" int mdb_destroy(mdb_handle h) {
    if (!h) return fail(MDB_ERR_INVALID, "mdb_destroy: null handle");
    free_all(h);
    return MDB_OK;
}

// ---- frames into ring slots (global frame index t0 .. t0+T-1), on stream `st` ----------------
static int copy_to_ring(mdb_detector *h, const uint8_t *frames, int first, int T, long long t0,
                        cudaMemcpyKind kind, cudaStream_t st) {
    long long t = t0 + first;
    int done = first;
    while (done < T) {
        const int slot = (int)(t % h->R);
        const int run = std::min(T - done, h->R - slot);
        CK(cudaMemcpyAsync(h->d_ring + (size_t)slot * h->HW, frames + (size_t)done * h->HW,
                           (size_t)run * h->HW, kind, st));
        done += run;
        t += run;
    }
    return MDB_OK;
}

static int launch_noise_thr(mdb_detector *h, BatchCtx &bc, const FrameSrc &src, int T, long long timer0,
                            cudaStream_t st) {
    const mdb_config &c = h->cfg;
    const int rh = c.roi[2] - c.roi[0], rw = c.roi[3] - c.roi[1];
    const long long std_interval = (long long)c.nz_interval * h->n;
    TL(bc, 0, st);
    CK(cudaMemsetAsync(h->d_noise, 0, (size_t)T * 2 * sizeof(unsigned long long), st));
    SampleList sl;
    sl.count = 0;
    for (int i = 0; i < T && sl.count >= 0; i++) {
        const long long tau = timer0 + i + 1;
        if ((tau > 1 && tau <= h->n) || (tau > h->n && std_interval > 0 && tau % std_interval == 0)) {
            if (sl.count < 63) sl.idx[sl.count++] = i;
            else sl.count = -1;  // too many to list: one grid row per frame
        }
    }
    if (sl.count != 0) launch_noise_samples(src, h->W, h->n, timer0, std_interval, c.roi, h->d_noise, 0, sl, T, st);
    threshold_kernel<<<1, 32, 0, st>>>(h->d_state, h->d_noise, T, timer0, h->n, std_interval,
                                             (long long)rh * rw, c.adaptive, c.sensitivity, bc.d_thr,
                                             bc.d_thrf, bc.d_snr);
    h->launches += 2;
    TL(bc, 1, st);
    CK(cudaGetLastError());
    return MDB_OK;
}

// fused mask chain for frames i = 0..T-1 of the batch (global index timer0 + i): temporal pass on the
// front stream, dst on the back stream
static int launch_fused(mdb_detector *h, BatchCtx &c, const FrameSrc &src, int T, long long timer0, long long dy0,
                        bool bits_ready = false) {
    // per-frame calls (single_stream) keep the whole chain on the front stream: no cross-stream hand-overs on the latency path
    cudaStream_t s2 = h->single_stream ? h->stream : h->stream2;
    CK(cudaMemsetAsync(c.d_npoints, 0, T * sizeof(unsigned), s2));
    CK(cudaEventRecord(c.ev_f0, h->stream));
    int nl = 0;
    if (h->cfg.detector == 1) {
        // ClassicDetector: bit planes on the front stream (after the thresholds), expansion on the back stream
        const size_t plane = (size_t)h->cfg.max_batch * h->H * h->Wb;
        uint32_t *ab = h->d_cbits, *bb = h->d_cbits + plane, *db = h->d_cbits + 2 * plane;
        const int nthreads = h->H * h->Wb;
        if (h->W % 16 == 0)
            classic_bits_kernel<true><<<(nthreads + 255) / 256, 256, 0, h->stream>>>(src, h->W, h->H, h->Wb, timer0, T, c.d_thr, ab, bb);
        else
            classic_bits_kernel<false><<<(nthreads + 255) / 256, 256, 0, h->stream>>>(src, h->W, h->H, h->Wb, timer0, T, c.d_thr, ab, bb);
        const int rows = 64, strips = (h->Wb + SP_USE - 1) / SP_USE, bands = (h->H + rows - 1) / rows;
        classic_spatial_kernel<<<dim3((strips * bands + SP_WARPS - 1) / SP_WARPS, T), SP_WARPS * 32, 0, h->stream>>>(
            ab, bb, h->H, h->Wb, rows, strips, bands, db);
        CK(cudaGetLastError());
        CK(cudaEventRecord(c.ev_f1, h->stream));
        CK(cudaStreamWaitEvent(s2, c.ev_f1, 0));
        CK(cudaEventRecord(c.ev_d0, s2));
        c.dst_dirty = true;
        if (h->W % 16 == 0)
            classic_expand_kernel<true><<<dim3((nthreads + 255) / 256, T), 256, 0, s2>>>(
                db, h->W, h->H, h->Wb, c.d_dst, c.d_npoints, c.d_points, MDB_POINT_CAP);
        else
            classic_expand_kernel<false><<<dim3((nthreads + 255) / 256, T), 256, 0, s2>>>(
                db, h->W, h->H, h->Wb, c.d_dst, c.d_npoints, c.d_points, MDB_POINT_CAP);
        nl = 3;
    } else if (h->use_stream_kernel && stream_kernel_supported(h->sk, T)) {
        if (c.dst_dirty) {  // the generic kernel wrote this buffer last: resynchronise buffer and bitmap
            CK(cudaMemsetAsync(c.d_dst, 0, (size_t)h->cfg.max_batch * h->HW, s2));
            CK(cudaMemsetAsync(c.d_dstbits, 0, (size_t)h->cfg.max_batch * h->H * h->Wb * sizeof(uint32_t), s2));
            CK(cudaMemsetAsync(c.d_wcount, 0, (size_t)h->cfg.max_batch * sizeof(unsigned), s2));
            c.dst_dirty = false;
        }
        SparseLists sl;
        sl.alist = c.d_alist; sl.acount = c.d_acount; sl.wlist = c.d_wlist; sl.wcount = c.d_wcount; sl.dense = c.d_dense;
        int rc = stream_kernel_launch(h->sk, src, timer0, dy0, T, h->cfg.dy_mask, c.d_thr, act_ring(h),
                                      c.d_dst, c.d_dstbits, c.d_npoints, c.d_points, MDB_POINT_CAP, sl, h->stream,
                                      s2, c.ev_f1, c.ev_d0, bits_ready ? 0 : (int)(c.bits_parity & 1), &nl, c.halo,
                                      bits_ready);
        if (rc != 0) return fail(MDB_ERR_CUDA, "stream kernel launch: %s", cudaGetErrorString(cudaGetLastError()));
    } else {
        // generic per-frame kernel: the whole chain runs on the back stream, after the front stream's thresholds
        CK(cudaEventRecord(c.ev_f1, h->stream));
        CK(cudaStreamWaitEvent(s2, c.ev_f1, 0));
        CK(cudaEventRecord(c.ev_d0, s2));
        c.dst_dirty = true;
        dim3 grid((h->W + V1_TW - 1) / V1_TW, (h->H + V1_TH - 1) / V1_TH);
        for (int i = 0; i < T; i++) {
            const long long t = timer0 + i;
            const int L = (int)std::min<long long>(h->n, t + 1);
            const int Ldy = (int)std::min<long long>(h->n, dy0 + i + 1);
            fused_frame_kernel<<<grid, 256, 0, s2>>>(
                src, h->W, h->H, h->n, t, L, dy0 + i, Ldy, h->cfg.dy_mask, c.d_thr + i, act_ring(h),
                c.d_dst + (size_t)i * h->HW, c.d_npoints + i, c.d_points + (size_t)i * MDB_POINT_CAP,
                MDB_POINT_CAP);
            nl++;
        }
    }
    CK(cudaEventRecord(c.ev_d1, s2));
    h->fused_launches = nl;
    h->launches += nl;
    CK(cudaGetLastError());
    return MDB_OK;
}

// the PPHT tiers over frames 0..T-1 whose on-pixel lists / masks start at the given pointers; `hp.max_lines` rows of
// `d_lines` per frame.  tl (may be null): batch context whose timeline events are recorded.
static int launch_hough_kernels(mdb_detector *h, BatchCtx *tl, const HoughParams &hp, int T, const unsigned *d_npoints,
                                const uint32_t *d_points, uint16_t *d_order, const uint8_t *d_dst, int32_t *d_lines,
                                int *d_nlines, unsigned *d_queue, cudaStream_t st) {
    CK(cudaMemsetAsync(d_queue, 0, 8 * sizeof(unsigned), st));
    if (tl) TL(*tl, 4, st);
    ppht_order_kernel<<<T, 32, HOUGH_ORDER_CAP * 2, st>>>(T, HOUGH_ORDER_CAP, d_npoints, d_order);
    // tier 1a: 2 CTAs/SM (2048 points, 90 KB table); tier 1b: 1 CTA/SM (4096 points, 184 KB table)
    // the CTAs take frames from a queue, so the grid only sets how much of the GPU the (latency-bound, shared-memory
    // hungry: a resident CTA displaces two of an SM's four temporal CTAs) PPHT occupies beside the next batch's
    // temporal pass.  Measured at 4K (profiles/sweep_hough_ctas3.txt): 0.7 of a CTA per SM costs 1 % of throughput
    // against the optimum (0.75 - 0.9 per SM) and keeps the mask chain at 0.6 of the HBM peak inside the live step
    // (0.5 at the optimum, 0.48 with two per SM); half a CTA per SM costs 8 %
    const int cap_a = h->hough_ctas > 0 ? h->hough_ctas : 7 * h->sm_count / 10, cap_b = h->hough_ctas > 0 ? h->hough_ctas : h->sm_count;
    hough_smem_kernel<<<std::min(T, cap_a), HOUGH_THREADS, HOUGH_SMEM_SMALL + HOUGH_TABLE_BYTES_SMALL, st>>>(
        hp, T, d_npoints, d_points, d_order, d_lines, d_nlines, d_queue, h->d_prof,
        HOUGH_CAP_SMALL, HOUGH_TABLE_BYTES_SMALL, 0);
    hough_smem_kernel<<<std::min(T, cap_b), HOUGH_THREADS, HOUGH_SMEM_LARGE + HOUGH_TABLE_BYTES, st>>>(
        hp, T, d_npoints, d_points, d_order, d_lines, d_nlines, d_queue + 1, h->d_prof,
        HOUGH_CAP_LARGE, HOUGH_TABLE_BYTES, 1);
    if (tl) TL(*tl, 5, st);
    hough_tier2_kernel<<<std::min(T, h->slots), HOUGH_THREADS, HOUGH_SMEM_BYTES, st>>>(
        hp, T, d_npoints, d_points, h->d_accum, d_lines, d_nlines, h->d_prof, d_queue + 7);
    if (!h->d_okeys) {  // tier-3 scratch, allocated once
        if (cudaMalloc((void **)&h->d_okeys, (size_t)h->slots3 * h->HW * sizeof(uint32_t)) != cudaSuccess ||
            cudaMalloc((void **)&h->d_oidx, (size_t)h->slots3 * h->HW * sizeof(uint32_t)) != cudaSuccess)
            return fail(MDB_ERR_NOMEM, "tier-3 scratch: %s", cudaGetErrorString(cudaGetLastError()));
    }
    hough_tier3_kernel<<<h->slots3, HOUGH_THREADS, H3_ORDER_SMEM, st>>>(hp, T, d_dst, h->d_okeys, h->d_oidx, h->d_accum,
                                                                        h->d_bitmap, h->d_walk, d_lines, d_nlines,
                                                                        d_queue + 2, h->d_prof);
    CK(cudaGetLastError());
    return MDB_OK;
}

static int launch_hough_and_copy(mdb_detector *h, BatchCtx &c, int T) {
    cudaStream_t s3 = h->single_stream ? h->stream : h->stream3;
    CK(cudaStreamWaitEvent(s3, c.ev_d1, 0));  // dst (stream2) has produced the on-pixel lists
    if (c.halo) {  // no results wanted: empty masks, no lines
        memset(c.h_tiers, 0, 4 * sizeof(unsigned));
        CK(cudaMemsetAsync(c.d_npoints, 0, T * sizeof(unsigned), s3));
        CK(cudaMemsetAsync(c.d_nlines, 0, T * sizeof(int), s3));
        CK(cudaMemsetAsync(c.d_queue, 0, 8 * sizeof(unsigned), s3));
        CK(cudaMemcpyAsync(c.h_res, c.d_res, c.res_head, cudaMemcpyDeviceToHost, s3));  // scalars only
        CK(cudaEventRecord(c.ev_done, s3));
        return MDB_OK;
    }
    int rc = launch_hough_kernels(h, &c, h->hp, T, c.d_npoints, c.d_points, c.d_order, c.d_dst, c.d_lines, c.d_nlines, c.d_queue,
                                  s3);
    if (rc) return rc;
    h->launches += 5;
    TL(c, 6, s3);
    CK(cudaGetLastError());
    // one copy brings back the scalars of the whole batch and the line rows of its T frames
    CK(cudaMemcpyAsync(c.h_res, c.d_res, c.res_head + (size_t)T * MDB_MAX_LINES * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost,
                       s3));
    CK(cudaEventRecord(c.ev_done, s3));
    TL(c, 7, s3);
    return MDB_OK;
}

// ---- lineset_nms on the host (MetLib/utils.py:780-839) --------------------------------------
static int nms_host(const int32_t *in, int n, int32_t *out, double *prob, const int32_t *given_order = nullptr,
                    int *ties = nullptr) {
    std::vector<long long> len2(n), A(n), B(n), C(n), cx(n), cy(n);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) {
        const long long x1 = in[4 * i], y1 = in[4 * i + 1], x2 = in[4 * i + 2], y2 = in[4 * i + 3];
        len2[i] = (y2 - y1) * (y2 - y1) + (x2 - x1) * (x2 - x1);
        A[i] = y2 - y1; B[i] = x1 - x2; C[i] = x2 * y1 - y2 * x1;
        // numpy floor division of the (non-negative) coordinate sums
        cx[i] = (x2 + x1) >= 0 ? (x2 + x1) / 2 : -((-(x2 + x1) + 1) / 2);
        cy[i] = (y2 + y1) >= 0 ? (y2 + y1) / 2 : -((-(y2 + y1) + 1) / 2);
        order[i] = i;
    }
    if (given_order) {
        for (int i = 0; i < n; i++) order[i] = given_order[i];
    } else {
        std::sort(order.begin(), order.end(), [&](int a, int b) {
            if (len2[a] != len2[b]) return len2[a] > len2[b];
            return a > b;
        });
    }
    if (ties) {
        *ties = 0;
        if (given_order) {
            std::vector<long long> l(len2);
            std::sort(l.begin(), l.end());
            for (int i = 1; i < n; i++) *ties |= l[i] == l[i - 1];
        } else {
            for (int i = 1; i < n; i++) *ties |= len2[order[i]] == len2[order[i - 1]];
        }
    }
    std::vector<char> taken(n, 0);
    int k = 0;
    for (int i = 0; i < n; i++) {
        const int a = order[i];
        if (taken[a]) continue;
        taken[a] = 1;
        long long w = 0;
        for (int j = i; j < n; j++) {
            const int b = order[j];
            if (taken[b]) continue;
            const long long dx = cx[a] - cx[b], dy = cy[a] - cy[b];
            const long long lim = len2[a] / 4;  // len2 >= 0: // == /
            if (dx * dx + dy * dy < lim) {
                taken[b] = 1;
                w = std::max(w, std::llabs(A[a] * cx[b] + B[a] * cy[b] + C[a]));
            }
        }
        memcpy(out + 4 * k, in + 4 * a, 4 * sizeof(int32_t));
        double p = (double)w / std::sqrt((double)(A[a] * A[a] + B[a] * B[a])) / std::sqrt((double)len2[a]) * 2;
        if (p > 1) p = 1;  // NaN (zero-length line) stays NaN, as in numpy
        prob[k] = p;
        k++;
    }
    return k;
}

extern "C" int mdb_lineset_nms(const int32_t *lines_in, int n, int32_t *lines_out, double *prob_out,
                               int32_t *n_out) {
    if (n < 0 || (n > 0 && (!lines_in || !lines_out || !prob_out)) || !n_out)
        return fail(MDB_ERR_INVALID, "mdb_lineset_nms: bad arguments");
    *n_out = n ? nms_host(lines_in, n, lines_out, prob_out) : 0;
    return MDB_OK;
}

extern "C" int mdb_lineset_nms_ordered(const int32_t *lines_in, int n, const int32_t *order, int32_t *lines_out,
                                       double *prob_out, int32_t *n_out) {
    if (n < 0 || (n > 0 && (!lines_in || !order || !lines_out || !prob_out)) || !n_out)
        return fail(MDB_ERR_INVALID, "mdb_lineset_nms_ordered: bad arguments");
    std::vector<char> seen(n, 0);
    for (int i = 0; i < n; i++) {
        if (order[i] < 0 || order[i] >= n || seen[order[i]])
            return fail(MDB_ERR_INVALID, "mdb_lineset_nms_ordered: order is not a permutation of 0..%d", n - 1);
        seen[order[i]] = 1;
    }
    *n_out = n ? nms_host(lines_in, n, lines_out, prob_out, order) : 0;
    return MDB_OK;
}

// lineset_nms redone for some frames of a finished batch with a caller-supplied visiting order per frame (numpy's
// argsort of the squared lengths on the caller's machine, utils.py:804): frames[j] indexes infos / raw_lines / lines /
// prob as mdb_collect_batch filled them; orders holds the permutations back to back, order_off[j] .. order_off[j+1].
extern "C" int mdb_lineset_nms_frames(int k, const int32_t *frames, const int32_t *orders, const int64_t *order_off,
                                      const int32_t *raw_lines, mdb_frame_info *infos, int32_t *lines, double *prob) {
    if (k < 0 || (k > 0 && (!frames || !orders || !order_off || !raw_lines || !infos || !lines || !prob)))
        return fail(MDB_ERR_INVALID, "mdb_lineset_nms_frames: bad arguments");
    std::vector<char> seen;
    for (int j = 0; j < k; j++) {
        const int i = frames[j];
        if (i < 0) return fail(MDB_ERR_INVALID, "mdb_lineset_nms_frames: negative frame index");
        const int n = infos[i].n_raw;
        const int32_t *ord = orders + order_off[j];
        if (order_off[j + 1] - order_off[j] != n || n < 1 || n > MDB_MAX_LINES)
            return fail(MDB_ERR_INVALID, "mdb_lineset_nms_frames: frame %d has %d raw segments, order has %lld entries", i, n,
                        (long long)(order_off[j + 1] - order_off[j]));
        seen.assign(n, 0);
        for (int q = 0; q < n; q++) {
            if (ord[q] < 0 || ord[q] >= n || seen[ord[q]])
                return fail(MDB_ERR_INVALID, "mdb_lineset_nms_frames: order of frame %d is not a permutation", i);
            seen[ord[q]] = 1;
        }
        infos[i].n_lines = nms_host(raw_lines + (size_t)i * MDB_MAX_LINES * 4, n, lines + (size_t)i * MDB_MAX_LINES * 4,
                                    prob + (size_t)i * MDB_MAX_LINES, ord);
    }
    return MDB_OK;
}

// fill infos / line outputs for frames 0..T-1 of a finished batch (pure host work)
static int finish_batch(mdb_detector *h, const BatchCtx &c, mdb_frame_info *infos, int32_t *lines,
                        double *prob, int32_t *raw_lines) {
    const int T = c.T;
    // FNV-1a over (threshold, on-pixel count, raw segment count, raw segments) of every frame: lets two runs of the same
    // frames be compared batch by batch without shipping the results anywhere (mdb_get_info "digest_hi" / "digest_lo")
    unsigned long long dg = 1469598103934665603ull;
    auto mix = [&dg](uint32_t w) { dg = (dg ^ w) * 1099511628211ull; };
    for (int i = 0; i < T; i++) {
        mdb_frame_info fi;
        memset(&fi, 0, sizeof fi);
        fi.timer = c.timer0 + i + 1;
        fi.bi_threshold = c.h_thr[i];
        fi.bi_threshold_float = c.h_thrf[i];
        fi.snr = c.h_snr[i];
        fi.n_on = (int32_t)c.h_npoints[i];
        // Detector.py:342-344 (host doubles; this TU is compiled with -ffp-contract=off)
        volatile double ds = (double)c.h_npoints[i] / h->hp.mask_area;
        ds = ds * 100.0;
        fi.dst_sum = ds;
        volatile double g = ds / 0.05;
        g = 1.0 - g;
        if (!(g > 0.0)) g = 0.0;
        fi.gap = g * (double)h->cfg.hough_max_gap;
        if (c.h_nlines[i] < 0) return fail(MDB_ERR_STATE, "internal: frame %d left unresolved by the Hough tiers", i);
        fi.lines_num = c.h_nlines[i];
        const int32_t *src = c.h_lines + (size_t)i * MDB_MAX_LINES * 4;
        int nraw = fi.lines_num > MDB_NUM_LINES_TOOMUCH ? 0 : fi.lines_num;
        if (h->cfg.detector == 1) nraw = std::min(fi.lines_num, MDB_MAX_LINES);  // ClassicDetector keeps every segment
        fi.n_raw = nraw;
        if (raw_lines && nraw) memcpy(raw_lines + (size_t)i * MDB_MAX_LINES * 4, src, (size_t)nraw * 16);
        mix((uint32_t)fi.bi_threshold); mix((uint32_t)fi.n_on); mix((uint32_t)fi.lines_num);
        for (int q = 0; q < 4 * nraw; q++) mix((uint32_t)src[q]);
        fi.n_lines = 0;
        if (nraw && lines && prob && h->cfg.detector == 0) {
            int ties = 0;
            fi.n_lines = nms_host(src, nraw, lines + (size_t)i * MDB_MAX_LINES * 4, prob + (size_t)i * MDB_MAX_LINES, nullptr, &ties);
            fi.len_ties = ties;
        }
        if (infos) infos[i] = fi;
    }
    h->last_digest = dg;
    return MDB_OK;
}

static int in_flight(const mdb_detector *h) { return (int)(h->submitted - h->collected); }

// the per-frame calls and the read-backs work on the front stream: order it after whatever the scalar
// stream still has queued from batched calls (history copy of the last batch)
static int front_after_scalar(mdb_detector *h) {
    CK(cudaEventRecord(h->ev_front, h->sstream));
    CK(cudaStreamWaitEvent(h->stream, h->ev_front, 0));
    return MDB_OK;
}

// ---- per-frame API ---------------------------------------------------------------------------
// the per-frame O(1) path serves M3 handles the streaming kernels serve (W % 32 == 0, 2 <= n <= 128)
static bool pf_usable(const mdb_detector *h) {
    return h->per_frame_fast && h->cfg.detector == 0 && h->sk.ok && h->use_stream_kernel;
}

static int pf_launch_suffix(mdb_detector *h) {
    // block b .. b+n-1 just ended (timer = b + n): SUF[j] = max(x[b+j .. b+n-1]), j = n-1 .. 1
    const size_t groups = h->HW / 16;
    const unsigned grid = (unsigned)((groups + PF_THREADS - 1) / PF_THREADS);
    const long long hi = h->timer - 1, lo = h->timer - h->n + 1;
    const FrameSrc src = frame_src(h, nullptr, 0);
    // on the back stream, behind the update kernel of the block's last frame (ev_front): it runs beside that frame's
    // chain and the next frame's copy; the next update kernel waits for ev_suffix
    CK(cudaEventRecord(h->ev_front, h->stream));
    CK(cudaStreamWaitEvent(h->stream2, h->ev_front, 0));
    if (src.mask) pf_suffix_kernel<true><<<grid, PF_THREADS, 0, h->stream2>>>(src, hi, lo, h->n - 1, h->d_pf_suf, groups);
    else pf_suffix_kernel<false><<<grid, PF_THREADS, 0, h->stream2>>>(src, hi, lo, h->n - 1, h->d_pf_suf, groups);
    CK(cudaEventRecord(h->ev_suffix, h->stream2));
    h->pf_suffix_wait = true;
    h->launches += 1;
    h->pf_suffix_pending = false;
    CK(cudaGetLastError());
    return MDB_OK;
}

static int pf_update(mdb_detector *h, const uint8_t *frame, int on_device) {
    const size_t groups = h->HW / 16;
    const unsigned grid = (unsigned)((groups + PF_THREADS - 1) / PF_THREADS);
    if (!h->d_pf_sum) {
        if (cudaMalloc((void **)&h->d_pf_sum, h->HW * sizeof(uint16_t)) != cudaSuccess ||
            cudaMalloc((void **)&h->d_pf_pmax, h->HW) != cudaSuccess ||
            cudaMalloc((void **)&h->d_pf_suf, (size_t)h->n * h->HW) != cudaSuccess ||
            cudaMalloc((void **)&h->d_pf_stage, h->HW) != cudaSuccess)
            return fail(MDB_ERR_NOMEM, "per-frame window state: %s", cudaGetErrorString(cudaGetLastError()));
        h->pf_timer = -1;
    }
    if (h->pf_suffix_pending && h->pf_timer == h->timer) {
        int rc = pf_launch_suffix(h);
        if (rc) return rc;
    }
    if (h->pf_timer != h->timer) {  // batched calls, reset or seek moved the stream on: rebuild from the ring
        if (h->pf_suffix_wait) {
            CK(cudaStreamWaitEvent(h->stream, h->ev_suffix, 0));
            h->pf_suffix_wait = false;
        }
        const FrameSrc src = frame_src(h, nullptr, 0);
        if (src.mask) pf_rebuild_kernel<true><<<grid, PF_THREADS, 0, h->stream>>>(src, h->timer, h->n, h->d_pf_sum, h->d_pf_pmax, h->d_pf_suf, groups);
        else pf_rebuild_kernel<false><<<grid, PF_THREADS, 0, h->stream>>>(src, h->timer, h->n, h->d_pf_sum, h->d_pf_pmax, h->d_pf_suf, groups);
        h->launches += 1;
        h->pf_suffix_pending = false;
        CK(cudaGetLastError());
    }
    const long long t = h->timer;
    // the frame arrives on the copy stream in two halves; the update kernel follows half by half on the front stream,
    // and on frames that are not noise-sample timers (nearly all) the threshold recurrence runs beside the copy
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const size_t gA = groups / 2, bytesA = gA * 16;
    CK(cudaStreamWaitEvent(h->cstream, h->ev_pf_done, 0));  // the staging buffer's last reader (previous update kernel)
    CK(cudaMemcpyAsync(h->d_pf_stage, frame, bytesA, kind, h->cstream));
    CK(cudaEventRecord(h->ev_pf_a, h->cstream));
    CK(cudaMemcpyAsync(h->d_pf_stage + bytesA, frame + bytesA, h->HW - bytesA, kind, h->cstream));
    CK(cudaEventRecord(h->ev_pf_b, h->cstream));
    const long long tau = t + 1, std_interval = (long long)h->cfg.nz_interval * h->n;
    const bool sample = (tau > 1 && tau <= h->n) || (tau > h->n && std_interval > 0 && tau % std_interval == 0);
    int rc = MDB_OK;
    if (!sample) rc = launch_noise_thr(h, h->ctx[0], frame_src(h, h->d_pf_stage, t), 1, t, h->stream);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(h->stream, h->ev_pf_a, 0));
    if (sample) {
        // noise sample + threshold of this frame: newest frame from the staging buffer, older ones from the ring
        CK(cudaStreamWaitEvent(h->stream, h->ev_pf_b, 0));
        rc = launch_noise_thr(h, h->ctx[0], frame_src(h, h->d_pf_stage, t), 1, t, h->stream);
        if (rc) return rc;
    }
    const int pos = (int)(t % h->n);
    const int L = (int)std::min<long long>(h->n, t + 1);
    uint8_t *slot = h->d_ring + (size_t)(t % h->R) * h->HW;
    const uint8_t *old = t >= h->n ? h->d_ring + (size_t)((t - h->n) % h->R) * h->HW : nullptr;
    // window = prefix of the current block + suffix [pos+1 ..] of the previous one (none for the block's last frame)
    const uint8_t *suf = (pos < h->n - 1) ? h->d_pf_suf + (size_t)(pos + 1) * h->HW : nullptr;
    uint16_t *bits = reinterpret_cast<uint16_t *>(h->sk.d_bits);
    if (h->pf_suffix_wait) {
        CK(cudaStreamWaitEvent(h->stream, h->ev_suffix, 0));
        h->pf_suffix_wait = false;
    }
    for (int half = 0; half < 2; half++) {
        const size_t g0 = half ? gA : 0, g1 = half ? groups : gA;
        if (half) CK(cudaStreamWaitEvent(h->stream, h->ev_pf_b, 0));
        if (g1 == g0) continue;
        const unsigned gh = (unsigned)((g1 - g0 + PF_THREADS - 1) / PF_THREADS);
        if (h->cfg.apply_mask)
            pf_update_kernel<true><<<gh, PF_THREADS, 0, h->stream>>>(h->d_pf_stage, slot, old, h->d_mask, h->d_pf_sum, h->d_pf_pmax, suf,
                                                                    pos == 0, L, h->ctx[0].d_thr, g0, g1, bits);
        else
            pf_update_kernel<false><<<gh, PF_THREADS, 0, h->stream>>>(h->d_pf_stage, slot, old, nullptr, h->d_pf_sum, h->d_pf_pmax, suf,
                                                                     pos == 0, L, h->ctx[0].d_thr, g0, g1, bits);
        h->launches += 1;
    }
    CK(cudaEventRecord(h->ev_pf_done, h->stream));
    CK(cudaGetLastError());
    h->timer += 1;
    h->pf_timer = h->pf_bits_timer = h->timer;
    h->pf_suffix_pending = pos == h->n - 1;
    h->front_dirty = true;
    return MDB_OK;
}

extern "C" int mdb_update(mdb_handle h, const uint8_t *frame, int on_device) {
    if (!h || !frame) return fail(MDB_ERR_INVALID, "mdb_update: null argument");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_update: a submitted batch has not been collected");
    CK(cudaSetDevice(h->cfg.device));
    int rc = front_after_scalar(h);
    if (rc) return rc;
    // O(1) path: the copy is asynchronous.  A pageable source has been consumed when cudaMemcpyAsync returns (the
    // runtime stages it); a pinned source must stay untouched until the next detect() / update() returns.
    if (pf_usable(h)) return pf_update(h, frame, on_device);
    rc = copy_to_ring(h, frame, 0, 1, h->timer, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream);
    if (rc) return rc;
    rc = launch_noise_thr(h, h->ctx[0], frame_src(h, nullptr, 0), 1, h->timer, h->stream);
    h->front_dirty = true;
    if (rc) return rc;
    h->timer += 1;
    if (!on_device) CK(cudaStreamSynchronize(h->stream));  // caller may reuse its buffer on return
    return MDB_OK;
}

extern "C" int mdb_detect(mdb_handle h, mdb_frame_info *info, int32_t *lines, double *nonline_prob,
                          int32_t *raw_lines) {
    if (!h) return fail(MDB_ERR_INVALID, "mdb_detect: null handle");
    if (h->timer == 0) return fail(MDB_ERR_STATE, "mdb_detect: no frame has been pushed yet");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_detect: a submitted batch has not been collected");
    CK(cudaSetDevice(h->cfg.device));
    BatchCtx &c = h->ctx[0];
    int rc = front_after_scalar(h);  // the history copy of the last batch runs on the scalar stream
    if (rc) return rc;
    c.halo = false;
    // update() of this frame already left its predicate bits behind (per-frame O(1) path): spatial passes only
    const bool bits_ready = pf_usable(h) && h->pf_bits_timer == h->timer;
    if (bits_ready && h->pf_suffix_pending && h->pf_timer == h->timer) {  // on the back stream, beside this frame's chain
        rc = pf_launch_suffix(h);
        if (rc) return rc;
    }
    h->single_stream = bits_ready;  // a frame whose bits are ready runs its spatial passes and PPHT on the front stream
    rc = launch_fused(h, c, frame_src(h, nullptr, 0), 1, h->timer - 1, h->dy_timer, bits_ready);
    if (!rc) rc = launch_hough_and_copy(h, c, 1);
    h->single_stream = false;
    if (rc) return rc;
    CK(cudaEventSynchronize(c.ev_done));
    {
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, c.ev_f0, c.ev_f1));
        CK(cudaEventElapsedTime(&b, c.ev_d0, c.ev_d1));
        h->fused_ms = a + b; h->temporal_ms = a; h->spatial_ms = b;
    }
    h->last_fused_launches = h->fused_launches;
    h->dy_timer += 1;
    h->last_T = 1;
    h->last_ctx = 0;
    c.T = 1;
    c.timer0 = h->timer - 1;
    rc = finish_batch(h, c, info, lines, nonline_prob, raw_lines);
    c.T = 0;
    return rc;
}

// ---- batched API -----------------------------------------------------------------------------
static int submit_impl(mdb_handle h, const uint8_t *frames, int T, int on_device, const int32_t *thr,
                       const double *thrf, const double *snr, int flags = 0) {
    if (!h || !frames) return fail(MDB_ERR_INVALID, "mdb_submit_batch: null argument");
    if (T < 1 || T > h->cfg.max_batch)
        return fail(MDB_ERR_INVALID, "mdb_submit_batch: T=%d outside 1..max_batch=%d", T, h->cfg.max_batch);
    if (in_flight(h) >= NCTX) return fail(MDB_ERR_STATE, "mdb_submit_batch: %d batches already in flight", NCTX);
    CK(cudaSetDevice(h->cfg.device));
    BatchCtx &c = h->ctx[h->submitted % NCTX];
    const long long timer0 = h->timer;
    FrameSrc src;
    if (h->front_dirty) {  // per-frame calls updated the scalar state on the front stream: order the scalar stream after them
        CK(cudaEventRecord(h->ev_front, h->stream));
        CK(cudaStreamWaitEvent(h->sstream, h->ev_front, 0));
        h->front_dirty = false;
    }
    BatchCtx &prev2 = h->ctx[(h->submitted + NCTX - 2) % NCTX];  // the batch before the previous one
    const bool have_prev2 = h->submitted >= 2;
    const uint8_t *dev_frames = frames;
    if (!on_device && T > 1) {
        // host frames are copied into one of two contiguous staging buffers and then take the zero-copy path: the
        // kernels read a batch's frames from one contiguous block, the ring only ever holds history.  The copy runs
        // on its own stream so that it overlaps the previous batch's kernels; the buffer it lands in belonged to
        // the batch before the previous one, all of whose readers must be done (temporal pass / generic kernels /
        // noise samples + history copy).
        uint8_t *&stage = h->d_stage[h->submitted & 1];
        if (!stage && cudaMalloc((void **)&stage, (size_t)h->cfg.max_batch * h->HW) != cudaSuccess) {
            cudaGetLastError();
            return fail(MDB_ERR_NOMEM, "staging buffer of %zu bytes for host frames", (size_t)h->cfg.max_batch * h->HW);
        }
        if (have_prev2) {
            CK(cudaStreamWaitEvent(h->cstream, prev2.ev_f1, 0));
            CK(cudaStreamWaitEvent(h->cstream, prev2.ev_d1, 0));
            CK(cudaStreamWaitEvent(h->cstream, prev2.ev_src, 0));
        }
        CK(cudaMemcpyAsync(stage, frames, (size_t)T * h->HW, cudaMemcpyHostToDevice, h->cstream));
        CK(cudaEventRecord(h->ev_copy, h->cstream));
        CK(cudaStreamWaitEvent(h->sstream, h->ev_copy, 0));
        dev_frames = stage;
    }
    if (on_device || T > 1) {
        src = frame_src(h, dev_frames, timer0);  // zero-copy: kernels read the batch from one contiguous block
    } else {
        // a single host frame goes straight into its ring slot
        if (have_prev2) {
            CK(cudaStreamWaitEvent(h->cstream, prev2.ev_f1, 0));
            CK(cudaStreamWaitEvent(h->cstream, prev2.ev_d1, 0));
        }
        int rc = copy_to_ring(h, frames, 0, T, timer0, cudaMemcpyHostToDevice, h->cstream);
        if (rc) return rc;
        CK(cudaEventRecord(h->ev_copy, h->cstream));
        CK(cudaStreamWaitEvent(h->sstream, h->ev_copy, 0));
        src = frame_src(h, nullptr, 0);
    }
    if (have_prev2) {
        // the act ring holds two batches + history: this batch's act frames land on those of the batch
        // before the previous one, whose dst pass (back stream) must have read them
        CK(cudaStreamWaitEvent(h->stream, prev2.ev_d1, 0));
    }
    // scalar stream: thresholds of this batch (noise samples + EMA recurrence) and the copy of its last n frames into
    // the ring (history of the next batch, and what mdb_get_stack reads).  None of it depends on the previous
    // batch's mask chain, so it runs beside that batch's temporal pass instead of in front of this one's.
    int rc;
    if (thr) {  // thresholds supplied by the caller (time-sharded streams): no local EMA recurrence
        CK(cudaMemcpyAsync(c.d_thr, thr, T * sizeof(int32_t), cudaMemcpyHostToDevice, h->sstream));
        CK(cudaMemcpyAsync(c.d_thrf, thrf, T * sizeof(double), cudaMemcpyHostToDevice, h->sstream));
        CK(cudaMemcpyAsync(c.d_snr, snr, T * sizeof(double), cudaMemcpyHostToDevice, h->sstream));
    } else {
        rc = launch_noise_thr(h, c, src, T, timer0, h->sstream);
        if (rc) return rc;
    }
    CK(cudaEventRecord(c.ev_thr, h->sstream));
    if (src.cur) {
        const int keep = std::min(T, h->n);
        // the slots they land on may still be read as history by the batch before the previous one: its temporal
        // pass (front stream) or, on the generic path, its per-frame kernels (back stream) must be done
        if (have_prev2) {
            CK(cudaStreamWaitEvent(h->sstream, prev2.ev_f1, 0));
            CK(cudaStreamWaitEvent(h->sstream, prev2.ev_d1, 0));
        }
        rc = copy_to_ring(h, dev_frames, T - keep, T, timer0, cudaMemcpyDeviceToDevice, h->sstream);
        if (rc) return rc;
    }
    CK(cudaEventRecord(c.ev_src, h->sstream));
    CK(cudaStreamWaitEvent(h->stream, c.ev_thr, 0));
    c.bits_parity = (int)(h->submitted & 1);
    c.halo = (flags & MDB_SUBMIT_HALO) != 0;
    rc = launch_fused(h, c, src, T, timer0, h->dy_timer);
    if (rc) return rc;
    rc = launch_hough_and_copy(h, c, T);
    if (rc) return rc;
    c.T = T;
    c.timer0 = timer0;
    c.seq = h->submitted;
    h->submitted += 1;
    h->timer += T;
    h->dy_timer += T;
    return MDB_OK;
}

extern "C" int mdb_submit_batch(mdb_handle h, const uint8_t *frames, int T, int on_device) {
    return submit_impl(h, frames, T, on_device, nullptr, nullptr, nullptr);
}

extern "C" int mdb_submit_batch_thr(mdb_handle h, const uint8_t *frames, int T, int on_device,
                                    const int32_t *thr, const double *thr_float, const double *snr) {
    if (!thr || !thr_float || !snr) return fail(MDB_ERR_INVALID, "mdb_submit_batch_thr: null threshold arrays");
    return submit_impl(h, frames, T, on_device, thr, thr_float, snr);
}

extern "C" int mdb_submit_batch_ex(mdb_handle h, const uint8_t *frames, int T, int on_device, const int32_t *thr,
                                   const double *thr_float, const double *snr, int flags) {
    if ((thr || thr_float || snr) && !(thr && thr_float && snr))
        return fail(MDB_ERR_INVALID, "mdb_submit_batch_ex: all three threshold arrays or none");
    if (flags & ~MDB_SUBMIT_HALO) return fail(MDB_ERR_INVALID, "mdb_submit_batch_ex: unknown flags 0x%x", flags);
    return submit_impl(h, frames, T, on_device, thr, thr_float, snr, flags);
}

extern "C" int mdb_seek(mdb_handle h, int64_t timer) {
    if (!h || timer < 0) return fail(MDB_ERR_INVALID, "mdb_seek: bad argument");
    if (h->timer != 0 || h->submitted != 0) return fail(MDB_ERR_STATE, "mdb_seek: only before the first frame");
    h->timer = h->dy_timer = timer;
    h->seek0 = timer;
    return MDB_OK;
}

extern "C" int mdb_reset(mdb_handle h) {
    if (!h) return fail(MDB_ERR_INVALID, "mdb_reset: null handle");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_reset: a submitted batch has not been collected");
    CK(cudaSetDevice(h->cfg.device));
    for (cudaStream_t st : {h->stream, h->stream2, h->stream3, h->cstream, h->sstream}) CK(cudaStreamSynchronize(st));
    CK(cudaMemsetAsync(h->d_ring, 0, (size_t)h->R * h->HW, h->stream));
    CK(cudaMemsetAsync(h->d_act, 0, (size_t)h->RA * h->H * h->Wb * sizeof(uint32_t), h->stream));
    const DevState st = initial_state(h);
    CK(cudaMemcpyAsync(h->d_state, &st, sizeof st, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->timer = h->dy_timer = h->seek0 = 0;
    h->submitted = h->collected = 0;
    h->last_T = 0;
    h->front_dirty = false;
    return MDB_OK;
}

// Noise sums of the sample timers among device frames, for several segments in one call, on the handle's scalar
// stream and buffers: what a rank of a time-sharded run computes for its chunk before the thresholds are replayed.
extern "C" int mdb_noise_sums_dev(mdb_handle h, int nseg, const uint8_t *const *frames, const int32_t *T, const int64_t *t0,
                                  uint64_t *sums) {
    if (!h || !frames || !T || !t0 || !sums || nseg < 1) return fail(MDB_ERR_INVALID, "mdb_noise_sums_dev: bad arguments");
    size_t total = 0;
    for (int k = 0; k < nseg; k++) {
        if (!frames[k] || T[k] < 1 || t0[k] < 0) return fail(MDB_ERR_INVALID, "mdb_noise_sums_dev: bad segment %d", k);
        total += (size_t)T[k];
    }
    CK(cudaSetDevice(h->cfg.device));
    if (h->noise2_cap < total) {  // grown generously: reallocating costs milliseconds (pinned memory, implicit syncs)
        if (h->d_noise2) CK(cudaFree(h->d_noise2));
        if (h->h_noise2) CK(cudaFreeHost(h->h_noise2));
        h->d_noise2 = h->h_noise2 = nullptr;
        h->noise2_cap = 0;
        const size_t cap = std::max<size_t>(2 * total, (size_t)1 << 20);
        CK(cudaMalloc((void **)&h->d_noise2, cap * 16));
        CK(cudaHostAlloc((void **)&h->h_noise2, cap * 16, cudaHostAllocDefault));
        h->noise2_cap = cap;
    }
    const mdb_config &c = h->cfg;
    const long long std_interval = (long long)c.nz_interval * h->n;
    CK(cudaMemsetAsync(h->d_noise2, 0, total * 16, h->sstream));
    size_t off = 0;
    for (int k = 0; k < nseg; k++) {
        const long long min_tau = t0[k] == 0 ? 0 : t0[k] + h->n;  // the whole window must lie inside the supplied frames
        // sample frames of the segment, 63 per launch
        int i = 0;
        while (i < T[k]) {
            SampleList sl;
            sl.count = 0;
            for (; i < T[k] && sl.count < 63; i++) {
                const long long tau = t0[k] + i + 1;
                if (tau >= min_tau && ((tau > 1 && tau <= h->n) || (tau > h->n && std_interval > 0 && tau % std_interval == 0)))
                    sl.idx[sl.count++] = i;
            }
            if (sl.count == 0) break;
            FrameSrc src;
            src.ring = nullptr; src.cur = frames[k]; src.t0 = t0[k]; src.mask = c.apply_mask ? h->d_mask : nullptr; src.R = 1; src.HW = h->HW;
            launch_noise_samples(src, h->W, h->n, t0[k], std_interval, c.roi, h->d_noise2 + 2 * off, min_tau, sl, T[k], h->sstream);
            CK(cudaGetLastError());
            h->launches += 1;
        }
        off += (size_t)T[k];
    }
    CK(cudaMemcpyAsync(h->h_noise2, h->d_noise2, total * 16, cudaMemcpyDeviceToHost, h->sstream));
    CK(cudaStreamSynchronize(h->sstream));
    memcpy(sums, h->h_noise2, total * 16);
    return MDB_OK;
}

// EMA.update (MetLib/utils.py:334-368) over the noise samples in timer order + LineDetector.update's threshold rule
// (MetLib/Detector.py:225-229): the scalar recurrence of threshold_kernel on the host (same IEEE double operations; this
// TU is compiled without FMA contraction), for frames 0 .. t_end-1; outputs for frames t_begin .. t_end-1.
extern "C" int mdb_replay_thresholds(int nsamples, const int64_t *timers, const uint64_t *sums, int64_t roi_pixels,
                                     int window, int nz_interval, int adaptive, int init_value, int sensitivity,
                                     int64_t t_begin, int64_t t_end, int32_t *thr, double *thr_float, double *snr) {
    if (nsamples < 0 || (nsamples && (!timers || !sums)) || !thr || !thr_float || !snr || t_begin < 0 || t_end < t_begin ||
        window < 1 || sensitivity < 0 || sensitivity > 2)
        return fail(MDB_ERR_INVALID, "mdb_replay_thresholds: bad arguments");
    const int n = window;
    const long long std_interval = (long long)nz_interval * n;
    volatile double ema_init_m = 1.0 - (double)nz_interval / 60.0;
    double ema_cur_m = ema_init_m, ema_warm = (double)n, ema_value = 0.0;
    long long ema_t = 0;
    static const int abs_sens[3] = {7, 5, 3};
    int bi = adaptive ? abs_sens[sensitivity] : init_value;
    double thrf = (double)bi;
    const double a = sensitivity == MDB_SENS_LOW ? 2.0 : (sensitivity == MDB_SENS_NORMAL ? 1.2 : 0.9);
    const double b = sensitivity == MDB_SENS_LOW ? 4.4 : (sensitivity == MDB_SENS_NORMAL ? 3.6 : 3.0);
    int k = 0;
    long long t = 0;
    auto emit = [&](long long upto) {  // frames t .. upto-1 carry the current values
        for (long long u = std::max<long long>(t, t_begin); u < upto; u++) {
            thr[u - t_begin] = bi; thr_float[u - t_begin] = thrf; snr[u - t_begin] = ema_value;
        }
        t = upto;
    };
    while (t < t_end) {
        // next sample timer at or after t+1
        while (k < nsamples && timers[k] < t + 1) k++;
        long long next_tau = t_end + 1;
        for (long long tau = t + 1; tau <= t_end; ) {
            const bool is = (tau > 1 && tau <= n) || (tau > n && std_interval > 0 && tau % std_interval == 0);
            if (is) { next_tau = tau; break; }
            tau = tau <= n ? tau + 1 : (std_interval > 0 ? (tau / std_interval + 1) * std_interval : t_end + 1);
        }
        if (next_tau > t_end) { emit(t_end); break; }
        emit(next_tau - 1);  // frames before the sample frame keep the old values
        if (k >= nsamples || timers[k] != next_tau)
            return fail(MDB_ERR_INVALID, "mdb_replay_thresholds: the noise sample of timer %lld is missing", next_tau);
        const int L = (int)(next_tau < n ? next_tau : n);
        volatile double N = (double)(L * roi_pixels);
        volatile double s1 = (double)sums[2 * k], s2 = (double)sums[2 * k + 1];
        volatile double mean = s1 / N;
        volatile double q = s2 / N;
        volatile double mm = mean * mean;
        volatile double var = q - mm;
        if (var < 0) var = 0;
        const double sigma = std::sqrt((double)var);
        if (ema_warm != 0.0) {
            volatile double one_m = 1.0 - ema_init_m;
            volatile double kk = (double)ema_t * one_m;
            kk = kk * ema_warm;
            if (kk < 1.0) {
                volatile double u = 1.0 - kk;
                volatile double uu = u * u;
                volatile double w = 1.0 - uu;
                ema_cur_m = ema_init_m * w;
            } else {
                ema_warm = 0.0;
                ema_cur_m = ema_init_m;
            }
        }
        volatile double p1 = ema_cur_m * ema_value;
        volatile double om = 1.0 - ema_cur_m;
        volatile double p2 = om * sigma;
        ema_value = p1 + p2;
        ema_t++;
        k++;
        if (adaptive && ema_value != 0.0) {
            volatile double x2 = ema_value * ema_value;
            volatile double ax = a * x2;
            thrf = ax + b;
            bi = (int)std::nearbyint(thrf);  // Python round(): half to even (default rounding mode)
        }
        emit(next_tau);  // the sample frame itself already carries the new values
    }
    return MDB_OK;
}

extern "C" int mdb_noise_sums(const uint8_t *frames, int T, int on_device, int64_t t0, int width, int height,
                              int window, int nz_interval, const int32_t *roi, const uint8_t *mask,
                              uint64_t *sums, int device) {
    if (!frames || !roi || !sums || T < 1 || width < 1 || height < 1 || window < 1 || t0 < 0)
        return fail(MDB_ERR_INVALID, "mdb_noise_sums: bad arguments");
    if (mdb_device_count() == 0) return fail(MDB_ERR_CUDA, "mdb_noise_sums: no CUDA device (no CPU fallback)");
    CK(cudaSetDevice(device));
    const size_t HW = (size_t)width * height;
    uint8_t *d_frames = nullptr, *d_mask = nullptr;
    unsigned long long *d_acc = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_acc, (size_t)T * 16);
    if (e == cudaSuccess) e = cudaMemset(d_acc, 0, (size_t)T * 16);
    if (e == cudaSuccess && !on_device) {
        e = cudaMalloc((void **)&d_frames, (size_t)T * HW);
        if (e == cudaSuccess) e = cudaMemcpy(d_frames, frames, (size_t)T * HW, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess && mask) {
        e = cudaMalloc((void **)&d_mask, HW);
        if (e == cudaSuccess) e = cudaMemcpy(d_mask, mask, HW, cudaMemcpyHostToDevice);
    }
    if (e == cudaSuccess) {
        FrameSrc src;
        src.ring = nullptr; src.cur = on_device ? frames : d_frames; src.t0 = t0; src.mask = d_mask; src.R = 1; src.HW = HW;
        // only samples whose whole window lies inside the supplied frames (or starts at global frame 0)
        const long long min_tau = t0 == 0 ? 0 : t0 + window;
        SampleList sl;
        sl.count = -1;
        launch_noise_samples(src, width, window, t0, (long long)nz_interval * window, roi, d_acc, min_tau, sl, T, 0);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpy(sums, d_acc, (size_t)T * 16, cudaMemcpyDeviceToHost);
    }
    if (d_frames) cudaFree(d_frames);
    if (d_mask) cudaFree(d_mask);
    if (d_acc) cudaFree(d_acc);
    if (e != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_noise_sums: %s", cudaGetErrorString(e));
    return MDB_OK;
}

extern "C" int mdb_collect_batch(mdb_handle h, mdb_frame_info *infos, int32_t *lines, double *nonline_prob,
                                 int32_t *raw_lines, uint8_t *dst_out, int dst_on_device) {
    if (!h) return fail(MDB_ERR_INVALID, "mdb_collect_batch: null handle");
    if (!in_flight(h)) return fail(MDB_ERR_STATE, "mdb_collect_batch: nothing submitted");
    CK(cudaSetDevice(h->cfg.device));
    BatchCtx &c = h->ctx[h->collected % NCTX];
    const int T = c.T;
    if (dst_out) {
        CK(cudaEventSynchronize(c.ev_done));
        CK(cudaMemcpy(dst_out, c.d_dst, (size_t)T * h->HW,
                      dst_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
    }
    CK(cudaEventSynchronize(c.ev_done));
    {
        float a = 0.f, b = 0.f;
        CK(cudaEventElapsedTime(&a, c.ev_f0, c.ev_f1));
        CK(cudaEventElapsedTime(&b, c.ev_d0, c.ev_d1));
        h->fused_ms = a + b;  // temporal (front stream) and act + dst (back stream), each bracketed by events
        h->temporal_ms = a; h->spatial_ms = b;
    }
    h->last_fused_launches = h->fused_launches;
    for (int k = 0; k < 4; k++) { h->tiers_last[k] = c.h_tiers[k]; h->tiers_total[k] += c.h_tiers[k]; }
    int rc = finish_batch(h, c, infos, lines, nonline_prob, raw_lines);
    h->last_T = T;
    h->last_ctx = (int)(h->collected % NCTX);
    c.T = 0;
    h->collected += 1;
    return rc;
}

extern "C" int mdb_detect_batch(mdb_handle h, const uint8_t *frames, int T, int on_device,
                                mdb_frame_info *infos, int32_t *lines, double *nonline_prob,
                                int32_t *raw_lines, uint8_t *dst_out, int dst_on_device) {
    if (h && in_flight(h)) return fail(MDB_ERR_STATE, "mdb_detect_batch: a submitted batch has not been collected");
    int rc = mdb_submit_batch(h, frames, T, on_device);
    if (rc) return rc;
    return mdb_collect_batch(h, infos, lines, nonline_prob, raw_lines, dst_out, dst_on_device);
}

extern "C" int mdb_get_dst(mdb_handle h, uint8_t *dst, int on_device) {
    if (!h || !dst) return fail(MDB_ERR_INVALID, "mdb_get_dst: null argument");
    if (h->last_T < 1) return fail(MDB_ERR_STATE, "mdb_get_dst: no detect has run yet");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_get_dst: a batch is in flight");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpy(dst, h->ctx[h->last_ctx].d_dst + (size_t)(h->last_T - 1) * h->HW, h->HW,
                  on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
    return MDB_OK;
}

extern "C" int mdb_get_dst_device(mdb_handle h, const uint8_t **ptr) {
    if (!h || !ptr) return fail(MDB_ERR_INVALID, "mdb_get_dst_device: null argument");
    *ptr = h->ctx[h->last_ctx].d_dst;
    return MDB_OK;
}

extern "C" int mdb_get_stack(mdb_handle h, uint8_t *max_out, uint8_t *mean_out, uint32_t *sum_out) {
    if (!h) return fail(MDB_ERR_INVALID, "mdb_get_stack: null handle");
    if (h->timer == 0) return fail(MDB_ERR_STATE, "mdb_get_stack: empty window");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_get_stack: a batch is in flight");
    CK(cudaSetDevice(h->cfg.device));
    uint8_t *dmx = nullptr, *dmean = nullptr;
    uint32_t *dsum = nullptr;
    CK(cudaMalloc((void **)&dmx, h->HW));
    CK(cudaMalloc((void **)&dmean, h->HW));
    CK(cudaMalloc((void **)&dsum, h->HW * 4));
    const long long t = h->timer - 1;
    const int L = (int)std::min<long long>(h->n, h->timer);
    {
        int rc = front_after_scalar(h);
        if (rc) { cudaFree(dmx); cudaFree(dmean); cudaFree(dsum); return rc; }
    }
    stack_readback_kernel<<<592, 256, 0, h->stream>>>(frame_src(h, nullptr, 0), h->HW, h->n, t, L, dmx, dmean, dsum);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && max_out) e = cudaMemcpyAsync(max_out, dmx, h->HW, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && mean_out) e = cudaMemcpyAsync(mean_out, dmean, h->HW, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && sum_out) e = cudaMemcpyAsync(sum_out, dsum, h->HW * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dmx); cudaFree(dmean); cudaFree(dsum);
    if (e != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_get_stack: %s", cudaGetErrorString(e));
    return MDB_OK;
}

// SlidingWindow.sliding_window (MetLib/utils.py:263-265): the n ring slots in the reference's own slot order (slot i holds
// the newest frame whose 0-based index is congruent to i modulo n; slots never written are zero).
extern "C" int mdb_get_window(mdb_handle h, uint8_t *out, int on_device) {
    if (!h || !out) return fail(MDB_ERR_INVALID, "mdb_get_window: null argument");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_get_window: a batch is in flight");
    CK(cudaSetDevice(h->cfg.device));
    int rc = front_after_scalar(h);
    if (rc) return rc;
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    uint8_t *tmp = nullptr;
    for (int i = 0; i < h->n; i++) {
        // newest frame index t <= timer-1 with t % n == i
        long long t = h->timer - 1 - (((h->timer - 1 - i) % h->n + h->n) % h->n);
        uint8_t *dst = out + (size_t)i * h->HW;
        if (h->timer == 0 || t < 0) {
            if (on_device) CK(cudaMemsetAsync(dst, 0, h->HW, h->stream));
            else memset(dst, 0, h->HW);
            continue;
        }
        const uint8_t *src = h->d_ring + (size_t)(t % h->R) * h->HW;
        if (h->cfg.apply_mask) {
            if (!tmp && !on_device) CK(cudaMalloc((void **)&tmp, h->HW));
            uint8_t *m = on_device ? dst : tmp;
            window_frame_kernel<<<592, 256, 0, h->stream>>>(src, h->d_mask, h->HW, m);
            h->launches += 1;
            src = m;
        }
        if (src != dst) {
            cudaError_t e = cudaMemcpyAsync(dst, src, h->HW, kind, h->stream);
            if (e != cudaSuccess) { cudaFree(tmp); return fail(MDB_ERR_CUDA, "mdb_get_window: %s", cudaGetErrorString(e)); }
            if (tmp) cudaStreamSynchronize(h->stream);  // tmp is reused by the next slot
        }
    }
    cudaError_t e = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_get_window: %s", cudaGetErrorString(e));
    return MDB_OK;
}

// SlidingWindow.std for the uint8 / force_int mode (MetLib/utils.py:309-321)
extern "C" int mdb_get_std(mdb_handle h, double *std_out) {
    if (!h || !std_out) return fail(MDB_ERR_INVALID, "mdb_get_std: null argument");
    if (h->timer == 0) return fail(MDB_ERR_STATE, "mdb_get_std: empty window");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_get_std: a batch is in flight");
    CK(cudaSetDevice(h->cfg.device));
    unsigned long long *d_tot = nullptr, tot = 0;
    CK(cudaMalloc((void **)&d_tot, sizeof tot));
    int rc = front_after_scalar(h);
    cudaError_t e = rc ? cudaErrorUnknown : cudaMemsetAsync(d_tot, 0, sizeof tot, h->stream);
    const int L = (int)std::min<long long>(h->n, h->timer);
    if (e == cudaSuccess) {
        // the reference's window holds unmasked frames unless the caller masked them: same source as mdb_get_stack
        stack_std_kernel<<<592, 256, 0, h->stream>>>(frame_src(h, nullptr, 0), h->HW, h->n, h->timer - 1, L, d_tot);
        h->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_tot);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_get_std: %s", cudaGetErrorString(e));
    volatile double mean = (double)tot / (double)h->HW;
    *std_out = std::sqrt(mean);
    return MDB_OK;
}

// every raw Hough segment of frame `frame` of the batch collected last.  The batched outputs hold MDB_MAX_LINES rows per
// frame; ClassicDetector returns every segment without a cap (Detector.py:282-292), so a frame with more is resolved
// here: its PPHT is run again (masks and on-pixel lists of the last batch are still on the device, the transform is
// deterministic) into a buffer of the size the first pass counted.
extern "C" int mdb_get_raw_lines(mdb_handle h, int frame, int32_t *out, int cap, int32_t *n_out) {
    if (!h || !n_out || cap < 0 || (cap > 0 && !out)) return fail(MDB_ERR_INVALID, "mdb_get_raw_lines: bad arguments");
    if (h->last_T < 1) return fail(MDB_ERR_STATE, "mdb_get_raw_lines: no detect has run yet");
    if (in_flight(h)) return fail(MDB_ERR_STATE, "mdb_get_raw_lines: a batch is in flight");
    if (frame < 0 || frame >= h->last_T) return fail(MDB_ERR_INVALID, "mdb_get_raw_lines: frame %d outside 0..%d", frame, h->last_T - 1);
    BatchCtx &c = h->ctx[h->last_ctx];
    const int nl = c.h_nlines[frame];
    *n_out = nl;
    if (nl <= 0 || cap == 0) return MDB_OK;
    if (cap < nl) return fail(MDB_ERR_INVALID, "mdb_get_raw_lines: %d segments, room for %d", nl, cap);
    if (nl <= MDB_MAX_LINES) {
        memcpy(out, c.h_lines + (size_t)frame * MDB_MAX_LINES * 4, (size_t)nl * 16);
        return MDB_OK;
    }
    CK(cudaSetDevice(h->cfg.device));
    int32_t *d_l = nullptr;
    int *d_n = nullptr;
    unsigned *d_q = nullptr;
    int got = 0;
    cudaError_t e = cudaMalloc((void **)&d_l, (size_t)nl * 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_n, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_q, 8 * sizeof(unsigned));
    int rc = MDB_OK;
    if (e == cudaSuccess) {
        HoughParams hp = h->hp;
        hp.max_lines = nl;
        rc = launch_hough_kernels(h, nullptr, hp, 1, c.d_npoints + frame, c.d_points + (size_t)frame * MDB_POINT_CAP,
                                  c.d_order + (size_t)frame * HOUGH_ORDER_CAP, c.d_dst + (size_t)frame * h->HW, d_l, d_n, d_q,
                                  h->stream3);
        h->launches += 5;
        if (!rc) e = cudaMemcpyAsync(out, d_l, (size_t)nl * 16, cudaMemcpyDeviceToHost, h->stream3);
        if (!rc && e == cudaSuccess) e = cudaMemcpyAsync(&got, d_n, sizeof got, cudaMemcpyDeviceToHost, h->stream3);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream3);
    }
    cudaFree(d_l); cudaFree(d_n); cudaFree(d_q);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_get_raw_lines: %s", cudaGetErrorString(e));
    if (got != nl) return fail(MDB_ERR_STATE, "internal: second PPHT pass found %d segments, first pass %d", got, nl);
    return MDB_OK;
}

extern "C" int mdb_get_stream(mdb_handle h, void **stream) {
    if (!h || !stream) return fail(MDB_ERR_INVALID, "mdb_get_stream: null argument");
    *stream = (void *)h->stream;
    return MDB_OK;
}

extern "C" int mdb_get_launch_count(mdb_handle h, int64_t *count) {
    if (!h || !count) return fail(MDB_ERR_INVALID, "mdb_get_launch_count: null argument");
    *count = h->launches;
    return MDB_OK;
}

extern "C" int mdb_get_fused_time(mdb_handle h, float *ms, int32_t *launches) {
    if (!h) return fail(MDB_ERR_INVALID, "mdb_get_fused_time: null handle");
    if (ms) *ms = h->fused_ms;
    if (launches) *launches = h->last_fused_launches;
    return MDB_OK;
}

// read-only counters and timings by name (tests, bench): see include/metdet_b200.h
extern "C" int mdb_get_info(mdb_handle h, const char *name, double *value) {
    if (!h || !name || !value) return fail(MDB_ERR_INVALID, "mdb_get_info: null argument");
    if (!strcmp(name, "temporal_generation")) { *value = h->sk.t_last; return MDB_OK; }
    if (!strcmp(name, "digest_hi")) { *value = (double)(h->last_digest >> 32); return MDB_OK; }
    if (!strcmp(name, "digest_lo")) { *value = (double)(h->last_digest & 0xffffffffull); return MDB_OK; }
    if (!strcmp(name, "temporal_ms")) { *value = h->temporal_ms; return MDB_OK; }
    if (!strcmp(name, "spatial_ms")) { *value = h->spatial_ms; return MDB_OK; }
    if (!strcmp(name, "stream_kernel")) { *value = h->sk.ok && h->use_stream_kernel; return MDB_OK; }
    {   // frames resolved by each PPHT tier: in the batch collected last / since creation ("..._total")
        static const char *tn[4] = {"hough_tier1a", "hough_tier1b", "hough_tier3", "hough_tier2"};
        for (int k = 0; k < 4; k++) {
            if (!strcmp(name, tn[k])) { *value = h->tiers_last[k]; return MDB_OK; }
            const size_t ln = strlen(tn[k]);
            if (!strncmp(name, tn[k], ln) && !strcmp(name + ln, "_total")) { *value = (double)h->tiers_total[k]; return MDB_OK; }
        }
    }
    return fail(MDB_ERR_INVALID, "mdb_get_info: unknown name %s", name);
}

// debug: ms offsets (from the moment the timeline was enabled) of the marks of the batch collected last:
// [0] front start, [1] thresholds done, [2] = ev_f0 (temporal start), [3] = ev_f1 (act done),
// [4] dst done (= back: order start), [5] tier-1 Hough done, [6] all Hough done, [7] results copied,
// [8] = ev_d0 (dst start)
extern "C" int mdb_debug_timeline(mdb_handle h, float *out) {
    if (!h || !out || !h->timeline) return fail(MDB_ERR_INVALID, "mdb_debug_timeline: not enabled");
    BatchCtx &c = h->ctx[h->last_ctx];  // the batch collected last / the per-frame context
    cudaEvent_t ev[9] = {c.tl[0], c.tl[1], c.ev_f0, c.ev_f1, c.tl[4], c.tl[5], c.tl[6], c.tl[7], c.ev_d0};
    for (int i = 0; i < 9; i++)
        if (cudaEventElapsedTime(&out[i], h->tl_base, ev[i]) != cudaSuccess) { cudaGetLastError(); out[i] = -1.f; }
    return MDB_OK;
}

extern "C" int mdb_debug_hough_profile(mdb_handle h, long long *out, int T) {
    if (!h || !out || !h->d_prof) return fail(MDB_ERR_INVALID, "mdb_debug_hough_profile: not enabled");
    CK(cudaMemcpy(out, h->d_prof, (size_t)T * 10 * sizeof(long long), cudaMemcpyDeviceToHost));
    return MDB_OK;
}

// ---- MFNR mix stacker (MetLib/stacker.py:296-403; kernels in mfnr.cuh) ------------------------------------------------
struct mdb_mfnr {
    int H = 0, W = 0, C = 0, device = 0, keep = 0;
    size_t E = 0, P = 0;
    long long n_frames = 0;
    uint8_t *d_max = nullptr;
    uint16_t *d_sum = nullptr;
    uint32_t *d_sq = nullptr;
    std::vector<uint8_t *> chunks;  // retained frames (keep): pointers into `blocks`; else one reusable staging buffer
    std::vector<int> counts;
    struct Block { uint8_t *p; size_t cap, used; };
    std::vector<Block> blocks;      // device memory behind the chunks (grown geometrically, or reserved up front)
    size_t stage_cap = 0;
    uint8_t *d_scratch = nullptr;   // scratch of mdb_mfnr_finish, one allocation, kept for the next call
    size_t scratch_cap = 0;
    cudaStream_t st = nullptr;
};

static void mfnr_free(mdb_mfnr *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->st) cudaStreamSynchronize(m->st);
    for (auto &b : m->blocks) cudaFree(b.p);
    cudaFree(m->d_scratch);
    cudaFree(m->d_max); cudaFree(m->d_sum); cudaFree(m->d_sq);
    if (m->st) cudaStreamDestroy(m->st);
    delete m;
}

extern "C" int mdb_mfnr_create(int height, int width, int channels, int keep_frames, int device, mdb_mfnr_handle *out) {
    if (!out) return fail(MDB_ERR_INVALID, "mdb_mfnr_create: null argument");
    *out = nullptr;
    if (height < 1 || width < 1 || channels < 1 || channels > 4)
        return fail(MDB_ERR_INVALID, "mdb_mfnr_create: unsupported frame shape %dx%dx%d", height, width, channels);
    const int ndev = mdb_device_count();
    if (ndev == 0) return fail(MDB_ERR_CUDA, "mdb_mfnr_create: no CUDA device -- this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(MDB_ERR_INVALID, "mdb_mfnr_create: device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    mdb_mfnr *m = new (std::nothrow) mdb_mfnr();
    if (!m) return fail(MDB_ERR_NOMEM, "mdb_mfnr_create: out of host memory");
    m->H = height; m->W = width; m->C = channels; m->device = device; m->keep = keep_frames;
    m->P = (size_t)height * width; m->E = m->P * channels;
    if (cudaStreamCreateWithFlags(&m->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc((void **)&m->d_max, m->E) != cudaSuccess || cudaMalloc((void **)&m->d_sum, m->E * 2) != cudaSuccess ||
        cudaMalloc((void **)&m->d_sq, m->E * 4) != cudaSuccess) {
        const int rc = fail(MDB_ERR_NOMEM, "mdb_mfnr_create: %s", cudaGetErrorString(cudaGetLastError()));
        mfnr_free(m);
        return rc;
    }
    *out = m;
    return MDB_OK;
}

extern "C" int mdb_mfnr_destroy(mdb_mfnr_handle m) {
    mfnr_free(m);
    return MDB_OK;
}

// room for `frames` more retained frames in one allocation (the loader knows how many frames the clip has)
extern "C" int mdb_mfnr_reserve(mdb_mfnr_handle m, int frames) {
    if (!m || frames < 0) return fail(MDB_ERR_INVALID, "mdb_mfnr_reserve: bad arguments");
    if (!m->keep || frames == 0) return MDB_OK;
    CK(cudaSetDevice(m->device));
    const size_t cap = (size_t)frames * ((m->E + 255) / 256 * 256);
    if (!m->blocks.empty() && m->blocks.back().cap - m->blocks.back().used >= cap) return MDB_OK;
    uint8_t *p = nullptr;
    if (cudaMalloc((void **)&p, cap) != cudaSuccess)
        return fail(MDB_ERR_NOMEM, "mdb_mfnr_reserve: %zu bytes for %d frames: %s", cap, frames, cudaGetErrorString(cudaGetLastError()));
    m->blocks.push_back({p, cap, 0});
    return MDB_OK;
}

extern "C" int mdb_mfnr_append(mdb_mfnr_handle m, const uint8_t *frames, int T, int on_device) {
    if (!m || !frames || T < 1) return fail(MDB_ERR_INVALID, "mdb_mfnr_append: bad arguments");
    CK(cudaSetDevice(m->device));
    const size_t bytes = (size_t)T * m->E;
    uint8_t *buf = nullptr;
    if (m->keep) {
        if (m->blocks.empty() || m->blocks.back().cap - m->blocks.back().used < bytes) {
            // no room in the last block: a new one, twice the size of the last (at least this chunk, at least 256 MB)
            size_t cap = std::max<size_t>(bytes, (size_t)256 << 20);
            if (!m->blocks.empty()) cap = std::max(cap, std::min<size_t>(2 * m->blocks.back().cap, (size_t)16 << 30));
            uint8_t *p = nullptr;
            if (cudaMalloc((void **)&p, cap) != cudaSuccess) {
                cudaGetLastError();
                cap = bytes;
                if (cudaMalloc((void **)&p, cap) != cudaSuccess)
                    return fail(MDB_ERR_NOMEM, "mdb_mfnr_append: %zu bytes for %d frames: %s", bytes, T, cudaGetErrorString(cudaGetLastError()));
            }
            m->blocks.push_back({p, cap, 0});
        }
        mdb_mfnr::Block &b = m->blocks.back();
        buf = b.p + b.used;
        b.used += (bytes + 255) / 256 * 256 <= b.cap - b.used ? (bytes + 255) / 256 * 256 : bytes;
        m->chunks.push_back(buf);
        m->counts.push_back(T);
    } else {
        if (m->blocks.empty() || m->stage_cap < bytes) {  // (re)allocate the staging buffer
            CK(cudaStreamSynchronize(m->st));
            for (auto &b : m->blocks) cudaFree(b.p);
            m->blocks.clear();
            uint8_t *p = nullptr;
            if (cudaMalloc((void **)&p, bytes) != cudaSuccess)
                return fail(MDB_ERR_NOMEM, "mdb_mfnr_append: %zu bytes for %d frames: %s", bytes, T, cudaGetErrorString(cudaGetLastError()));
            m->blocks.push_back({p, bytes, bytes});
            m->stage_cap = bytes;
        }
        buf = m->blocks[0].p;
    }
    CK(cudaMemcpyAsync(buf, frames, bytes, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, m->st));
    const int first = m->n_frames == 0;
    // (16 elements per thread were measured 2.4x slower: registers, strided element stores)
    if (m->E % 4 == 0) {
        const size_t thr = m->E / 4;
        mfnr_accum_kernel<4><<<(unsigned)((thr + MF_THREADS - 1) / MF_THREADS), MF_THREADS, 0, m->st>>>(buf, T, m->E, m->d_max, m->d_sum, m->d_sq, first);
    } else {
        mfnr_accum_kernel<1><<<(unsigned)((m->E + MF_THREADS - 1) / MF_THREADS), MF_THREADS, 0, m->st>>>(buf, T, m->E, m->d_max, m->d_sum, m->d_sq, first);
    }
    CK(cudaGetLastError());
    if (!on_device) CK(cudaStreamSynchronize(m->st));  // the caller may reuse its buffer
    m->n_frames += T;
    return MDB_OK;
}

// the running statistics as the reference's containers hold them: MaxImgContainer.container (uint8 max),
// FastGaussianContainer.container.sum_mu (uint16) / .square_sum (uint32); any of the three may be NULL
extern "C" int mdb_mfnr_stats(mdb_mfnr_handle m, uint8_t *max_out, uint16_t *sum_out, uint32_t *sq_out, int64_t *n_frames) {
    if (!m) return fail(MDB_ERR_INVALID, "mdb_mfnr_stats: null handle");
    if (n_frames) *n_frames = m->n_frames;
    if (m->n_frames == 0) return (max_out || sum_out || sq_out) ? fail(MDB_ERR_STATE, "mdb_mfnr_stats: no frame appended yet") : MDB_OK;
    CK(cudaSetDevice(m->device));
    if (max_out) CK(cudaMemcpyAsync(max_out, m->d_max, m->E, cudaMemcpyDeviceToHost, m->st));
    if (sum_out) CK(cudaMemcpyAsync(sum_out, m->d_sum, m->E * 2, cudaMemcpyDeviceToHost, m->st));
    if (sq_out) CK(cudaMemcpyAsync(sq_out, m->d_sq, m->E * 4, cudaMemcpyDeviceToHost, m->st));
    CK(cudaStreamSynchronize(m->st));
    return MDB_OK;
}

extern "C" int mdb_mfnr_finish(mdb_mfnr_handle m, const mdb_mfnr_params *prm, uint8_t *out, int out_on_device, double *stats) {
    if (!m || !prm || !out) return fail(MDB_ERR_INVALID, "mdb_mfnr_finish: null argument");
    if (m->n_frames < 2) return fail(MDB_ERR_STATE, "mdb_mfnr_finish: %lld frames appended, at least 2 are needed", m->n_frames);
    if (m->n_frames > 32767) return fail(MDB_ERR_INVALID, "mdb_mfnr_finish: %lld frames: the reference's int16 frame count wraps beyond 32767", m->n_frames);
    if (prm->blur_ksize < 1 || prm->blur_ksize % 2 == 0 || prm->blur_ksize > 255)
        return fail(MDB_ERR_INVALID, "mdb_mfnr_finish: blur_ksize %d must be odd and in 1..255", prm->blur_ksize);
    if (prm->bg_algorithm < 0 || prm->bg_algorithm > 3)
        return fail(MDB_ERR_INVALID, "mdb_mfnr_finish: bg_algorithm %d (0 = mean, 1 = sigma-clipping, 2 = median, 3 = med-of-med)", prm->bg_algorithm);
    if (prm->bg_algorithm != 0 && !m->keep)
        return fail(MDB_ERR_STATE, "mdb_mfnr_finish: this background algorithm needs the frames (create with keep_frames = 1)");
    const bool med = prm->bg_algorithm >= 2;
    // stacker.py:343-347: "median", or "med-of-med" on at most 16 frames -> np.median over all frames
    int med_block = 0;
    if (prm->bg_algorithm == 3 && m->n_frames > 16) {
        med_block = prm->med_block_size > 0 ? prm->med_block_size : (int)std::sqrt((double)m->n_frames);
        if (med_block < 1 || (m->n_frames - 1) / med_block + 1 > MF_MAX_BLOCKS)
            return fail(MDB_ERR_INVALID, "mdb_mfnr_finish: block size %d gives more than %d blocks", med_block, MF_MAX_BLOCKS);
    }
    CK(cudaSetDevice(m->device));
    const int N = (int)m->n_frames, ks = prm->blur_ksize;
    const size_t E = m->E, P = m->P;
    // scratch (clipped sums, partials, mask, blur planes, kernel taps, output, chunk table): ONE allocation, kept in the handle
    const bool clip = prm->bg_algorithm == 1;
    size_t off = 0;
    auto carve = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
    const size_t o_pv = carve(MF_PARTS * sizeof(double)), o_pc = carve(MF_PARTS * sizeof(unsigned long long)),
                 o_tot = carve(sizeof(double)), o_cnt = carve(sizeof(unsigned long long)), o_fg = carve(P),
                 o_row = carve(P * sizeof(double)), o_blur = carve(P * sizeof(double)), o_k = carve(ks * sizeof(double)),
                 o_out = carve(out_on_device ? 0 : E), o_sum2 = carve(clip ? E * 2 : 0), o_sq2 = carve(clip ? E * 4 : 0),
                 o_n2 = carve(clip ? E * 4 : 0), o_cptr = carve(clip ? m->chunks.size() * sizeof(uint8_t *) : 0),
                 o_ccnt = carve(clip ? m->chunks.size() * sizeof(int) : 0), o_mu = carve(med ? E * sizeof(float) : 0),
                 o_fptr = carve(med ? (size_t)N * sizeof(uint8_t *) : 0);
    if (m->scratch_cap < off) {
        CK(cudaStreamSynchronize(m->st));
        cudaFree(m->d_scratch);
        m->d_scratch = nullptr; m->scratch_cap = 0;
        if (cudaMalloc((void **)&m->d_scratch, off) != cudaSuccess)
            return fail(MDB_ERR_NOMEM, "mdb_mfnr_finish: %zu bytes of scratch: %s", off, cudaGetErrorString(cudaGetLastError()));
        m->scratch_cap = off;
    }
    uint8_t *sc = m->d_scratch;
    double *d_pv = (double *)(sc + o_pv), *d_tot = (double *)(sc + o_tot), *d_row = (double *)(sc + o_row),
           *d_blur = (double *)(sc + o_blur), *d_k = (double *)(sc + o_k);
    unsigned long long *d_pc = (unsigned long long *)(sc + o_pc), *d_cnt = (unsigned long long *)(sc + o_cnt);
    uint8_t *d_fg = sc + o_fg, *d_out = sc + o_out;
    uint16_t *d_sum2 = (uint16_t *)(sc + o_sum2);
    uint32_t *d_sq2 = (uint32_t *)(sc + o_sq2);
    int32_t *d_n2 = (int32_t *)(sc + o_n2);
    const uint8_t **d_cptr = (const uint8_t **)(sc + o_cptr);
    int *d_ccnt = (int *)(sc + o_ccnt);
    float *d_mu = med ? (float *)(sc + o_mu) : nullptr;
    const uint8_t **d_fptr = (const uint8_t **)(sc + o_fptr);
#define MF_TRY(expr)                                                                      \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            cudaStreamSynchronize(m->st);                                                 \
            return fail(MDB_ERR_CUDA, "mdb_mfnr_finish: %s", cudaGetErrorString(e_));      \
        }                                                                                 \
    } while (0)
    const unsigned gE = (unsigned)((E + MF_THREADS - 1) / MF_THREADS), gP = (unsigned)((P + MF_THREADS - 1) / MF_THREADS);
    const uint16_t *sum = m->d_sum;
    const int32_t *n_arr = nullptr;
    if (clip) {
        MF_TRY(cudaMemcpyAsync(d_cptr, m->chunks.data(), m->chunks.size() * sizeof(uint8_t *), cudaMemcpyHostToDevice, m->st));
        MF_TRY(cudaMemcpyAsync(d_ccnt, m->counts.data(), m->counts.size() * sizeof(int), cudaMemcpyHostToDevice, m->st));
        MfnrChunks ch;
        ch.ptr = d_cptr; ch.count = d_ccnt; ch.n = (int)m->chunks.size();
        // stacker.py:333-336 passes sigma_high = sigma_low = 3.0 whatever the configuration holds; the caller decides
        if (E % 4 == 0)
            mfnr_sigma_kernel<4><<<(unsigned)((E / 4 + MF_THREADS - 1) / MF_THREADS), MF_THREADS, 0, m->st>>>(
                ch, E, N, prm->sigma_high, prm->sigma_low, m->d_sum, m->d_sq, d_sum2, d_sq2, d_n2);
        else
            mfnr_sigma_kernel<1><<<gE, MF_THREADS, 0, m->st>>>(ch, E, N, prm->sigma_high, prm->sigma_low, m->d_sum, m->d_sq, d_sum2,
                                                               d_sq2, d_n2);
        MF_TRY(cudaGetLastError());
        sum = d_sum2;
        n_arr = d_n2;
    }
    if (med) {
        std::vector<const uint8_t *> fp;  // one pointer per frame, in arrival order
        for (size_t c = 0; c < m->chunks.size(); c++)
            for (int t = 0; t < m->counts[c]; t++) fp.push_back(m->chunks[c] + (size_t)t * E);
        MF_TRY(cudaMemcpyAsync(d_fptr, fp.data(), fp.size() * sizeof(uint8_t *), cudaMemcpyHostToDevice, m->st));
        MF_TRY(cudaStreamSynchronize(m->st));  // fp is a local
        if (E % 4 == 0)
            mfnr_median_kernel<4><<<(unsigned)((E / 4 + MF_THREADS - 1) / MF_THREADS), MF_THREADS, 0, m->st>>>(d_fptr, E, N, med_block, d_mu);
        else
            mfnr_median_kernel<1><<<gE, MF_THREADS, 0, m->st>>>(d_fptr, E, N, med_block, d_mu);
        MF_TRY(cudaGetLastError());
    }
    const uint32_t *sq = prm->bg_algorithm == 1 ? d_sq2 : m->d_sq;
    double tot = 0.0;
    unsigned long long cnt = 0;
    mfnr_sqrtvar_kernel<<<MF_PARTS, MF_THREADS, 0, m->st>>>(E, N, sum, sq, n_arr, d_pv, d_pc);
    mfnr_final_reduce_kernel<<<1, 32, 0, m->st>>>(MF_PARTS, d_pv, d_pc, d_tot, d_cnt);
    MF_TRY(cudaGetLastError());
    MF_TRY(cudaMemcpyAsync(&tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, m->st));
    MF_TRY(cudaStreamSynchronize(m->st));
    const double est_bg_var = tot / (double)E;
    // get_gumbel_mean (stacker.py:118-126): the caller may hand in the value its own numpy computed
    double g = prm->gumbel_mean;
    if (!(g > 0.0)) {
        const double s2 = std::sqrt(2.0 * std::log((double)N));
        g = s2 - (std::log(std::log((double)N)) + std::log(4.0 * 3.141592653589793)) / (2.0 * s2) + 0.5772 / s2;
    }
    volatile double c2v = est_bg_var * g;            // (est_bg_var * gumble_mean)
    volatile double c1v = c2v * prm->bg_fix_factor;  // est_bg_var * gumble_mean * bg_fix_factor, left to right
    const double c1 = c1v, c2 = c2v;
    mfnr_diffpos_kernel<<<MF_PARTS, MF_THREADS, 0, m->st>>>(E, N, c1, m->d_max, sum, n_arr, d_mu, d_pv, d_pc);
    mfnr_final_reduce_kernel<<<1, 32, 0, m->st>>>(MF_PARTS, d_pv, d_pc, d_tot, d_cnt);
    MF_TRY(cudaGetLastError());
    MF_TRY(cudaMemcpyAsync(&tot, d_tot, sizeof tot, cudaMemcpyDeviceToHost, m->st));
    MF_TRY(cudaMemcpyAsync(&cnt, d_cnt, sizeof cnt, cudaMemcpyDeviceToHost, m->st));
    MF_TRY(cudaStreamSynchronize(m->st));
    const double avg = tot / (double)cnt;  // np.average of an empty selection is NaN there too
    std::vector<double> taps(ks);
    {   // cv2.getGaussianKernel(ksize, sigma, CV_64F), computed branch
        const double sigma = prm->blur_sigma > 0 ? prm->blur_sigma : 3.0;
        double ssum = 0.0;
        for (int i = 0; i < ks; i++) {
            const double x = i - (ks - 1) * 0.5;
            taps[i] = std::exp(-(x * x) / (2.0 * sigma * sigma));
            ssum += taps[i];
        }
        for (int i = 0; i < ks; i++) taps[i] = taps[i] / ssum;
    }
    MF_TRY(cudaMemcpyAsync(d_k, taps.data(), ks * sizeof(double), cudaMemcpyHostToDevice, m->st));
    volatile double hl = 255.0 * prm->highlight_preserve, omh = 1.0 - prm->highlight_preserve;
    mfnr_mask_kernel<<<gP, MF_THREADS, 0, m->st>>>(P, m->C, N, c1, avg, hl, m->d_max, sum, n_arr, d_mu, d_fg);
    mfnr_blur_row_kernel<<<gP, MF_THREADS, 0, m->st>>>(m->H, m->W, ks, d_k, d_fg, d_row);
    mfnr_blur_col_kernel<<<gP, MF_THREADS, 0, m->st>>>(m->H, m->W, ks, d_k, d_row, d_blur);
    uint8_t *dst = out_on_device ? out : d_out;
    mfnr_mix_kernel<<<gE, MF_THREADS, 0, m->st>>>(E, m->C, N, c2, prm->highlight_preserve, omh, m->d_max, sum, n_arr, d_mu, d_blur, dst);
    MF_TRY(cudaGetLastError());
    if (!out_on_device) MF_TRY(cudaMemcpyAsync(out, d_out, E, cudaMemcpyDeviceToHost, m->st));
    MF_TRY(cudaStreamSynchronize(m->st));
#undef MF_TRY
    if (stats) { stats[0] = est_bg_var; stats[1] = g; stats[2] = avg; stats[3] = (double)cnt; }
    return MDB_OK;
}

// internal knobs used by the parity tests (force the generic per-frame kernel) and debugging
extern "C" int mdb_set_option(mdb_handle h, const char *name, int value) {
    if (!h || !name) return fail(MDB_ERR_INVALID, "mdb_set_option: null argument");
    if (!strcmp(name, "hough_profile")) {
        if (value && !h->d_prof) {
            CK(cudaMalloc((void **)&h->d_prof, (size_t)h->cfg.max_batch * 10 * sizeof(long long)));
            CK(cudaMemset(h->d_prof, 0, (size_t)h->cfg.max_batch * 10 * sizeof(long long)));
        }
        return MDB_OK;
    }
    if (!strcmp(name, "stream_kernel")) { h->use_stream_kernel = value; return MDB_OK; }
    if (!strcmp(name, "hough_ctas")) { h->hough_ctas = value; return MDB_OK; }
    if (!strcmp(name, "per_frame_fast")) { h->per_frame_fast = value; h->pf_timer = h->pf_bits_timer = -1; return MDB_OK; }
    if (!strcmp(name, "timeline")) {
        if (value && !h->tl_base) {
            for (BatchCtx &c : h->ctx)
                for (cudaEvent_t &e : c.tl) CK(cudaEventCreate(&e));
            CK(cudaEventCreate(&h->tl_base));
            CK(cudaEventRecord(h->tl_base, h->stream));
        }
        h->timeline = value;
        return MDB_OK;
    }
    if (!strcmp(name, "temporal_nt")) {
        if (!h->sk.ok || stream_state_config(h->sk, h->sk.t_wpt, value) != 0)
            return fail(MDB_ERR_INVALID, "mdb_set_option: temporal_nt=%d not possible here", value);
        return MDB_OK;
    }
    if (!strcmp(name, "temporal_kdiv")) {
        h->sk.t_kdiv_req = value;
        if (!h->sk.ok || stream_state_config(h->sk, h->sk.t_wpt) != 0)
            return fail(MDB_ERR_INVALID, "mdb_set_option: temporal_kdiv=%d not possible here", value);
        return MDB_OK;
    }
    if (!strcmp(name, "temporal_version")) { h->sk.t_version = value == 2 ? 2 : 3; return MDB_OK; }
    if (!strcmp(name, "t3_variant")) { h->sk.t3_variant = value; return MDB_OK; }
    if (!strcmp(name, "force_dense")) { h->sk.force_dense = value; return MDB_OK; }
    if (!strcmp(name, "force_strip")) { h->sk.force_strip = value; return MDB_OK; }
    if (!strcmp(name, "sp_rows")) { h->sk.sp_rows = value; return MDB_OK; }
    if (!strcmp(name, "dst_rows")) { h->sk.dst_rows = value; return MDB_OK; }
    if (!strcmp(name, "sp_rows_single")) { h->sk.sp_rows_single = value > 0 ? value : 8; return MDB_OK; }
    if (!strcmp(name, "single_dense")) { h->sk.single_dense = value; return MDB_OK; }
    if (!strcmp(name, "temporal_wpt")) {
        if (!h->sk.ok || stream_state_config(h->sk, value) != 0)
            return fail(MDB_ERR_INVALID, "mdb_set_option: temporal_wpt=%d not possible here", value);
        return MDB_OK;
    }
    return fail(MDB_ERR_INVALID, "mdb_set_option: unknown option %s", name);
}

// ---- max stack --------------------------------------------------------------------------------
extern "C" int mdb_max_stack(const uint8_t *frames, int T, size_t frame_bytes, uint8_t *out,
                             int frames_on_device, int out_on_device, int device) {
    if (!frames || !out || T < 1 || frame_bytes == 0)
        return fail(MDB_ERR_INVALID, "mdb_max_stack: bad arguments");
    if (mdb_device_count() == 0) return fail(MDB_ERR_CUDA, "mdb_max_stack: no CUDA device (no CPU fallback)");
    CK(cudaSetDevice(device));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint8_t *d_out = out, *d_chunk = nullptr;
    cudaError_t e = cudaSuccess;
    if (!out_on_device) e = cudaMalloc((void **)&d_out, frame_bytes);
    const int chunk = frames_on_device ? T : (int)std::max<size_t>(1, std::min<size_t>(T, (512ull << 20) / frame_bytes));
    if (e == cudaSuccess && !frames_on_device) e = cudaMalloc((void **)&d_chunk, (size_t)chunk * frame_bytes);
    const int grid = (int)std::min<size_t>((frame_bytes / 16 + 255) / 256 + 1, 148 * 16);
    for (int t0 = 0; e == cudaSuccess && t0 < T; t0 += chunk) {
        const int c = std::min(chunk, T - t0);
        const uint8_t *src = frames + (size_t)t0 * frame_bytes;
        if (!frames_on_device) {
            e = cudaMemcpyAsync(d_chunk, src, (size_t)c * frame_bytes, cudaMemcpyHostToDevice, st);
            src = d_chunk;
        }
        if (e == cudaSuccess) {
            max_stack_kernel<<<grid, 256, 0, st>>>(src, c, frame_bytes, d_out, t0 > 0);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && !frames_on_device) e = cudaStreamSynchronize(st);
    }
    if (e == cudaSuccess && !out_on_device) e = cudaMemcpyAsync(out, d_out, frame_bytes, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (d_chunk) cudaFree(d_chunk);
    if (!out_on_device && d_out) cudaFree(d_out);
    cudaStreamDestroy(st);
    if (e != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_max_stack: %s", cudaGetErrorString(e));
    return MDB_OK;
}

// ---- Gaussian stack (sum, sum of squares) ------------------------------------------------------
extern "C" int mdb_gauss_stack(const uint8_t *frames, int T, size_t frame_bytes, uint16_t *sum_out, uint32_t *sq_out,
                               int frames_on_device, int out_on_device, int accumulate, int device) {
    if (!frames || !sum_out || !sq_out || T < 1 || frame_bytes == 0)
        return fail(MDB_ERR_INVALID, "mdb_gauss_stack: bad arguments");
    if (mdb_device_count() == 0) return fail(MDB_ERR_CUDA, "mdb_gauss_stack: no CUDA device (no CPU fallback)");
    CK(cudaSetDevice(device));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint16_t *d_sum = sum_out;
    uint32_t *d_sq = sq_out;
    uint8_t *d_chunk = nullptr;
    cudaError_t e = cudaSuccess;
    if (!out_on_device) {
        e = cudaMalloc((void **)&d_sum, frame_bytes * 2);
        if (e == cudaSuccess) e = cudaMalloc((void **)&d_sq, frame_bytes * 4);
        if (e == cudaSuccess && accumulate) e = cudaMemcpyAsync(d_sum, sum_out, frame_bytes * 2, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess && accumulate) e = cudaMemcpyAsync(d_sq, sq_out, frame_bytes * 4, cudaMemcpyHostToDevice, st);
    }
    const int chunk = frames_on_device ? T : (int)std::max<size_t>(1, std::min<size_t>(T, (512ull << 20) / frame_bytes));
    if (e == cudaSuccess && !frames_on_device) e = cudaMalloc((void **)&d_chunk, (size_t)chunk * frame_bytes);
    const int grid = (int)std::min<size_t>((frame_bytes / 16 + 255) / 256 + 1, 148 * 32);
    for (int t0 = 0; e == cudaSuccess && t0 < T; t0 += chunk) {
        const int c = std::min(chunk, T - t0);
        const uint8_t *src = frames + (size_t)t0 * frame_bytes;
        if (!frames_on_device) {
            e = cudaMemcpyAsync(d_chunk, src, (size_t)c * frame_bytes, cudaMemcpyHostToDevice, st);
            src = d_chunk;
        }
        if (e == cudaSuccess) {
            gauss_stack_kernel<<<grid, 256, 0, st>>>(src, c, frame_bytes, d_sum, d_sq, accumulate || t0 > 0);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && !frames_on_device) e = cudaStreamSynchronize(st);
    }
    if (e == cudaSuccess && !out_on_device) {
        e = cudaMemcpyAsync(sum_out, d_sum, frame_bytes * 2, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(sq_out, d_sq, frame_bytes * 4, cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (d_chunk) cudaFree(d_chunk);
    if (!out_on_device) { if (d_sum) cudaFree(d_sum); if (d_sq) cudaFree(d_sq); }
    cudaStreamDestroy(st);
    if (e != cudaSuccess) return fail(MDB_ERR_CUDA, "mdb_gauss_stack: %s", cudaGetErrorString(e));
    return MDB_OK;
}

// ---- loader preprocessing ---------------------------------------------------------------------
struct mdb_preproc {
    PreParams P;
    int device = 0, max_out = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    PreTap *d_xt = nullptr, *d_yt = nullptr;
    uint8_t *d_mask = nullptr, *d_out = nullptr, *d_in = nullptr;
    size_t in_cap = 0;
    float last_ms = 0.f;
};

// OpenCV's 8-bit INTER_LINEAR taps: float32 position, 11-bit weights (checked against the test-suite's CPU checker)
static void pre_axis_taps(int dst, int src, bool clamp_fraction, int stride, std::vector<PreTap> &taps) {
    taps.resize(dst);
    const double scale = (double)src / (double)dst;
    for (int d = 0; d < dst; d++) {
        volatile float f = (float)(((double)d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        volatile float fr = f - (float)s;
        if (clamp_fraction) {
            if (s < 0) { fr = 0.f; s = 0; }
            if (s >= src - 1) { fr = 0.f; s = src - 1; }
        }
        volatile float one_minus = 1.0f - fr;
        volatile float p1 = fr * 2048.0f, p0 = one_minus * 2048.0f;
        PreTap t;
        t.w1 = (int)nearbyintf(p1);  // cvRound: half to even (default rounding mode)
        t.w0 = (int)nearbyintf(p0);
        t.s0 = std::min(std::max(s, 0), src - 1) * stride;
        t.s1 = std::min(std::max(s + 1, 0), src - 1) * stride;
        taps[d] = t;
    }
}

extern "C" int mdb_preproc_axis_taps(int dst, int src, int clamp_fraction, int32_t *s0, int32_t *s1, int32_t *w0,
                                     int32_t *w1) {
    if (dst < 1 || src < 1 || !s0 || !s1 || !w0 || !w1) return fail(MDB_ERR_INVALID, "mdb_preproc_axis_taps: bad arguments");
    std::vector<PreTap> t;
    pre_axis_taps(dst, src, clamp_fraction != 0, 1, t);
    for (int d = 0; d < dst; d++) { s0[d] = t[d].s0; s1[d] = t[d].s1; w0[d] = t[d].w0; w1[d] = t[d].w1; }
    return MDB_OK;
}

static void preproc_free(mdb_preproc *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void *dev[] = {h->d_xt, h->d_yt, h->d_mask, h->d_out, h->d_in};
    for (void *p : dev)
        if (p) cudaFree(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

extern "C" int mdb_preproc_create(int src_w, int src_h, int channels, int rgb_order, int dst_w, int dst_h,
                                  const uint8_t *mask, int exp_frame, int max_out, int device,
                                  mdb_preproc_handle *out) {
    if (!out) return fail(MDB_ERR_INVALID, "mdb_preproc_create: null argument");
    *out = nullptr;
    if (src_w < 1 || src_h < 1 || dst_w < 1 || dst_h < 1 || src_w > 65535 || src_h > 65535 || dst_w > 65535 || dst_h > 65535)
        return fail(MDB_ERR_INVALID, "mdb_preproc_create: bad sizes %dx%d -> %dx%d", src_w, src_h, dst_w, dst_h);
    if (channels != 1 && channels != 3)
        return fail(MDB_ERR_INVALID, "mdb_preproc_create: channels must be 1 or 3 (got %d)", channels);
    if (exp_frame < 1 || max_out < 1) return fail(MDB_ERR_INVALID, "mdb_preproc_create: exp_frame and max_out must be >= 1");
    const int ndev = mdb_device_count();
    if (ndev == 0) return fail(MDB_ERR_CUDA, "mdb_preproc_create: no CUDA device -- this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(MDB_ERR_INVALID, "mdb_preproc_create: device %d of %d", device, ndev);
    CK(cudaSetDevice(device));
    mdb_preproc *h = new (std::nothrow) mdb_preproc();
    if (!h) return fail(MDB_ERR_NOMEM, "mdb_preproc_create: out of host memory");
    h->device = device; h->max_out = max_out;
    PreParams &P = h->P;
    P.src_w = src_w; P.src_h = src_h; P.channels = channels; P.dst_w = dst_w; P.dst_h = dst_h;
    P.resize = (src_w != dst_w || src_h != dst_h) ? 1 : 0;
    P.rgb = rgb_order ? 1 : 0; P.exp_frame = exp_frame;
    P.xt = P.yt = nullptr; P.mask = nullptr;
    std::vector<PreTap> xt, yt;
    pre_axis_taps(dst_w, src_w, true, channels, xt);
    pre_axis_taps(dst_h, src_h, false, 1, yt);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_xt, xt.size() * sizeof(PreTap));
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_yt, yt.size() * sizeof(PreTap));
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_out, (size_t)max_out * dst_w * dst_h);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_xt, xt.data(), xt.size() * sizeof(PreTap), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_yt, yt.data(), yt.size() * sizeof(PreTap), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && mask) {
        e = cudaMalloc((void **)&h->d_mask, (size_t)dst_w * dst_h);
        if (e == cudaSuccess) e = cudaMemcpy(h->d_mask, mask, (size_t)dst_w * dst_h, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        int rc = fail(MDB_ERR_CUDA, "mdb_preproc_create: %s", cudaGetErrorString(e));
        preproc_free(h);
        return rc;
    }
    P.xt = h->d_xt; P.yt = h->d_yt; P.mask = h->d_mask;
    *out = h;
    return MDB_OK;
}

extern "C" int mdb_preproc_run(mdb_preproc_handle h, const uint8_t *frames, int T, int frames_on_device,
                               uint8_t *out, int out_on_device, int32_t *n_out) {
    if (!h || !frames) return fail(MDB_ERR_INVALID, "mdb_preproc_run: null argument");
    if (T < 1) return fail(MDB_ERR_INVALID, "mdb_preproc_run: T=%d", T);
    const PreParams &P = h->P;
    const int G = (T + P.exp_frame - 1) / P.exp_frame;
    if (G > h->max_out) return fail(MDB_ERR_INVALID, "mdb_preproc_run: %d output frames exceed max_out=%d", G, h->max_out);
    CK(cudaSetDevice(h->device));
    const size_t in_bytes = (size_t)T * P.src_w * P.src_h * P.channels;
    const uint8_t *d_frames = frames;
    if (!frames_on_device) {
        if (h->in_cap < in_bytes) {
            if (h->d_in) CK(cudaFree(h->d_in));
            h->d_in = nullptr; h->in_cap = 0;
            if (cudaMalloc((void **)&h->d_in, in_bytes) != cudaSuccess) {
                cudaGetLastError();
                return fail(MDB_ERR_NOMEM, "mdb_preproc_run: cudaMalloc(%zu bytes) for the source frames", in_bytes);
            }
            h->in_cap = in_bytes;
        }
        CK(cudaMemcpyAsync(h->d_in, frames, in_bytes, cudaMemcpyHostToDevice, h->stream));
        d_frames = h->d_in;
    }
    dim3 grid((P.dst_w + 255) / 256, P.dst_h, G);
    CK(cudaEventRecord(h->ev0, h->stream));
    if (P.channels == 3) preproc_kernel<3><<<grid, 256, 0, h->stream>>>(P, d_frames, T, h->d_out);
    else preproc_kernel<1><<<grid, 256, 0, h->stream>>>(P, d_frames, T, h->d_out);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev1, h->stream));
    if (out)
        CK(cudaMemcpyAsync(out, h->d_out, (size_t)G * P.dst_w * P.dst_h,
                           out_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    if (n_out) *n_out = G;
    return MDB_OK;
}

extern "C" int mdb_preproc_output(mdb_preproc_handle h, const uint8_t **device_ptr) {
    if (!h || !device_ptr) return fail(MDB_ERR_INVALID, "mdb_preproc_output: null argument");
    *device_ptr = h->d_out;
    return MDB_OK;
}

extern "C" int mdb_preproc_time(mdb_preproc_handle h, float *ms) {
    if (!h || !ms) return fail(MDB_ERR_INVALID, "mdb_preproc_time: null argument");
    *ms = h->last_ms;
    return MDB_OK;
}

extern "C" int mdb_preproc_destroy(mdb_preproc_handle h) {
    if (!h) return fail(MDB_ERR_INVALID, "mdb_preproc_destroy: null handle");
    preproc_free(h);
    return MDB_OK;
}

extern "C" int mdb_alloc_pinned(size_t bytes, void **ptr) {
    if (!ptr) return fail(MDB_ERR_INVALID, "mdb_alloc_pinned: null argument");
    CK(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return MDB_OK;
}
extern "C" int mdb_free_pinned(void *ptr) {
    CK(cudaFreeHost(ptr));
    return MDB_OK;
}
