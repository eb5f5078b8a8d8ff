// Shared declarations for libmetdet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/metdet_b200.h"

#define MDB_HOUGH_ANGLES 180
#define MDB_POINT_CAP 16384    // per-frame on-pixel list capacity (= most points the tier-2 PPHT kernel takes)
#define HOUGH_CAP_LARGE 4096   // most points of the shared-memory PPHT tier 1b
#define HOUGH_ORDER_CAP 4096   // the PPHT visiting order is precomputed per batch for frames up to this size

// Where frame `t` (0-based global frame index) lives: frames of the batch in flight may be read
// straight from the caller's device buffer (`cur`, frame t0 at offset 0: zero-copy), everything
// older -- and host-fed batches -- from slot t % R of the device frame ring.
struct FrameSrc {
    const uint8_t *ring;
    const uint8_t *cur;   // nullptr: every frame is in the ring
    const uint8_t *mask;  // {0,1} bytes, or nullptr when frames arrive already masked
    long long t0;         // global index of cur's first frame
    int R;
    size_t HW;
    __device__ __forceinline__ const uint8_t *frame(long long t) const {
        return (cur && t >= t0) ? cur + (size_t)(t - t0) * HW : ring + (size_t)(t % R) * HW;
    }
    __device__ __forceinline__ unsigned px(long long t, size_t p) const {
        unsigned v = frame(t)[p];
        return mask ? v * mask[p] : v;
    }
};

// `act` bit-frames (mask after the close, before the dynamic mask; Detector.py:335) of the last n-1
// detects plus the current batch: the dynamic mask of Detector.py:234-242 is "not on in ALL of the
// last L act frames", so this ring replaces the reference's second n-frame SlidingWindow.
struct ActRing {
    uint32_t *base;  // [RA][H][Wb] 32-pixel words, row stride Wb = ceil(W/32)
    int RA, Wb;
    size_t frame_words;
    __device__ __forceinline__ uint32_t *frame(long long d) const {
        return base + (size_t)(d % RA) * frame_words;
    }
};

// Scalar detector state that lives on the device so that batches never round-trip to the host.
struct DevState {
    double ema_value;   // EMA.cur_value           utils.py:343
    double ema_cur_m;   // EMA.cur_momentum
    double ema_init_m;  // EMA.init_momentum
    double ema_warm;    // EMA.warmup_speed (0 once warm-up ended)
    long long ema_t;    // EMA.t
    double thr_float;   // LineDetector.bi_threshold_float
    int bi_threshold;   // LineDetector.bi_threshold
    int pad;
};

struct HoughParams {
    int W, H, numrho;
    int threshold, min_len, max_gap;
    double mask_area;
    int cap;        // point-list stride / smem capacity
    int max_lines;  // rows stored per frame
    int walk_cap;   // capacity of the per-slot line-pixel scratch
    int fixed_gap;  // >= 0: maxLineGap is this constant (ClassicDetector); < 0: M3's adaptive gap
};
