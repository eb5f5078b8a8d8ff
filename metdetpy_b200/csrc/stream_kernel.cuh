// Time-tiled streaming fused kernel (placeholder until the persistent kernel lands).
#pragma once
#include "common.cuh"

struct StreamState {
    int ok = 0;
};
static inline int stream_state_init(StreamState &, int, int, int, int) { return 0; }
static inline void stream_state_free(StreamState &) {}
static inline bool stream_kernel_supported(const StreamState &, int) { return false; }
static inline int stream_kernel_launch(StreamState &, FrameSrc, long long, long long, int, int, const int *,
                                       uint8_t *, uint8_t *, unsigned *, uint32_t *, int, cudaStream_t,
                                       int *) { return -1; }
