// MFNR mix stacker on the device (second half): mfnr_mix_stacker, MetLib/stacker.py:296-403, with
// connect_lines off and the background algorithms "mean" (:339-342), "sigma-clipping" (:333-338, :94-115), "median" and
// "med-of-med" (:343-349, :62-78).
// All frames of a clip (colour, full resolution) are accumulated as they arrive: per element the max (MaxImgContainer,
// :43-49) and the uint16 sum / uint32 sum of squares of FastGaussianParam (utils.py:435-493, wrapping like numpy);
// sigma clipping needs a second pass over the frames, which therefore stay resident in HBM (a 300-frame 4K colour
// clip is 7.5 GB).  The finishing passes are float64 like the reference (B200 has real FP64 units; the passes are
// HBM-bound): two global means (deterministic two-level reductions), the foreground mask, a separable 31-tap Gaussian
// of that mask (cv2.GaussianBlur on float64: sequential row taps, symmetric column taps, BORDER_REFLECT_101) and the mix.
// The library is built with --fmad=false: no contraction anywhere below.
#pragma once
#include "common.cuh"

#define MF_THREADS 256
#define MF_PARTS 1024  // blocks of the reduction kernels = partial sums

// V consecutive bytes at p (V = 1, 4 or 16; p aligned to V) as V unsigned values
template <int V>
__device__ __forceinline__ void mf_load(const uint8_t *p, unsigned (&x)[V]) {
    if (V == 16) {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(p));
        const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int v = 0; v < V; v++) x[v] = (ws[v >> 2] >> (8 * (v & 3))) & 0xffu;
    } else if (V == 4) {
        const unsigned w = __ldg(reinterpret_cast<const unsigned *>(p));
#pragma unroll
        for (int v = 0; v < V; v++) x[v] = (w >> (8 * v)) & 0xffu;
    } else {
        x[0] = __ldg(p);
    }
}

struct MfnrChunks {  // frames retained on the device, chunk by chunk
    const uint8_t *const *ptr;
    const int *count;
    int n;
};

// per element: max, sum (u16 wrap), sum of squares (u32 wrap) over T frames, continuing from the stored values
template <int V>
__global__ void __launch_bounds__(MF_THREADS)
mfnr_accum_kernel(const uint8_t *__restrict__ frames, int T, size_t E, uint8_t *mx, uint16_t *sum, uint32_t *sq, int first) {
    const size_t i0 = (blockIdx.x * (size_t)MF_THREADS + threadIdx.x) * V;
    if (i0 >= E) return;
    unsigned m[V], s[V], q[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
        m[v] = first ? 0u : mx[i0 + v];
        s[v] = first ? 0u : sum[i0 + v];
        q[v] = first ? 0u : sq[i0 + v];
    }
#pragma unroll 4
    for (int t = 0; t < T; t++) {
        unsigned x[V];
        mf_load<V>(frames + (size_t)t * E + i0, x);
#pragma unroll
        for (int v = 0; v < V; v++) {
            m[v] = max(m[v], x[v]);
            s[v] += x[v];
            q[v] += x[v] * x[v];
        }
    }
#pragma unroll
    for (int v = 0; v < V; v++) {
        mx[i0 + v] = (uint8_t)m[v];
        sum[i0 + v] = (uint16_t)s[v];  // uint16 wrap-around, as the reference's numpy adds
        sq[i0 + v] = q[v];
    }
}

// FastGaussianParam.mu / .var (utils.py:454-465) for sums s (uint16), q (uint32) and a count n
__device__ __forceinline__ double mf_mu(unsigned s, int n) { return rint((double)s / (double)n); }
__device__ __forceinline__ double mf_var(unsigned s, unsigned q, int n) {
    const unsigned s2 = s * s;  // np.square on the uint32 copy of sum_mu
    return ((double)q - (double)s2 / (double)n) / (double)(n - 1);
}
__device__ __forceinline__ uint8_t mf_u8(double v) {  // np.round(v).clip(0, 255).astype(uint8)
    const double r = fmin(fmax(rint(v), 0.0), 255.0);
    return (uint8_t)r;
}

// single_sigma_clipping (stacker.py:94-115): subtract the clipped frames' contributions; n becomes per-element.
// V = 4: four elements per thread, one 32-bit load per frame (E and every chunk base are multiples of 4 / 256-byte aligned)
template <int V>
__global__ void __launch_bounds__(MF_THREADS)
mfnr_sigma_kernel(MfnrChunks ch, size_t E, int N, double sigma_high, double sigma_low, const uint16_t *sum, const uint32_t *sq,
                  uint16_t *sum_out, uint32_t *sq_out, int32_t *n_out) {
    const size_t i0 = (blockIdx.x * (size_t)MF_THREADS + threadIdx.x) * V;
    if (i0 >= E) return;
    const int n16 = (int)(int16_t)N;  // n is an int16 array in the reference (utils.py:450-451)
    unsigned s[V], q[V], hi[V], lo[V], cs[V], cq[V], cn[V];
#pragma unroll
    for (int v = 0; v < V; v++) {
        s[v] = sum[i0 + v]; q[v] = sq[i0 + v];
        const double mu = mf_mu(s[v], n16), sd = sqrt(mf_var(s[v], q[v], n16));
        hi[v] = mf_u8(mu + sigma_high * sd);
        lo[v] = mf_u8(mu - sigma_low * sd);
        cs[v] = cq[v] = cn[v] = 0;
    }
    for (int c = 0; c < ch.n; c++) {
        const uint8_t *p = ch.ptr[c] + i0;
        const int cnt = ch.count[c];
#pragma unroll 4
        for (int t = 0; t < cnt; t++, p += E) {
            unsigned x[V];
            mf_load<V>(p, x);
#pragma unroll
            for (int v = 0; v < V; v++)
                if (x[v] > hi[v] || x[v] < lo[v]) { cs[v] += x[v]; cq[v] += x[v] * x[v]; cn[v] += 1; }
        }
    }
#pragma unroll
    for (int v = 0; v < V; v++) {
        sum_out[i0 + v] = (uint16_t)(s[v] - (uint16_t)cs[v]);
        sq_out[i0 + v] = q[v] - cq[v];
        n_out[i0 + v] = n16 - (int)(uint16_t)cn[v];  // int16 - uint16 -> int32 in numpy
    }
}

// ---- median backgrounds (stacker.py:343-349, median_of_medians :62-78) -------------------------------------------------
// np.median over frames f0 .. f1-1 of one element, V elements per thread: the k-th smallest by bisection on the value
// (8 counting passes over the frames, no per-element storage), the upper middle element of an even count by one more pass.
#define MF_MAX_BLOCKS 192  // med-of-med: block_size = int(sqrt(N)) -> at most 182 blocks for N <= 32767

template <int V>
__device__ __forceinline__ void mf_median_range(const uint8_t *const *fptr, size_t i0, int f0, int f1, float (&med)[V]) {
    const int m = f1 - f0, k = (m - 1) >> 1;  // lower middle (0-based)
    unsigned lo[V], hi[V];
#pragma unroll
    for (int v = 0; v < V; v++) { lo[v] = 0; hi[v] = 255; }
    for (int it = 0; it < 8; it++) {
        unsigned mid[V], c[V];
#pragma unroll
        for (int v = 0; v < V; v++) { mid[v] = (lo[v] + hi[v]) >> 1; c[v] = 0; }
#pragma unroll 4
        for (int f = f0; f < f1; f++) {
            unsigned x[V];
            mf_load<V>(fptr[f] + i0, x);
#pragma unroll
            for (int v = 0; v < V; v++) c[v] += x[v] <= mid[v];
        }
#pragma unroll
        for (int v = 0; v < V; v++) {
            if (lo[v] < hi[v]) {
                if (c[v] >= (unsigned)(k + 1)) hi[v] = mid[v]; else lo[v] = mid[v] + 1;
            }
        }
    }
    if (m & 1) {
#pragma unroll
        for (int v = 0; v < V; v++) med[v] = (float)lo[v];
        return;
    }
    unsigned c[V], nx[V];
#pragma unroll
    for (int v = 0; v < V; v++) { c[v] = 0; nx[v] = 256; }
    for (int f = f0; f < f1; f++) {
        unsigned x[V];
        mf_load<V>(fptr[f] + i0, x);
#pragma unroll
        for (int v = 0; v < V; v++) {
            c[v] += x[v] <= lo[v];
            if (x[v] > lo[v]) nx[v] = min(nx[v], x[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < V; v++) {
        const unsigned b = c[v] >= (unsigned)(k + 2) ? lo[v] : nx[v];  // the (k+1)-th smallest
        med[v] = (float)(lo[v] + b) * 0.5f;
    }
}

// block_size == 0: np.median over all N frames; else median_of_medians with that block size.  mu_out: float plane
// (values are multiples of 0.25: exact)
template <int V>
__global__ void __launch_bounds__(MF_THREADS)
mfnr_median_kernel(const uint8_t *const *fptr, size_t E, int N, int block_size, float *mu_out) {
    const size_t i0 = (blockIdx.x * (size_t)MF_THREADS + threadIdx.x) * V;
    if (i0 >= E) return;
    float res[V];
    if (block_size <= 0) {
        mf_median_range<V>(fptr, i0, 0, N, res);
    } else {
        const int nb = (N - 1) / block_size + 1;
        float meds[V][MF_MAX_BLOCKS];
        for (int b = 0; b < nb; b++) {
            float mb[V];
            mf_median_range<V>(fptr, i0, b * block_size, min((b + 1) * block_size, N), mb);
#pragma unroll
            for (int v = 0; v < V; v++) {  // insertion into the sorted prefix
                int j = b;
                while (j > 0 && meds[v][j - 1] > mb[v]) { meds[v][j] = meds[v][j - 1]; j--; }
                meds[v][j] = mb[v];
            }
        }
#pragma unroll
        for (int v = 0; v < V; v++)
            res[v] = (nb & 1) ? meds[v][nb >> 1] : (meds[v][(nb >> 1) - 1] + meds[v][nb >> 1]) * 0.5f;
    }
#pragma unroll
    for (int v = 0; v < V; v++) mu_out[i0 + v] = res[v];
}

// block-level deterministic sum of (value, count): partial[blockIdx.x]
__device__ __forceinline__ void mf_block_reduce(double v, unsigned long long c, double *part_v, unsigned long long *part_c) {
    __shared__ double sv[MF_THREADS / 32];
    __shared__ unsigned long long sc[MF_THREADS / 32];
    for (int o = 16; o; o >>= 1) {
        v += __shfl_down_sync(0xffffffffu, v, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sv[w] = v; sc[w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < MF_THREADS / 32; k++) { v += sv[k]; c += sc[k]; }
        part_v[blockIdx.x] = v;
        part_c[blockIdx.x] = c;
    }
}

// sum over elements of sqrt(var): est_bg_var = mean(sqrt(var)) (stacker.py:338, :342).  n_arr == nullptr: n = N everywhere
__global__ void __launch_bounds__(MF_THREADS)
mfnr_sqrtvar_kernel(size_t E, int N, const uint16_t *sum, const uint32_t *sq, const int32_t *n_arr, double *part_v,
                    unsigned long long *part_c) {
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)MF_THREADS + threadIdx.x; i < E; i += (size_t)gridDim.x * MF_THREADS) {
        const int n = n_arr ? n_arr[i] : (int)(int16_t)N;
        acc += sqrt(mf_var(sum[i], sq[i], n));
    }
    mf_block_reduce(acc, 0ull, part_v, part_c);
}

// max_bias_diff = max - (mu + c1) (stacker.py:351-354); sum and count of its positive entries (:356-357)
__global__ void __launch_bounds__(MF_THREADS)
mfnr_diffpos_kernel(size_t E, int N, double c1, const uint8_t *mx, const uint16_t *sum, const int32_t *n_arr,
                    const float *mu_arr, double *part_v, unsigned long long *part_c) {
    double acc = 0.0;
    unsigned long long cnt = 0;
    for (size_t i = blockIdx.x * (size_t)MF_THREADS + threadIdx.x; i < E; i += (size_t)gridDim.x * MF_THREADS) {
        const int n = n_arr ? n_arr[i] : (int)(int16_t)N;
        const double d = (double)mx[i] - ((mu_arr ? (double)mu_arr[i] : mf_mu(sum[i], n)) + c1);
        if (d > 0.0) { acc += d; cnt += 1; }
    }
    mf_block_reduce(acc, cnt, part_v, part_c);
}

// one warp, fixed order: lane l sums partials l, l+32, ...; then a shuffle tree
__global__ void mfnr_final_reduce_kernel(int parts, const double *part_v, const unsigned long long *part_c, double *out_v,
                                         unsigned long long *out_c) {
    if (blockIdx.x || threadIdx.x >= 32) return;
    double v = 0.0;
    unsigned long long c = 0;
    for (int k = threadIdx.x; k < parts; k += 32) { v += part_v[k]; c += part_c[k]; }
    for (int o = 16; o; o >>= 1) {
        v += __shfl_down_sync(0xffffffffu, v, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if (threadIdx.x == 0) { *out_v = v; *out_c = c; }
}

// fg_mask (stacker.py:358-365): a pixel is foreground when any of its channels is an outlier or a highlight
__global__ void __launch_bounds__(MF_THREADS)
mfnr_mask_kernel(size_t P, int C, int N, double c1, double avg, double hl, const uint8_t *mx, const uint16_t *sum,
                 const int32_t *n_arr, const float *mu_arr, uint8_t *fg) {
    const size_t p = blockIdx.x * (size_t)MF_THREADS + threadIdx.x;
    if (p >= P) return;
    int on = 0;
    for (int c = 0; c < C; c++) {
        const size_t i = p * C + c;
        const int n = n_arr ? n_arr[i] : (int)(int16_t)N;
        const double d = (double)mx[i] - ((mu_arr ? (double)mu_arr[i] : mf_mu(sum[i], n)) + c1);
        on |= (d > avg) | ((double)mx[i] > hl);
    }
    fg[p] = (uint8_t)on;
}

__device__ __forceinline__ int mf_reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// cv2.GaussianBlur(float64): row pass, taps accumulated in order (RowFilter<double, double>)
__global__ void __launch_bounds__(MF_THREADS)
mfnr_blur_row_kernel(int H, int W, int ksize, const double *__restrict__ k, const uint8_t *__restrict__ fg, double *row) {
    const size_t p = blockIdx.x * (size_t)MF_THREADS + threadIdx.x;
    if (p >= (size_t)H * W) return;
    const int y = (int)(p / W), x = (int)(p % W), r = ksize / 2;
    const uint8_t *src = fg + (size_t)y * W;
    double s;
    if (x >= r && x + r < W) {  // interior: no border arithmetic (same taps, same order)
        const uint8_t *q = src + x - r;
        s = k[0] * (double)q[0];
        for (int j = 1; j < ksize; j++) s += k[j] * (double)q[j];
    } else {
        s = k[0] * (double)src[mf_reflect101(x - r, W)];
        for (int j = 1; j < ksize; j++) s += k[j] * (double)src[mf_reflect101(x - r + j, W)];
    }
    row[p] = s;
}

// column pass, symmetric form (SymmColumnFilter): centre tap, then pairs
__global__ void __launch_bounds__(MF_THREADS)
mfnr_blur_col_kernel(int H, int W, int ksize, const double *__restrict__ k, const double *__restrict__ row, double *out) {
    const size_t p = blockIdx.x * (size_t)MF_THREADS + threadIdx.x;
    if (p >= (size_t)H * W) return;
    const int y = (int)(p / W), x = (int)(p % W), r = ksize / 2;
    double s = k[r] * row[p];
    if (y >= r && y + r < H) {
        for (int j = 1; j <= r; j++) s += k[r + j] * (row[p + (size_t)j * W] + row[p - (size_t)j * W]);
    } else {
        for (int j = 1; j <= r; j++)
            s += k[r + j] * (row[(size_t)mf_reflect101(y + j, H) * W + x] + row[(size_t)mf_reflect101(y - j, H) * W + x]);
    }
    out[p] = s;
}

// highlight fix + mix (stacker.py:383-397)
__global__ void __launch_bounds__(MF_THREADS)
mfnr_mix_kernel(size_t E, int C, int N, double c2, double hp, double one_minus_hp, const uint8_t *mx, const uint16_t *sum,
                const int32_t *n_arr, const float *mu_arr, const double *__restrict__ blur, uint8_t *out) {
    const size_t i = blockIdx.x * (size_t)MF_THREADS + threadIdx.x;
    if (i >= E) return;
    const int n = n_arr ? n_arr[i] : (int)(int16_t)N;
    const double m = (double)mx[i], mu = mu_arr ? (double)mu_arr[i] : mf_mu(sum[i], n), b = blur[i / C];
    const double hff = 1.0 - (fmin(fmax(m / 255.0 - hp, 0.0), 1.0) / one_minus_hp);
    double fixed = m - (c2 * hff);
    fixed = fmin(fmax(fixed, 0.0), 255.0);
    const double v = rint(fixed * b + mu * (1.0 - b));
    out[i] = (uint8_t)(long long)v;
}
