// temporal2_kernel: stack -> diff -> threshold fused (SlidingWindow.max / .mean, utils.py:269-307;
// M3Detector.detect, Detector.py:327-332), one predicate bit per pixel, every frame read once.
//
// Same algorithm as the first-generation kernel it replaced in round 1 -- per-thread shared-memory ring fed by
// cp.async, van Herk / Gil-Werman sliding max, running window sums, integer predicate
// max*L - sum > thr*L -- restructured for the instruction issue limit, which is what bounds this
// pass on sm_100a (the ALU pipe issues one warp instruction every two cycles per SM sub-partition):
//
//   * "high-byte form": VIMNMX.U16x2 compares 16-bit lanes, and the high byte of a lane decides, so
//     the max chain of the ODD pixels runs on the packed u8x4 words as they are, and the EVEN pixels'
//     chain on  x << 8  (one IMAD on the FMA pipe).  Low bytes carry garbage that never reaches a
//     high byte.  No unpacking of the new frame or of the suffix-max words; one PRMT per lane pair
//     extracts the clean window max at the end.
//   * window sums: W = sum of the raw packed words (mod 2^32) and O = sum of the odd pixels (u16x2);
//     the even sums are W - 256*O (one IMAD).  Only the odd lanes of new / evicted words are unpacked.
//   * sub-blocked van Herk: with the window split into k blocks of b = n/k frames (k divides n) a window is the
//     suffix of its oldest block + (k-1) whole blocks + the prefix of the current block.  Only the OLDEST block
//     needs a suffix-max array (b slots instead of n; it is rebuilt from the raw ring at every block end), the
//     whole blocks' maxima sit in registers, and VIMNMX3 combines prefix, middle maximum and suffix in the one
//     instruction the two-block scheme needs as well.  Shared memory per pixel: n + K + n/k bytes instead of
//     2n + K, i.e. half again as many resident warps at n = 30.
//   * the frame loop is cut into runs inside which no ring pointer wraps, L is constant and the
//     prefetch predicate does not change, so a frame costs no pointer selects or compares; shared
//     memory slots are addressed with immediate offsets (CTA size is a template parameter).
#pragma once
#include "common.cuh"

#define T2_K 8  // frames in flight per thread (cp.async groups)

__device__ __forceinline__ void t2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void t2_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ unsigned t2_prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// bytes 1 and 3 of x as clean u16x2 lanes
__device__ __forceinline__ unsigned t2_hi(unsigned x) { return t2_prmt(x, 0u, 0x4341u); }

template <int WPT>
__device__ __forceinline__ void t2_cp(uint32_t saddr, const void *g) {
    if (WPT == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(g) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(g) : "memory");
}
template <int WPT>
__device__ __forceinline__ void t2_lds(unsigned (&w)[WPT], uint32_t saddr) {
    if (WPT == 4) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2 % WPT]), "=r"(w[3 % WPT]) : "r"(saddr));
    else asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(saddr));
}
template <int WPT>
__device__ __forceinline__ void t2_sts(uint32_t saddr, const unsigned (&w)[WPT]) {
    if (WPT == 4) asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(w[0]), "r"(w[1]), "r"(w[2 % WPT]), "r"(w[3 % WPT]) : "memory");
    else asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(saddr), "r"(w[0]), "r"(w[1]) : "memory");
}

// per-thread running state of the window
template <int WPT>
struct T2State {
    unsigned pO[WPT], pE[WPT];  // prefix max of the current block, high-byte form (odd / even pixels)
    unsigned Wd[WPT];           // sum of the raw words of the window, mod 2^32
    unsigned Od[WPT];           // sum of the odd pixels of the window, u16x2
    unsigned mO[WPT], mE[WPT];  // max over the whole blocks inside the window (high-byte form; 0 when k == 1)
};

#define T2_KMAX 6  // most sub-blocks per window
// maxima of the last T2_KMAX-1 finished blocks, oldest first (high-byte form)
template <int WPT>
struct T2Fifo {
    unsigned fO[T2_KMAX - 1][WPT], fE[T2_KMAX - 1][WPT];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int q = 0; q < T2_KMAX - 1; q++)
#pragma unroll
            for (int k = 0; k < WPT; k++) fO[q][k] = fE[q][k] = 0u;
    }
    __device__ __forceinline__ void push(const unsigned (&bO)[WPT], const unsigned (&bE)[WPT]) {
#pragma unroll
        for (int q = 0; q + 1 < T2_KMAX - 1; q++)
#pragma unroll
            for (int k = 0; k < WPT; k++) { fO[q][k] = fO[q + 1][k]; fE[q][k] = fE[q + 1][k]; }
#pragma unroll
        for (int k = 0; k < WPT; k++) { fO[T2_KMAX - 2][k] = bO[k]; fE[T2_KMAX - 2][k] = bE[k]; }
    }
    // max over the newest `cnt` entries (cnt = k-1 whole blocks inside the window)
    __device__ __forceinline__ void newest_max(int cnt, unsigned (&mO)[WPT], unsigned (&mE)[WPT]) const {
#pragma unroll
        for (int k = 0; k < WPT; k++) mO[k] = mE[k] = 0u;
#pragma unroll
        for (int q = 0; q < T2_KMAX - 1; q++) {
            const bool use = q >= T2_KMAX - 1 - cnt;
#pragma unroll
            for (int k = 0; k < WPT; k++) {
                mO[k] = __vmaxu2(mO[k], use ? fO[q][k] : 0u);
                mE[k] = __vmaxu2(mE[k], use ? fE[q][k] : 0u);
            }
        }
    }
};

// Operands of one frame: its raw word(s), frame t-n's, the suffix-max word(s) of its position.
template <int WPT>
struct T2Ops {
    unsigned xw[WPT], ow[WPT], mw[WPT];
};

// Load stage of one frame.  ac / ao / asx: shared addresses of the frame's slot, of frame t-n's slot
// (which then receives the prefetch of frame t+K), of the suffix-max word for this position.
template <bool MASKED, int WPT>
__device__ __forceinline__ void t2_load(T2Ops<WPT> &q, const unsigned (&mk)[WPT], uint32_t ac, uint32_t ao,
                                        uint32_t asx) {
    t2_lds<WPT>(q.xw, ac);
    t2_lds<WPT>(q.ow, ao);
    t2_lds<WPT>(q.mw, asx);
    if (MASKED) {
#pragma unroll
        for (int k = 0; k < WPT; k++) q.xw[k] &= mk[k];
        t2_sts<WPT>(ac, q.xw);
    }
}
// base + idx * stride as one IMAD.WIDE on the FMA pipe (no 64-bit pointer increments on the ALU pipe)
__device__ __forceinline__ const uint8_t *t2_addr(const uint8_t *base, unsigned idx, unsigned stride) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(idx), "r"(stride), "l"((unsigned long long)base));
    return reinterpret_cast<const uint8_t *>(r);
}

// Compute stage.  Lu = window length, cpk = per-lane bias (0x7fff - thr*L) * 0x10001.
// SUB: the window holds whole blocks between its oldest block and the current one (sub-blocked scheme).
template <int WPT, bool SUB>
__device__ __forceinline__ void t2_compute(T2State<WPT> &st, const T2Ops<WPT> &q, unsigned Lu, unsigned cpk,
                                           uint8_t *bp) {
    unsigned M[WPT];
#pragma unroll
    for (int k = 0; k < WPT; k++) {
        const unsigned x = q.xw[k], x8 = x << 8, m = q.mw[k], m8 = m << 8;
        st.pO[k] = __vmaxu2(st.pO[k], x);
        st.pE[k] = __vmaxu2(st.pE[k], x8);
        // window max = prefix of this block, whole blocks in between, suffix of the oldest block (one VIMNMX3)
        const unsigned wO = t2_hi(SUB ? __vimax3_u16x2(st.pO[k], m, st.mO[k]) : __vmaxu2(st.pO[k], m));    // odd pixels, clean u16x2
        const unsigned wE = t2_hi(SUB ? __vimax3_u16x2(st.pE[k], m8, st.mE[k]) : __vmaxu2(st.pE[k], m8));  // ... even pixels
        st.Wd[k] = st.Wd[k] + x - q.ow[k];
        st.Od[k] = st.Od[k] + t2_hi(x) - t2_hi(q.ow[k]);
        const unsigned sE = st.Wd[k] - (st.Od[k] << 8);
        // per lane: max*L - sum + 0x7fff - thr*L ; bit 15 set <=> max*L - sum > thr*L
        const unsigned vO = wO * Lu + cpk - st.Od[k];
        const unsigned vE = wE * Lu + cpk - sE;
        M[k] = t2_prmt(vE, vO, 0xFBD9u);  // sign-replicate bytes 1,5,3,7 -> 0x00/0xff per pixel
    }
    const unsigned q01 = (M[0] & 0x08040201u) | (M[1] & 0x80402010u);
    const unsigned r01 = q01 * 0x01010101u;
    if (WPT == 4) {
        const unsigned q23 = (M[2 % WPT] & 0x08040201u) | (M[3 % WPT] & 0x80402010u);
        const unsigned r23 = q23 * 0x01010101u;
        *reinterpret_cast<uint16_t *>(bp) = (uint16_t)t2_prmt(r01, r23, 0x4473u);
    } else {
        *bp = (uint8_t)(r01 >> 24);
    }
}

// Backward scan over `cnt` raw frames ending at slot a (descending slots, no wrap inside): suffix max
// by position, stored packed at so, so - S, ...
template <int WPT, uint32_t S>
__device__ __forceinline__ void t2_scan_run(unsigned (&aO)[WPT], unsigned (&aE)[WPT], uint32_t a, uint32_t so,
                                            int cnt) {
#pragma unroll 2
    for (int u = 0; u < cnt; u++) {
        unsigned w[WPT], o[WPT];
        t2_lds<WPT>(w, a);
#pragma unroll
        for (int k = 0; k < WPT; k++) {
            aO[k] = __vmaxu2(aO[k], w[k]);
            aE[k] = __vmaxu2(aE[k], w[k] << 8);
            o[k] = t2_prmt(aE[k], aO[k], 0x7351u);  // high bytes: E0 O0 E1 O1 = pixel order
        }
        t2_sts<WPT>(so, o);
        a -= S;
        so -= S;
    }
}

template <bool MASKED, int WPT, int NT, bool SUB>
__global__ void __launch_bounds__(NT)
temporal2_kernel(FrameSrc src, long long t0, int T, int n, int kdiv, int HWG, const int *__restrict__ thr,
                 uint8_t *__restrict__ bits) {
    constexpr int VB = WPT * 4;           // bytes (= pixels) per thread per frame
    constexpr uint32_t S = NT * VB;       // bytes between consecutive slots
    extern __shared__ uint4 t_smem[];
    const int tid = threadIdx.x;
    const int R = n + T2_K;
    if (!SUB) kdiv = 1;
    const int b = n / kdiv;  // block length (kdiv divides n; kdiv == 1: the classic two-block scheme)
    // shared memory: ring [R][NT] raw frames, smx [b][NT] suffix max of the window's oldest block by position
    // (slot p-1 = position p; slot b-1 stays zero), thr_s [T]
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(t_smem);
    const uint32_t ring_s = smem0 + tid * VB;
    const uint32_t smx_s = ring_s + R * S;
    // per-frame lane bias 0x7fff - thr*L (L = SlidingWindow.length of that frame), u16
    uint16_t *thr_s = reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(t_smem) + (size_t)(R + b) * S);
    for (int i = tid; i < T; i += NT) {
        const long long Li = t0 + i + 1 < n ? t0 + i + 1 : n;
        thr_s[i] = (uint16_t)(0x7fff - min(max(thr[i], 0), 255) * (int)Li);
    }
    __syncthreads();
    const int g = blockIdx.x * NT + tid;
    if (g >= HWG) return;
    const uint8_t *gbase = (src.cur ? src.cur : src.ring) + (size_t)g * VB;
    const int Rw = src.cur ? 0x7fffffff : src.R;  // the caller's buffer does not wrap
    const unsigned HWu = (unsigned)src.HW;

    unsigned mk[WPT], zero[WPT];
#pragma unroll
    for (int k = 0; k < WPT; k++) { mk[k] = ~0u; zero[k] = 0u; }
    if (MASKED) {
        const unsigned *m = reinterpret_cast<const unsigned *>(src.mask + (size_t)g * VB);
#pragma unroll
        for (int k = 0; k < WPT; k++) mk[k] = m[k] * 0xffu;  // {0,1} -> {0x00,0xff}
    }

    // ---- history: frames t0-n+1 .. t0-1 -> slots 0 .. n-2 ; slot R-1 = zeros ("frame t0-n") ----
    t2_sts<WPT>(ring_s + (R - 1) * S, zero);
    for (int p = 1; p < n; p++) {
        const long long th = t0 - n + p;
        if (th >= 0) t2_cp<WPT>(ring_s + (p - 1) * S, src.frame(th) + (size_t)g * VB);
        else t2_sts<WPT>(ring_s + (p - 1) * S, zero);
    }
    t2_commit();
    // ---- prime the pipeline: frames 0 .. K-1 of the batch -> slots n-1 .. n+K-2 -------------------
    int pf_slot = src.cur ? (int)(t0 - src.t0) : (int)(t0 % src.R);  // source slot of the next frame to fetch
    for (int i = 0; i < T2_K; i++) {
        if (i < T) t2_cp<WPT>(ring_s + (n - 1 + i) * S, gbase + (size_t)pf_slot * HWu);
        t2_commit();
        if (++pf_slot == Rw) pf_slot = 0;
    }
    t2_wait<T2_K>();  // history landed

    T2State<WPT> st;
    T2Fifo<WPT> fifo;
    if (SUB) fifo.clear();
    {
        // history position p (1 .. n-1) = frame t0-n+p lives in ring slot p-1; block q = p / b (0 = oldest, only its
        // positions 1 .. b-1 exist), position inside the block p % b
#pragma unroll
        for (int k = 0; k < WPT; k++) st.Wd[k] = st.Od[k] = 0u;
        for (int q = 1; SUB && q < kdiv; q++) {  // whole blocks between the oldest one and the batch: their maxima
            unsigned bO[WPT], bE[WPT];
#pragma unroll
            for (int k = 0; k < WPT; k++) bO[k] = bE[k] = 0u;
            for (int p = q * b; p < (q + 1) * b; p++) {
                unsigned w[WPT];
                t2_lds<WPT>(w, ring_s + (p - 1) * S);
                if (MASKED) {
#pragma unroll
                    for (int k = 0; k < WPT; k++) w[k] &= mk[k];
                    t2_sts<WPT>(ring_s + (p - 1) * S, w);
                }
#pragma unroll
                for (int k = 0; k < WPT; k++) {
                    st.Wd[k] += w[k];
                    st.Od[k] += t2_hi(w[k]);
                    bO[k] = __vmaxu2(bO[k], w[k]);
                    bE[k] = __vmaxu2(bE[k], w[k] << 8);
                }
            }
            fifo.push(bO, bE);
        }
        unsigned aO[WPT], aE[WPT];
#pragma unroll
        for (int k = 0; k < WPT; k++) aO[k] = aE[k] = 0u;
        for (int p = b - 1; p >= 1; p--) {  // oldest block: suffix max by position
            unsigned w[WPT], o[WPT];
            t2_lds<WPT>(w, ring_s + (p - 1) * S);
            if (MASKED) {
#pragma unroll
                for (int k = 0; k < WPT; k++) w[k] &= mk[k];
                t2_sts<WPT>(ring_s + (p - 1) * S, w);
            }
#pragma unroll
            for (int k = 0; k < WPT; k++) {
                st.Wd[k] += w[k];
                st.Od[k] += t2_hi(w[k]);
                aO[k] = __vmaxu2(aO[k], w[k]);
                aE[k] = __vmaxu2(aE[k], w[k] << 8);
                o[k] = t2_prmt(aE[k], aO[k], 0x7351u);
            }
            t2_sts<WPT>(smx_s + (p - 1) * S, o);
        }
        if (SUB) fifo.newest_max(kdiv - 1, st.mO, st.mE);
    }
    t2_sts<WPT>(smx_s + (b - 1) * S, zero);

    const size_t bstride = (size_t)HWG * WPT / 2;          // WPT*4 bits per thread and frame
    uint8_t *bout = bits + (size_t)g * WPT / 2;
    int c = n - 1;   // ring slot of the current frame
    int o = R - 1;   // ring slot of frame t-n (then: destination of the prefetch of frame t+K)
    int i = 0;
    while (i < T) {
        const int nb = min(b, T - i);  // frames of this block
#pragma unroll
        for (int k = 0; k < WPT; k++) st.pO[k] = st.pE[k] = 0u;  // 0 = identity of max
        int j = 0;
        while (j < nb) {
            // a run: no ring pointer wraps, constant L, constant prefetch predicate
            const long long tg = t0 + i;
            const bool warm = tg + 1 < n;
            const unsigned Lu = (unsigned)(warm ? tg + 1 : n);  // SlidingWindow.length
            const bool pf = i + T2_K < T;
            int run = min(min(nb - j, R - c), min(R - o, Rw - pf_slot));
            if (pf) run = min(run, T - T2_K - i);
            if (warm) run = 1;
            uint32_t ac = ring_s + c * S, ao = ring_s + o * S, asx = smx_s + j * S;
            const uint8_t *gp = gbase + (size_t)pf_slot * HWu;
            uint8_t *bp = bout + (size_t)i * bstride;
            const uint16_t *tp = thr_s + i;
            const uint8_t *gp1 = gp + HWu;
            uint8_t *bp1 = bp + bstride;
            const unsigned bs32 = (unsigned)bstride;
            // frames are taken two at a time: both frames' shared-memory loads first, then both
            // prefetches back to back (ptxas pads every LDS -> LDGSTS transition with three dummy LDS)
            T2Ops<WPT> qa, qb;
            unsigned u = 0;
            for (; u + 2 <= (unsigned)run; u += 2) {
                t2_wait<T2_K - 2>();  // frames u and u+1 have landed
                t2_load<MASKED, WPT>(qa, mk, ac, ao, asx);
                t2_load<MASKED, WPT>(qb, mk, ac + S, ao + S, asx + S);
                if (pf) t2_cp<WPT>(ao, t2_addr(gp, u, HWu));  // slot of frame t-n is free: fetch frame t+K
                t2_commit();
                if (pf) t2_cp<WPT>(ao + S, t2_addr(gp1, u, HWu));
                t2_commit();
                t2_compute<WPT, SUB>(st, qa, Lu, (unsigned)tp[u] * 0x00010001u, (uint8_t *)t2_addr(bp, u, bs32));
                t2_compute<WPT, SUB>(st, qb, Lu, (unsigned)tp[u + 1] * 0x00010001u, (uint8_t *)t2_addr(bp1, u, bs32));
                ac += 2 * S; ao += 2 * S; asx += 2 * S;
            }
            if (u < (unsigned)run) {  // odd run: one frame left
                t2_wait<T2_K - 1>();
                t2_load<MASKED, WPT>(qa, mk, ac, ao, asx);
                if (pf) t2_cp<WPT>(ao, t2_addr(gp, u, HWu));
                t2_commit();
                t2_compute<WPT, SUB>(st, qa, Lu, (unsigned)tp[u] * 0x00010001u, (uint8_t *)t2_addr(bp, u, bs32));
            }
            i += run; j += run;
            c += run; if (c == R) c = 0;
            o += run; if (o == R) o = 0;
            pf_slot += run; if (pf_slot >= Rw) pf_slot = 0;
        }
        if (nb == b && i < T) {
            // block complete and more frames follow: it joins the whole blocks of the window (maximum = its final
            // prefix), and the window's new oldest block -- the b oldest frames still in the raw ring, positions
            // 0 .. b-1 at slots c+K .. c+K+b-1 (mod R) -- gets its suffix max by position 1 .. b-1
            if (SUB) {
                fifo.push(st.pO, st.pE);
                fifo.newest_max(kdiv - 1, st.mO, st.mE);
            }
            unsigned aO[WPT], aE[WPT];
#pragma unroll
            for (int k = 0; k < WPT; k++) aO[k] = aE[k] = 0u;
            int a = c + T2_K + b - 1;  // slot of the oldest block's last frame
            if (a >= R) a -= R;
            int p = b - 1;
            while (p >= 1) {
                const int cnt = min(p, a + 1);
                t2_scan_run<WPT, S>(aO, aE, ring_s + a * S, smx_s + (p - 1) * S, cnt);
                p -= cnt;
                a -= cnt;
                if (a < 0) a = R - 1;
            }
        }
    }
    t2_wait<0>();
}
