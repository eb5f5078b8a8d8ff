// temporal3_kernel: stack -> diff -> threshold fused (SlidingWindow.max / .mean, MetLib/utils.py:269-307;
// M3Detector.detect, MetLib/Detector.py:327-332), one predicate bit per pixel, every frame read once.
//
// Third generation of the temporal pass.  The first two (the second lives on in temporal_kernel.cuh) keep each
// thread's last n frames in a shared-memory ring fed by cp.async and are bound by instruction issue (10.3
// thread-instructions per pixel-frame, ALU pipe 65 %, 12 warps/SM because of 68 B of shared memory per pixel).
// Here the ring lives in REGISTERS:
//
//   * a thread owns 8 consecutive pixels (two packed u8x4 words) for the whole batch and keeps its last U raw
//     frames in 2*U registers.  The frame loop is unrolled U times, so every ring access is a register name:
//     no shared-memory ring traffic, no LDGSTS / DEPBAR / address arithmetic, no pointer wraps.  Windows of
//     n = 2*U frames keep the younger half of the ring in a shared-memory page (one LDS + one STS per frame).
//   * frames are fetched with plain 8-byte global loads K frames ahead into K*2 registers (a warp reads 256
//     contiguous bytes per frame).  Batch frames are always contiguous in memory (`src.cur`), so a frame's
//     address is  base + immediate * HW  (one IMAD.WIDE).
//   * sliding max = sub-blocked van Herk / Gil-Werman on CLEAN u16x2 lanes (even pixels A = x & 0x00ff00ff,
//     odd pixels B = bytes 1,3): window = prefix of the current block (registers) + whole blocks in between
//     (their maxima in a small shared-memory FIFO, folded once per block) + suffix of the oldest block (BL-1
//     shared-memory slots, rebuilt from the ring registers at every block end); one VIMNMX3 combines them.
//     Clean lanes cost one PRMT + one IMAD per word but save the two clean-up PRMTs and the << 8 shifts of the
//     "high-byte form", and they are shared with the window sums.
//   * window sums by eviction: W = sum of raw words (mod 2^32), SB = sum of odd pixels (u16x2), even sums
//     = W - 256*SB.  The evicted frame is a register.
//   * predicate  max*L - sum > thr*L  per u16 lane, as before:  v = max*L + (0x7fff - thr*L) - sum, bit 15.
//
// Per 8 pixels and frame: ~27 ALU-pipe + ~17 FMA-pipe + ~9 load/store instructions (second generation: 36.5 / 18 / 21).
#pragma once
#include "common.cuh"

#if defined(__CUDA_ARCH__) || !defined(T3_HOST_EMU)
#define T3_HD __device__ __forceinline__
#else
#define T3_HD inline
#endif

#define T3_NT 128  // threads per CTA; a thread owns 8 pixels

namespace t3 {

#if defined(__CUDA_ARCH__)
T3_HD unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
T3_HD unsigned vmax2(unsigned a, unsigned b) { return __vmaxu2(a, b); }
T3_HD unsigned vmax3(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
T3_HD uint2 ldg8(const uint8_t *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
// the frame prefetch: loads 8 bytes at p if idx < rem, otherwise leaves the outputs undefined (the caller never uses
// them).  Predicated inside the asm so that the compiler does not preserve the registers' old contents with two
// moves per frame; volatile so that ptxas keeps the load where it is written, K frames ahead of its use (it sinks
// plain loads towards their consumers in bursts to shorten live ranges).
T3_HD void ldg8_if(unsigned &x, unsigned &y, const uint8_t *p, int idx, int rem) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.lt.s32 p, %3, %4;\n"
        "@p ld.volatile.global.v2.u32 {%0,%1}, [%2];\n"
        "}\n" : "=r"(x), "=r"(y) : "l"(p), "r"(idx), "r"(rem));
}
// base + idx * stride as ONE IMAD.WIDE.U32 on the FMA pipe (idx is a literal after unrolling)
T3_HD const uint8_t *addr(const uint8_t *base, unsigned idx, unsigned stride) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(idx), "r"(stride), "l"((unsigned long long)base));
    return reinterpret_cast<const uint8_t *>(r);
}
// eight 0x00/0xff pixel bytes (two words, pixel order) -> one byte of bits: two IDP.4A with signed weights -2^k
// (0xff = -1 as s8), instead of two LOP3 + multiply + shift on the ALU pipe
T3_HD unsigned bits8(unsigned m0, unsigned m1) {
    return (unsigned)__dp4a((int)m1, (int)0x80C0E0F0, __dp4a((int)m0, (int)0xF8FCFEFF, 0));
}
// the same with per-thread weights (a weight of 0 drops a masked-out pixel)
T3_HD unsigned bits8w(unsigned m0, unsigned m1, unsigned w0, unsigned w1) {
    return (unsigned)__dp4a((int)m1, (int)w1, __dp4a((int)m0, (int)w0, 0));
}
// ---- bulk-copy feed (FEED = 1): cp.async.bulk global -> shared, completion on an mbarrier (UBLKCP + SYNCS in SASS)
T3_HD uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
T3_HD void mbar_init(uint32_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
T3_HD void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
T3_HD void mbar_expect_tx(uint32_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
T3_HD void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
T3_HD void mbar_wait(uint32_t bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "T3_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra T3_WAIT_%=;\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// one thread asks the TMA unit to pull `bytes` (multiple of 16) at src into L2; no destination, no completion
T3_HD void bulk_prefetch_l2(const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
T3_HD void bulk_g2s(uint32_t dst, const void *src, unsigned bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
#else   // host emulation (tests/t3_host_emu.cu): same arithmetic, one thread at a time
T3_HD unsigned prmt(unsigned a, unsigned b, unsigned sel) {
    const unsigned long long v = ((unsigned long long)b << 32) | a;
    unsigned d = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned s = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xff;
        if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        d |= byte << (8 * i);
    }
    return d;
}
T3_HD unsigned vmax2(unsigned a, unsigned b) {
    const unsigned lo = (a & 0xffff) > (b & 0xffff) ? (a & 0xffff) : (b & 0xffff);
    const unsigned hi = (a >> 16) > (b >> 16) ? (a >> 16) : (b >> 16);
    return (hi << 16) | lo;
}
T3_HD unsigned vmax3(unsigned a, unsigned b, unsigned c) { return vmax2(vmax2(a, b), c); }
T3_HD uint2 ldg8(const uint8_t *p) { uint2 v; memcpy(&v, p, 8); return v; }
T3_HD void ldg8_if(unsigned &x, unsigned &y, const uint8_t *p, int idx, int rem) {
    if (idx < rem) { const uint2 v = ldg8(p); x = v.x; y = v.y; }
}
T3_HD const uint8_t *addr(const uint8_t *base, unsigned idx, unsigned stride) { return base + (size_t)idx * stride; }
T3_HD unsigned bits8(unsigned m0, unsigned m1) {
    unsigned r = 0;
    for (int k = 0; k < 4; k++) {
        if ((m0 >> (8 * k)) & 0x80) r |= 1u << k;
        if ((m1 >> (8 * k)) & 0x80) r |= 16u << k;
    }
    return r;
}
T3_HD unsigned bits8w(unsigned m0, unsigned m1, unsigned w0, unsigned w1) {
    unsigned r = 0;
    for (int k = 0; k < 4; k++) {
        if (((m0 >> (8 * k)) & 0x80) && ((w0 >> (8 * k)) & 0xff)) r |= 1u << k;
        if (((m1 >> (8 * k)) & 0x80) && ((w1 >> (8 * k)) & 0xff)) r |= 16u << k;
    }
    return r;
}
// the bulk-copy feed is not emulated: the host build reads the frames straight from memory
T3_HD uint32_t smem_u32(const void *) { return 0; }
T3_HD void mbar_init(uint32_t, unsigned) {}
T3_HD void mbar_init_fence() {}
T3_HD void mbar_expect_tx(uint32_t, unsigned) {}
T3_HD void mbar_arrive(uint32_t) {}
T3_HD void mbar_wait(uint32_t, unsigned) {}
T3_HD void bulk_g2s(uint32_t, const void *, unsigned, uint32_t) {}
T3_HD void bulk_prefetch_l2(const void *, unsigned) {}
#endif

// bytes 1 and 3 of x as clean u16x2 lanes (the odd pixels)
T3_HD unsigned hi2(unsigned x) { return prmt(x, 0u, 0x4341u); }

template <int U, int BL, int P, int FEED = 0, int K = 1>
struct Layout {
    static constexpr int N = U * P;     // window length
    static constexpr int KD = N / BL;   // blocks per window
    static constexpr int NB = U / BL;   // blocks per unrolled body
    static_assert(U % BL == 0 && BL >= 2 && P >= 1 && P <= 4, "bad temporal3 shape");
    // shared memory of a CTA: suffix max [BL][NT] uint4 | block maxima [KD][NT] uint4 | page [(P-1)*U][NT] uint2 | table [T] uint2
    static constexpr size_t suf_bytes = (size_t)BL * T3_NT * 16;
    static constexpr size_t fifo_bytes = (size_t)KD * T3_NT * 16;
    static constexpr size_t page_bytes = (size_t)(P - 1) * U * T3_NT * 8;
    // FEED = 1: the frames of one unrolled body are staged in shared memory by bulk copies, K frames per mbarrier stage:
    // stage [U][NT] uint2 | full[U/K], empty[U/K] mbarriers.  The per-frame table then lives in global memory.
    static constexpr int NS = U / K;
    static constexpr size_t stage_off = suf_bytes + fifo_bytes + page_bytes;
    static constexpr size_t stage_bytes = FEED == 1 ? (size_t)U * T3_NT * 8 : 0;
    static constexpr size_t bar_off = stage_off + stage_bytes;
    static constexpr size_t bar_bytes = FEED == 1 ? (size_t)2 * NS * 8 : 0;
    static constexpr size_t tab_off = bar_off + bar_bytes;
    // the table is padded to whole blocks (frames past T inside the last block are computed, see thread_main)
    static size_t smem_bytes(int T) { return tab_off + (FEED == 1 ? 0 : (size_t)((T + BL - 1) / BL * BL) * 8); }
};

// per-frame table entry: x = (0x7fff - thr*L) in both u16 lanes, y = L = SlidingWindow.length (utils.py:288-296)
T3_HD uint2 table_entry(int thr, long long t, int n) {
    const long long L = t + 1 < n ? t + 1 : n;
    const int th = thr < 0 ? 0 : (thr > 255 ? 255 : thr);
    return make_uint2((unsigned)(0x7fff - th * (int)L) * 0x00010001u, (unsigned)L);
}

struct Acc {  // u16x2 lanes of the thread's two words: even pixels (A) and odd pixels (B)
    unsigned a0, b0, a1, b1;
};

// One thread's whole batch.  g = pixel group (8 px) of the thread, smem = the CTA's shared memory.
// FEED = 0: frames by 8-byte global loads K frames ahead (registers), table in shared memory;
// FEED = 1: frames by bulk copies into a shared-memory stage ring (K frames per mbarrier, U frames deep), table `gtab`
// in global memory.  cta_threads = threads of this CTA that own pixels (the last CTA of a frame may be partial).
template <int U, int BL, int P, int K, bool MASKED, int FEED>
T3_HD void thread_main(const FrameSrc &src, long long t0, int T, int g, int tid, unsigned char *smem,
                       uint8_t *bits, size_t bstride, const uint2 *gtab = nullptr, int cta_threads = T3_NT) {
    typedef Layout<U, BL, P, FEED, K> LY;
    constexpr int N = LY::N, KD = LY::KD, NB = LY::NB;
    static_assert(U % K == 0, "prefetch depth must divide the unroll length");
    uint4 *suf = reinterpret_cast<uint4 *>(smem) + tid;                      // slot s (= position s+1) at suf[s * NT]
    uint4 *fifo = reinterpret_cast<uint4 *>(smem + LY::suf_bytes) + tid;     // block maxima, slot r at fifo[r * NT]
    uint2 *page = reinterpret_cast<uint2 *>(smem + LY::suf_bytes + LY::fifo_bytes) + tid;
    const uint2 *tab = FEED == 1 ? gtab : reinterpret_cast<const uint2 *>(smem + LY::tab_off);
    const size_t off = (size_t)g * 8;
    const size_t HW = src.HW;

    // frame * mask (imgproc.py:96-101) with mask in {0,1}: a masked-out pixel has max = mean = 0 over any window, so its
    // predicate max*L - sum > thr*L is false (thr >= 0) whatever the window holds -- the mask is applied to the eight
    // output bits instead of to every loaded frame
    // ... by zeroing that pixel's weight in the two dot products that gather the eight bits (bits8w): no per-frame cost
    unsigned w0 = 0xF8FCFEFFu, w1 = 0x80C0E0F0u;
    if (MASKED) {
        const uint2 m = ldg8(src.mask + off);
        w0 &= m.x * 0xffu;  // {0,1} -> {0x00,0xff}
        w1 &= m.y * 0xffu;
    }

    // ---- history: ring position p (1 .. N-1) = frame t0-N+p; position 0 ("frame t0-N") counts as zeros -------
    unsigned r0[U], r1[U];  // the oldest U frames of the ring, raw
    unsigned W0 = 0, W1 = 0, SB0 = 0, SB1 = 0;
    Acc mid = {0, 0, 0, 0};
    {
        long long th = t0 - N;  // frame of position 0
        const int slot = th >= 0 ? (int)(th % src.R) : 0;  // ring slot of position 0 (frames are stored at t % R)
#pragma unroll
        for (int r = 0; r < KD; r++) {
            unsigned h0[BL], h1[BL];
#pragma unroll
            for (int jj = 0; jj < BL; jj++) {
                const int p = r * BL + jj;
                h0[jj] = h1[jj] = 0u;
                if (p > 0 && th + p >= 0) {
                    int s;
                    if (th >= 0) { s = slot + p; if (s >= src.R) s -= src.R; }  // N <= R: one wrap at most
                    else s = (int)(th + p);                                      // frame th+p < N <= R
                    const uint2 v = ldg8(src.ring + (size_t)s * HW + off);
                    h0[jj] = v.x;
                    h1[jj] = v.y;
                }
            }
            Acc bm = {0, 0, 0, 0};
#pragma unroll
            for (int jj = 0; jj < BL; jj++) {
                const int p = r * BL + jj;
                const unsigned b0 = hi2(h0[jj]), b1 = hi2(h1[jj]);
                const unsigned a0 = h0[jj] - (b0 << 8), a1 = h1[jj] - (b1 << 8);
                W0 += h0[jj]; W1 += h1[jj];
                SB0 += b0; SB1 += b1;
                bm.a0 = vmax2(bm.a0, a0); bm.b0 = vmax2(bm.b0, b0);
                bm.a1 = vmax2(bm.a1, a1); bm.b1 = vmax2(bm.b1, b1);
                if (p < U) { r0[p] = h0[jj]; r1[p] = h1[jj]; }
                else page[(p - U) * T3_NT] = make_uint2(h0[jj], h1[jj]);
            }
            if (r >= 1) {
                fifo[r * T3_NT] = make_uint4(bm.a0, bm.b0, bm.a1, bm.b1);
                mid.a0 = vmax2(mid.a0, bm.a0); mid.b0 = vmax2(mid.b0, bm.b0);
                mid.a1 = vmax2(mid.a1, bm.a1); mid.b1 = vmax2(mid.b1, bm.b1);
            }
        }
    }
    // suffix max of the oldest block by position, with the whole blocks of the window folded in: slot p-1 = max over
    // positions p .. BL-1 and over the blocks in between; slot BL-1 = those blocks alone
    {
        Acc a = mid;
        suf[(BL - 1) * T3_NT] = make_uint4(a.a0, a.b0, a.a1, a.b1);
#pragma unroll
        for (int p = BL - 1; p >= 1; p--) {
            const unsigned b0 = hi2(r0[p]), b1 = hi2(r1[p]);
            a.a0 = vmax2(a.a0, r0[p] - (b0 << 8)); a.b0 = vmax2(a.b0, b0);
            a.a1 = vmax2(a.a1, r1[p] - (b1 << 8)); a.b1 = vmax2(a.b1, b1);
            suf[(p - 1) * T3_NT] = make_uint4(a.a0, a.b0, a.a1, a.b1);
        }
    }

    // ---- batch frames: contiguous at src.cur ---------------------------------------------------------------
    // 32-bit strides: a frame's address is  base + immediate * HW  (one IMAD.WIDE.U32); (U + K) * HW < 2^32
    const unsigned HWu = (unsigned)HW;
    const uint8_t *gcur = src.cur + off;
    constexpr int KR = FEED == 1 ? 1 : K;
    // FEED = 2: register feed as FEED = 0, plus thread 0 of the CTA asks the TMA unit to pull the CTA's pixels of the
    // frames PFD .. PFD+K-1 ahead into L2 (cp.async.bulk.prefetch.L2, one per frame and CTA): the register loads K
    // frames ahead then hit L2 instead of waiting for HBM.
    constexpr int PFD = 4 * K;
    unsigned pf0[KR], pf1[KR];
    // FEED = 1: stage ring.  Stage s = frames [s*K, s*K+K) of a body; full[s] completes when their bytes have landed
    // (phase = body number), empty[s] when every thread of the CTA has read them.  Thread 0 is the producer: after
    // it has finished stage s it refills stage s-1 (free for one stage time already, so the wait rarely blocks).
    constexpr int NS = LY::NS;
    const uint2 *stg = reinterpret_cast<const uint2 *>(smem + LY::stage_off) + tid;
    const uint32_t stage_a = smem_u32(smem + LY::stage_off);
    const uint32_t full_a = smem_u32(smem + LY::bar_off), empty_a = full_a + NS * 8;
    const unsigned cta_bytes = (unsigned)cta_threads * 8u;
    auto issue_stage = [&](int st, int f0) {  // frames f0 .. f0+K-1 of the batch (those below T) -> stage st
        const int cnt = T - f0 < K ? T - f0 : K;
        mbar_expect_tx(full_a + st * 8, (unsigned)cnt * cta_bytes);
        for (int m = 0; m < cnt; m++)
            bulk_g2s(stage_a + (unsigned)(st * K + m) * (T3_NT * 8), gcur + (size_t)(f0 + m) * HW, cta_bytes, full_a + st * 8);
    };
    if (FEED == 1) {
        if (tid == 0) {
#pragma unroll 1
            for (int st = 0; st < NS; st++)
                if (st * K < T) issue_stage(st, st * K);
        }
    } else {
        if (FEED == 2 && tid == 0) {
#pragma unroll 1
            for (int m = K; m < PFD + K && m < T; m++) bulk_prefetch_l2(gcur + (size_t)m * HW, cta_bytes);
        }
#pragma unroll
        for (int m = 0; m < KR; m++) {
            pf0[m] = pf1[m] = 0u;
            if (m < T) { const uint2 v = ldg8(gcur + (size_t)((unsigned)m * HWu)); pf0[m] = v.x; pf1[m] = v.y; }
        }
    }
    Acc pm = {0, 0, 0, 0};  // prefix max of the current block (0 = identity)
    int fs = 0;             // FIFO slot of the current block (= slot of the window's oldest block)
    const uint8_t *gp = gcur + (size_t)K * HW;  // FEED 0, 2: the next frame to fetch (K ahead of the one being processed)
    const uint8_t *bo = bits + g;               // this frame's byte of predicate bits
    const uint2 *tb = tab;
    // frames left, kept in a vector register (nz is zero, but not provably): the per-frame "is there a frame K
    // ahead" test is then one ISETP instead of a uniform compare + a predicate transfer
    const int nz = g >> 31;

    unsigned parity = 0;  // FEED 1: body number & 1
    int fbase = 0;        // FEED 1: first frame of the body
    // P > 1: the page holds the (P-1)*U frames younger than the registers' as P-1 sub-pages of U frames; the sub-page of
    // body b mod (P-1) holds the oldest of them, which this body moves into the registers and replaces by its own frames
    uint2 *subpage = page;
    int sp = 0;
    for (int rem = T + nz; rem > 0; rem -= U) {
#pragma unroll
        for (int j = 0; j < U; j++) {
            const int jj = j % BL;
            // frames past the end of the batch inside the last block are computed on stale data and land in the
            // slack planes behind the bit buffer (the table is padded likewise): one exit test per block
            if (jj == 0 && j >= rem) return;
            unsigned x0, x1;
            if (FEED == 2 && j % K == 0 && tid == 0) {
#pragma unroll 1
                for (int m = j + PFD + K; m < j + PFD + 2 * K && m < rem; m++)
                    bulk_prefetch_l2(gcur + (size_t)(fbase + m) * HW, cta_bytes);
            }
            if (FEED == 1) {
#if defined(__CUDA_ARCH__)
                if (j % K == 0 && j < rem) mbar_wait(full_a + (j / K) * 8, parity);
                const uint2 v = stg[j * T3_NT];
#else
                uint2 v = make_uint2(0u, 0u);
                if (j < rem) v = ldg8(gcur + (size_t)(fbase + j) * HW);
#endif
                x0 = v.x; x1 = v.y;
            } else {
                x0 = pf0[j % KR]; x1 = pf1[j % KR];
                ldg8_if(pf0[j % KR], pf1[j % KR], gp, j + K, rem);
                gp += HW;
            }
            const unsigned o0 = r0[j], o1 = r1[j];  // frame t-N leaves the window
            if (P == 1) { r0[j] = x0; r1[j] = x1; }
            else {
                const uint2 pg = subpage[j * T3_NT];   // frame t-(P-1)*U moves from the page into the registers
                r0[j] = pg.x; r1[j] = pg.y;
                subpage[j * T3_NT] = make_uint2(x0, x1);
            }
            const uint4 s = suf[jj * T3_NT];  // oldest block, positions jj+1 .. BL-1, and the whole blocks in between
            const uint2 tl = tb[j];
            const unsigned cpk = tl.x, Lu = tl.y;
            // clean lanes of the new and the evicted frame
            const unsigned xb0 = hi2(x0), xb1 = hi2(x1);
            const unsigned xa0 = x0 - (xb0 << 8), xa1 = x1 - (xb1 << 8);
            const unsigned ob0 = hi2(o0), ob1 = hi2(o1);
            W0 = W0 + x0 - o0; W1 = W1 + x1 - o1;
            SB0 = SB0 + xb0 - ob0; SB1 = SB1 + xb1 - ob1;
            const unsigned SA0 = W0 - (SB0 << 8), SA1 = W1 - (SB1 << 8);
            pm.a0 = vmax2(pm.a0, xa0); pm.b0 = vmax2(pm.b0, xb0);
            pm.a1 = vmax2(pm.a1, xa1); pm.b1 = vmax2(pm.b1, xb1);
            const unsigned wa0 = vmax2(pm.a0, s.x), wb0 = vmax2(pm.b0, s.y);
            const unsigned wa1 = vmax2(pm.a1, s.z), wb1 = vmax2(pm.b1, s.w);
            // per lane: max*L + 0x7fff - thr*L - sum ; bit 15 set <=> max*L - sum > thr*L
            const unsigned va0 = wa0 * Lu + cpk - SA0, vb0 = wb0 * Lu + cpk - SB0;
            const unsigned va1 = wa1 * Lu + cpk - SA1, vb1 = wb1 * Lu + cpk - SB1;
            const unsigned M0 = prmt(va0, vb0, 0xFBD9u);  // sign-replicate bytes 1,5,3,7 -> 0x00/0xff per pixel
            const unsigned M1 = prmt(va1, vb1, 0xFBD9u);
            *const_cast<uint8_t *>(bo) = (uint8_t)(MASKED ? bits8w(M0, M1, w0, w1) : bits8(M0, M1));
            bo += bstride;
            if (FEED == 1 && j % K == K - 1) {
                // stage consumed.  Its data was read into registers above; the arrive orders those reads before the
                // producer's next bulk copy into the stage.
                const int st = j / K;
                mbar_arrive(empty_a + st * 8);
                if (tid == 0) {
                    // refill the stage before this one: for the next body, or (st == 0) stage NS-1 for this body,
                    // which the prologue has already filled for the first body
                    const int r = (st + NS - 1) % NS;
                    const int f0 = fbase + r * K + (st >= 1 ? U : 0);
                    const bool first = st == 0 && fbase == 0;
                    if (!first && f0 < T) {
                        mbar_wait(empty_a + r * 8, st >= 1 ? parity : parity ^ 1u);
                        issue_stage(r, f0);
                    }
                }
            }
            if (jj == BL - 1 && j + 1 < rem) {
                // block complete and more frames follow: it joins the whole blocks of the window, and the block that
                // the next one overwrites becomes the window's oldest block
                Acc a = {0, 0, 0, 0};  // max over the whole blocks of the next window
                if (KD > 1) {
                    fifo[fs * T3_NT] = make_uint4(pm.a0, pm.b0, pm.a1, pm.b1);
                    fs = fs + 1 == KD ? 0 : fs + 1;
                    int sl = fs;
#pragma unroll
                    for (int r = 1; r < KD; r++) {
                        sl = sl + 1 == KD ? 0 : sl + 1;
                        const uint4 f = fifo[sl * T3_NT];
                        a.a0 = vmax2(a.a0, f.x); a.b0 = vmax2(a.b0, f.y);
                        a.a1 = vmax2(a.a1, f.z); a.b1 = vmax2(a.b1, f.w);
                    }
                }
                pm.a0 = pm.b0 = pm.a1 = pm.b1 = 0u;
                const int qn = ((j / BL) + 1) % NB;  // ring registers of that block (static)
                suf[(BL - 1) * T3_NT] = make_uint4(a.a0, a.b0, a.a1, a.b1);
#pragma unroll
                for (int p = BL - 1; p >= 1; p--) {
                    const unsigned w0 = r0[qn * BL + p], w1 = r1[qn * BL + p];
                    const unsigned b0 = hi2(w0), b1 = hi2(w1);
                    a.a0 = vmax2(a.a0, w0 - (b0 << 8)); a.b0 = vmax2(a.b0, b0);
                    a.a1 = vmax2(a.a1, w1 - (b1 << 8)); a.b1 = vmax2(a.b1, b1);
                    suf[(p - 1) * T3_NT] = make_uint4(a.a0, a.b0, a.a1, a.b1);
                }
            }
        }
        tb += U;
        parity ^= 1u;
        fbase += U;
        if (P > 2) {
            sp = sp + 1 == P - 1 ? 0 : sp + 1;
            subpage = page + sp * (U * T3_NT);
        }
    }
}

}  // namespace t3

#if defined(__CUDACC__)
template <int U, int BL, int P, int K, bool MASKED, int MINB, int FEED>
__global__ void __launch_bounds__(T3_NT, MINB)
temporal3_kernel(FrameSrc src, long long t0, int T, int HWG, const int *__restrict__ thr, const uint2 *__restrict__ gtab,
                 uint8_t *__restrict__ bits) {
    extern __shared__ uint4 t3_smem[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(t3_smem);
    typedef t3::Layout<U, BL, P, FEED, K> LY;
    const int g = blockIdx.x * T3_NT + threadIdx.x;
    const int cta_threads = min(T3_NT, HWG - (int)blockIdx.x * T3_NT);
    if (FEED == 1) {
        if (threadIdx.x == 0) {
            const uint32_t full_a = t3::smem_u32(smem + LY::bar_off);
            for (int s = 0; s < LY::NS; s++) {
                t3::mbar_init(full_a + s * 8, 1);
                t3::mbar_init(full_a + (LY::NS + s) * 8, (unsigned)cta_threads);
            }
            t3::mbar_init_fence();
        }
    } else {
        uint2 *tab = reinterpret_cast<uint2 *>(smem + LY::tab_off);
        const int Tp = (T + BL - 1) / BL * BL;
        for (int i = threadIdx.x; i < Tp; i += T3_NT) tab[i] = t3::table_entry(thr[i < T ? i : T - 1], t0 + i, LY::N);
    }
    __syncthreads();
    if (g >= HWG) return;
    t3::thread_main<U, BL, P, K, MASKED, FEED>(src, t0, T, g, threadIdx.x, smem, bits, (size_t)HWG, gtab, cta_threads);
}

// per-frame table in global memory for the FEED = 1 kernels (padded to Tp entries)
__global__ void t3_table_kernel(const int *__restrict__ thr, long long t0, int T, int Tp, int n, uint2 *__restrict__ tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Tp) tab[i] = t3::table_entry(thr[i < T ? i : T - 1], t0 + i, n);
}
#endif
