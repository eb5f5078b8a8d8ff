// Host stand-ins for the inline-PTX helpers of csrc/temporal_kernel.cuh (temporal2_kernel), used by the CPU emulation only:
// tests/emu_build.py removes the PTX definitions from a scratch copy of the kernel header and includes this file instead.
// "Shared addresses" are byte offsets into the emulated dynamic shared memory t_smem (the harness defines
// __cvta_generic_to_shared(t_smem) = 0); cp.async copies at once, commit / wait are no-ops (a legal timing: every group has
// landed when it is waited for).  Test infrastructure.
#pragma once
extern uint4 t_smem[];
static inline uint8_t *t2_sm(uint32_t saddr) { return reinterpret_cast<uint8_t *>(t_smem) + saddr; }
static inline void t2_commit() {}
template <int N>
static inline void t2_wait() {}
// PRMT.B32 in its default mode: result byte i = byte (sel_i & 7) of {b:a}, or that byte's sign replicated when sel_i & 8
static inline unsigned t2_prmt(unsigned a, unsigned b, unsigned sel) {
    const unsigned long long src = ((unsigned long long)b << 32) | a;
    unsigned d = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned s = (sel >> (4 * i)) & 0xfu;
        unsigned byte = (unsigned)(src >> (8 * (s & 7))) & 0xffu;
        if (s & 8) byte = (byte & 0x80u) ? 0xffu : 0u;
        d |= byte << (8 * i);
    }
    return d;
}
template <int WPT>
static inline void t2_cp(uint32_t saddr, const void *g) { memcpy(t2_sm(saddr), g, WPT * 4); }
template <int WPT>
static inline void t2_lds(unsigned (&w)[WPT], uint32_t saddr) { memcpy(w, t2_sm(saddr), WPT * 4); }
template <int WPT>
static inline void t2_sts(uint32_t saddr, const unsigned (&w)[WPT]) { memcpy(t2_sm(saddr), w, WPT * 4); }
static inline const uint8_t *t2_addr(const uint8_t *base, unsigned idx, unsigned stride) { return base + (unsigned long long)idx * stride; }
