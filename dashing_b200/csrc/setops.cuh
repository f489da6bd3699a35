// setops.cuh — register-array set operations behind `dashing union` and `dashing fold`.
// Reference paths are relative to /root/reference.
#pragma once
#include "common.cuh"

namespace db200 {

// ---------------------------------------------------------------------------------------------
// union: hll_t::operator+= (bonsai/hll/include/sketch/hll.h:958-992) folded over n sketches, as union_core does
// (src/union.cpp:33-58).  Element-wise byte maximum; max is associative and commutative, so any schedule gives the
// reference's bytes.  HBM-bound streaming: n * 2^p bytes read once.  CTA (bx, by) owns a 4 KiB column slab (256 threads x
// 16 bytes) and the row chunk `by`; its running maximum is merged into `out` (zero-initialised) with a packed-byte CAS.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_vmaxu4(uint32_t *addr, uint32_t v) {
    uint32_t old = *addr;
    while (true) {
        const uint32_t want = __vmaxu4(old, v);
        if (want == old) return;
        const uint32_t seen = atomicCAS(addr, old, want);
        if (seen == old) return;
        old = seen;
    }
}

__global__ void __launch_bounds__(256) union_kernel(const uint4 *__restrict__ regs16, uint64_t n, uint32_t m16, uint64_t rows_per_cta,
                                                    uint32_t *__restrict__ out32, int merge) {
    const uint32_t col = blockIdx.x * 256u + threadIdx.x;
    if (col >= m16) return;
    const uint64_t r0 = (uint64_t)blockIdx.y * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    uint4 acc = make_uint4(0u, 0u, 0u, 0u);
    const uint4 *src = regs16 + r0 * m16 + col;
    uint64_t r = r0;
    for (; r + 4 <= r1; r += 4, src += (uint64_t)4 * m16) {   // four independent loads in flight per thread
        const uint4 a = __ldg(src), b = __ldg(src + m16), c = __ldg(src + 2 * (uint64_t)m16), d = __ldg(src + 3 * (uint64_t)m16);
        acc.x = __vmaxu4(__vmaxu4(acc.x, a.x), __vmaxu4(__vmaxu4(b.x, c.x), d.x));
        acc.y = __vmaxu4(__vmaxu4(acc.y, a.y), __vmaxu4(__vmaxu4(b.y, c.y), d.y));
        acc.z = __vmaxu4(__vmaxu4(acc.z, a.z), __vmaxu4(__vmaxu4(b.z, c.z), d.z));
        acc.w = __vmaxu4(__vmaxu4(acc.w, a.w), __vmaxu4(__vmaxu4(b.w, c.w), d.w));
    }
    for (; r < r1; ++r, src += m16) {
        const uint4 a = __ldg(src);
        acc.x = __vmaxu4(acc.x, a.x); acc.y = __vmaxu4(acc.y, a.y); acc.z = __vmaxu4(acc.z, a.z); acc.w = __vmaxu4(acc.w, a.w);
    }
    uint32_t *dst = out32 + (uint64_t)col * 4;
    if (!merge) { dst[0] = acc.x; dst[1] = acc.y; dst[2] = acc.z; dst[3] = acc.w; return; }
    atomic_vmaxu4(dst + 0, acc.x); atomic_vmaxu4(dst + 1, acc.y); atomic_vmaxu4(dst + 2, acc.z); atomic_vmaxu4(dst + 3, acc.w);
}

// ---------------------------------------------------------------------------------------------
// fold: hll_t::compress(new_np) (hll.h:903-924).  New register i looks at the `ratio` = 2^(p - new_p) old registers
// [i*ratio, (i+1)*ratio): with j the offset of the first non-zero one,
//     all zero -> 0;   j == 0 -> min(q' + 1, old[b] + diff);   j > 0 -> min(q' + 1, clz(j) + 1)
// where q' = 64 - new_p and clz is the 64-bit count the reference's `clz(size_t)` overload yields — i.e. the code as
// written (64 - floor(log2 j), which saturates at q' + 1 for every p the CLI accepts), not Ertl's Algorithm 3.
// One thread per new register.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compress_kernel(const uint8_t *__restrict__ regs, uint64_t n, int p, int new_p,
                                                       uint8_t *__restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * 256u + threadIdx.x;
    const uint64_t newm = 1ull << new_p;
    if (t >= n * newm) return;
    const int diff = p - new_p;
    const uint64_t ratio = 1ull << diff, s = t >> new_p, i = t & (newm - 1);
    const uint8_t *src = regs + (s << p) + i * ratio;
    uint64_t j = 0;
    while (j < ratio && src[j] == 0) ++j;
    uint32_t v = 0;
    if (j != ratio) {
        const uint32_t cap = (uint32_t)(64 - new_p) + 1u;
        const uint32_t cand = j ? (uint32_t)__clzll((long long)j) + 1u : (uint32_t)src[0] + (uint32_t)diff;
        v = cand < cap ? cand : cap;
    }
    out[t] = (uint8_t)v;
}

} // namespace db200
