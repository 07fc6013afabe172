// Shared device helpers of the B200 matching-pursuit engine (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <limits.h>

#include "../../include/hsc_b200.h"

namespace hsc {

// Centre tap of a filter of L taps (hsc/utils.py:83-99): L/2-1 for even L, L/2 for odd L.
__host__ __device__ inline int centre_offset(int L) { return (L % 2 == 0) ? (L / 2 - 1) : (L / 2); }

// np.pad(..., mode='reflect') index for a slice [lo, hi] extended arbitrarily far on both sides
// (hsc/modeling.py:1046).  Periodic with period 2(n-1); n == 1 repeats the single sample.
__host__ __device__ inline long long reflect_index(long long i, long long lo, long long hi) {
    if (i >= lo && i <= hi) return i;        // in-slice samples: no 64-bit modulo (it made an edge atom cost ~1 ms)
    long long n = hi - lo + 1;
    if (n <= 1) return lo;
    long long P = 2 * (n - 1);
    long long j = (i - lo) % P;
    if (j < 0) j += P;
    if (j >= n) j = P - j;
    return lo + j;
}

// (value, index) candidates of the argmax hierarchy: larger value wins, ties go to the LOWER index,
// which is np.argmax's first-occurrence rule on the row-major [T,K] map (hsc/modeling.py:967).
template <typename real>
struct Cand {
    real v;
    int i;
};

template <typename real>
__device__ __forceinline__ bool better(real v, int i, real bv, int bi) {
    return (v > bv) || (v == bv && i < bi);
}

template <typename real>
__device__ __forceinline__ void take_better(real& bv, int& bi, real v, int i) {
    if (better(v, i, bv, bi)) {
        bv = v;
        bi = i;
    }
}

__device__ __forceinline__ float shfl_xor(float v, int m, unsigned mask = 0xffffffffu) { return __shfl_xor_sync(mask, v, m); }
__device__ __forceinline__ double shfl_xor(double v, int m, unsigned mask = 0xffffffffu) { return __shfl_xor_sync(mask, v, m); }
__device__ __forceinline__ int shfl_xor(int v, int m, unsigned mask = 0xffffffffu) { return __shfl_xor_sync(mask, v, m); }

// Running best of one lane over candidates visited in INCREASING index order: strict '>' keeps the
// first occurrence, which is the lowest index.  Sentinel = (0, INT_MAX): scores are |.| >= 0.
template <typename real>
__device__ __forceinline__ void take_first_max(real& bv, int& bi, real v, int i) {
    if (v > bv) {
        bv = v;
        bi = i;
    }
}

// Argmax over the `width` consecutive lanes (power of two <= 32) that hold one row/group; every lane
// of the group ends up with the result.  double: butterfly shuffles.
__device__ __forceinline__ void group_argmax(double& v, int& i, int width) {
    for (int m = width >> 1; m > 0; m >>= 1) {
        double ov = shfl_xor(v, m);
        int oi = shfl_xor(i, m);
        take_better(v, i, ov, oi);
    }
}

// float: scores are non-negative, so their bit patterns order like unsigned integers -> two REDUX
// instructions (max of the value bits, then min index among the lanes that hold the max).
__device__ __forceinline__ void group_argmax(float& v, int& i, int width) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = width >= 32 ? 0xffffffffu : (((1u << width) - 1u) << (lane & ~(unsigned)(width - 1)));
    const unsigned ub = __float_as_uint(v);
    const unsigned mx = __reduce_max_sync(gmask, ub);
    const int cand = (ub == mx) ? i : INT_MAX;
    i = __reduce_min_sync(gmask, cand);
    v = __uint_as_float(mx);
}

__device__ __forceinline__ double warp_sum(double v) {
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

template <typename real>
__device__ __forceinline__ real warp_max(real v) {
    for (int m = 16; m > 0; m >>= 1) {
        real o = shfl_xor(v, m);
        v = o > v ? o : v;
    }
    return v;
}

template <typename real> __device__ __forceinline__ real rabs(real x);
template <> __device__ __forceinline__ float rabs<float>(float x) { return fabsf(x); }
template <> __device__ __forceinline__ double rabs<double>(double x) { return fabs(x); }

template <typename real> __device__ __forceinline__ real rlog10(real x);
template <> __device__ __forceinline__ float rlog10<float>(float x) { return log10f(x); }
template <> __device__ __forceinline__ double rlog10<double>(double x) { return log10(x); }

// r - c*d with the product rounded before the add, like NumPy's `signal += -c*element`
// (hsc/utils.py:119, hsc/modeling.py:1003); an FMA here would change the residual's last bit.
__device__ __forceinline__ float sub_scaled(float r, float c, float d) { return __fadd_rn(r, -__fmul_rn(c, d)); }
__device__ __forceinline__ double sub_scaled(double r, double c, double d) { return __dadd_rn(r, -__dmul_rn(c, d)); }

__host__ __device__ inline int pow2_at_least(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

// ---- bulk asynchronous copies (TMA, non-tensor form) + mbarrier helpers -----------------------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarrier_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbarrier_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarrier_wait_parity(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tHSCW_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HSCW_DONE_%=;\n\tbra HSCW_WAIT_%=;\n\tHSCW_DONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}
// global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
// same, with an L2 eviction-priority hint (createpolicy): the Gram tensor is re-read by every atom and should stay
// in L2 (evict_last), the map streams through (evict_first)
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void bulk_load_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 :: "r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store_s2g_hint(void* dst, uint32_t src_smem, uint32_t bytes, unsigned long long pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(dst), "r"(src_smem), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// One lane of a converged warp (always the same one for the full mask).  A branch on this predicate is known to the
// assembler to run a single thread, so the uniform-datapath bulk-copy instructions under it need no per-lane replay loop.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

}  // namespace hsc
