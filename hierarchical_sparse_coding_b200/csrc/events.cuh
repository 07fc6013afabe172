// Event-buffer compaction: the pursuit kernels write each signal's atoms into its own slice [s][capacity] of the event
// buffers; what leaves the device (device-to-host read of the codes, NCCL gather of the sparse codes) is the
// flat list of the n_buffered[s] atoms of every signal in signal order, 12-16 bytes per atom instead of the padded slices.
#pragma once
#include "common.cuh"

namespace hsc {
namespace events {

// offsets[s] = sum_{s' < s} n_buffered[s'], offsets[S] = total.  One block; S <= 65535 signals.
__global__ void __launch_bounds__(1024) offsets_kernel(const hsc_signal_state* __restrict__ states, int S, long long* __restrict__ offsets) {
    __shared__ long long warp_tot[32];
    __shared__ long long carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < S; base += 1024) {
        const int s = base + tid;
        const long long n = s < S ? states[s].n_buffered : 0;
        long long incl = n;                                       // inclusive scan inside the warp
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += o;
        }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = warp_tot[lane];
            for (int d = 1; d < 32; d <<= 1) {
                const long long o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += o;
            }
            warp_tot[lane] = w;                                   // inclusive totals of the warps
        }
        __syncthreads();
        const long long before = carry + (warp > 0 ? warp_tot[warp - 1] : 0) + (incl - n);
        if (s < S) offsets[s] = before;
        __syncthreads();
        if (tid == 1023) carry = before + n;
        __syncthreads();
    }
    if (tid == 0) offsets[S] = carry;
}

// One block per signal: its events to their place in the flat arrays (atoms past out_capacity are dropped; the caller
// compares offsets[S] with out_capacity).
template <typename real>
__global__ void __launch_bounds__(256) compact_kernel(const int* __restrict__ evp, const int* __restrict__ evi, const real* __restrict__ evc,
                                                      long long cap, const long long* __restrict__ offsets, int* __restrict__ pos,
                                                      int* __restrict__ idx, real* __restrict__ coef, long long out_capacity) {
    const long long s = blockIdx.x;
    const long long o = offsets[s];
    long long n = offsets[s + 1] - o;
    if (o + n > out_capacity) n = out_capacity > o ? out_capacity - o : 0;
    const int* p = evp + s * cap;
    const int* i = evi + s * cap;
    const real* c = evc + s * cap;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) {
        pos[o + e] = p[e];
        idx[o + e] = i[e];
        coef[o + e] = c[e];
    }
}

// Level hand-off of the hierarchical encoder (hsc/modeling.py:1489, `input = levelCoefficients.todense()`): the
// accumulated code of every signal of the encode in flight as a dense float64 map out[s][T][K] (zeroed by the caller).
// Same arithmetic as the reference's bookkeeping: the events of one (t, k) are summed in float64 in selection order
// (`coefficients[t,k] += c` on a float64 LIL matrix, :992), sums below minCoefficients are dropped (:1171-1177).
// One block per signal; the first event of every distinct (t, k) sums the later ones, so the result does not depend on
// the thread schedule (no atomics).  The event keys are staged through shared memory in tiles.
template <typename real>
__global__ void __launch_bounds__(256) events_to_dense_kernel(const hsc_signal_state* __restrict__ states, const int* __restrict__ evp,
                                                              const int* __restrict__ evi, const real* __restrict__ evc, long long cap,
                                                              int T, int K, double min_coef, double* __restrict__ out) {
    constexpr int TILE = 1024;
    __shared__ long long keys[TILE];
    const long long s = blockIdx.x;
    const long long n = states[s].n_buffered;
    const int* p = evp + s * cap;
    const int* i = evi + s * cap;
    const real* c = evc + s * cap;
    double* o = out + s * (long long)T * K;
    for (long long e0 = 0; e0 < n; e0 += blockDim.x) {
        const long long e = e0 + threadIdx.x;
        const long long key = e < n ? (long long)p[e] * K + i[e] : -1;
        bool first = e < n;
        double sum = e < n ? (double)c[e] : 0.0;
        // pass over ALL events in tiles: an earlier one with the same key -> this thread is not the owner; later ones add up
        for (long long j0 = 0; j0 < n; j0 += TILE) {
            __syncthreads();
            for (long long j = j0 + threadIdx.x; j < min(j0 + TILE, n); j += blockDim.x) keys[j - j0] = (long long)p[j] * K + i[j];
            __syncthreads();
            if (e < n) {
                const long long jn = min((long long)TILE, n - j0);
                for (long long j = 0; j < jn; ++j) {
                    if (keys[j] == key) {
                        const long long jj = j0 + j;
                        if (jj < e) first = false;
                        else if (jj > e) sum += (double)c[jj];          // ascending jj: selection order
                    }
                }
            }
        }
        if (first && sum != 0.0 && !(min_coef >= 0.0 && fabs(sum) < min_coef)) o[key] = sum;
    }
}

}  // namespace events
}  // namespace hsc
