// Assignment step of the convolutional k-means dictionary learner (ConvolutionalDictionaryLearner._train_kmean,
// hsc/modeling.py:455-480; Dundar et al. 2016): every training window (twice the filter length) is correlated with
// the centroids at its 'valid' positions (K1 does that), the position / centroid of maximum |similarity| is found
// (np.argmax over the flattened [position][filter] scores: first occurrence, :459-460), and the L2-normalised patch
// at that position is added to its centroid's sum (the "cosine mean" of :480 is sum / count, taken by the host).
#pragma once
#include "common.cuh"

namespace hsc {
namespace kmeans {

// One CTA per window: argmax of |map[b][r][k]| over rows r in [row_lo, row_hi] and all k, lowest (r, k) on ties.
// pos[b] = r - row_lo (the 'valid' index = first sample of the patch), idx[b] = k.
template <typename real>
__global__ void __launch_bounds__(256) assign_kernel(const real* __restrict__ map, int Tw, int K, int row_lo, int row_hi,
                                                     int* __restrict__ pos, int* __restrict__ idx) {
    const long long b = blockIdx.x;
    const real* m = map + b * (long long)Tw * K;
    const int n = (row_hi - row_lo + 1) * K;
    real bv = (real)-1;
    int bi = INT_MAX;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {          // increasing e per thread: strict '>' keeps the first
        const real v = rabs<real>(m[(long long)row_lo * K + e]);
        if (v > bv) { bv = v; bi = e; }
    }
    __shared__ real s_v[256];
    __shared__ int s_i[256];
    s_v[threadIdx.x] = bv;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if ((int)threadIdx.x < h) {
            const real ov = s_v[threadIdx.x + h];
            const int oi = s_i[threadIdx.x + h];
            if (ov > s_v[threadIdx.x] || (ov == s_v[threadIdx.x] && oi < s_i[threadIdx.x])) { s_v[threadIdx.x] = ov; s_i[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int e = s_i[0] == INT_MAX ? 0 : s_i[0];
        pos[b] = e / K;
        idx[b] = e - (e / K) * K;
    }
}

// One warp per window: sums[idx[b]][q] += patch[q] / ||patch|| (zero-norm patches are added as they are, like
// normalize(), hsc/utils.py:67-74), counts[idx[b]] += 1.  float64 accumulation.
template <typename real>
__global__ void __launch_bounds__(256) accumulate_kernel(const real* __restrict__ x, long long B, int Tw, int LF, int F,
                                                         const int* __restrict__ pos, const int* __restrict__ idx,
                                                         double* __restrict__ sums, int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long b = warp; b < B; b += nwarps) {
        const real* p = x + (b * Tw + pos[b]) * (long long)F;
        double ss = 0.0;
        for (int q = lane; q < LF; q += 32) { const double v = (double)p[q]; ss = fma(v, v, ss); }
        ss = warp_sum(ss);
        const double nrm = sqrt(ss);
        const double sc = nrm > 0.0 ? 1.0 / nrm : 1.0;
        double* dst = sums + (long long)idx[b] * LF;
        for (int q = lane; q < LF; q += 32) atomicAdd(dst + q, (double)p[q] * sc);
        if (lane == 0) atomicAdd(counts + idx[b], 1);
    }
}

}  // namespace kmeans
}  // namespace hsc
