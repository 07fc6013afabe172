// K1 (general shapes, fp32/fp64 SIMT): initial residual-by-dictionary cross-correlation,
//   c[s][t][k] = sum_{j<L} sum_{f<F} xz[s][t-off+j][f] * D[k][j][f]      ('same', zero padded)
// i.e. convolve1d(x, D, padding='same') of the reference (hsc/modeling.py:149-188, called at :1077)
// WITHOUT the im2col copy the reference materialises (:186): each CTA stages the (BT+L-1) x F slab
// of the signal once in shared memory and reads the Toeplitz rows out of it.
// Also the Gram tensor builder used by the local update.
#pragma once
#include "common.cuh"

namespace hsc {

// Output tile: BT = TY*TM rows (time) x BK = TX*TN columns (filters); thread (ty,tx) owns rows
// ty + i*TY and columns tx + j*TX so that shared-memory reads are broadcast (slab) or unit-stride
// (dictionary chunk) and global stores are TX-wide runs.
template <typename real, int TX, int TY, int TM, int TN, int QC>
__global__ void __launch_bounds__(TX* TY)
correlate_same_kernel(const real* __restrict__ x, const real* __restrict__ D, real* __restrict__ map,
                      int T, int K, int L, int F, int off) {
    constexpr int BT = TY * TM;
    constexpr int BK = TX * TN;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    real* slab = reinterpret_cast<real*>(smem_raw);          // [(BT+L-1)*F]
    const int slab_rows = BT + L - 1;
    real* Dt = slab + (size_t)slab_rows * F;                 // [QC][BK+1]
    const int LF = L * F;
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int nthreads = TX * TY;
    const long long s = blockIdx.z;
    const int t0 = blockIdx.x * BT;
    const int k0 = blockIdx.y * BK;
    const real* xs = x + s * (long long)T * F;

    // stage the slab: rows t0-off .. t0-off+slab_rows-1, zero outside [0,T)
    for (int e = tid; e < slab_rows * F; e += nthreads) {
        int r = e / F;
        int gt = t0 - off + r;
        slab[e] = (gt >= 0 && gt < T) ? xs[(long long)gt * F + (e - r * F)] : (real)0;
    }

    real acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = (real)0;

    for (int q0 = 0; q0 < LF; q0 += QC) {
        __syncthreads();
        // dictionary chunk, transposed: Dt[qq][kk] = D[k0+kk][q0+qq]  (q fastest in global memory)
        for (int e = tid; e < QC * BK; e += nthreads) {
            int kk = e / QC;
            int qq = e - kk * QC;
            int gk = k0 + kk, gq = q0 + qq;
            Dt[qq * (BK + 1) + kk] = (gk < K && gq < LF) ? D[(long long)gk * LF + gq] : (real)0;
        }
        __syncthreads();
        const int qn = min(QC, LF - q0);
        for (int qq = 0; qq < qn; ++qq) {
            real xv[TM], dv[TN];
            const int q = q0 + qq;
#pragma unroll
            for (int i = 0; i < TM; ++i) xv[i] = slab[(threadIdx.y + i * TY) * F + q];
#pragma unroll
            for (int j = 0; j < TN; ++j) dv[j] = Dt[qq * (BK + 1) + threadIdx.x + j * TX];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fma(xv[i], dv[j], acc[i][j]);
        }
    }

    real* ms = map + s * (long long)T * K;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int t = t0 + threadIdx.y + i * TY;
        if (t >= T) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            int k = k0 + threadIdx.x + j * TX;
            if (k < K) ms[(long long)t * K + k] = acc[i][j];
        }
    }
}

template <typename real, int TM, int TN>
cudaError_t launch_correlate_cfg(const real* x, const real* D, real* map, long long S, int T, int K, int L, int F,
                                 cudaStream_t stream, bool* fits) {
    constexpr int TX = 16, TY = 16, QC = 32;
    constexpr int BT = TY * TM, BK = TX * TN;
    size_t smem = ((size_t)(BT + L - 1) * F + (size_t)QC * (BK + 1)) * sizeof(real);
    *fits = smem <= 200 * 1024;
    if (!*fits) return cudaSuccess;
    auto kern = correlate_same_kernel<real, TX, TY, TM, TN, QC>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    dim3 grid((T + BT - 1) / BT, (K + BK - 1) / BK, (unsigned)S);
    dim3 block(TX, TY);
    kern<<<grid, block, smem, stream>>>(x, D, map, T, K, L, F, centre_offset(L));
    return cudaGetLastError();
}

// Picks the widest tile the filter count can fill and the tallest the slab lets fit in shared memory.
template <typename real>
cudaError_t launch_correlate(const real* x, const real* D, real* map, long long S, int T, int K, int L, int F,
                             cudaStream_t stream, bool* ok) {
    bool fits = false;
    cudaError_t err = cudaSuccess;
    *ok = true;
#define HSC_TRY(TM, TN)                                                                  \
    err = launch_correlate_cfg<real, TM, TN>(x, D, map, S, T, K, L, F, stream, &fits);   \
    if (err != cudaSuccess || fits) return err;
    if (K >= 96) {
        HSC_TRY(8, 8) HSC_TRY(4, 8) HSC_TRY(2, 8) HSC_TRY(1, 8)
    } else if (K >= 48) {
        HSC_TRY(8, 4) HSC_TRY(4, 4) HSC_TRY(2, 4) HSC_TRY(1, 4)
    } else if (K >= 24) {
        HSC_TRY(8, 2) HSC_TRY(4, 2) HSC_TRY(2, 2) HSC_TRY(1, 2)
    } else {
        HSC_TRY(8, 1) HSC_TRY(4, 1) HSC_TRY(2, 1) HSC_TRY(1, 1)
    }
#undef HSC_TRY
    *ok = false;
    return cudaSuccess;
}

// Shift Gram tensor of the dictionary, the operand of the local map update:
//   G[k][i][k'] = sum_{j,f} D[k][j+tau][f] * D[k'][j][f],  tau = i-(L-1) in [-(L-1), L-1],
// so that subtracting c*D[k] centred at p changes map row p+tau, filter k' by -c*G[k][i][k'].
// (The reference's _precomputeGramMatrixForShifts, hsc/modeling.py:1204-1219, is dead code and masks
// instead of shifting; this is the shifted product.)  Accumulated in double, rounded once.
template <typename real>
__global__ void gram_kernel(const real* __restrict__ D, real* __restrict__ G, int K, int L, int F) {
    const int W = 2 * L - 1;
    long long total = (long long)K * W * K;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int k2 = (int)(e % K);
        long long r = e / K;
        int i = (int)(r % W);
        int k = (int)(r / W);
        int tau = i - (L - 1);
        int jlo = tau < 0 ? -tau : 0;
        int jhi = tau > 0 ? L - tau : L;
        const real* a = D + ((long long)k * L + tau) * F;   // a[j*F+f] = D[k][j+tau][f]
        const real* b = D + (long long)k2 * L * F;
        double acc = 0.0;
        for (int q = jlo * F; q < jhi * F; ++q) acc = fma((double)a[q], (double)b[q], acc);
        G[e] = (real)acc;
    }
}

}  // namespace hsc
