// Probe (GPU box only): where do the cycles of the K2 map-window update go?  One CTA per SM replays
// the interior update (127 rows x 256 floats: map RMW + Gram slice read + per-row argmax) at random
// windows of a private 64 MB map, with pieces switched off.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/win_probe tools/win_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <limits.h>
#include <vector>

constexpr int K = 256, L = 64, W = 2 * L - 1, T = 65536, NT = 256;

template <int MODE, int R>
__global__ void __launch_bounds__(NT) probe(float* __restrict__ map, const float* __restrict__ G, float* __restrict__ v1,
                                            int* __restrict__ i1, const int* __restrict__ pos, const int* __restrict__ kidx,
                                            int natoms, long long* cycles) {
    float* map_s = map + (size_t)blockIdx.x * T * K;
    float* v1s = v1 + (size_t)blockIdx.x * T;
    int* i1s = i1 + (size_t)blockIdx.x * T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long t0 = clock64();
    for (int a = 0; a < natoms; ++a) {
        const int t = pos[blockIdx.x * natoms + a];
        const float* Gk = G + (size_t)kidx[blockIdx.x * natoms + a] * W * K;
        const float ncoef = -0.37f;
        for (int base = 0; base < W; base += 8 * R) {
            float4 m[R][2], g[R][2];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = base + r * 8 + warp;
                if (i < W) {
                    const float4* mrow = reinterpret_cast<const float4*>(map_s + (size_t)(t - (L - 1) + i) * K);
                    const float4* grow = reinterpret_cast<const float4*>(Gk + (size_t)i * K);
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        m[r][p] = (MODE == 4) ? mrow[lane + 32 * p] : __ldcs(mrow + lane + 32 * p);
                        if (MODE != 3) g[r][p] = __ldg(grow + lane + 32 * p); else g[r][p] = make_float4(1, 2, 3, 4);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = base + r * 8 + warp;
                const int tr = t - (L - 1) + i;
                float bv = 0.f;
                int bi = INT_MAX;
                if (i < W) {
                    float4* mrow = reinterpret_cast<float4*>(map_s + (size_t)tr * K);
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        float pm[4] = {m[r][p].x, m[r][p].y, m[r][p].z, m[r][p].w};
                        const float pg[4] = {g[r][p].x, g[r][p].y, g[r][p].z, g[r][p].w};
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            pm[c] = fmaf(ncoef, pg[c], pm[c]);
                            const float sc = fabsf(pm[c]);
                            if (sc > bv) { bv = sc; bi = (lane + 32 * p) * 4 + c; }
                        }
                        if (MODE != 2) {
                            if (MODE == 4) mrow[lane + 32 * p] = make_float4(pm[0], pm[1], pm[2], pm[3]);
                            else __stcs(mrow + lane + 32 * p, make_float4(pm[0], pm[1], pm[2], pm[3]));
                        }
                    }
                }
                if (MODE != 1) {
                    const unsigned ub = __float_as_uint(bv);
                    const unsigned mx = __reduce_max_sync(0xffffffffu, ub);
                    const int cand = (ub == mx) ? bi : INT_MAX;
                    bi = __reduce_min_sync(0xffffffffu, cand);
                    bv = __uint_as_float(mx);
                } else if (MODE == 1) {
                    bv += __shfl_xor_sync(0xffffffffu, bv, 1);
                }
                if (i < W && lane == 0) { v1s[tr] = bv; i1s[tr] = bi; }
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

template <int MODE, int R>
void run(const char* name, float* map, float* G, float* v1, int* i1, int* pos, int* kidx, int natoms, long long* cyc, int nblk) {
    probe<MODE, R><<<nblk, NT>>>(map, G, v1, i1, pos, kidx, natoms, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE, R><<<nblk, NT>>>(map, G, v1, i1, pos, kidx, natoms, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(nblk);
    cudaMemcpy(h.data(), cyc, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0; for (auto c : h) mean += (double)c; mean /= nblk;
    printf("%-44s blocks=%d R=%d: %8.0f cycles/atom  (%.2f ms, %s)  %.0f GB/s map+gram+store\n", name, nblk, R, mean / natoms, ms,
           cudaGetErrorString(err), (double)nblk * natoms * W * K * 4 * 3 / (ms * 1e6));
}

int main() {
    const int nblk = 148, natoms = 300;
    float *map, *G, *v1; int *i1, *pos, *kidx; long long* cyc;
    cudaMalloc(&map, (size_t)nblk * 3 * T * K * 4);
    cudaMalloc(&G, (size_t)K * W * K * 4);
    cudaMalloc(&v1, (size_t)nblk * 3 * T * 4); cudaMalloc(&i1, (size_t)nblk * 3 * T * 4);
    cudaMalloc(&pos, nblk * 3 * natoms * 4); cudaMalloc(&kidx, nblk * 3 * natoms * 4); cudaMalloc(&cyc, nblk * 3 * 8);
    cudaMemset(map, 0, (size_t)nblk * 3 * T * K * 4); cudaMemset(G, 0, (size_t)K * W * K * 4);
    std::vector<int> hp(nblk * 3 * natoms), hk(nblk * 3 * natoms);
    srand(7);
    for (auto& p : hp) p = L + rand() % (T - 2 * L);
    for (auto& k : hk) k = rand() % K;
    cudaMemcpy(pos, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(kidx, hk.data(), hk.size() * 4, cudaMemcpyHostToDevice);
    run<0, 1>("full", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<0, 2>("full", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<0, 4>("full", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<1, 1>("no redux argmax", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<2, 1>("no map stores", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<3, 1>("no gram loads", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<4, 1>("default cache policy (no .cs)", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<4, 2>("default cache policy (no .cs)", map, G, v1, i1, pos, kidx, natoms, cyc, nblk);
    run<0, 2>("full, 3 CTAs/SM", map, G, v1, i1, pos, kidx, natoms, cyc, nblk * 3);
    run<4, 2>("default policy, 3 CTAs/SM", map, G, v1, i1, pos, kidx, natoms, cyc, nblk * 3);
    run<0, 4>("full, 3 CTAs/SM", map, G, v1, i1, pos, kidx, natoms, cyc, nblk * 3);
    return 0;
}
