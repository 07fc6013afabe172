// Probe (GPU box only): the reflect-padded EDGE re-correlation of K2 (pursuit.cuh: edge_recorrelate) costs ~0.5 ms per
// edge atom on the config-4 shape.  Every CTA recomputes NR rows x K filters of dot products of length L*F between a
// materialised residual slice `ext` and the dictionary D:  out[r][k] = sum_j D[k][j] * ext[r*F + j].
// Variants:
//   A  the shipped loop: 8 rows x 8 taps per step, one scalar load of ext per (row, tap), double accumulators
//   B  register window: 4 rows x 8 taps per step, the (R-1)*F+U distinct ext values of a step loaded once
//   C  register window: 8 rows x 4 taps per step
//   D  B with float accumulators (shows what the double FMAs cost; not a parity candidate)
//   E  B with ext staged in shared memory first;  F  8 x 8 window, ext in shared memory
//   G-J  F-like with the dictionary taps prefetched 1-4 steps ahead (edge_p)
//   K-N  ext converted once to the accumulator type in shared memory, operands straight from there (edge_s)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/edge_probe tools/edge_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

constexpr int K = 256, L = 64, F = 4, LF = L * F, NR = 2 * L - 1, NT = 256;
constexpr int EXT = (NR + L - 1) * F;      // samples of the slice
constexpr int EXT_STRIDE = 3 * L * F;      // per CTA, as the engine lays it out

__global__ void __launch_bounds__(NT, 4) edge_a(const float* __restrict__ D, const float* __restrict__ ext_all, float* __restrict__ out_all) {
    const float* ext = ext_all + (size_t)blockIdx.x * EXT_STRIDE;
    float* out = out_all + (size_t)blockIdx.x * NR * K;
    constexpr int R = 8, U = 8;
    const int kk = threadIdx.x;
    const float* dd = D + (size_t)kk * LF;
    const int nchunks = (NR + R - 1) / R;
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * R;
        const float* e0 = ext + (size_t)r0 * F;
        int roff[R];
#pragma unroll
        for (int r = 0; r < R; ++r) roff[r] = (min(r0 + r, NR - 1) - r0) * F;
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
        for (int q0 = 0; q0 < LF; q0 += U) {
            float dv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) dv[u] = dd[q0 + u];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma((double)e0[roff[r] + q0 + u], (double)dv[u], acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r0 + r < NR) out[(size_t)(r0 + r) * K + kk] = (float)acc[r];
    }
}

// Register-window variant: rows r0..r0+R-1 and taps q0..q0+U-1 touch ext[(r0*F + q0) + r*F + u]: (R-1)*F+U distinct
// values, loaded once per step (ext is padded so that reads past the last row stay inside the slice buffer).
template <int R, int U, typename ACC, bool SMEM>
__global__ void __launch_bounds__(NT, 4) edge_w(const float* __restrict__ D, const float* __restrict__ ext_all, float* __restrict__ out_all) {
    __shared__ float s_ext[SMEM ? EXT_STRIDE : 1];
    const float* ext = ext_all + (size_t)blockIdx.x * EXT_STRIDE;
    if (SMEM) {
        for (int i = threadIdx.x; i < EXT_STRIDE; i += NT) s_ext[i] = ext[i];
        __syncthreads();
        ext = s_ext;
    }
    float* out = out_all + (size_t)blockIdx.x * NR * K;
    constexpr int WN = (R - 1) * F + U;
    const int kk = threadIdx.x;
    const float4* dd = reinterpret_cast<const float4*>(D + (size_t)kk * LF);
    const int nchunks = (NR + R - 1) / R;
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * R;
        const float* e0 = ext + (size_t)r0 * F;
        ACC acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = (ACC)0;
        for (int q0 = 0; q0 < LF; q0 += U) {
            float dv[U], ew[WN];
#pragma unroll
            for (int u = 0; u < U / 4; ++u) {
                const float4 v = __ldg(dd + q0 / 4 + u);
                dv[4 * u] = v.x; dv[4 * u + 1] = v.y; dv[4 * u + 2] = v.z; dv[4 * u + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < WN / 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(e0 + q0 + 4 * i);
                ew[4 * i] = v.x; ew[4 * i + 1] = v.y; ew[4 * i + 2] = v.z; ew[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (sizeof(ACC) == 8) acc[r] = (ACC)fma((double)ew[r * F + u], (double)dv[u], (double)acc[r]);
                    else acc[r] = (ACC)fmaf(ew[r * F + u], dv[u], (float)acc[r]);
                }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r0 + r < NR) out[(size_t)(r0 + r) * K + kk] = (float)acc[r];
    }
}

// Window variant with the dictionary taps PREFETCHED P steps ahead (every chunk re-reads the same row of D from tap 0,
// so the prefetch simply wraps around): the variants above expose one L2 round trip per step.
template <int R, int U, int P>
__global__ void __launch_bounds__(NT, 4) edge_p(const float* __restrict__ D, const float* __restrict__ ext_all, float* __restrict__ out_all) {
    __shared__ float s_ext[EXT_STRIDE];
    {
        const float* ext = ext_all + (size_t)blockIdx.x * EXT_STRIDE;
        for (int i = threadIdx.x; i < EXT_STRIDE; i += NT) s_ext[i] = ext[i];
        __syncthreads();
    }
    float* out = out_all + (size_t)blockIdx.x * NR * K;
    constexpr int WN = (R - 1) * F + U, V = U / 4, STEPS = LF / U;
    const int kk = threadIdx.x;
    const float4* dd = reinterpret_cast<const float4*>(D + (size_t)kk * LF);
    const int nchunks = (NR + R - 1) / R;
    float4 pf[P][V];
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int v = 0; v < V; ++v) pf[p][v] = __ldg(dd + ((p % STEPS) * V) + v);
    int nxt = P % STEPS;                      // step (within a chunk) of the next prefetch
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * R;
        const float* e0 = s_ext + r0 * F;
        double acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0;
#pragma unroll 1
        for (int q0 = 0; q0 < LF; q0 += U) {
            float dv[U], ew[WN];
#pragma unroll
            for (int v = 0; v < V; ++v) { dv[4 * v] = pf[0][v].x; dv[4 * v + 1] = pf[0][v].y; dv[4 * v + 2] = pf[0][v].z; dv[4 * v + 3] = pf[0][v].w; }
#pragma unroll
            for (int p = 0; p + 1 < P; ++p)
#pragma unroll
                for (int v = 0; v < V; ++v) pf[p][v] = pf[p + 1][v];
#pragma unroll
            for (int v = 0; v < V; ++v) pf[P - 1][v] = __ldg(dd + nxt * V + v);
            nxt = (nxt + 1 == STEPS) ? 0 : nxt + 1;
#pragma unroll
            for (int i = 0; i < WN / 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(e0 + q0 + 4 * i);
                ew[4 * i] = v.x; ew[4 * i + 1] = v.y; ew[4 * i + 2] = v.z; ew[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma((double)ew[r * F + u], (double)dv[u], acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r0 + r < NR) out[(size_t)(r0 + r) * K + kk] = (float)acc[r];
    }
}

// ext converted ONCE to the accumulator type in shared memory and used straight from there (one broadcast LDS per
// FMA, no per-use float->double conversion: the conversions of variant F cost as much fp64-pipe time as its FMAs).
template <int R, int U, typename ACC>
__global__ void __launch_bounds__(NT, 4) edge_s(const float* __restrict__ D, const float* __restrict__ ext_all, float* __restrict__ out_all) {
    __shared__ ACC s_ext[EXT_STRIDE];
    {
        const float* ext = ext_all + (size_t)blockIdx.x * EXT_STRIDE;
        for (int i = threadIdx.x; i < EXT_STRIDE; i += NT) s_ext[i] = (ACC)ext[i];
        __syncthreads();
    }
    float* out = out_all + (size_t)blockIdx.x * NR * K;
    const int kk = threadIdx.x;
    const float4* dd = reinterpret_cast<const float4*>(D + (size_t)kk * LF);
    const int nchunks = (NR + R - 1) / R;
    for (int c = 0; c < nchunks; ++c) {
        const int r0 = c * R;
        const ACC* e0 = s_ext + r0 * F;
        ACC acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = (ACC)0;
#pragma unroll 1
        for (int q0 = 0; q0 < LF; q0 += U) {
            ACC dv[U];
#pragma unroll
            for (int v = 0; v < U / 4; ++v) {
                const float4 t = __ldg(dd + q0 / 4 + v);
                dv[4 * v] = (ACC)t.x; dv[4 * v + 1] = (ACC)t.y; dv[4 * v + 2] = (ACC)t.z; dv[4 * v + 3] = (ACC)t.w;
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < R; ++r) acc[r] = fma(e0[q0 + r * F + u], dv[u], acc[r]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r0 + r < NR) out[(size_t)(r0 + r) * K + kk] = (float)acc[r];
    }
}

template <typename Fn>
static float time_ms(Fn launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const int max_ctas = 148 * 4;
    std::vector<float> hD((size_t)K * LF), hext((size_t)max_ctas * EXT_STRIDE);
    srand(1);
    for (auto& v : hD) v = (float)rand() / RAND_MAX - 0.5f;
    for (auto& v : hext) v = (float)rand() / RAND_MAX - 0.5f;
    float *D, *ext, *out, *out_ref;
    cudaMalloc(&D, hD.size() * 4); cudaMalloc(&ext, hext.size() * 4);
    cudaMalloc(&out, (size_t)max_ctas * NR * K * 4); cudaMalloc(&out_ref, (size_t)max_ctas * NR * K * 4);
    cudaMemcpy(D, hD.data(), hD.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ext, hext.data(), hext.size() * 4, cudaMemcpyHostToDevice);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    std::vector<float> ha((size_t)NR * K), hb((size_t)NR * K);
    for (int grid : {148, 148 * 4}) {
        edge_a<<<grid, NT>>>(D, ext, out_ref);
        cudaDeviceSynchronize();
        cudaMemcpy(ha.data(), out_ref, ha.size() * 4, cudaMemcpyDeviceToHost);
        auto report = [&](const char* name, float ms) {
            cudaMemcpy(hb.data(), out, hb.size() * 4, cudaMemcpyDeviceToHost);
            double md = 0.0;
            for (size_t i = 0; i < ha.size(); ++i) md = fmax(md, fabs((double)ha[i] - (double)hb[i]));
            printf("grid %4d  %-28s %8.3f ms  (%.0f kcycles at %d MHz)  max|diff vs A| %.2e  %s\n", grid, name, ms, ms * clk_khz / 1e3,
                   clk_khz / 1000, md, cudaGetErrorString(cudaGetLastError()));
        };
        float ms = time_ms([&] { edge_a<<<grid, NT>>>(D, ext, out); }, 5);
        report("A shipped (8x8, scalar ext)", ms);
        ms = time_ms([&] { edge_w<8, 8, double, true><<<grid, NT>>>(D, ext, out); }, 5);
        report("F window 8 x 8, ext in smem", ms);
        ms = time_ms([&] { edge_s<8, 4, double><<<grid, NT>>>(D, ext, out); }, 5);
        report("K 8 x 4, ext as double in smem", ms);
        ms = time_ms([&] { edge_s<4, 8, double><<<grid, NT>>>(D, ext, out); }, 5);
        report("L 4 x 8, ext as double in smem", ms);
        ms = time_ms([&] { edge_s<8, 8, float><<<grid, NT>>>(D, ext, out); }, 5);
        report("M 8 x 8, float, ext in smem", ms);
        ms = time_ms([&] { edge_s<16, 4, float><<<grid, NT>>>(D, ext, out); }, 5);
        report("N 16 x 4, float, ext in smem", ms);
    }
    return 0;
}
