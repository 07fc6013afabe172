// Probe (GPU box only): does tcgen05.mma kind::tf32 accept an OVERLAPPING K-major no-swizzle A
// descriptor (LBO = 16 B, SBO = 128 B), i.e. can the Toeplitz operand of the 1-D correlation be read
// straight out of the raw signal slab?  Compares (1) a canonical expanded A tile and (2) the slab
// trick against a CPU product on tf32-exact inputs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu && tools/tc_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

constexpr int M = 128, N = 64, KD = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;                 // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int m, int n) {
    uint32_t d = 0;
    d |= 1u << 4;              // c_format = F32
    d |= 2u << 7;              // a_format = TF32
    d |= 2u << 10;             // b_format = TF32
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(m >> 4) << 24;
    return d;                  // a_major = b_major = K
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        :: "r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\tbra WAIT_LOOP;\n\tDONE:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(128) probe(const float* __restrict__ Adense, const float* __restrict__ slab_in,
                                             const float* __restrict__ B, float* __restrict__ C, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sA = reinterpret_cast<float*>(smem);                   // 128*64*4 = 32 KB (mode 0) or slab (mode 1)
    float* sB = reinterpret_cast<float*>(smem + 32768);           // 64*64*4 = 16 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // operands -> shared memory
    uint32_t a_lbo, a_sbo;
    if (mode == 0) {   // canonical: chunk kc (4 floats) of row r at kc*2048 + r*16
        for (int e = tid; e < M * KD; e += 128) {
            int r = e / KD, c = e % KD;
            sA[(c / 4) * (M * 4) + r * 4 + (c % 4)] = Adense[e];
        }
        a_lbo = M * 16; a_sbo = 128;
    } else {           // Toeplitz slab: row r = slab[4r .. 4r+KD)
        for (int e = tid; e < 4 * (M - 1) + KD; e += 128) sA[e] = slab_in[e];
        a_lbo = 16; a_sbo = 128;
    }
    for (int e = tid; e < N * KD; e += 128) {
        int n = e / KD, c = e % KD;
        sB[(c / 4) * (N * 4) + n * 4 + (c % 4)] = B[e];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(M, N);
        for (int kk = 0; kk < KD / 8; ++kk) {
            uint64_t da = make_desc(smem_u32(sA) + kk * 2 * a_lbo, a_lbo, a_sbo);
            uint64_t db = make_desc(smem_u32(sB) + kk * 2 * (N * 16), N * 16, 128);
            mma_tf32(tmem_base, da, db, idesc, kk > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bar)) : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + half * 32;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
                     "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                       "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                       "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                       "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + lane;
        for (int j = 0; j < 32; ++j) C[row * N + half * 32 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(64));
}

int main() {
    const int slab_n = 4 * (M - 1) + KD;
    std::vector<float> slab(slab_n), A(M * KD), B(N * KD), Cref(M * N), C(M * N);
    srand(1);
    for (auto& v : slab) v = (float)((rand() % 17) - 8) / 8.0f;        // tf32-exact values
    for (auto& v : B) v = (float)((rand() % 13) - 6) / 4.0f;
    for (int r = 0; r < M; ++r) for (int c = 0; c < KD; ++c) A[r * KD + c] = slab[4 * r + c];
    for (int r = 0; r < M; ++r) for (int n = 0; n < N; ++n) {
        double acc = 0; for (int c = 0; c < KD; ++c) acc += (double)A[r * KD + c] * B[n * KD + c];
        Cref[r * N + n] = (float)acc;
    }
    float *dA, *dS, *dB, *dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dS, slab.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dC, C.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dS, slab.data(), slab.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024);
    int rc = 0;
    for (int mode = 0; mode < 2; ++mode) {
        cudaMemset(dC, 0, C.size() * 4);
        probe<<<1, 128, 49152 + 1024>>>(dA, dS, dB, dC, mode);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(err)); return 2; }
        cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int i = 0; i < M * N; ++i) { double e = fabs((double)C[i] - Cref[i]); if (e > maxerr) maxerr = e; if (e > 1e-4) ++bad; }
        printf("mode %d (%s): max abs err %.3e, mismatches %d / %d  C[0]=%f ref=%f C[last]=%f ref=%f\n", mode,
               mode == 0 ? "canonical A" : "Toeplitz slab A (LBO=16B)", maxerr, bad, M * N, C[0], Cref[0], C[M * N - 1], Cref[M * N - 1]);
        if (bad) rc = 1;
    }
    return rc;
}
