// K1 on the 5th-generation tensor cores: the initial residual-by-dictionary cross-correlation
// (convolve1d 'same', hsc/modeling.py:149-188, called at :1077) as an IMPLICIT GEMM,
//
//     c[t][k] = sum_q A[t][q] * B[k][q],   A[t][q] = xz[(t-off)*F + q],  q = j*F + f  (Toeplitz view)
//
// with tcgen05.mma kind::tf32 (fp32 accumulators in TMEM) and the 3xTF32 split
//     x = x_hi + x_lo  =>  c ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi          (fp32-grade accuracy),
// so that atom selection and coefficients keep the reference's float32 precision.
//
// No im2col, not even in shared memory.  With F' = 4 floats per (super-)row the Toeplitz operand IS a
// canonical K-major / no-swizzle UMMA layout of the raw signal slab: core matrix = 8 rows x 16 B with
// a 16-byte row pitch, K-chunk stride LBO = 16 B, 8-row-group stride SBO = 128 B -- overlapping core
// matrices, which the tensor core reads like any other (probed on B200: tools/tc_probe.cu).  One
// 128-row A tile is therefore a (127*4 + Kd)-float slab (3 KB for L*F = 256) instead of 128 KB, and
// shared memory holds the whole dictionary slice B (hi and lo parts) for the life of the CTA.
//
//   F = 4: rows are time steps.          F = 2 / F = 1: s = 4/F time steps are grouped in one
//   super-row m (t = s*m + i); column (i,k) of the product uses the dictionary shifted by i*F
//   floats, B'[(i,k)][q] = D[k][q - i*F]; the output [T/s][s*K] row-major is the same memory as
//   [T][K].  Kd = roundup((L+s-1)*F, 8), Ntot = s*K.
//
// CTA roles (192 threads): warp 0 stages + splits the signal slab (4-stage ring), warp 1 allocates TMEM
// and issues the MMAs (one elected lane), warps 2-5 drain the accumulators (tcgen05.ld 32x32b) to HBM
// (2-deep TMEM ring, so the epilogue of tile i overlaps the MMAs of tile i+1).  Persistent grid:
// each CTA owns one N-slice of NS columns (its B slice never leaves shared memory) and strides over
// the (signal, M-tile) list.
#pragma once
#include <string.h>
#include <vector>
#include "common.cuh"

namespace hsc {
namespace tc {

constexpr int kThreads = 192;
constexpr int kTileM = 128;
constexpr int kStages = 4;      // slab ring depth (a slab is ~3 KB per part)

struct Plan {            // host-side geometry of one dictionary
    int s;               // time steps per super-row (4 / F)
    int Kd;              // padded reduction length
    int Ntot;            // s * K, unpadded
    int NS;              // columns per slice (multiple of 32, <= 256)
    int nslices;
    int slab_floats;     // 4*(kTileM-1) + Kd
    size_t smem_bytes;
    bool ok;
};

inline Plan make_plan(int K, int L, int F) {
    Plan p{};
    p.ok = false;
    if (!(F == 1 || F == 2 || F == 4)) return p;
    p.s = 4 / F;
    p.Kd = (((L + p.s - 1) * F) + 7) / 8 * 8;
    p.Ntot = p.s * K;
    const int npad = (p.Ntot + 31) / 32 * 32;
    int best = 0;
    for (int ns = 256; ns >= 32; ns -= 32) {
        if (npad % ns) continue;
        size_t b = (size_t)2 * ns * p.Kd * 4;
        if (b <= 160 * 1024) { best = ns; break; }
    }
    if (!best) return p;
    p.NS = best;
    p.nslices = npad / best;
    if (p.nslices > 148) return p;
    p.slab_floats = 4 * (kTileM - 1) + p.Kd;
    p.smem_bytes = (size_t)2 * p.NS * p.Kd * 4 + (size_t)2 * kStages * ((p.slab_floats + 3) / 4 * 4) * 4 + 1024 + 256;
    p.ok = true;
    return p;
}

// fp32 -> tf32, round to nearest, ties away from zero (what cvt.rna.tf32.f32 does), host side.
inline float tf32_rna_host(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return x;
    u = (u + 0x1000u) & 0xffffe000u;
    float r;
    memcpy(&r, &u, 4);
    return r;
}

// Expanded, shifted, split dictionary in the per-slice canonical layout the CTA copies verbatim:
//   out[((slice*(Kd/4) + kc)*NS + n)*4 + j] = part(B'[slice*NS + n][kc*4 + j])
inline void build_b_operand(const float* D, int K, int L, int F, const Plan& p, std::vector<float>& hi, std::vector<float>& lo) {
    const size_t n = (size_t)p.nslices * p.NS * p.Kd;
    hi.assign(n, 0.f);
    lo.assign(n, 0.f);
    const int LF = L * F;
    for (int row = 0; row < p.Ntot; ++row) {
        const int i = row / K, k = row % K;
        const int sl = row / p.NS, nn = row % p.NS;
        for (int q = 0; q < p.Kd; ++q) {
            const int src = q - i * F;
            if (src < 0 || src >= LF) continue;
            const float v = D[(size_t)k * LF + src];
            const float h = tf32_rna_host(v);
            const float l = tf32_rna_host(v - h);
            const size_t o = (((size_t)sl * (p.Kd / 4) + q / 4) * p.NS + nn) * 4 + (q % 4);
            hi[o] = h;
            lo[o] = l;
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version 1 (sm_100); no swizzle, base offset 0
    return d;
}

__device__ __forceinline__ uint32_t idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        :: "r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tHSC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra HSC_DONE_%=;\n\tbra HSC_WAIT_%=;\n\tHSC_DONE_%=:\n\t}\n" :: "r"(bar), "r"(parity) : "memory");
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
                 "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr));
}

struct Args {
    const float* x;       // [S][T][F]
    const float* b_hi;    // [nslices][Kd/4][NS][4]
    const float* b_lo;
    float* map;           // [S][T][K]
    int S, T, F, K, off;
    int s, Kd, Ntot, NS, nslices, slab_floats;
    int tmem_cols;        // power of two >= 2*NS
};

__global__ void __launch_bounds__(kThreads, 1) correlate_tc_kernel(Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = a.NS, Kd = a.Kd;
    const int slab_stride = (a.slab_floats + 3) / 4 * 4;                 // floats, keeps 16-byte alignment
    float* sBhi = reinterpret_cast<float*>(smem_raw);
    float* sBlo = sBhi + (size_t)NS * Kd;
    float* sA = sBlo + (size_t)NS * Kd;                                  // [stage][part][slab_stride]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 2 * kStages * slab_stride);  // 2*kStages + 4 mbarriers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    const uint32_t bar_slab_full = smem_u32(bars + 0), bar_slab_empty = smem_u32(bars + kStages);
    const uint32_t bar_acc_full = smem_u32(bars + 2 * kStages), bar_acc_empty = smem_u32(bars + 2 * kStages + 2);

    const int slice = blockIdx.x % a.nslices;
    const int cta_m = blockIdx.x / a.nslices;
    const int ctas_per_slice = gridDim.x / a.nslices;
    const int Ts = (a.T + a.s - 1) / a.s;                                // super-rows per signal
    const int MT = (Ts + kTileM - 1) / kTileM;
    const long long ntiles = (long long)a.S * MT;

    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(bar_slab_full + 8 * i, 1);
            mbar_init(bar_slab_empty + 8 * i, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_acc_full + 8 * i, 1);
            mbar_init(bar_acc_empty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    // dictionary slice (already split and laid out): straight 16-byte copy
    {
        const float4* ghi = reinterpret_cast<const float4*>(a.b_hi + (size_t)slice * NS * Kd);
        const float4* glo = reinterpret_cast<const float4*>(a.b_lo + (size_t)slice * NS * Kd);
        float4* shi = reinterpret_cast<float4*>(sBhi);
        float4* slo = reinterpret_cast<float4*>(sBlo);
        const int n4 = NS * Kd / 4;
        for (int e = tid; e < n4; e += kThreads) {
            shi[e] = __ldg(ghi + e);
            slo[e] = __ldg(glo + e);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ slab producer
        int stage = 0, phase = 0;
        for (long long tile = cta_m; tile < ntiles; tile += ctas_per_slice) {
            const int sig = (int)(tile / MT);
            const int m0 = (int)(tile % MT) * kTileM;
            const float* xs = a.x + (long long)sig * a.T * a.F;
            const long long g0 = ((long long)a.s * m0 - a.off) * a.F;    // flat index of slab element 0
            const long long gmax = (long long)a.T * a.F;
            mbar_wait(bar_slab_empty + 8 * stage, phase ^ 1);
            float* shi = sA + (size_t)(stage * 2 + 0) * slab_stride;
            float* slo = sA + (size_t)(stage * 2 + 1) * slab_stride;
            // all loads of a batch are issued before the first conversion (one DRAM round trip per 8 x 32 floats)
            for (int e0 = 0; e0 < a.slab_floats; e0 += 8 * 32) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int e = e0 + j * 32 + lane;
                    const long long gi = g0 + e;
                    v[j] = (e < a.slab_floats && gi >= 0 && gi < gmax) ? __ldg(xs + gi) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int e = e0 + j * 32 + lane;
                    if (e < a.slab_floats) {
                        const float h = to_tf32(v[j]);
                        shi[e] = h;
                        slo[e] = to_tf32(v[j] - h);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_slab_full + 8 * stage);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (one lane)
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(kTileM, NS);
            const uint32_t b_lbo = (uint32_t)NS * 16;
            const uint32_t bhi0 = smem_u32(sBhi), blo0 = smem_u32(sBlo);
            int stage = 0, phase = 0, acc = 0, acc_phase = 0;
            const int nk = Kd / 8;
            for (long long tile = cta_m; tile < ntiles; tile += ctas_per_slice) {
                mbar_wait(bar_slab_full + 8 * stage, phase);
                mbar_wait(bar_acc_empty + 8 * acc, acc_phase ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t ahi0 = smem_u32(sA + (size_t)(stage * 2 + 0) * slab_stride);
                const uint32_t alo0 = smem_u32(sA + (size_t)(stage * 2 + 1) * slab_stride);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NS);
                uint32_t accum = 0;
#pragma unroll 1
                for (int pass = 0; pass < 3; ++pass) {          // lo*hi, hi*lo, hi*hi (small terms first)
                    // descriptors advance by a constant in their 14-bit address field: one add per operand per MMA
                    uint64_t da = smem_desc(pass == 0 ? alo0 : ahi0, 16, 128);                // Toeplitz slab: LBO 16 B
                    uint64_t db = smem_desc(pass == 1 ? blo0 : bhi0, b_lbo, 128);
                    const uint64_t da_step = 32 >> 4, db_step = (2 * b_lbo) >> 4;
#pragma unroll 8
                    for (int kk = 0; kk < nk; ++kk) {
                        mma_tf32(d_tmem, da, db, idesc, accum);
                        accum = 1;
                        da += da_step;
                        db += db_step;
                    }
                }
                umma_commit(bar_slab_empty + 8 * stage);
                umma_commit(bar_acc_full + 8 * acc);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: TMEM -> HBM
        const int quarter = warp & 3;                     // TMEM lane quarter this warp may read
        int acc = 0, acc_phase = 0;
        const long long row_pitch = (long long)a.Ntot;    // floats per super-row of the output
        const long long map_elems = (long long)a.T * a.K;
        for (long long tile = cta_m; tile < ntiles; tile += ctas_per_slice) {
            const int sig = (int)(tile / MT);
            const int m0 = (int)(tile % MT) * kTileM;
            float* ms = a.map + (long long)sig * map_elems;
            mbar_wait(bar_acc_full + 8 * acc, acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int row = m0 + quarter * 32 + lane;
            const uint32_t t0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * NS);
            for (int c0 = 0; c0 < NS; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(t0 + c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c0 + 32 >= NS) {                      // last read of this accumulator: hand it back
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
                }
                const int col0 = slice * NS + c0;
                if (row < Ts && col0 < a.Ntot) {
                    const long long o = (long long)row * row_pitch + col0;
                    if (col0 + 32 <= a.Ntot && o + 32 <= map_elems && (row_pitch % 4) == 0 && (map_elems % 4) == 0) {
                        float4* dst = reinterpret_cast<float4*>(ms + o);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            __stcs(dst + j, make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                        __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (col0 + j < a.Ntot && o + j < map_elems) ms[o + j] = __uint_as_float(v[j]);
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(a.tmem_cols));
}

}  // namespace tc
}  // namespace hsc
