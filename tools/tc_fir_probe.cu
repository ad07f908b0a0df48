// tc_fir_probe.cu -- stand-alone probe: the STHT quadrature FIR as a tcgen05 Toeplitz GEMM.
//
// Question answered on the GPU (no documentation available offline): may the A operand of tcgen05.mma be a
// HANKEL matrix read straight from a time-contiguous fp16 sample array, i.e. a K-major no-swizzle shared-memory
// descriptor whose core matrices OVERLAP (row r of a core matrix = 8 samples starting at 8 r; leading byte offset
// 16 B)?  Then one M=128 x N=16 x K=16 instruction computes 8 output phases (x taps hi | taps lo) of 128 sliding
// windows and nothing is replicated in shared memory.
//
//   D[m][a]     = sum_e A[m][e] * B[a][e],   A[m][e] = u[8 m - LAG + e],   B[a][e] = g[a + LAG - e]
//   y[8 m + a]  = D[m][a] + D[m][8 + a]      (taps split hi + lo; data split hi + lo = two accumulating MMAs)
//
// Prints the maximum relative error against a float64 FIR and the cycles per 32-MMA batch.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_fir_probe tc_fir_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kTapsN = 240;           // dense polyphase taps
constexpr int kLag = 240;             // window start = 8 m - kLag
constexpr int kKSteps = 16;           // K = 256 >= 248
constexpr int kStreams = 16;          // row groups of one M = 128 instruction
constexpr int kRingLen = 384;         // samples per stream in shared memory (linear here)
constexpr int kPitchH = kRingLen + 8; // halves per stream row (pitch 784 B)
constexpr int kN = 16;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}

// mode 0: LBO = 16 (overlapping core matrices along K), SBO = stream pitch
// mode 1: the same with the two offsets swapped (in case the fields mean the opposite)
__global__ void __launch_bounds__(128) k_probe(const float *__restrict__ u, const float *__restrict__ g, float *__restrict__ out,
                                               long long *__restrict__ cycles, int mode, int reps, int n_start) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half *ring_hi = reinterpret_cast<__half *>(smem);                       // [kStreams][kPitchH]
    __half *ring_lo = ring_hi + kStreams * kPitchH;
    __half *tapsB = ring_lo + kStreams * kPitchH;                             // canonical [N=16][K=256]
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // data: stream i, position p holds u_i[p - kLag] (zeros before the start); u_i = u shifted by 37 i samples
    for (int e = tid; e < kStreams * kRingLen; e += 128) {
        const int i = e / kRingLen, p = e % kRingLen;
        const int q = p - kLag;
        const float v = q >= 0 ? u[q + 37 * i] : 0.f;
        const __half h = __float2half_rn(v);
        ring_hi[i * kPitchH + p] = h;
        ring_lo[i * kPitchH + p] = __float2half_rn(v - __half2float(h));
    }
    // taps: B[n][e], n = 8 set + a; set 0 = hi, set 1 = lo;  B[a][e] = g[a + kLag - e]
    // canonical K-major no-swizzle: element (n, e) at ((n / 8) * 32 + e / 8) * 128 B + (n % 8) * 16 B + (e % 8) * 2 B
    for (int e = tid; e < kN * 256; e += 128) {
        const int n = e / 256, kk = e % 256;
        const int a = n & 7, j = a + kLag - kk;
        float v = 0.f;
        if (j >= 0 && j < kTapsN) {
            const float gs = g[j] * 16384.f;
            const __half h = __float2half_rn(gs);
            v = (n >> 3) == 0 ? __half2float(h) : gs - __half2float(h);
        }
        tapsB[(((n >> 3) * 32 + (kk >> 3)) * 64) + (n & 7) * 8 + (kk & 7)] = __float2half_rn(v);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t parity = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            const uint32_t a_hi = smem_u32(ring_hi) + 2u * n_start, a_lo = smem_u32(ring_lo) + 2u * n_start;
            const uint32_t b0 = smem_u32(tapsB);
            const uint32_t pitch = kPitchH * 2;
#pragma unroll 1
            for (int piece = 0; piece < 2; ++piece) {
#pragma unroll
                for (int ks = 0; ks < kKSteps; ++ks) {
                    const uint32_t abase = (piece ? a_lo : a_hi) + 32u * ks;
                    const uint64_t ad = mode == 0 ? make_desc(abase, 16, pitch) : make_desc(abase, pitch, 16);
                    const uint64_t bd = make_desc(b0 + 256u * ks, 128, 4096);
                    mma_f16(tmem, ad, bd, idesc, (piece | ks) ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        mbar_wait(smem_u32(&mbar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t1 = clock64();
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0) {
        const int m = 32 * warp + lane;
#pragma unroll
        for (int c = 0; c < 16; ++c) out[m * 16 + c] = __uint_as_float(r[c]);
        if (tid == 0) cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 200;
    const int NU = 4096;
    std::vector<float> u(NU), g(kTapsN);
    srand(7);
    for (int i = 0; i < NU; ++i) u[i] = 12000.f * sinf(0.23f * i) + 3000.f * ((rand() % 2001) / 1000.f - 1.f);
    for (int j = 0; j < kTapsN; ++j) {   // polyphase Hilbert taps of K = 480: h[2 j + 1] = (2/K) cot(pi (2 j + 1 - 240) / K)
        const int n = 2 * j + 1 - 240;
        g[j] = (float)((2.0 / 480.0) / tan(M_PI * n / 480.0));
    }
    float *d_u, *d_g, *d_out; long long *d_cyc;
    cudaMalloc(&d_u, NU * 4); cudaMalloc(&d_g, kTapsN * 4); cudaMalloc(&d_out, 128 * 16 * 4); cudaMalloc(&d_cyc, 8);
    cudaMemcpy(d_u, u.data(), NU * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_g, g.data(), kTapsN * 4, cudaMemcpyHostToDevice);
    const int smem = (2 * kStreams * kPitchH + kN * 256) * 2 + 256;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 2; ++mode) {
        for (int n_start = 0; n_start <= 64; n_start += 64) {
            cudaMemset(d_out, 0, 128 * 16 * 4);
            k_probe<<<1, 128, smem>>>(d_u, d_g, d_out, d_cyc, mode, 1, n_start);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
            std::vector<float> out(128 * 16);
            cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; ++m) {
                const int i = m >> 3, r = m & 7;
                for (int a = 0; a < 8; ++a) {
                    const int n = n_start + 8 * r + a;      // stream output index
                    double ref = 0;
                    for (int j = 0; j < kTapsN; ++j) { const int q = n - j; if (q >= 0) ref += (double)g[j] * (double)u[q + 37 * i]; }
                    const double got = ((double)out[m * 16 + a] + (double)out[m * 16 + 8 + a]) / 16384.0;
                    maxerr = fmax(maxerr, fabs(got - ref)); maxref = fmax(maxref, fabs(ref));
                }
            }
            printf("mode %d n_start %d: max |err| %.3e  max |ref| %.3e  rel %.3e\n", mode, n_start, maxerr, maxref, maxerr / maxref);
        }
    }
    // timing: one CTA, then one CTA per SM
    for (int grid : {1, 148, 296}) {
        k_probe<<<grid, 128, smem>>>(d_u, d_g, d_out, d_cyc, 0, reps, 0);
        cudaDeviceSynchronize();
        long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("grid %d: %.1f cycles per batch of 32 MMAs (M128 N16 K16, SS), %.1f per MMA\n", grid, (double)cyc / reps, (double)cyc / reps / 32);
    }
    return 0;
}
