// tc_ts_probe.cu -- stand-alone probe: tcgen05.mma with the A operand in TENSOR MEMORY (kind::f16, M = 128).
// Answers (no documentation offline): how are the K = 16 fp16 elements of a row laid out in the 8 columns of its
// lane (which half of a 32-bit column is the even k), and what does one M128 x N32 x K16 instruction cost when only
// B (1 KB) comes from shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_ts_probe tc_ts_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kN = 32;
constexpr int kKSteps = 23;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
}

// A[m][k] (k < 16 * kKSteps) = ((m * 3 + k * 7) % 11) - 5;  B[n][k] = ((n * 5 + k) % 7) - 3 : exact in fp16
__device__ __host__ inline float a_val(int m, int k) { return (float)((m * 3 + k * 7) % 11 - 5); }
__device__ __host__ inline float b_val(int n, int k) { return (float)((n * 5 + k) % 7 - 3); }

__global__ void __launch_bounds__(128) k_probe(float *__restrict__ out, long long *__restrict__ cycles, int order, int reps) {
    extern __shared__ __align__(128) unsigned char smem[];
    __half *Bm = reinterpret_cast<__half *>(smem);     // canonical K-major no-swizzle [N = 32][K = 16 kKSteps]
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Kp = 16 * kKSteps;
    for (int e = tid; e < kN * Kp; e += 128) {
        const int n = e / Kp, k = e % Kp;
        Bm[((n >> 3) * (2 * kKSteps) + (k >> 3)) * 64 + (n & 7) * 8 + (k & 7)] = __float2half_rn(b_val(n, k));
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    // A into tensor memory: lane m = 32 warp + lane, columns 32 + 8 ks + c hold k = 16 ks + 2 c (+1)
    {
        const int m = 32 * warp + lane;
        for (int ks = 0; ks < kKSteps; ++ks) {
            uint32_t r[8];
            for (int c = 0; c < 8; ++c) {
                const __half e0 = __float2half_rn(a_val(m, 16 * ks + 2 * c)), e1 = __float2half_rn(a_val(m, 16 * ks + 2 * c + 1));
                const uint32_t lo = order == 0 ? __half_as_ushort(e0) : __half_as_ushort(e1);
                const uint32_t hi = order == 0 ? __half_as_ushort(e1) : __half_as_ushort(e0);
                r[c] = lo | (hi << 16);
            }
            const uint32_t taddr = tmem + 32u + 8u * ks + ((uint32_t)(32 * warp) << 16);
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                         ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t parity = 0;
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            const uint32_t b0 = smem_u32(Bm);
#pragma unroll 1
            for (int pass = 0; pass < 4; ++pass)            // 4 x 23 = 92 MMAs: the count of one 256-frame tile
#pragma unroll
                for (int ks = 0; ks < kKSteps; ++ks) {
                    const uint64_t bd = make_desc(b0 + 256u * ks, 128, 2 * kKSteps * 128);
                    const uint32_t accum = (pass | ks) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                                 ::"r"(tmem), "r"(tmem + 32u + 8u * ks), "l"(bd), "r"(idesc), "r"(accum) : "memory");
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
        }
        mbar_wait(smem_u32(&mbar), parity);
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    long long t1 = clock64();
    uint32_t r[32];
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (blockIdx.x == 0) {
        const int m = 32 * warp + lane;
        for (int c = 0; c < 32; ++c) out[m * 32 + c] = __uint_as_float(r[c]);
        if (tid == 0) cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main(int argc, char **argv) {
    const int reps = argc > 1 ? atoi(argv[1]) : 100;
    float *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 128 * 32 * 4); cudaMalloc(&d_cyc, 8);
    const int smem = kN * 16 * kKSteps * 2 + 256;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int order = 0; order < 2; ++order) {
        cudaMemset(d_out, 0, 128 * 32 * 4);
        k_probe<<<1, 128, smem>>>(d_out, d_cyc, order, 1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("order %d: CUDA error %s\n", order, cudaGetErrorString(e)); return 1; }
        std::vector<float> out(128 * 32);
        cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
                double ref = 0;
                for (int k = 0; k < 16 * kKSteps; ++k) ref += 4.0 * a_val(m, k) * b_val(n, k);     // 4 passes
                const double d = out[m * 32 + n] - ref;
                maxerr = d < 0 ? (-d > maxerr ? -d : maxerr) : (d > maxerr ? d : maxerr);
            }
        printf("A in TMEM, half order %d (0 = even k in the low half): max |err| %.3f\n", order, maxerr);
    }
    for (int grid : {1, 148}) {
        k_probe<<<grid, 128, smem>>>(d_out, d_cyc, 0, reps);
        cudaDeviceSynchronize();
        long long cyc; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("grid %d: %.1f cycles per chain of 92 MMAs (M128 N32 K16, A in TMEM), %.1f per MMA\n", grid, (double)cyc / reps, (double)cyc / reps / 92);
    }
    return 0;
}
