// Does compute-sanitizer synccheck accept partial named barriers (bar.sync id, count < blockDim) of warp-specialised
// kernels?  Two groups of 128 threads meet at their own barrier a different number of times; no thread of a warp ever
// diverges at a barrier.  Variant 1 (argv[1] = 1) is k_fused_tc's shape: the group that runs out of work waits at the
// CTA-wide barrier 0 (__syncthreads) while the other group still meets at its own named barrier -- legal PTX (separate
// barrier resources).  A report here is the tool's model, not a defect (tools/gpu_run_sanitize.sh, profiles/).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out, int n0, int n1, int join) {
    const int grp = threadIdx.x >> 7;
    const int n = grp ? n1 : n0;
    int acc = 0;
    for (int i = 0; i < n; ++i) {
        acc += i;
        __syncwarp();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    }
    if (join) __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
#include <cstdlib>
int main(int argc, char **argv) {
    const int join = argc > 1 ? atoi(argv[1]) : 0;
    int *d; cudaMalloc(&d, 4 * 256 * sizeof(int));
    k<<<4, 256>>>(d, join ? 0 : 100, 1000, join);
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
