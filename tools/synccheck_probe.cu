// Does compute-sanitizer synccheck accept partial named barriers (bar.sync id, count < blockDim) of warp-specialised
// kernels?  Two groups of 128 threads meet at their own barrier a different number of times; no thread of a warp ever
// diverges at a barrier.  A report here is the tool's model, not a defect (tools/gpu_run_sanitize.sh, profiles/).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out, int n0, int n1) {
    const int grp = threadIdx.x >> 7;
    const int n = grp ? n1 : n0;
    int acc = 0;
    for (int i = 0; i < n; ++i) {
        acc += i;
        __syncwarp();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    int *d; cudaMalloc(&d, 4 * 256 * sizeof(int));
    k<<<4, 256>>>(d, 100, 1000);
    printf("sync: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
