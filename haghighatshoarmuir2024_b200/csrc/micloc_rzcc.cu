// micloc_rzcc.cu -- ZeroCrossingSpikeEncoder.evolve on arbitrary float64 input with
// exact scipy.signal.find_peaks(distance=) semantics and no bound on cluster size
// (micloc/spike_encoder.py:115-137).
//
//   k_rzcc_scan: one thread per (clip, channel): sequential float64 cumsum (the same
//                left-to-right sum as np.cumsum, hence bit-identical), strict local
//                maxima / minima with scipy's plateau-midpoint rule -> candidate flags.
//   k_rzcc_nms : one CTA per (clip, channel, polarity): parallel fixed point of the
//                greedy rule of _select_by_peak_distance:
//                  keep a candidate when no undecided/kept neighbour nearer than w
//                  has higher priority; drop it when a kept neighbour is nearer than w.
//                Decisions only ever depend on higher-priority candidates, so the
//                fixed point equals the sequential greedy result.
#include <cuda_runtime.h>

#include "micloc_common.h"

namespace micloc {

enum : uint8_t { kNone = 0, kUndecided = 1, kKept = 2, kDropped = 3 };

__global__ void __launch_bounds__(128)
k_rzcc_scan(const double *__restrict__ sig, double *__restrict__ csum, uint8_t *__restrict__ flag,
            long long B, long long T, int C, int bipolar) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * C) return;
    const long long b = idx / C;
    const int c = (int)(idx % C);
    const double *x = sig + b * T * C + c;
    double *cs = csum + idx * T;
    uint8_t *fp = flag + (idx * 2 + 0) * T;  // peaks
    uint8_t *fv = flag + (idx * 2 + 1) * T;  // valleys
    double acc = 0.0, prev = 0.0;
    long long rise = -1, fall = -1;
    for (long long t = 0; t < T; ++t) {
        acc += x[t * C];
        cs[t] = acc;
        fp[t] = kNone; fv[t] = kNone;
        if (t >= 1) {
            if (acc > prev) {
                if (bipolar && fall >= 0) { fv[(fall + t - 1) >> 1] = kUndecided; fall = -1; }
                rise = t;
            } else if (acc < prev) {
                if (rise >= 0) { fp[(rise + t - 1) >> 1] = kUndecided; rise = -1; }
                fall = t;
            }
        }
        prev = acc;
    }
}

__global__ void __launch_bounds__(256)
k_rzcc_nms(const double *__restrict__ csum, uint8_t *__restrict__ flag, int8_t *__restrict__ spikes,
           long long T, int C, int w, int bipolar) {
    const long long chan = blockIdx.x;       // b*C + c
    const int pol = blockIdx.y;              // 0 peaks, 1 valleys
    if (pol == 1 && !bipolar) return;
    const double sgn = pol ? -1.0 : 1.0;
    const double *cs = csum + chan * T;
    volatile uint8_t *f = flag + (chan * 2 + pol) * T;
    for (;;) {
        int pending = 0;
        for (long long t = threadIdx.x; t < T; t += blockDim.x) {
            if (f[t] != kUndecided) continue;
            const double h = sgn * cs[t];
            bool kept_near = false, higher_near = false;
            const long long lo = t - (w - 1) < 0 ? 0 : t - (w - 1);
            const long long hi = t + (w - 1) > T - 1 ? T - 1 : t + (w - 1);
            for (long long u = lo; u <= hi; ++u) {
                if (u == t) continue;
                const uint8_t fu = f[u];
                if (fu == kKept) kept_near = true;
                else if (fu == kUndecided) {
                    const double hu = sgn * cs[u];
                    if (hu > h || (hu == h && u > t)) higher_near = true;
                }
            }
            if (kept_near) f[t] = kDropped;
            else if (!higher_near) f[t] = kKept;
            else pending = 1;
        }
        if (!__syncthreads_or(pending)) break;
    }
    // peaks write +1; the valley CTA of the same channel writes -1 to disjoint positions
    const long long b = chan / C;
    const int c = (int)(chan % C);
    for (long long t = threadIdx.x; t < T; t += blockDim.x)
        if (f[t] == kKept) spikes[(b * T + t) * C + c] = pol ? -1 : 1;
}

}  // namespace micloc

using namespace micloc;

extern "C" int micloc_rzcc_encode_f64(const double *sig_dev, int64_t B, int64_t T, int32_t C,
                                      int32_t robust_width, int32_t bipolar, int8_t *spikes_dev,
                                      int device, void *stream) {
    if (!sig_dev || !spikes_dev) return set_error(MICLOC_ERR_SHAPE, "null pointer");
    if (B < 1 || T < 1 || C < 1) return set_error(MICLOC_ERR_SHAPE, "empty input (B=%lld T=%lld C=%d)", (long long)B, (long long)T, C);
    if (robust_width < 1) return set_error(MICLOC_ERR_CONFIG, "`distance` must be greater or equal to 1");
    if (B * C > 0x7fffffffll) return set_error(MICLOC_ERR_UNSUPPORTED, "too many channels");
    MICLOC_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)B * C * T;
    double *csum = nullptr; uint8_t *flag = nullptr;
    MICLOC_CUDA(cudaMallocAsync((void **)&csum, n * sizeof(double), st));
    MICLOC_CUDA(cudaMallocAsync((void **)&flag, n * 2, st));
    MICLOC_CUDA(cudaMemsetAsync(spikes_dev, 0, n, st));
    k_rzcc_scan<<<(unsigned)((B * C + 127) / 128), 128, 0, st>>>(sig_dev, csum, flag, B, T, C, bipolar);
    dim3 grid((unsigned)(B * C), 2);
    k_rzcc_nms<<<grid, 256, 0, st>>>(csum, flag, spikes_dev, T, C, robust_width, bipolar);
    count_launch(2);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(csum, st);
    cudaFreeAsync(flag, st);
    if (e != cudaSuccess) return set_error(MICLOC_ERR_CUDA, "rzcc launch failed: %s", cudaGetErrorString(e));
    return MICLOC_OK;
}
