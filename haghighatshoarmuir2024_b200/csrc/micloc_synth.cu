// micloc_synth.cu -- Monte-Carlo input synthesis on the device.
//
// Reference sites restated:
//   SNNBeamformer.apply_to_template     micloc/snn_beamformer.py:243-275   (mode 0)
//       delays = -r cos(theta_m - doa) / c, minus their minimum; x[t][m] = interp(t - delay_m) clamped at t_min;
//       AWGN sigma = sqrt(mean(x^2)) / sqrt(snr) over the whole clip
//   signal_multiple_targets             paper_plots/multiple_targets_snn.py:87-159   (mode 1)
//       x[t][m] = sum_k gain_k interp(t + delay_{k,m})  (un-normalised delays, np.interp clamps at both ends)
// The source is either a table sampled on the clip's own grid (chirp, filtered noise, speech-shaped noise: built once
// on the host or by the caller, shared by all clips or one row per clip) or an analytic sine (the test clips of
// paper_plots/target_snn_localization.py:439-441), interpolated linearly between its samples exactly as np.interp
// does.  Noise is Philox4x32-10 + Box-Muller (cuRAND device API), one counter block per 4 samples.
//
// Three small kernels: clean clip + sum of squares per clip; noise (+ largest magnitude per clip); optional int16
// quantisation (peak / max|x| per clip: the int16 wire format of the host path).
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include "micloc_common.h"

namespace micloc {

struct SynthParams {
    int M, n_targets, mode, kind;
    long long T;
    double fs, f0, speed;
    float r[128], th[128];
};

__device__ __forceinline__ float synth_source(const SynthParams &sp, const float *__restrict__ src, double pos) {
    // linear interpolation between the samples floor(pos) and floor(pos) + 1 of the source (np.interp on the clip grid)
    const long long last = sp.T - 1;
    if (pos <= 0.0) pos = 0.0;
    if (pos >= (double)last) pos = (double)last;
    const long long i0 = (long long)pos;
    const float fr = (float)(pos - (double)i0);
    const long long i1 = i0 < last ? i0 + 1 : last;
    float s0, s1;
    if (sp.kind == 1) {
        // sine of frequency f0 sampled at fs: the phase is reduced in float64 turns, the sine taken in float32
        const double w = sp.f0 / sp.fs;
        double t0 = w * (double)i0, t1 = w * (double)i1;
        t0 -= floor(t0); t1 -= floor(t1);
        s0 = sinpif(2.f * (float)t0);
        s1 = sinpif(2.f * (float)t1);
    } else {
        s0 = src[i0];
        s1 = src[i1];
    }
    return fmaf(fr, s1 - s0, s0);
}

// grid (tiles over T, B); thread = one frame (all microphones)
__global__ void __launch_bounds__(256)
k_synth_clean(const __grid_constant__ SynthParams sp, const float *__restrict__ src, const int32_t *__restrict__ src_index,
              const double *__restrict__ doa, const float *__restrict__ gain, float *__restrict__ out,
              double *__restrict__ sumsq) {
    const long long b = blockIdx.y;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float *s = sp.kind == 0 ? src + (src_index ? (long long)src_index[b] : 0) * sp.T : nullptr;
    double acc = 0.0;
    if (t < sp.T) {
        float *o = out + (b * sp.T + t) * sp.M;
        for (int m = 0; m < sp.M; ++m) o[m] = 0.f;
        for (int k = 0; k < sp.n_targets; ++k) {
            const double th = doa[b * sp.n_targets + k];
            const float gk = gain ? gain[b * sp.n_targets + k] : 1.f;
            double dmin = 0.0;
            if (sp.mode == 0) {
                dmin = 1e300;
                for (int m = 0; m < sp.M; ++m) dmin = fmin(dmin, -(double)sp.r[m] * cos((double)sp.th[m] - th) / sp.speed);
            }
            for (int m = 0; m < sp.M; ++m) {
                const double d = -(double)sp.r[m] * cos((double)sp.th[m] - th) / sp.speed;
                const double pos = sp.mode == 0 ? (double)t - (d - dmin) * sp.fs : (double)t + d * sp.fs;
                o[m] += gk * synth_source(sp, s, pos);
            }
        }
        for (int m = 0; m < sp.M; ++m) acc += (double)o[m] * (double)o[m];
    }
    // block sum of squares -> one atomic per block
    __shared__ double red[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        atomicAdd(sumsq + b, tot);
    }
}

// x += sigma_b * N(0, 1); amax[b] = max |x| (bit pattern of the float); thread = 4 consecutive elements of one clip
__global__ void __launch_bounds__(256)
k_synth_noise(float *__restrict__ x, const double *__restrict__ sumsq, const float *__restrict__ snr_lin,
              unsigned int *__restrict__ amax, long long n_per_clip, unsigned long long seed) {
    const long long b = blockIdx.y;
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float mx = 0.f;
    if (i4 < n_per_clip) {
        float sigma = 0.f;
        if (snr_lin) sigma = (float)(sqrt(sumsq[b] / (double)n_per_clip) / sqrt((double)snr_lin[b]));
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        if (sigma != 0.f) {
            curandStatePhilox4_32_10_t st;
            curand_init(seed, (unsigned long long)(b * ((n_per_clip + 3) / 4) + i4 / 4), 0, &st);
            z = curand_normal4(&st);
        }
        float *p = x + b * n_per_clip + i4;
        const float zz[4] = {z.x, z.y, z.z, z.w};
        for (int e = 0; e < 4 && i4 + e < n_per_clip; ++e) {
            const float v = fmaf(sigma, zz[e], p[e]);
            p[e] = v;
            mx = fmaxf(mx, fabsf(v));
        }
    }
    if (amax) {
        unsigned int mb = __float_as_uint(mx);
        for (int o = 16; o > 0; o >>= 1) { const unsigned int ot = __shfl_xor_sync(0xffffffffu, mb, o); mb = ot > mb ? ot : mb; }
        if ((threadIdx.x & 31) == 0 && mb) atomicMax(amax + b, mb);
    }
}

__global__ void __launch_bounds__(256)
k_synth_quantize(const float *__restrict__ x, const unsigned int *__restrict__ amax, int16_t *__restrict__ out,
                 long long n_per_clip, float peak) {
    const long long b = blockIdx.y;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_per_clip) return;
    const float a = __uint_as_float(amax[b]);
    const float sc = a > 0.f ? peak / a : 0.f;
    out[b * n_per_clip + i] = (int16_t)__float2int_rn(x[b * n_per_clip + i] * sc);
}

}  // namespace micloc

using namespace micloc;

extern "C" int micloc_synth_clips(const micloc_synth_config *cfg, int64_t B, const float *src_dev,
                                  const int32_t *src_index_dev, const double *doa_dev, const float *gain_dev,
                                  const float *snr_lin_dev, uint64_t seed, float *out_f32_dev, int16_t *out_i16_dev,
                                  float int16_peak, double *scratch_dev, int device, void *stream) {
    if (!cfg || !doa_dev || !out_f32_dev || !scratch_dev) return set_error(MICLOC_ERR_SHAPE, "null argument");
    if (cfg->num_mic < 1 || cfg->num_mic > 128) return set_error(MICLOC_ERR_CONFIG, "num_mic %d out of range [1, 128]", cfg->num_mic);
    if (B < 1 || cfg->clip_len < 2) return set_error(MICLOC_ERR_SHAPE, "empty batch (B=%lld, T=%lld)", (long long)B, (long long)cfg->clip_len);
    if (B > 65535) return set_error(MICLOC_ERR_UNSUPPORTED, "synthesis takes at most 65535 clips per call");
    if (cfg->n_targets < 1) return set_error(MICLOC_ERR_CONFIG, "n_targets must be >= 1");
    if (cfg->source_kind == 0 && !src_dev) return set_error(MICLOC_ERR_SHAPE, "table source needs src_dev");
    if (cfg->source_kind != 0 && cfg->source_kind != 1) return set_error(MICLOC_ERR_CONFIG, "source_kind must be 0 (table) or 1 (sine)");
    if (!cfg->r_vec || !cfg->theta_vec) return set_error(MICLOC_ERR_CONFIG, "null geometry");
    MICLOC_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    SynthParams sp{};
    sp.M = cfg->num_mic; sp.n_targets = cfg->n_targets; sp.mode = cfg->mode; sp.kind = cfg->source_kind;
    sp.T = cfg->clip_len; sp.fs = cfg->fs; sp.f0 = cfg->sine_freq; sp.speed = cfg->speed > 0 ? cfg->speed : 340.0;
    for (int m = 0; m < sp.M; ++m) { sp.r[m] = (float)cfg->r_vec[m]; sp.th[m] = (float)cfg->theta_vec[m]; }
    // scratch: [B] float64 sums of squares, then [B] uint32 largest magnitudes
    double *sumsq = scratch_dev;
    unsigned int *amax = reinterpret_cast<unsigned int *>(scratch_dev + B);
    MICLOC_CUDA(cudaMemsetAsync(scratch_dev, 0, (size_t)B * (sizeof(double) + sizeof(unsigned int)), st));
    const long long n_per_clip = cfg->clip_len * cfg->num_mic;
    dim3 g1((unsigned)((cfg->clip_len + 255) / 256), (unsigned)B);
    k_synth_clean<<<g1, 256, 0, st>>>(sp, src_dev, src_index_dev, doa_dev, gain_dev, out_f32_dev, sumsq);
    dim3 g2((unsigned)(((n_per_clip + 3) / 4 + 255) / 256), (unsigned)B);
    k_synth_noise<<<g2, 256, 0, st>>>(out_f32_dev, sumsq, snr_lin_dev, out_i16_dev ? amax : nullptr, n_per_clip, seed);
    count_launch(2);
    if (out_i16_dev) {
        dim3 g3((unsigned)((n_per_clip + 255) / 256), (unsigned)B);
        k_synth_quantize<<<g3, 256, 0, st>>>(out_f32_dev, amax, out_i16_dev, n_per_clip, int16_peak > 0.f ? int16_peak : 12000.f);
        count_launch(1);
    }
    MICLOC_CUDA(cudaGetLastError());
    return MICLOC_OK;
}
