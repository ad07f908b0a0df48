// micloc_stream.cu -- stateful frames of the float SNN chain + the Envelope tracker.
//
// Reference sites:
//   live frame loop             micloc/localization_demo_snn.py:125-193 (one recording = one frame; int32 `T x 8` wav
//                               frames, last channel dropped: micloc/record.py:54-75, localization_demo_snn.py:145)
//   chain                       micloc/snn_beamformer.py:283-370
//   Envelope                    micloc/utils.py:15-81; per-sample argmax tests/test_snn_hilbert_localization.py:284-293
//
// The reference restarts every filter from zero state at each frame.  A stream keeps the state instead, so that N
// pushed frames give exactly what ONE clip of their concatenation gives: the K-1 newest audio frames (STHT history and
// the K/2 in-phase delay), the band-pass biquad state, the running sum and the open candidate clusters of the RZCC
// encoder, the neuron recurrences and the envelope state all carry over.  Differences to a one-shot clip, by nature of
// a stream: the in-phase branch is the CAUSAL delay x[t - K/2] (zeros before the stream starts; np.roll's wrap-around
// needs the clip's end), and results lag the input by D = rzcc_lag(robust_width) samples -- a spike at p is only final
// once the distance rule has seen D more samples.  micloc_snn_stream_push returns the n_out samples that became final;
// micloc_snn_stream_flush closes the stream like a clip end and returns the rest.
//
// The kernels are the staged ones (micloc_staged.cuh) with their state in global memory: a stream is one clip wide
// (7 microphones = 14 sequential channels), i.e. latency matters here, not throughput.
#include <cuda_runtime.h>

#include <vector>

#include "micloc_common.h"
#include "micloc_staged.cuh"

namespace micloc {

struct ChanState {            // one per real channel (in-phase | quadrature)
    BiquadState bq;
    RzccState rz;
    int cl_pos[2 * kClusterMax];
    float cl_h[2 * kClusterMax];
    NeuronState nr;
};

// frame (any integer / float PCM, `in_ch` >= M interleaved channels, extra channels dropped) -> float rows behind
// the K-1 history rows of xbuf
template <typename IN_T>
__global__ void __launch_bounds__(256)
k_stream_ingest(const IN_T *__restrict__ frame, float *__restrict__ xbuf, long long n, int in_ch, int M, int hist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * M) return;
    const long long t = i / M;
    const int m = (int)(i % M);
    xbuf[(hist + t) * M + m] = (float)frame[t * in_ch + m];
}

// the K-1 newest rows move to the front (history of the next frame)
__global__ void __launch_bounds__(256)
k_stream_shift(float *__restrict__ xbuf, float *__restrict__ tmp, long long n, int M, int hist, int phase) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)hist * M) return;
    if (phase == 0) tmp[i] = xbuf[n * M + i]; else xbuf[i] = tmp[i];
}

// band-pass + RZCC of the new samples [N0, N0 + n), one thread per channel, state carried; spikes go to the ring
// (int8 [R][C2], R a power of two) at their absolute position.  `final` closes the open clusters (stream end).
__global__ void __launch_bounds__(32)
k_stream_chain(const float *__restrict__ xbuf, const float *__restrict__ q, ChanState *__restrict__ state,
               int8_t *__restrict__ ring, int ring_mask, int32_t *__restrict__ flags,
               const __grid_constant__ ChainParams p, long long N0, long long n, int hist, int final) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C2) return;
    ChanState &cs = state[c];
    BiquadState bq = cs.bq;
    RzccState rz = cs.rz;
    const RzccStore store{cs.cl_pos, cs.cl_h, 1};
    auto emit = [&](int pos, int sign) { ring[(long long)(pos & ring_mask) * p.C2 + c] = (int8_t)sign; };
    const bool inphase = c < p.M;
    for (long long i = 0; i < n; ++i) {
        const long long t = N0 + i;
        // in-phase: x[t - K/2] (row hist + i - K/2 of xbuf); quadrature: q of the extended clip at row hist + i
        const float x = inphase ? xbuf[(hist + i - p.half) * p.M + c] : q[(hist + i) * p.M + (c - p.M)];
        const float z = biquad_step(p.sos, p.nsec, bq, x);
        ring[(t & ring_mask) * p.C2 + c] = 0;
        rzcc_detect(rz, store, p.bipolar, p.w, (int)t, z, emit);
        if ((t & (kSeg - 1)) == kSeg - 1) rzcc_close(rz, store, p.w, (int)t, false, emit);
    }
    if (final && N0 + n > 0) rzcc_close(rz, store, p.w, (int)(N0 + n - 1), true, emit);
    cs.bq = bq;
    cs.rz = rz;
    if (rz.overflow && flags) atomicOr(flags, 1);
}

// neuron filter of the final samples [F0, F0 + m): spikes from the ring -> int8 spikes out + membrane rows
__global__ void __launch_bounds__(32)
k_stream_neuron(const int8_t *__restrict__ ring, int ring_mask, ChanState *__restrict__ state,
                int8_t *__restrict__ spikes_out, float *__restrict__ vmem, const __grid_constant__ ChainParams p,
                long long F0, long long m) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C2) return;
    NeuronState nr = state[c].nr;
    for (long long i = 0; i < m; ++i) {
        const long long t = F0 + i;
        const int8_t s8 = ring[(t & ring_mask) * p.C2 + c];
        const float sd = t >= p.nL ? (float)ring[((t - p.nL) & ring_mask) * p.C2 + c] : 0.f;
        if (spikes_out) spikes_out[i * p.C2 + c] = s8;
        vmem[i * p.C2 + c] = neuron_step(p, nr, (float)s8, sd);
    }
    state[c].nr = nr;
}

}  // namespace micloc

// Envelope.evolve (micloc/utils.py:49-81) over rows of |x|, one thread per channel, state carried between calls:
//   first row ever: state = |x[0]|, out[0] = state;  afterwards  rise = |x| >= state,  win = rise ? int(fs rise_time)
//   : int(fs fall_time),  state = (1 - 1/win) state + (1/win) |x| rise,  out = state
namespace micloc {
__global__ void __launch_bounds__(128)
k_envelope(const float *__restrict__ x, float *__restrict__ env, float *__restrict__ state, int32_t *__restrict__ started,
           long long T, int C, float inv_fall, float inv_rise) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = state[c];
    long long t = 0;
    if (!*started && T > 0) { s = fabsf(x[c]); env[c] = s; t = 1; }
    for (; t < T; ++t) {
        const float a = fabsf(x[t * C + c]);
        const bool rise = a >= s;
        const float iw = rise ? inv_rise : inv_fall;
        s = (1.f - iw) * s + (rise ? iw * a : 0.f);
        env[t * C + c] = s;
    }
    state[c] = s;
}
__global__ void k_envelope_mark(int32_t *started, long long T) { if (T > 0) *started = 1; }

// first argmax over the C channels of every row (np.argmax(env, axis=1)); one warp per row
__global__ void __launch_bounds__(256)
k_argmax_rows(const float *__restrict__ env, int32_t *__restrict__ idx, long long T, int C) {
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    float best = -1.f; int bi = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
        const float v = env[t * C + c];
        if (v > best) { best = v; bi = c; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) idx[t] = bi;
}
}  // namespace micloc

using namespace micloc;

struct micloc_stream {
    micloc_snn_params prm{};        // copy of what the kernels need from the parent context
    int device = 0;
    long long max_frame = 0;
    int hist = 0;                   // K - 1
    int D = 0;                      // output latency in samples
    int ring_len = 0;
    long long N = 0;                // samples pushed
    long long F = 0;                // samples returned (final)
    bool flushed = false;
    float *xbuf = nullptr, *tmp = nullptr, *q = nullptr, *vmem = nullptr, *y = nullptr, *env = nullptr, *env_state = nullptr;
    int8_t *ring = nullptr;
    ChanState *state = nullptr;
    int32_t *flags = nullptr, *env_started = nullptr;
    double *gram = nullptr;
    float rise_time = 10e-3f, fall_time = 100e-3f, fs = 48000.f;
};

extern "C" int micloc_snn_stream_destroy(micloc_stream *s) {
    if (!s) return MICLOC_OK;
    cudaSetDevice(s->device);
    cudaFree(s->xbuf); cudaFree(s->tmp); cudaFree(s->q); cudaFree(s->vmem); cudaFree(s->y); cudaFree(s->env);
    cudaFree(s->env_state); cudaFree(s->ring); cudaFree(s->state); cudaFree(s->flags); cudaFree(s->env_started);
    cudaFree(s->gram);
    delete s;
    return MICLOC_OK;
}

static int stream_reset_state(micloc_stream *s, cudaStream_t st) {
    const ChainParams &p = s->prm.chain;
    MICLOC_CUDA(cudaMemsetAsync(s->xbuf, 0, (size_t)(s->hist + s->max_frame) * p.M * sizeof(float), st));
    MICLOC_CUDA(cudaMemsetAsync(s->ring, 0, (size_t)s->ring_len * p.C2, st));
    MICLOC_CUDA(cudaMemsetAsync(s->env_state, 0, (size_t)p.G * sizeof(float), st));
    MICLOC_CUDA(cudaMemsetAsync(s->env_started, 0, sizeof(int32_t), st));
    MICLOC_CUDA(cudaMemsetAsync(s->flags, 0, sizeof(int32_t), st));
    std::vector<ChanState> init((size_t)p.C2);
    for (auto &c : init) {
        for (int k = 0; k < kMaxSections; ++k) { c.bq.s1[k] = 0.f; c.bq.s2[k] = 0.f; }
        c.rz.csum = 0.f; c.rz.r = -1; c.rz.sgn = 0; c.rz.n0 = c.rz.n1 = 0; c.rz.last0 = c.rz.last1 = 0; c.rz.overflow = 0;
        c.nr.p1 = c.nr.p2 = c.nr.q1 = c.nr.q2 = 0.f;
        for (int i = 0; i < 2 * kClusterMax; ++i) { c.cl_pos[i] = 0; c.cl_h[i] = 0.f; }
    }
    MICLOC_CUDA(cudaMemcpyAsync(s->state, init.data(), init.size() * sizeof(ChanState), cudaMemcpyHostToDevice, st));
    MICLOC_CUDA(cudaStreamSynchronize(st));      // `init` is a local vector
    s->N = 0; s->F = 0; s->flushed = false;
    return MICLOC_OK;
}

extern "C" int micloc_snn_stream_create(micloc_snn *ctx, int64_t max_frame_len, double fs, double rise_time,
                                        double fall_time, micloc_stream **out) {
    if (!ctx || !out) return set_error(MICLOC_ERR_CONFIG, "null argument");
    *out = nullptr;
    if (max_frame_len < 1 || max_frame_len > (1ll << 24)) return set_error(MICLOC_ERR_SHAPE, "max_frame_len out of range");
    if (rise_time > fall_time)
        return set_error(MICLOC_ERR_CONFIG, "for proper functioning, an envelope estimator should have a larger fall time!");
    if ((int)(fs * rise_time) < 1 || (int)(fs * fall_time) < 1) return set_error(MICLOC_ERR_CONFIG, "envelope windows shorter than one sample");
    micloc_snn_params prm;
    MICLOC_TRY(micloc_snn_get_params(ctx, &prm));
    const ChainParams &p = prm.chain;
    MICLOC_CUDA(cudaSetDevice(prm.device));
    micloc_stream *s = new micloc_stream();
    s->prm = prm; s->device = prm.device; s->max_frame = max_frame_len;
    s->hist = p.K - 1;
    s->D = rzcc_lag(p.w);
    int need = (int)max_frame_len + s->D + p.nL + 2 * kSeg;
    s->ring_len = 1;
    while (s->ring_len < need) s->ring_len <<= 1;
    s->fs = (float)fs; s->rise_time = (float)rise_time; s->fall_time = (float)fall_time;
    const size_t rows = (size_t)s->hist + max_frame_len;
    const size_t fin = (size_t)max_frame_len + s->D;                // most samples one call can finalise
    bool ok = cudaMalloc(&s->xbuf, rows * p.M * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&s->tmp, (size_t)s->hist * p.M * sizeof(float) + 16) == cudaSuccess &&
              cudaMalloc(&s->q, rows * p.M * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&s->vmem, fin * p.C2 * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&s->y, fin * p.G * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&s->env, fin * p.G * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&s->env_state, (size_t)p.G * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&s->ring, (size_t)s->ring_len * p.C2) == cudaSuccess &&
              cudaMalloc(&s->state, (size_t)p.C2 * sizeof(ChanState)) == cudaSuccess &&
              cudaMalloc(&s->flags, sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(&s->env_started, sizeof(int32_t)) == cudaSuccess &&
              cudaMalloc(&s->gram, (size_t)p.C2 * p.C2 * sizeof(double)) == cudaSuccess;
    if (!ok) { micloc_snn_stream_destroy(s); return set_error(MICLOC_ERR_CUDA, "cudaMalloc(stream) failed"); }
    const int rc = stream_reset_state(s, nullptr);
    if (rc) { micloc_snn_stream_destroy(s); return rc; }
    *out = s;
    return MICLOC_OK;
}

extern "C" int micloc_snn_stream_reset(micloc_stream *s, void *stream) {
    if (!s) return set_error(MICLOC_ERR_CONFIG, "null stream");
    MICLOC_CUDA(cudaSetDevice(s->device));
    return stream_reset_state(s, (cudaStream_t)stream);
}

extern "C" int micloc_snn_stream_latency(micloc_stream *s) { return s ? s->D : 0; }

// shared tail of push / flush: the samples [F, F_new) are final -> neuron, spikes, Gram power, envelope, argmax
static int stream_finalise(micloc_stream *s, long long F_new, int8_t *spikes_dev, float *power_dev, int32_t *doa_dev,
                           float *env_dev, int32_t *doa_t_dev, int64_t *n_out, cudaStream_t st) {
    const ChainParams &p = s->prm.chain;
    const long long m = F_new - s->F;
    if (n_out) *n_out = m;
    if (m <= 0) return MICLOC_OK;
    k_stream_neuron<<<(p.C2 + 31) / 32, 32, 0, st>>>(s->ring, s->ring_len - 1, s->state, spikes_dev, s->vmem, p, s->F, m);
    count_launch(1);
    if (power_dev || doa_dev) {
        dim3 gg(1, (unsigned)((p.C2 * p.C2 + 255) / 256));
        k_gram<<<gg, 256, 0, st>>>(s->vmem, s->gram, p.C2, m, 0);
        const size_t smem = (size_t)p.C2 * p.C2 * sizeof(double);
        MICLOC_CUDA(cudaFuncSetAttribute(k_power_argmax, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_power_argmax<<<1, 256, smem, st>>>(s->gram, (const double *)s->prm.bf_f64_dev, power_dev, doa_dev, p.C2, p.G, 1.0 / (double)m, 1, nullptr, nullptr);
        count_launch(2);
    }
    if (env_dev || doa_t_dev) {
        const int slab = 256;
        dim3 grid((unsigned)((p.G + 127) / 128), (unsigned)((m + slab - 1) / slab), 1);
        if (p.C2 <= 16) k_dense<16><<<grid, 128, 0, st>>>(s->vmem, (const float *)s->prm.bf_f32_dev, s->y, p.C2, p.G, m, slab);
        else k_dense<0><<<grid, 128, 0, st>>>(s->vmem, (const float *)s->prm.bf_f32_dev, s->y, p.C2, p.G, m, slab);
        float *env = env_dev ? env_dev : s->env;
        k_envelope<<<(p.G + 127) / 128, 128, 0, st>>>(s->y, env, s->env_state, s->env_started, m, p.G,
                                                      1.f / (float)(int)(s->fs * s->fall_time), 1.f / (float)(int)(s->fs * s->rise_time));
        k_envelope_mark<<<1, 1, 0, st>>>(s->env_started, m);
        count_launch(3);
        if (doa_t_dev) {
            k_argmax_rows<<<(unsigned)((m * 32 + 255) / 256), 256, 0, st>>>(env, doa_t_dev, m, p.G);
            count_launch(1);
        }
    }
    MICLOC_CUDA(cudaGetLastError());
    s->F = F_new;
    return MICLOC_OK;
}

extern "C" int micloc_snn_stream_push(micloc_stream *s, const void *frame_dev, int dtype, int64_t n, int32_t in_channels,
                                      int8_t *spikes_dev, float *power_dev, int32_t *doa_dev, float *env_dev,
                                      int32_t *doa_t_dev, int64_t *n_out, void *stream) {
    if (!s || !frame_dev) return set_error(MICLOC_ERR_SHAPE, "null argument");
    const ChainParams &p = s->prm.chain;
    if (s->flushed) return set_error(MICLOC_ERR_CONFIG, "stream was flushed: reset it before pushing again");
    if (n < 1 || n > s->max_frame) return set_error(MICLOC_ERR_SHAPE, "frame of %lld samples (stream takes 1..%lld)", (long long)n, s->max_frame);
    if (in_channels < p.M)
        return set_error(MICLOC_ERR_SHAPE, "number of channels in the input siganl %d should be the same as the number of microphones %d!", in_channels, p.M);
    if (s->N + n >= (1ll << 31) - 4096) return set_error(MICLOC_ERR_SHAPE, "stream position exceeds 2^31 samples: reset the stream");
    MICLOC_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned gi = (unsigned)((n * p.M + 255) / 256);
    if (dtype == MICLOC_F32) k_stream_ingest<float><<<gi, 256, 0, st>>>((const float *)frame_dev, s->xbuf, n, in_channels, p.M, s->hist);
    else if (dtype == MICLOC_I16) k_stream_ingest<int16_t><<<gi, 256, 0, st>>>((const int16_t *)frame_dev, s->xbuf, n, in_channels, p.M, s->hist);
    else if (dtype == MICLOC_I32) k_stream_ingest<int32_t><<<gi, 256, 0, st>>>((const int32_t *)frame_dev, s->xbuf, n, in_channels, p.M, s->hist);
    else return set_error(MICLOC_ERR_SHAPE, "dtype must be MICLOC_F32, MICLOC_I16 or MICLOC_I32");
    count_launch(1);
    // STHT of the extended clip [history | frame]: rows >= hist see their whole window
    MICLOC_TRY(launch_stht_any(p, (const float *)s->prm.taps_dev, s->xbuf, MICLOC_F32, s->q, 1, s->hist + n, st));
    k_stream_chain<<<(p.C2 + 31) / 32, 32, 0, st>>>(s->xbuf, s->q, s->state, s->ring, s->ring_len - 1, s->flags, p, s->N, n, s->hist, 0);
    const unsigned gs = (unsigned)(((long long)s->hist * p.M + 255) / 256);
    k_stream_shift<<<gs, 256, 0, st>>>(s->xbuf, s->tmp, n, p.M, s->hist, 0);
    k_stream_shift<<<gs, 256, 0, st>>>(s->xbuf, s->tmp, n, p.M, s->hist, 1);
    count_launch(3);
    MICLOC_CUDA(cudaGetLastError());
    s->N += n;
    const long long F_new = s->N - s->D > s->F ? s->N - s->D : s->F;
    return stream_finalise(s, F_new, spikes_dev, power_dev, doa_dev, env_dev, doa_t_dev, n_out, st);
}

extern "C" int micloc_snn_stream_flush(micloc_stream *s, int8_t *spikes_dev, float *power_dev, int32_t *doa_dev,
                                       float *env_dev, int32_t *doa_t_dev, int64_t *n_out, int32_t *flags_host,
                                       void *stream) {
    if (!s) return set_error(MICLOC_ERR_CONFIG, "null stream");
    const ChainParams &p = s->prm.chain;
    MICLOC_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (!s->flushed) {
        k_stream_chain<<<(p.C2 + 31) / 32, 32, 0, st>>>(s->xbuf, s->q, s->state, s->ring, s->ring_len - 1, s->flags, p, s->N, 0, s->hist, 1);
        count_launch(1);
        s->flushed = true;
    }
    MICLOC_TRY(stream_finalise(s, s->N, spikes_dev, power_dev, doa_dev, env_dev, doa_t_dev, n_out, st));
    if (flags_host) {
        MICLOC_CUDA(cudaMemcpyAsync(flags_host, s->flags, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        MICLOC_CUDA(cudaStreamSynchronize(st));
    }
    return MICLOC_OK;
}

// Envelope.evolve on a whole device array [T][C] (fresh state) + optional per-row argmax
extern "C" int micloc_envelope(const float *x_dev, int64_t T, int32_t C, double fs, double rise_time, double fall_time,
                               float *env_dev, int32_t *argmax_dev, int device, void *stream) {
    if (!x_dev || !env_dev || T < 1 || C < 1) return set_error(MICLOC_ERR_SHAPE, "bad envelope arguments");
    if (rise_time > fall_time)
        return set_error(MICLOC_ERR_CONFIG, "for proper functioning, an envelope estimator should have a larger fall time!");
    const int wf = (int)(fs * fall_time), wr = (int)(fs * rise_time);
    if (wf < 1 || wr < 1) return set_error(MICLOC_ERR_CONFIG, "envelope windows shorter than one sample");
    MICLOC_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    float *state = nullptr; int32_t *started = nullptr;
    MICLOC_CUDA(cudaMallocAsync(&state, (size_t)C * sizeof(float), st));
    MICLOC_CUDA(cudaMallocAsync(&started, sizeof(int32_t), st));
    MICLOC_CUDA(cudaMemsetAsync(state, 0, (size_t)C * sizeof(float), st));
    MICLOC_CUDA(cudaMemsetAsync(started, 0, sizeof(int32_t), st));
    k_envelope<<<(C + 127) / 128, 128, 0, st>>>(x_dev, env_dev, state, started, T, C, 1.f / (float)wf, 1.f / (float)wr);
    count_launch(1);
    if (argmax_dev) {
        k_argmax_rows<<<(unsigned)((T * 32 + 255) / 256), 256, 0, st>>>(env_dev, argmax_dev, T, C);
        count_launch(1);
    }
    MICLOC_CUDA(cudaGetLastError());
    MICLOC_CUDA(cudaFreeAsync(state, st));
    MICLOC_CUDA(cudaFreeAsync(started, st));
    return MICLOC_OK;
}
