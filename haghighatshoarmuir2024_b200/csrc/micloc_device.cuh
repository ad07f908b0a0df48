// micloc_device.cuh -- device-side building blocks shared by the staged and the
// fused kernels of the SNN-localisation hot path (sm_100a).
//
// Reference sites restated (paths under /root/reference):
//   STHT FIR                    micloc/snn_beamformer.py:325-327
//   band-pass (as SOS cascade)  micloc/snn_beamformer.py:330-331
//   RZCC                        micloc/spike_encoder.py:115-137 (+ scipy find_peaks)
//   neuron alpha kernel         micloc/snn_beamformer.py:342-364
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace micloc {

constexpr int kMaxSections = 4;   // biquads per band-pass
constexpr int kClusterMax = 8;    // RZCC candidates buffered per open cluster
constexpr int kFirR = 16;         // consecutive outputs per thread in the FIR
constexpr int kFirJB = 8;         // taps per register block

// Constants of one SNN chain; lives in __grid_constant__ kernel parameters.
struct ChainParams {
    int M;            // microphones
    int C2;           // 2*M real channels (in-phase | quadrature)
    int K;            // STHT kernel length
    int half;         // K/2: circular shift of the in-phase branch
    int tap_stride;   // 1 (dense) or 2 (Hilbert: every other tap is zero)
    int tap_first;    // index k0 of the first kept tap
    int n_taps;       // kept taps, padded with zeros to a multiple of kFirJB
    int span;         // largest look-back = tap_first + tap_stride*(n_taps-1)
    int nsec;         // biquad sections
    float sos[kMaxSections][5];  // b0 b1 b2 a1 a2
    int w;            // RZCC distance (>=1)
    int bipolar;
    float na;         // neuron decay a
    float nc;         // neuron scale c
    float ncT;        // c * a^L
    float nLf;        // (float)L
    int nL;           // neuron kernel length L
    int G;
};

// ---------------------------------------------------------------------------
// input sample fetch (float32 or int16 PCM; int16 is used at face value)
// ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<int16_t>(int16_t v) { return (float)v; }

// padded index inside one mic row of the FIR tile: 4 extra floats per 32 so
// that lanes reading float4 at a 16-float pitch hit distinct bank groups.
__host__ __device__ __forceinline__ int fir_pad(int l) { return l + ((l >> 5) << 2); }

// ---------------------------------------------------------------------------
// STHT FIR building blocks (used by k_stht and by the fused kernel)
// ---------------------------------------------------------------------------
template <int STRIDE> struct FirGeom {
    static constexpr int WIN = kFirR + STRIDE * (kFirJB - 1);
    static constexpr int NV = (WIN + 3) / 4;
};

__host__ __device__ inline int fir_row_pitch(int TT, int span) {
    const int lrow = TT + span + 8;
    return (lrow + ((lrow >> 5) << 2) + 4 + 3) & ~3;
}

template <int STRIDE>
__device__ __forceinline__ void fir_accumulate(const float *__restrict__ row, const float *__restrict__ taps_s,
                                               int n_taps, int span, int k0, int chunk,
                                               float (&acc)[kFirR]) {
    constexpr int WIN = FirGeom<STRIDE>::WIN;
    constexpr int NV = FirGeom<STRIDE>::NV;
#pragma unroll
    for (int i = 0; i < kFirR; ++i) acc[i] = 0.f;
    const int nblk = n_taps / kFirJB;
    // l index of the window start for tap block 0; it moves back by STRIDE*JB per block
    int lmin = span + kFirR * chunk - k0 - STRIDE * (kFirJB - 1);
#pragma unroll 1
    for (int jb = 0; jb < nblk; ++jb, lmin -= STRIDE * kFirJB) {
        float g[kFirJB];
        {
            const float4 g0 = *reinterpret_cast<const float4 *>(taps_s + jb * kFirJB);
            const float4 g1 = *reinterpret_cast<const float4 *>(taps_s + jb * kFirJB + 4);
            g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w;
            g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        }
        float wv[NV * 4];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float4 x4 = *reinterpret_cast<const float4 *>(row + fir_pad(lmin + 4 * v));
            wv[4 * v + 0] = x4.x; wv[4 * v + 1] = x4.y; wv[4 * v + 2] = x4.z; wv[4 * v + 3] = x4.w;
        }
#pragma unroll
        for (int jj = 0; jj < kFirJB; ++jj)
#pragma unroll
            for (int i = 0; i < kFirR; ++i)
                acc[i] = fmaf(g[jj], wv[i + STRIDE * (kFirJB - 1 - jj)], acc[i]);
        (void)WIN;
    }
}

// fill rows[mm][fir_pad(l)] = x[b][t0 - span + l][m0 + mm] for l < lrow (zero outside the clip)
template <typename IN_T>
__device__ __forceinline__ void fir_fill_rows(float *rows, int pitch, const IN_T *__restrict__ clip,
                                              long long T, int M, int m0, int mg, long long t0,
                                              int span, int lrow) {
    for (int l = threadIdx.x; l < lrow; l += blockDim.x) {
        const long long t = t0 - span + l;
        const bool ok = (t >= 0) && (t < T);
        const IN_T *src = clip + (ok ? t : 0) * M + m0;
        const int pl = fir_pad(l);
        for (int mm = 0; mm < mg; ++mm) rows[mm * pitch + pl] = ok ? to_f32<IN_T>(src[mm]) : 0.f;
    }
}

// ---------------------------------------------------------------------------
// biquad cascade, direct form II transposed, one sample
// ---------------------------------------------------------------------------
struct BiquadState { float s1[kMaxSections]; float s2[kMaxSections]; };

__device__ __forceinline__ void biquad_reset(BiquadState &st) {
#pragma unroll
    for (int k = 0; k < kMaxSections; ++k) { st.s1[k] = 0.f; st.s2[k] = 0.f; }
}

__device__ __forceinline__ float biquad_step(const float (&sos)[kMaxSections][5], int nsec,
                                             BiquadState &st, float x) {
#pragma unroll
    for (int k = 0; k < kMaxSections; ++k) {
        if (k < nsec) {
            const float y = fmaf(sos[k][0], x, st.s1[k]);
            st.s1[k] = fmaf(sos[k][1], x, fmaf(-sos[k][3], y, st.s2[k]));
            st.s2[k] = fmaf(sos[k][2], x, -sos[k][4] * y);
            x = y;
        }
    }
    return x;
}

// ---------------------------------------------------------------------------
// RZCC: streaming find_peaks(cumsum(z), distance=w) for one channel.
//
// Candidates: a peak of the cumulative sum sits where z goes + -> (zeros) -> -,
// at the midpoint of the flat top (scipy _local_maxima_1d); first and last
// sample are never peaks.  Candidates closer than w samples form a cluster;
// a cluster is closed once w samples pass without a new candidate and is then
// resolved by scipy's greedy rule (_select_by_peak_distance: highest cumsum
// first, it removes every candidate nearer than w; ties -> later position).
// Valleys are the same on -cumsum, independently (spike_encoder.py:131-135).
// A cluster larger than kClusterMax sets `overflow` (the caller reruns the clip
// through the unbounded encoder, micloc_rzcc_encode_f64 semantics).
// ---------------------------------------------------------------------------
struct RzccCluster {
    int n;
    int last;                 // position of the newest candidate
    int pos[kClusterMax];
    double hgt[kClusterMax];
};

struct RzccState {
    double csum;              // running cumulative sum (float64 like np.cumsum)
    int rise;                 // index of the active '+' sample, -1 if none
    int fall;                 // index of the active '-' sample, -1 if none
    double rise_h, fall_h;    // cumsum at rise / -cumsum at fall
    int overflow;
    RzccCluster pk, vl;
};

__device__ __forceinline__ void rzcc_reset(RzccState &s) {
    s.csum = 0.0; s.rise = -1; s.fall = -1; s.rise_h = 0.0; s.fall_h = 0.0; s.overflow = 0;
    s.pk.n = 0; s.pk.last = 0; s.vl.n = 0; s.vl.last = 0;
}

template <typename Emit>
__device__ __forceinline__ void rzcc_resolve(RzccCluster &cl, int w, int sign, Emit &&emit) {
    if (cl.n == 1) { emit(cl.pos[0], sign); cl.n = 0; return; }
    unsigned und = (1u << cl.n) - 1u;
    while (und) {
        int best = -1;
        for (int i = 0; i < cl.n; ++i)
            if (((und >> i) & 1u) && (best < 0 || cl.hgt[i] >= cl.hgt[best])) best = i;
        emit(cl.pos[best], sign);
        for (int i = 0; i < cl.n; ++i) {
            int d = cl.pos[i] - cl.pos[best];
            d = d < 0 ? -d : d;
            if (d < w) und &= ~(1u << i);
        }
    }
    cl.n = 0;
}

template <typename Emit>
__device__ __forceinline__ void rzcc_push(RzccState &s, RzccCluster &cl, int w, int sign, int pos,
                                          double h, Emit &&emit) {
    if (cl.n > 0 && pos - cl.last >= w) rzcc_resolve(cl, w, sign, emit);
    if (cl.n == kClusterMax) { s.overflow = 1; return; }
    cl.pos[cl.n] = pos; cl.hgt[cl.n] = h; cl.last = pos; ++cl.n;
}

// feed sample z at time t; emit(pos, sign) is called for every final spike,
// always with pos < t and pos > t - kClusterMax*w - 1.
template <typename Emit>
__device__ __forceinline__ void rzcc_step(RzccState &s, int w, int bipolar, int t, float z,
                                          Emit &&emit) {
    s.csum += (double)z;
    if (z > 0.f) {
        if (bipolar && s.fall >= 0) {
            rzcc_push(s, s.vl, w, -1, (s.fall + t - 1) >> 1, s.fall_h, emit);
            s.fall = -1;
        }
        if (t >= 1) { s.rise = t; s.rise_h = s.csum; }
    } else if (z < 0.f) {
        if (s.rise >= 0) {
            rzcc_push(s, s.pk, w, +1, (s.rise + t - 1) >> 1, s.rise_h, emit);
            s.rise = -1;
        }
        if (t >= 1) { s.fall = t; s.fall_h = -s.csum; }
    }
    if (s.pk.n > 0 && t - s.pk.last >= w) rzcc_resolve(s.pk, w, +1, emit);
    if (bipolar && s.vl.n > 0 && t - s.vl.last >= w) rzcc_resolve(s.vl, w, -1, emit);
}

template <typename Emit>
__device__ __forceinline__ void rzcc_finish(RzccState &s, int w, int bipolar, Emit &&emit) {
    if (s.pk.n > 0) rzcc_resolve(s.pk, w, +1, emit);
    if (bipolar && s.vl.n > 0) rzcc_resolve(s.vl, w, -1, emit);
}

// ---------------------------------------------------------------------------
// neuron (synapse + membrane) alpha kernel h[n] = c*n*a^n, n < L, as two
// two-state recurrences: the full alpha response minus its tail, the tail being
// the same response driven by the spike train delayed by L.
// ---------------------------------------------------------------------------
struct NeuronState { float p1, p2, q1, q2; };

__device__ __forceinline__ void neuron_reset(NeuronState &n) { n.p1 = n.p2 = n.q1 = n.q2 = 0.f; }

// s = spike at t, sd = spike at t-L (0 before the clip starts)
__device__ __forceinline__ float neuron_step(const ChainParams &p, NeuronState &n, float s, float sd) {
    n.p2 = p.na * (n.p2 + n.p1);
    n.p1 = fmaf(p.na, n.p1, s);
    n.q2 = p.na * (n.q2 + n.q1);
    n.q1 = fmaf(p.na, n.q1, sd);
    const float tail = fmaf(p.nLf, n.q1, n.q2);
    return fmaf(-p.ncT, tail, p.nc * n.p2);
}

}  // namespace micloc
