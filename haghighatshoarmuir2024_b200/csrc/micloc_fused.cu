// micloc_fused.cu -- the fused hot-path kernel: raw audio in, spikes + per-DoA
// power + DoA index out; nothing else touches HBM.
//
//   audio tile --(smem, mic-major, padded)--> STHT FIR (FP32 FFMA, register window)
//     --> [in-phase | quadrature] --> SOS band-pass --> RZCC (cluster NMS)
//     --> spike ring (smem) --> alpha-kernel neuron recurrences --> vmem tile (smem)
//     --> Gram accumulation  C += v v^T   (registers, float64 across tiles)
//   clip end:  power[g] = w_g^T C w_g / T (float64), DoA = first argmax.
//
// Reference sites: micloc/snn_beamformer.py:283-370 and the callers' power/argmax
// paper_plots/target_snn_localization.py:462-464.
//
// One CTA owns one clip at a time (persistent grid-stride over clips) and walks
// it in time tiles; filter/RZCC/neuron state is carried in registers by one
// thread per channel, spikes of the last RING steps live in a shared-memory
// ring so the neuron stage can run D = kClusterMax*w steps behind the encoder
// (the latency of the exact find_peaks(distance=w) decision).
#include <cuda_runtime.h>

#include "micloc_common.h"

namespace micloc {

struct FusedGeom {
    int TT;        // time tile
    int pitch_x;   // floats per mic row of the input tile
    int pitch_q;   // floats per mic row of the quadrature tile
    int CP;        // padded channel count of the vmem tile (multiple of 4)
    int ring;      // spike ring length (power of two)
    int D;         // neuron-stage lag behind the encoder
    int NP;        // 4x4 Gram blocks (upper triangle)
    int NS;        // time slices per Gram block
    int off_x, off_q, off_vm, off_ring;  // smem offsets in floats
    int smem_bytes;
};

template <typename IN_T, int STRIDE, int NB>
__global__ void __launch_bounds__(128, 4)
k_fused(const IN_T *__restrict__ audio, const float *__restrict__ taps, const double *__restrict__ Wd,
        int8_t *__restrict__ spikes, float *__restrict__ power, int32_t *__restrict__ doa,
        int32_t *__restrict__ flags, const __grid_constant__ ChainParams p,
        const __grid_constant__ FusedGeom g, long long B, long long T) {
    extern __shared__ __align__(16) float smem[];
    float *taps_s = smem;
    float *xs = smem + g.off_x;
    float *qs = smem + g.off_q;
    float *vm = smem + g.off_vm;
    int8_t *ring = reinterpret_cast<int8_t *>(smem + g.off_ring);
    __shared__ double red_v[128];
    __shared__ int red_i[128];

    const int tid = threadIdx.x;
    const int C2 = p.C2, M = p.M, TT = g.TT, CP = g.CP;
    const int rmask = g.ring - 1;
    const bool chain_thread = tid < C2;
    const bool inphase = tid < M;

    // Gram role: (block pair, time slice)
    const bool gram_thread = tid < g.NP * g.NS;
    int bi = 0, bj = 0;
    const int gslice = tid / g.NP;
    {
        int pr = tid % g.NP;  // enumerate upper-triangular (bi <= bj) block pairs
        for (int r = 0; r < NB; ++r) {
            const int len = NB - r;
            if (pr < len) { bi = r; bj = r + pr; break; }
            pr -= len;
        }
    }

    for (int i = tid; i < p.n_taps; i += blockDim.x) taps_s[i] = taps[i];

    for (long long b = blockIdx.x; b < B; b += gridDim.x) {
        const IN_T *clip = audio + b * T * M;
        int8_t *spk_out = spikes ? spikes + b * T * C2 : nullptr;

        // ---- per-clip state ----
        BiquadState bq; biquad_reset(bq);
        RzccState rz; rzcc_reset(rz);
        NeuronState nr; neuron_reset(nr);
        double acc64[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) acc64[k] = 0.0;
        for (int i = tid; i < g.ring * C2; i += blockDim.x) ring[i] = 0;
        for (int i = tid; i < TT * CP; i += blockDim.x) vm[i] = 0.f;
        long long src = ((-(long long)p.half) % T + T) % T;  // wrapped source index of the in-phase branch
        __syncthreads();

        auto emit = [&](int pos, int sign) { ring[(pos & rmask) * C2 + tid] = (int8_t)sign; };

        for (long long t0 = 0; t0 < T + g.D; t0 += TT) {
            // ---- phase 1+2: tile fill and STHT FIR ----
            if (t0 < T) {
                fir_fill_rows<IN_T>(xs, g.pitch_x, clip, T, M, 0, M, t0, p.span, TT + p.span + 8);
                __syncthreads();
                const int chunks = TT / kFirR;
                for (int item = tid; item < M * chunks; item += blockDim.x) {
                    const int mm = item / chunks, chunk = item % chunks;
                    float acc[kFirR];
                    fir_accumulate<STRIDE>(xs + mm * g.pitch_x, taps_s, p.n_taps, p.span, p.tap_first, chunk, acc);
                    float *dst = qs + mm * g.pitch_q + fir_pad(chunk * kFirR);
#pragma unroll
                    for (int v = 0; v < kFirR / 4; ++v)
                        *reinterpret_cast<float4 *>(dst + 4 * v) =
                            make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
                }
            }
            __syncthreads();

            // ---- phase 3: per-channel sequential chain ----
            if (chain_thread) {
                const float *xrow = inphase ? xs + tid * g.pitch_x : qs + (tid - M) * g.pitch_q;
                for (int i = 0; i < TT; ++i) {
                    const long long t = t0 + i;
                    if (t < T) {
                        float x;
                        if (inphase) {
                            if (t < p.half || p.span < p.half) x = to_f32<IN_T>(clip[src * M + tid]);
                            else x = xrow[fir_pad(p.span + i - p.half)];
                            if (++src == T) src = 0;
                        } else {
                            x = xrow[fir_pad(i)];
                        }
                        const float z = biquad_step(p.sos, p.nsec, bq, x);
                        ring[((int)t & rmask) * C2 + tid] = 0;
                        rzcc_step(rz, p.w, p.bipolar, (int)t, z, emit);
                    } else if (t == T) {
                        rzcc_finish(rz, p.w, p.bipolar, emit);
                    }
                    const long long u = t - g.D;
                    float v = 0.f;
                    if (u >= 0 && u < T) {
                        const float s = (float)ring[((int)u & rmask) * C2 + tid];
                        const float sd = u >= p.nL ? (float)ring[((int)(u - p.nL) & rmask) * C2 + tid] : 0.f;
                        v = neuron_step(p, nr, s, sd);
                    }
                    vm[i * CP + tid] = v;
                }
            }
            __syncthreads();

            // ---- phase 4: Gram accumulation over the tile + spike write-out ----
            if (gram_thread) {
                float acc[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[k] = 0.f;
                for (int i = gslice; i < TT; i += g.NS) {
                    const float4 a = *reinterpret_cast<const float4 *>(vm + i * CP + 4 * bi);
                    const float4 c = *reinterpret_cast<const float4 *>(vm + i * CP + 4 * bj);
                    const float av[4] = {a.x, a.y, a.z, a.w};
                    const float cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int s = 0; s < 4; ++s) acc[4 * r + s] = fmaf(av[r], cv[s], acc[4 * r + s]);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) acc64[k] += (double)acc[k];
            }
            if (spk_out) {
                const long long u0 = t0 - g.D;
                for (int r = tid; r < TT; r += blockDim.x) {
                    const long long u = u0 + r;
                    if (u >= 0 && u < T) {
                        const int8_t *srow = ring + ((int)u & rmask) * C2;
                        int8_t *drow = spk_out + u * C2;
                        for (int c = 0; c < C2; ++c) drow[c] = srow[c];
                    }
                }
            }
            // the next tile's fill/FIR touch xs/qs only; vm and ring are rewritten after two barriers
        }
        __syncthreads();

        // ---- clip epilogue: reduce Gram partials (fixed order), power, argmax ----
        double *part = reinterpret_cast<double *>(xs);            // [NP*NS][16]
        double *Cd = part + g.NP * g.NS * 16;                     // [CP][CP]
        if (gram_thread) {
#pragma unroll
            for (int k = 0; k < 16; ++k) part[tid * 16 + k] = acc64[k];
        }
        __syncthreads();
        for (int e = tid; e < g.NP * 16; e += blockDim.x) {
            const int pr = e / 16, k = e % 16;
            double s = 0.0;
            for (int sl = 0; sl < g.NS; ++sl) s += part[(sl * g.NP + pr) * 16 + k];
            int r = 0, rem = pr, pbi = 0, pbj = 0;
            for (r = 0; r < NB; ++r) {
                const int len = NB - r;
                if (rem < len) { pbi = r; pbj = r + rem; break; }
                rem -= len;
            }
            const int row = 4 * pbi + k / 4, col = 4 * pbj + k % 4;
            Cd[row * CP + col] = s;
            if (pbi != pbj) Cd[col * CP + row] = s;
        }
        __syncthreads();
        double best = -1.0; int besti = 0x7fffffff;
        const double inv_T = 1.0 / (double)T;
        for (int gg = tid; gg < p.G; gg += blockDim.x) {
            double w[4 * NB];
#pragma unroll
            for (int c = 0; c < 4 * NB; ++c) w[c] = c < C2 ? Wd[(long long)c * p.G + gg] : 0.0;
            double accp = 0.0;
#pragma unroll
            for (int r = 0; r < 4 * NB; ++r) {
                double rr = 0.0;
#pragma unroll
                for (int c = 0; c < 4 * NB; ++c) rr = fma(Cd[r * CP + c], w[c], rr);
                accp = fma(w[r], rr, accp);
            }
            accp *= inv_T;
            if (power) power[b * p.G + gg] = (float)accp;
            if (accp > best) { best = accp; besti = gg; }
        }
        red_v[tid] = best; red_i[tid] = besti;
        __syncthreads();
        for (int s = blockDim.x / 2; s > 0; s >>= 1) {
            if (tid < s) {
                const double ov = red_v[tid + s]; const int oi = red_i[tid + s];
                if (ov > red_v[tid] || (ov == red_v[tid] && oi < red_i[tid])) { red_v[tid] = ov; red_i[tid] = oi; }
            }
            __syncthreads();
        }
        if (tid == 0 && doa) doa[b] = red_i[0];
        if (chain_thread && rz.overflow && flags) atomicOr(flags + b, 1);
        __syncthreads();
    }
}

static int next_pow2(int v) { int r = 1; while (r < v) r <<= 1; return r; }

template <typename IN_T, int STRIDE, int NB>
static int launch_fused_t(const ChainParams &p, const FusedGeom &g, const float *d_taps, const double *d_Wd,
                          const IN_T *audio, long long B, long long T, int8_t *spikes, float *power, int32_t *doa,
                          int32_t *flags, int sm_count, cudaStream_t st) {
    auto kern = k_fused<IN_T, STRIDE, NB>;
    MICLOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
    int per_sm = 1;
    MICLOC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, g.smem_bytes));
    if (per_sm < 1) return set_error(MICLOC_ERR_UNSUPPORTED, "fused kernel does not fit (smem %d B)", g.smem_bytes);
    long long grid = (long long)sm_count * per_sm;
    if (grid > B) grid = B;
    kern<<<(unsigned)grid, 128, g.smem_bytes, st>>>(audio, d_taps, d_Wd, spikes, power, doa, flags, p, g, B, T);
    count_launch(1);
    MICLOC_CUDA(cudaGetLastError());
    return MICLOC_OK;
}

int launch_fused(const ChainParams &p, const float *d_taps, const double *d_Wd, const void *audio, int dtype,
                 long long B, long long T, int8_t *spikes, float *power, int32_t *doa, int32_t *flags,
                 int sm_count, cudaStream_t st) {
    if (p.C2 > 32)
        return set_error(MICLOC_ERR_UNSUPPORTED, "fused kernel supports up to 16 microphones (got %d); use the staged path", p.M);
    FusedGeom g{};
    g.TT = 256;
    g.D = ((kClusterMax * p.w + 7) / 8) * 8;
    g.ring = next_pow2(g.TT + g.D + p.nL + 8);
    g.pitch_x = fir_row_pitch(g.TT, p.span);
    g.pitch_q = (fir_pad(g.TT) + 4 + 3) & ~3;
    const int NB = p.C2 <= 16 ? 4 : 8;
    g.CP = 4 * NB;
    g.NP = NB * (NB + 1) / 2;
    g.NS = 128 / g.NP;
    int off = (p.n_taps + 3) & ~3;
    g.off_x = off;
    int x_floats = p.M * g.pitch_x;
    const int epi_floats = (g.NP * g.NS * 16 + g.CP * g.CP) * 2;  // doubles reuse the xs region at clip end
    if (x_floats < epi_floats) x_floats = epi_floats;
    off += (x_floats + 3) & ~3;
    g.off_q = off; off += p.M * g.pitch_q;
    g.off_vm = off; off += g.TT * g.CP;
    g.off_ring = off; off += (g.ring * p.C2 + 3) / 4;
    g.smem_bytes = off * (int)sizeof(float);
    if (g.smem_bytes > 227 * 1024)
        return set_error(MICLOC_ERR_UNSUPPORTED, "fused kernel needs %d B of shared memory; use the staged path", g.smem_bytes);
    if (T + g.D >= (1ll << 31)) return set_error(MICLOC_ERR_SHAPE, "T too large for the fused kernel");

#define MICLOC_FUSED_CASE(IN, S, N)                                                                          \
    return launch_fused_t<IN, S, N>(p, g, d_taps, d_Wd, (const IN *)audio, B, T, spikes, power, doa, flags, \
                                    sm_count, st)
    const bool i16 = dtype == MICLOC_I16;
    if (p.tap_stride == 2) {
        if (NB == 4) { if (i16) MICLOC_FUSED_CASE(int16_t, 2, 4); else MICLOC_FUSED_CASE(float, 2, 4); }
        else         { if (i16) MICLOC_FUSED_CASE(int16_t, 2, 8); else MICLOC_FUSED_CASE(float, 2, 8); }
    } else {
        if (NB == 4) { if (i16) MICLOC_FUSED_CASE(int16_t, 1, 4); else MICLOC_FUSED_CASE(float, 1, 4); }
        else         { if (i16) MICLOC_FUSED_CASE(int16_t, 1, 8); else MICLOC_FUSED_CASE(float, 1, 8); }
    }
#undef MICLOC_FUSED_CASE
}

}  // namespace micloc
