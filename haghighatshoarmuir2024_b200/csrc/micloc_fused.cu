// micloc_fused.cu -- the fused hot-path kernel: raw audio in, spikes + per-DoA
// power + DoA index out; nothing else touches HBM.
//
// Reference sites: micloc/snn_beamformer.py:283-370 and the callers' power/argmax
// paper_plots/target_snn_localization.py:462-464.
//
// One persistent CTA of four warps owns kSlots = 2 clips at a time and walks them
// in time tiles of kTile = 64 samples.  The warps are specialised and run as a
// software pipeline, one barrier per tile (iteration k):
//
//   back warp   tile k+1   audio (HBM) -> mic-major ring in shared memory
//   FIR warps   tile k     STHT quadrature FIR, one warp per clip: every lane owns 16
//                          consecutive outputs of one microphone and walks the 240
//                          non-zero Hilbert taps in blocks of 8 with a sliding register
//                          window; the multiply-adds are packed FFMA2 (fma.rn.f32x2)
//   front warp  tile k-1   one lane per (clip, channel): SOS band-pass recurrence, running
//                          sum, sign / zero bit masks; every 32 samples the masks are
//                          turned into RZCC candidates and resolved (find_peaks distance
//                          rule) into a bit-packed spike ring
//   back warp   tile k-4   (the latency of the exact find_peaks decision) one lane per
//                          (clip, channel): alpha-kernel neuron recurrences driven by the
//                          final spike bits -> membrane tile; then Gram accumulation
//                          C += v v^T of that tile with FFMA2, and the int8 spike raster
//                          of the tile -> HBM
//   clip end               power[g] = w_g^T C w_g / T (float64), DoA = first argmax.
#include <cuda_runtime.h>

#include "micloc_common.h"

namespace micloc {

constexpr int kTile = 64;      // samples per pipeline step
constexpr int kSlots = 2;      // clips per CTA
constexpr int kRows = 8;       // most microphones per clip the lane maps cover
constexpr int kQPitch = kTile + 4;
constexpr int kVmPitch = 32;   // floats per time step in the membrane tile: [slot][16]
constexpr int kRingWords = 16; // spike-bit ring: 16 words of 32 samples per channel and polarity

struct FusedGeom {
    int ring_x;      // audio ring length in samples (multiple of 32)
    int pitch_x;     // floats per ring row; pitch_x / 4 is odd (conflict-free LDS.128 across microphones)
    int shift;       // ring coordinate of sample t is (t + shift) mod ring_x
    int nblk;        // FIR tap blocks of 8 (multiple of 3)
    int dtile;       // the back warp runs dtile tiles behind the pipeline step (RZCC decision latency)
    int tiles_is;    // tiles whose in-phase input comes from the clip tail (t < K/2)
    int off_x, off_q, off_vm, off_is, off_cs, off_clus, off_bits, off_stage;   // byte offsets in dynamic smem
    int smem_bytes;
};

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// acc.xy += w.xy * g   (one FFMA2)
__device__ __forceinline__ void ffma2(unsigned long long &acc, unsigned long long w, unsigned long long g2) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(w), "l"(g2));
}

struct Chunk { unsigned long long p[8]; };   // 16 consecutive samples as 8 float pairs

__device__ __forceinline__ void load_chunk(Chunk &c, const float *row, int coord) {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(row + coord);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const ulonglong2 u = src[v];
        c.p[2 * v] = u.x; c.p[2 * v + 1] = u.y;
    }
}

// 8 taps x 16 outputs: acc[ip] += g[jj] * W[ip + 7 - jj], W = lo pairs 0..7 | hi pairs 0..6
__device__ __forceinline__ void fir_block(unsigned long long (&acc)[8], const Chunk &lo, const Chunk &hi,
                                          const float *__restrict__ taps8) {
    const float4 g0 = *reinterpret_cast<const float4 *>(taps8);
    const float4 g1 = *reinterpret_cast<const float4 *>(taps8 + 4);
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const unsigned long long g2 = pack2(g[jj], g[jj]);
#pragma unroll
        for (int ip = 0; ip < 8; ++ip) {
            const int idx = ip + 7 - jj;
            ffma2(acc[ip], idx < 8 ? lo.p[idx] : hi.p[idx - 8], g2);
        }
    }
}

// All four warps meet here once per pipeline step (the roles run different code).
__device__ __forceinline__ void tile_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

struct FusedSmem {
    float *taps, *xs, *qs, *vms, *is_s, *cs;
    int *clus;
    unsigned int *bits;     // [2 polarities][kRingWords][32 lanes]
    int8_t *stage;          // [kSlots][kTile][C2]
    double *gram;
};

// ============================== STHT FIR warp (one per clip slot) ==============================
template <int MM>
__device__ __forceinline__ void fir_role(const FusedSmem &sm, const ChainParams &p, const FusedGeom &g, int slot,
                                         bool clip_ok, int lane, int NT, int k_last) {
    const int M = MM ? MM : p.M;
    const int f_chunk = lane >> 3, f_mic = lane & 7;     // lane = chunk * 8 + mic
    const bool work = clip_ok && f_mic < M;
    const float *row = sm.xs + (slot * M + f_mic) * g.pitch_x;
    for (int k = -1; k <= k_last; ++k) {
        if (work && k >= 0 && k < NT) {
            unsigned long long acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0ull;
            // ring coordinate of the window of tap block 0 (a multiple of 16 by the choice of shift)
            const int c0 = (k * kTile + 16 * f_chunk - p.tap_first - 14 + g.shift + 2 * g.ring_x) % g.ring_x;
            Chunk A, Bq, Cq;
            { int ch = c0 + 16; if (ch >= g.ring_x) ch -= g.ring_x; load_chunk(Bq, row, ch); }
            load_chunk(A, row, c0);
            int cn = c0 - 16; if (cn < 0) cn += g.ring_x;
            const float *tp = sm.taps;
#pragma unroll 1
            for (int jb = 0; jb < g.nblk; jb += 3) {
                load_chunk(Cq, row, cn); cn -= 16; if (cn < 0) cn += g.ring_x;
                fir_block(acc, A, Bq, tp);
                load_chunk(Bq, row, cn); cn -= 16; if (cn < 0) cn += g.ring_x;
                fir_block(acc, Cq, A, tp + 8);
                load_chunk(A, row, cn); cn -= 16; if (cn < 0) cn += g.ring_x;
                fir_block(acc, Bq, Cq, tp + 16);
                tp += 24;
            }
            float *dst = sm.qs + (((k & 1) * kSlots + slot) * M + f_mic) * kQPitch + 16 * f_chunk;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float4 o;
                unpack2(acc[2 * v], o.x, o.y);
                unpack2(acc[2 * v + 1], o.z, o.w);
                reinterpret_cast<float4 *>(dst)[v] = o;
            }
        }
        tile_barrier();
    }
}

// ============ front warp: band-pass + RZCC -> spike bits, lane = slot*16 + channel ============
// two biquads, direct form II transposed, coefficients in registers
struct Sos2 { float b0[2], b1[2], b2[2], a1[2], a2[2]; };
__device__ __forceinline__ float biquad2_step(const Sos2 &c, BiquadState &st, float x) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float y = fmaf(c.b0[k], x, st.s1[k]);
        st.s1[k] = fmaf(c.b1[k], x, fmaf(-c.a1[k], y, st.s2[k]));
        st.s2[k] = fmaf(c.b2[k], x, -c.a2[k] * y);
        x = y;
    }
    return x;
}

template <typename IN_T, int MM>
__device__ __forceinline__ void front_role(const FusedSmem &sm, const ChainParams &p, const FusedGeom &g,
                                           const IN_T *__restrict__ audio, int32_t *__restrict__ flags,
                                           long long clip0, long long B, long long T64, int lane, int k_last) {
    const int M = MM ? MM : p.M, C2 = 2 * M;
    const int T = (int)T64;
    const int c_slot = lane >> 4, c_ch = lane & 15;
    const bool slot_ok = clip0 + c_slot < B;
    const bool c_valid = c_ch < C2 && slot_ok;
    const bool c_inphase = c_ch < M;
    const int w = p.w, bipolar = p.bipolar;
    const RzccStore store{sm.clus + lane, reinterpret_cast<float *>(sm.clus + 2 * kClusterMax * 32) + lane, 32};
    unsigned int *bits = sm.bits + lane;
    // a final spike: set its bit in this channel's ring word (only this lane ever writes these words)
    auto emit = [&](int pos, int sign) {
        unsigned int *wd = bits + ((sign > 0 ? kRingWords : 0) + ((pos >> 5) & (kRingWords - 1))) * 32;
        *wd |= 1u << (pos & 31);
    };
    const IN_T *clip_audio = audio + (slot_ok ? clip0 + c_slot : clip0) * T64 * M;
    float *cs = sm.cs + lane;
    Sos2 sos;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        sos.b0[k] = p.sos[k][0]; sos.b1[k] = p.sos[k][1]; sos.b2[k] = p.sos[k][2];
        sos.a1[k] = p.sos[k][3]; sos.a2[k] = p.sos[k][4];
    }
    BiquadState bq; biquad_reset(bq);
    RzccState rz; rzcc_reset(rz);

    for (int k = -1; k <= k_last; ++k) {
        const int kc = k - 1;
        const int t0 = kc * kTile;
        if (kc >= 0 && t0 < T) {
            const bool from_is = kc < g.tiles_is;
            if (from_is) {
                // in-phase input of the first K/2 samples is the clip's tail (np.roll, snn_beamformer.py:325)
                if (slot_ok) {
                    float *dsti = sm.is_s + c_slot * kTile * M;
                    for (int e = c_ch; e < kTile * M; e += 16) {
                        const int t = t0 + e / M;
                        float v = 0.f;
                        if (t < T) {
                            int src = (t - p.half) % T;
                            if (src < 0) src += T;
                            v = to_f32<IN_T>(clip_audio[(long long)src * M + (e % M)]);
                        }
                        dsti[e] = v;
                    }
                }
                __syncwarp();
            }
            if (c_valid) {
#pragma unroll 1
                for (int sg = 0; sg < kTile / kSeg; ++sg) {
                    const int ts = t0 + sg * kSeg;            // first sample of this segment
                    if (ts >= T) break;
                    const float *xp;
                    int stride = 1, wrap_at = kSeg;
                    if (!c_inphase) {
                        xp = sm.qs + (((kc & 1) * kSlots + c_slot) * M + (c_ch - M)) * kQPitch + sg * kSeg;
                    } else if (from_is) {
                        xp = sm.is_s + c_slot * kTile * M + sg * kSeg * M + c_ch;
                        stride = M;
                    } else {
                        const int cin = (ts - p.half + g.shift) % g.ring_x;
                        xp = sm.xs + (c_slot * M + c_ch) * g.pitch_x + cin;
                        wrap_at = g.ring_x - cin;
                    }
                    // this segment's words of the spike-bit ring start empty
                    bits[((ts >> 5) & (kRingWords - 1)) * 32] = 0u;
                    bits[(kRingWords + ((ts >> 5) & (kRingWords - 1))) * 32] = 0u;
                    // (warp-uniform) no lane wraps around the audio ring inside this segment
                    const int cin_u = (ts - p.half + g.shift + g.ring_x) % g.ring_x;
                    const bool fast = !from_is && ts + kSeg <= T && cin_u + kSeg <= g.ring_x;
                    const float carry = rz.csum;
                    unsigned int neg = 0u, zero = 0u;
                    float csum = carry;
                    int nvalid = kSeg;
                    if (fast) {
#pragma unroll
                        for (int i = 0; i < kSeg; ++i) {
                            const float z = biquad2_step(sos, bq, xp[i]);
                            csum += z;
                            cs[i * 32] = csum;
                            neg = __funnelshift_l(__float_as_uint(z), neg, 1);
                            zero = __funnelshift_l(z == 0.f ? 0x80000000u : 0u, zero, 1);
                        }
                    } else {
                        nvalid = T - ts < kSeg ? T - ts : kSeg;
#pragma unroll 1
                        for (int i = 0; i < nvalid; ++i) {
                            if (i == wrap_at) xp -= g.ring_x;
                            const float z = biquad2_step(sos, bq, xp[i * stride]);
                            csum += z;
                            cs[i * 32] = csum;
                            neg |= (__float_as_uint(z) >> 31) << (31 - i);
                            zero |= (z == 0.f ? 1u : 0u) << (31 - i);
                        }
                    }
                    rz.csum = csum;
                    rzcc_segment_masks(rz, store, bipolar, w, ts, nvalid, neg, zero, cs, 32, carry, emit);
                    const bool last = ts + kSeg >= T;
                    rzcc_close(rz, store, w, last ? T - 1 : ts + kSeg - 1, last, emit);
                }
            }
        }
        tile_barrier();
    }
    if (c_valid && rz.overflow && flags) atomicOr(flags + clip0 + c_slot, 1);
}

// ==== back warp: audio tile k+1 -> ring; neuron + Gram + spike write-out of tile k - dtile ====
template <typename IN_T, int MM>
__device__ __forceinline__ void back_role(const FusedSmem &sm, const ChainParams &p, const FusedGeom &g,
                                          const IN_T *__restrict__ audio, int8_t *__restrict__ spikes,
                                          long long clip0, long long B, long long T64, int lane, int NT, int k_last) {
    const int M = MM ? MM : p.M, C2 = 2 * M;
    const int T = (int)T64;
    // neuron-lane geometry: lane = slot * 16 + channel
    const int c_slot = lane >> 4, c_ch = lane & 15;
    const bool c_valid = c_ch < C2 && clip0 + c_slot < B;
    const unsigned int *bits = sm.bits + lane;
    int8_t *stg = sm.stage + c_slot * kTile * C2 + c_ch;
    float *vmo = sm.vms + lane;
    const float na = p.na, nc = p.nc, ncT = p.ncT, nLf = p.nLf;
    const int nL = p.nL;
    NeuronState nr; neuron_reset(nr);
    // Gram-lane geometry: lane = slot * 10 + upper-triangular 4x4 block pair
    const int g_slot = lane / 10;
    int g_bi = 0, g_bj = 0;
    {
        int pr = lane % 10;
        for (int r = 0; r < 4; ++r) {
            const int len = 4 - r;
            if (pr < len) { g_bi = r; g_bj = r + pr; break; }
            pr -= len;
        }
    }
    const bool g_lane = lane < 10 * kSlots;
    double acc64[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc64[i] = 0.0;

    for (int k = -1; k <= k_last; ++k) {
        // (a) audio tile k+1 -> mic-major ring: every lane moves whole frames (all microphones of one
        //     sample); the loads of a tile are all issued before the first store
        const int kf = k + 1;
        if (kf < NT) {
            constexpr int NF = kSlots * (kTile / 32);
            float v[NF][kRows];
            int coord[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const int s = f / (kTile / 32), h = f % (kTile / 32);
                const int t = kf * kTile + h * 32 + lane;
                coord[f] = (t + g.shift) % g.ring_x;
                const bool ok = t < T && clip0 + s < B;
                const IN_T *fr = audio + ((clip0 + (ok ? s : 0)) * T64 + (ok ? t : 0)) * M;
#pragma unroll
                for (int m = 0; m < kRows; ++m) v[f][m] = (ok && m < M) ? to_f32<IN_T>(fr[m]) : 0.f;
            }
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                const int s = f / (kTile / 32);
                float *rows = sm.xs + s * M * g.pitch_x + coord[f];
#pragma unroll
                for (int m = 0; m < kRows; ++m)
                    if (m < M) rows[m * g.pitch_x] = v[f][m];
            }
        }
        // (b) neuron filter of tile j = k - dtile from the final spike bits
        const int j = k - g.dtile;
        const int u0 = j * kTile;
        if (j >= 0 && u0 < T) {
            if (c_valid) {
#pragma unroll 1
                for (int sg = 0; sg < kTile / kSeg; ++sg) {
                    const int us = u0 + sg * kSeg;
                    const int wi = (us >> 5) & (kRingWords - 1);
                    unsigned int P = bits[(kRingWords + wi) * 32], Nn = bits[wi * 32];
                    if (us >= T) { P = 0u; Nn = 0u; }
                    // the same bits nL samples earlier (funnel over two ring words; zero before the clip start)
                    const int d0 = us - nL;
                    const int wd = (d0 >> 5) & (kRingWords - 1), wd1 = (wd + 1) & (kRingWords - 1), sh = d0 & 31;
                    unsigned int PD = __funnelshift_r(bits[(kRingWords + wd) * 32], bits[(kRingWords + wd1) * 32], sh);
                    unsigned int ND = __funnelshift_r(bits[wd * 32], bits[wd1 * 32], sh);
                    if (d0 < 0) {
                        const unsigned int keep = d0 <= -32 ? 0u : (0xffffffffu << (-d0));
                        PD &= keep; ND &= keep;
                    }
                    const int nvalid = T - us < kSeg ? (T - us > 0 ? T - us : 0) : kSeg;
                    if (nvalid < kSeg) {
                        const unsigned int keep = nvalid <= 0 ? 0u : (0xffffffffu >> (32 - nvalid));
                        P &= keep; Nn &= keep; PD &= keep; ND &= keep;
                    }
                    float *vseg = vmo + sg * kSeg * kVmPitch;
                    int8_t *sseg = stg + sg * kSeg * C2;
#pragma unroll
                    for (int i = 0; i < kSeg; ++i) {
                        // neuron_step with s, sd in {-1, 0, +1} given as bits
                        nr.p2 = na * (nr.p2 + nr.p1);
                        float a1 = na * nr.p1;
                        if (P & (1u << i)) a1 += 1.f;
                        if (Nn & (1u << i)) a1 -= 1.f;
                        nr.p1 = a1;
                        nr.q2 = na * (nr.q2 + nr.q1);
                        float b1 = na * nr.q1;
                        if (PD & (1u << i)) b1 += 1.f;
                        if (ND & (1u << i)) b1 -= 1.f;
                        nr.q1 = b1;
                        const float tail = fmaf(nLf, nr.q1, nr.q2);
                        float v = fmaf(-ncT, tail, nc * nr.p2);
                        if (i >= nvalid) v = 0.f;
                        vseg[i * kVmPitch] = v;
                        sseg[i * C2] = (int8_t)(((P >> i) & 1u) - ((Nn >> i) & 1u));
                    }
                }
            }
            __syncwarp();
            // (c) Gram of the membrane tile: C += v v^T, 4x4 blocks, FFMA2
            if (g_lane) {
                const float *vm = sm.vms + g_slot * 16;
                unsigned long long a2[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a2[i] = 0ull;
#pragma unroll 4
                for (int i = 0; i < kTile; ++i) {
                    const float4 a = *reinterpret_cast<const float4 *>(vm + i * kVmPitch + 4 * g_bi);
                    const ulonglong2 c = *reinterpret_cast<const ulonglong2 *>(vm + i * kVmPitch + 4 * g_bj);
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const unsigned long long ar = pack2(av[r], av[r]);
                        ffma2(a2[2 * r], c.x, ar);
                        ffma2(a2[2 * r + 1], c.y, ar);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float lo, hi;
                    unpack2(a2[i], lo, hi);
                    acc64[2 * i] += (double)lo;
                    acc64[2 * i + 1] += (double)hi;
                }
            }
            // (d) int8 spike raster of the tile -> HBM (contiguous [kTile][C2] in both places)
            if (spikes) {
                const int nrow = T - u0 < kTile ? T - u0 : kTile;
                for (int s = 0; s < kSlots; ++s) {
                    if (clip0 + s >= B) continue;
                    const int8_t *src = sm.stage + s * kTile * C2;
                    int8_t *dst = spikes + ((clip0 + s) * T64 + u0) * C2;
                    const int nbytes = nrow * C2;
                    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (nbytes & 15) == 0) {
                        for (int v = lane; v < nbytes / 16; v += 32)
                            reinterpret_cast<int4 *>(dst)[v] = reinterpret_cast<const int4 *>(src)[v];
                    } else {
                        for (int e = lane; e < nbytes; e += 32) dst[e] = src[e];
                    }
                }
            }
            __syncwarp();
        }
        tile_barrier();
    }
    // hand the Gram matrices to the epilogue (they reuse the audio rings, dead by now)
    if (g_lane) {
        double *Cd = sm.gram + g_slot * 256;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int row = 4 * g_bi + k / 4, col = 4 * g_bj + k % 4;
            Cd[row * 16 + col] = acc64[k];
            if (g_bi != g_bj) Cd[col * 16 + row] = acc64[k];
        }
    }
}

template <typename IN_T, int MM>
__global__ void __launch_bounds__(128, 3)
k_fused(const IN_T *__restrict__ audio, const float *__restrict__ taps, const double *__restrict__ Wd,
        int8_t *__restrict__ spikes, float *__restrict__ power, int32_t *__restrict__ doa,
        int32_t *__restrict__ flags, unsigned int *__restrict__ sm_slots,
        const __grid_constant__ ChainParams p, const __grid_constant__ FusedGeom g, long long B, long long T) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem sm;
    sm.taps = reinterpret_cast<float *>(smem_raw);
    sm.xs = reinterpret_cast<float *>(smem_raw + g.off_x);       // [kSlots*M][pitch_x]
    sm.qs = reinterpret_cast<float *>(smem_raw + g.off_q);       // [2][kSlots*M][kQPitch]
    sm.vms = reinterpret_cast<float *>(smem_raw + g.off_vm);     // [kTile][kVmPitch]
    sm.is_s = reinterpret_cast<float *>(smem_raw + g.off_is);    // [kSlots][kTile][M]
    sm.cs = reinterpret_cast<float *>(smem_raw + g.off_cs);      // [kSeg][32] running sums of the open segment
    sm.clus = reinterpret_cast<int *>(smem_raw + g.off_clus);    // RZCC cluster buffers, interleaved over 32 lanes
    sm.bits = reinterpret_cast<unsigned int *>(smem_raw + g.off_bits);   // [2][kRingWords][32]
    sm.stage = reinterpret_cast<int8_t *>(smem_raw + g.off_stage);       // [kSlots][kTile][C2]
    sm.gram = reinterpret_cast<double *>(smem_raw + g.off_x);    // [kSlots][16][16], clip epilogue only
    double *red_v = reinterpret_cast<double *>(smem_raw + g.off_q);      // [128], clip epilogue only
    int *red_i = reinterpret_cast<int *>(smem_raw + g.off_q + 128 * sizeof(double));
    __shared__ unsigned int s_rot;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = MM ? MM : p.M, C2 = 2 * M;

    // Warp w of a CTA lands on SM sub-partition w % 4; rotate the roles per co-resident
    // CTA so that every sub-partition gets its share of FIR warps.
    if (tid == 0) {
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned int k = atomicAdd(sm_slots + (smid & 255u), 1u) & 3u;
        s_rot = (0x3120u >> (4 * k)) & 3u;   // 0, 2, 1, 3
    }
    for (int i = tid; i < 8 * g.nblk; i += blockDim.x) sm.taps[i] = i < p.n_taps ? taps[i] : 0.f;
    __syncthreads();
    const int role = (warp + (int)s_rot) & 3;     // 0, 1: FIR of clip slot 0 / 1; 2: front; 3: back

    const int NT = (int)((T + kTile - 1) / kTile);
    const int k_last = NT - 1 + g.dtile;
    const long long npairs = (B + kSlots - 1) / kSlots;

    for (long long pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
        const long long clip0 = pair * kSlots;
        {   // zero the audio rings: samples before the clip start are zeros (lfilter's zero state)
            float4 *x4 = reinterpret_cast<float4 *>(sm.xs);
            const int n4 = kSlots * M * g.pitch_x / 4;
            for (int i = tid; i < n4; i += blockDim.x) x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            // no spikes before the clip start; membrane columns of unused lanes stay zero
            for (int i = tid; i < 2 * kRingWords * 32; i += blockDim.x) sm.bits[i] = 0u;
            for (int i = tid; i < kTile * kVmPitch; i += blockDim.x) sm.vms[i] = 0.f;
        }
        __syncthreads();

        if (role < 2) fir_role<MM>(sm, p, g, role, clip0 + role < B, lane, NT, k_last);
        else if (role == 2) front_role<IN_T, MM>(sm, p, g, audio, flags, clip0, B, T, lane, k_last);
        else back_role<IN_T, MM>(sm, p, g, audio, spikes, clip0, B, T, lane, NT, k_last);
        __syncthreads();

        // ---- clip epilogue: power[g] = w_g^T C w_g / T (float64), DoA = first argmax ----
        const double inv_T = 1.0 / (double)T;
        for (int s = 0; s < kSlots; ++s) {
            const long long clip = clip0 + s;
            if (clip >= B) break;
            const double *Cd = sm.gram + s * 256;
            double best = -1.0; int besti = 0x7fffffff;
            for (int gg = tid; gg < p.G; gg += blockDim.x) {
                double accp = 0.0;
                for (int r = 0; r < C2; ++r) {
                    double rr = 0.0;
                    for (int c = 0; c < C2; ++c) rr = fma(Cd[r * 16 + c], Wd[(long long)c * p.G + gg], rr);
                    accp = fma(Wd[(long long)r * p.G + gg], rr, accp);
                }
                accp *= inv_T;
                if (power) power[clip * p.G + gg] = (float)accp;
                if (accp > best) { best = accp; besti = gg; }
            }
            red_v[tid] = best; red_i[tid] = besti;
            __syncthreads();
            for (int st = blockDim.x / 2; st > 0; st >>= 1) {
                if (tid < st) {
                    const double ov = red_v[tid + st]; const int oi = red_i[tid + st];
                    if (ov > red_v[tid] || (ov == red_v[tid] && oi < red_i[tid])) { red_v[tid] = ov; red_i[tid] = oi; }
                }
                __syncthreads();
            }
            if (tid == 0 && doa) doa[clip] = red_i[0];
            __syncthreads();
        }
    }
}

static int next_pow2(int v) { int r = 1; while (r < v) r <<= 1; return r; }

bool fused_supported(const ChainParams &p) {
    return p.tap_stride == 2 && p.M <= kRows && p.nsec == 2 && (p.n_taps % 8) == 0;
}

template <typename IN_T, int MM>
static int launch_fused_t(const ChainParams &p, const FusedGeom &g, const float *d_taps, const double *d_Wd,
                          const IN_T *audio, long long B, long long T, int8_t *spikes, float *power, int32_t *doa,
                          int32_t *flags, unsigned int *sm_slots, int sm_count, cudaStream_t st) {
    auto kern = k_fused<IN_T, MM>;
    MICLOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
    int per_sm = 1;
    MICLOC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, g.smem_bytes));
    if (per_sm < 1) return set_error(MICLOC_ERR_UNSUPPORTED, "fused kernel does not fit (smem %d B)", g.smem_bytes);
    long long grid = (long long)sm_count * per_sm;
    const long long npairs = (B + kSlots - 1) / kSlots;
    if (grid > npairs) grid = npairs;
    kern<<<(unsigned)grid, 128, g.smem_bytes, st>>>(audio, d_taps, d_Wd, spikes, power, doa, flags, sm_slots, p, g, B, T);
    count_launch(1);
    MICLOC_CUDA(cudaGetLastError());
    return MICLOC_OK;
}

int launch_fused(const ChainParams &p, const float *d_taps, const double *d_Wd, const void *audio, int dtype,
                 long long B, long long T, int8_t *spikes, float *power, int32_t *doa, int32_t *flags,
                 unsigned int *sm_slots, int sm_count, cudaStream_t st) {
    if (!fused_supported(p))
        return set_error(MICLOC_ERR_UNSUPPORTED,
                         "fused kernel covers Hilbert-type STHT kernels (every other tap zero), a 2-section band-pass "
                         "and up to %d microphones; use the staged path", kRows);
    FusedGeom g{};
    // FIR tap blocks: groups of three blocks of 8 (zero taps appended by setup_stht up to a multiple of 8)
    g.nblk = (p.n_taps / 8 + 2) / 3 * 3;
    const int lookback = p.tap_first + 14 + 16 * (g.nblk - 1);      // oldest sample a tile's FIR windows load
    g.ring_x = ((lookback + 2 * kTile) + 31) / 32 * 32;             // history + current tile + tile being filled
    g.pitch_x = g.ring_x + 4;
    g.shift = ((p.tap_first + 14) % 16 + 16) % 16;
    // a spike at p is final once the front warp passed p + rzcc_lag(w) - 1; the back warp works on
    // tile k - dtile while the front warp has completed tile k - 2
    g.dtile = 1 + (rzcc_lag(p.w) + 63 + kTile - 1) / kTile;
    g.tiles_is = (p.half + kTile - 1) / kTile;
    int off = (8 * g.nblk * (int)sizeof(float) + 15) & ~15;
    g.off_x = off; off += kSlots * p.M * g.pitch_x * (int)sizeof(float);
    g.off_q = off; off += 2 * kSlots * p.M * kQPitch * (int)sizeof(float);
    g.off_vm = off; off += kTile * kVmPitch * (int)sizeof(float);
    g.off_is = off; off += kSlots * kTile * p.M * (int)sizeof(float);
    g.off_cs = off; off += kSeg * 32 * (int)sizeof(float);
    g.off_clus = off; off += 4 * kClusterMax * 32 * (int)sizeof(int);
    g.off_bits = off; off += 2 * kRingWords * 32 * (int)sizeof(int);
    g.off_stage = off; off += (kSlots * kTile * p.C2 + 15) & ~15;
    g.smem_bytes = off;
    // the spike-bit ring must hold the back warp's oldest read and the front warp's newest write
    if (kTile * (g.dtile + 1) + p.nL + kSeg > kRingWords * 32)
        return set_error(MICLOC_ERR_UNSUPPORTED, "robust_width %d / neuron length %d exceed the fused kernel's spike ring; "
                         "use the staged path", p.w, p.nL);
    if (kSlots * 256 * (int)sizeof(double) > kSlots * p.M * g.pitch_x * (int)sizeof(float) ||
        128 * 12 > 2 * kSlots * p.M * kQPitch * (int)sizeof(float))
        return set_error(MICLOC_ERR_UNSUPPORTED, "shared-memory tiles too small for the epilogue");
    if (g.smem_bytes > 227 * 1024)
        return set_error(MICLOC_ERR_UNSUPPORTED, "fused kernel needs %d B of shared memory; use the staged path", g.smem_bytes);
    if (T + 16 * kTile >= (1ll << 31)) return set_error(MICLOC_ERR_SHAPE, "T too large for the fused kernel");
#define MICLOC_FUSED_CASE(IN, MMV)                                                                        \
    return launch_fused_t<IN, MMV>(p, g, d_taps, d_Wd, (const IN *)audio, B, T, spikes, power, doa, flags, \
                                   sm_slots, sm_count, st)
    const bool i16 = dtype == MICLOC_I16;
    if (p.M == 7) { if (i16) MICLOC_FUSED_CASE(int16_t, 7); else MICLOC_FUSED_CASE(float, 7); }
    if (i16) MICLOC_FUSED_CASE(int16_t, 0); else MICLOC_FUSED_CASE(float, 0);
#undef MICLOC_FUSED_CASE
}

}  // namespace micloc
