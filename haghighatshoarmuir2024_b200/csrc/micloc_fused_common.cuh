// micloc_fused_common.cuh -- what the two fused kernels (micloc_fused_tc.cu: STHT on the tensor cores, the default;
// micloc_fused.cu: STHT on the FP32 FMA pipe) share: tile geometry, packed-FMA helpers, the per-tile barrier and the optional role timers, the
// band-pass biquad pair, the RZCC and neuron warp roles and the tensor-core MMA of the Gram warp.
//
// Reference sites: micloc/snn_beamformer.py:283-370 (see the kernels' own headers).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "micloc_common.h"

namespace micloc {

constexpr int kTile = 64;      // samples per pipeline step
constexpr int kSlots = 2;      // clips per group
constexpr int kQPitch = kTile + 4;
constexpr int kVmRows = 16 * kSlots;    // membrane tile rows: [slot][16 channels] (channels 14, 15 stay zero)
constexpr int kVmPitch = kTile + 8;     // halves per row of a membrane tile (channel-major: ldmatrix rows of 8 samples)
constexpr float kVmScale = 16384.f;     // membrane values are stored x 2^14 (|v| <= 1: the alpha kernel sums to 1) as fp16 hi + lo
constexpr int kWarps = 8;      // warps of one clip-pair group
constexpr int kThreads = kWarps * 32;
constexpr int kGramFlush = 2;  // tiles of float32 Gram accumulation (inside the tensor cores: truncating adds) between two folds into float64
constexpr int kSegsPerTile = kTile / kSeg;
constexpr int kRoleBandpass = 4, kRoleRzcc = 5, kRoleNeuron = 6, kRoleGram = 7;    // roles 0..3 belong to the FIR side

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

// The eight warps of one clip-pair group meet here once per pipeline step (the roles run different code);
// every group of a CTA owns one named barrier.
// Named barrier of a clip-pair group.  `bar.sync` is `barrier.sync.aligned`: PTX wants the whole warp converged on one
// barrier instruction.  compute-sanitizer synccheck reports role warps that arrive split (lanes without a channel leave the
// step's loops early and wait at the WARPSYNC.ALL ptxas puts in front of the barrier; the tool sees the two halves pass it
// separately).  Both PTX forms compile to the SAME instruction, BAR.SYNC.DEFER_BLOCKING, which counts threads, so the
// hardware behaviour is that of the unaligned form either way; what differs is the code ptxas builds around it: with
// `barrier.sync` (no .aligned) it duplicates the paths into the barrier and the kernel runs 6 % slower (260 k vs 277 k
// clips/s).  -DMICLOC_UNALIGNED_BARRIERS selects that form (tools/build_rt.sh; synccheck-clean, same results bit for bit).
#ifdef MICLOC_UNALIGNED_BARRIERS
#define MICLOC_BAR_SYNC "barrier.sync"
#else
#define MICLOC_BAR_SYNC "bar.sync"
#endif
__device__ __forceinline__ void tile_barrier(int id, int nthreads = kThreads) {
    asm volatile(MICLOC_BAR_SYNC " %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Called by every serial role at the top of pipeline step k, before its own work: the tensor-core kernel
// (micloc_fused_tc.cu) moves the finished STHT tile out of tensor memory there; the FFMA kernel passes this no-op.
struct NoStepHook { __device__ __forceinline__ void operator()(int) const {} };

// Optional role timing (MICLOC_ROLE_TIMING): busy cycles of each warp role between barriers, summed into
// the 64-bit counters at sm_slots[kSlotDbg] (busy of roles 0..7, then the number of warps that reported
// each) by lane 0; read back by micloc_snn_debug_counters.
#ifdef MICLOC_ROLE_TIMING
__device__ __forceinline__ long long rt_clock() {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    return t;
}
struct RoleTimer {
    long long t0, busy;
    __device__ __forceinline__ void start() { t0 = rt_clock(); busy = 0; }
    __device__ __forceinline__ void before_barrier() { busy += rt_clock() - t0; }
    __device__ __forceinline__ void after_barrier() {
        // BAR.SYNC only blocks at the next consumer: touch shared memory, then read the clock
        unsigned int v;
        long long t;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(0) : "memory");
        asm volatile("{ .reg .u32 d; mov.u32 d, %1; mov.u64 %0, %%clock64; }" : "=l"(t) : "r"(v));
        t0 = t;
    }
    __device__ __forceinline__ void flush(unsigned int *sm_slots, int role, int lane, int rec) {
        if (lane == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(sm_slots + kSlotDbg) + role, (unsigned long long)busy);
            atomicAdd(reinterpret_cast<unsigned long long *>(sm_slots + kSlotDbg) + 8 + role, 1ull);
            if (rec < 512)
                (reinterpret_cast<unsigned long long *>(sm_slots + kSlotCta) + 16 * rec)[4 + role] = (unsigned long long)busy;
        }
    }
};
// phase timers inside a role: cycles summed into debug counters 18 + i (i < 14), read back by micloc_snn_debug_counters
#define PH_DECL long long ph_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pht_ = 0
#define PH_START() pht_ = rt_clock()
#define PH_END(i) do { const long long n_ = rt_clock(); ph_[i] += n_ - pht_; pht_ = n_; } while (0)
#define PH_FLUSH(dbg, base, n) do { if (lane == 0) for (int i_ = 0; i_ < (n); ++i_) \
        atomicAdd(reinterpret_cast<unsigned long long *>((dbg) + kSlotDbg) + 18 + (base) + i_, (unsigned long long)ph_[i_]); } while (0)
#define ROLE_TIMER_DECL RoleTimer rt_; rt_.start()
#define ROLE_BARRIER() do { rt_.before_barrier(); tile_barrier(sm.bar_id, sm.bar_threads); rt_.after_barrier(); } while (0)
#define ROLE_TIMER_FLUSH(role) rt_.flush(sm.dbg, role, lane, sm.rec)
#else
#define PH_DECL
#define PH_START()
#define PH_END(i)
#define PH_FLUSH(dbg, base, n)
#define ROLE_TIMER_DECL
#define ROLE_BARRIER() tile_barrier(sm.bar_id, sm.bar_threads)
#define ROLE_TIMER_FLUSH(role)
#endif

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

// ============ RZCC warp: masks of tile k-2 -> candidates -> clusters -> spike bits, lane = slot*16 + channel ============
template <typename SMEM, int kRingWords, typename HOOK = NoStepHook>
__device__ __forceinline__ void rzcc_role(const SMEM &sm, const ChainParams &p, int32_t *__restrict__ flags,
                                          long long clip0, long long B, long long T64, int MMv, int lane, int k_last,
                                          HOOK hook = HOOK()) {
    const int C2 = 2 * MMv;
    const int T = (int)T64;
    const int c_slot = lane >> 4, c_ch = lane & 15;
    const bool c_valid = c_ch < C2 && clip0 + c_slot < B;
    const int w = p.w, bipolar = p.bipolar;
    const RzccStore store{sm.clus + lane, reinterpret_cast<float *>(sm.clus + 2 * kClusterMax * 32) + lane, 32};
    unsigned int *bits = sm.bits + lane;
    // a final spike: set its bit in this channel's ring word (only this lane ever writes these words)
    auto emit = [&](int pos, int sign) {
        unsigned int *wd = bits + ((sign > 0 ? kRingWords : 0) + ((pos >> 5) & (kRingWords - 1))) * 32;
        *wd |= 1u << (pos & 31);
    };
    RzccState rz; rzcc_reset(rz);
    ROLE_TIMER_DECL;

    for (int k = sm.k_first; k <= k_last; ++k) {
        hook(k);
        const int kr = k - 3;
        const int t0 = kr * kTile;
        if (kr >= 0 && t0 < T && c_valid) {
#pragma unroll 1
            for (int sg = 0; sg < kSegsPerTile; ++sg) {
                const int ts = t0 + sg * kSeg;
                if (ts >= T) break;
                // this segment's words of the spike-bit ring start empty
                bits[((ts >> 5) & (kRingWords - 1)) * 32] = 0u;
                bits[(kRingWords + ((ts >> 5) & (kRingWords - 1))) * 32] = 0u;
                const float *cs = sm.cs + ((kr & 1) * kSegsPerTile + sg) * kSeg * 32 + lane;
                const unsigned int *sgm = sm.seg + ((kr & 1) * kSegsPerTile + sg) * 3 * 32 + lane;
                const unsigned int neg = sgm[0], zero = sgm[32];
                const float carry = __uint_as_float(sgm[64]);
                const int nvalid = T - ts < kSeg ? T - ts : kSeg;
                rzcc_segment_masks(rz, store, bipolar, w, ts, nvalid, neg, zero, cs, 32, carry, emit);
                const bool last = ts + kSeg >= T;
                rzcc_close(rz, store, w, last ? T - 1 : ts + kSeg - 1, last, emit);
            }
        }
        ROLE_BARRIER();
    }
    ROLE_TIMER_FLUSH(kRoleRzcc);
    if (c_valid && rz.overflow && flags) atomicOr(flags + clip0 + c_slot, 1);
}

// ==== neuron warp: alpha-kernel recurrences of tile k - dtile -> membrane tile + int8 spike tile, lane = slot*16 + channel ====
template <typename SMEM, typename GEOM, int kRingWords, typename HOOK = NoStepHook>
__device__ __forceinline__ void neuron_role(const SMEM &sm, const ChainParams &p, const GEOM &g,
                                            long long clip0, long long B, long long T64, int MMv, int lane, int k_last,
                                            HOOK hook = HOOK()) {
    const int C2 = 2 * MMv;
    const int T = (int)T64;
    const int c_slot = lane >> 4, c_ch = lane & 15;
    const bool c_valid = c_ch < C2 && clip0 + c_slot < B;
    const unsigned int *bits = sm.bits + lane;
    const float na = p.na, nc = p.nc * kVmScale, ncT = p.ncT * kVmScale, nLf = p.nLf;     // membrane values x 2^14
    const int nL = p.nL;
    NeuronState nr; neuron_reset(nr);
    ROLE_TIMER_DECL;

    for (int k = sm.k_first; k <= k_last; ++k) {
        hook(k);
        const int j = k - g.dtile;
        const int u0 = j * kTile;
        if (j >= 0 && u0 < T && c_valid) {
            int8_t *stg = sm.stage + ((j & 1) * kSlots + c_slot) * kTile * C2 + c_ch;
            __half *vmo = sm.vms + ((j & 1) * kVmRows + lane) * kVmPitch;      // hi tile; the lo tile follows all hi tiles
#pragma unroll 1
            for (int sg = 0; sg < kSegsPerTile; ++sg) {
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
                __half *vseg = vmo + sg * kSeg;
                int8_t *sseg = stg + sg * kSeg * C2;
#pragma unroll 1
                for (int o = 0; o < kSeg / 8; ++o) {
                    float vq[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // neuron_step with s, sd in {-1, 0, +1} given as bits; a (p2 + p1) as a p2 + (a p1): the
                        // product a p1 is needed for p1 anyway
                        float a1 = na * nr.p1;
                        nr.p2 = fmaf(na, nr.p2, a1);
                        if (P & (1u << i)) a1 += 1.f;
                        if (Nn & (1u << i)) a1 -= 1.f;
                        nr.p1 = a1;
                        float b1 = na * nr.q1;
                        nr.q2 = fmaf(na, nr.q2, b1);
                        if (PD & (1u << i)) b1 += 1.f;
                        if (ND & (1u << i)) b1 -= 1.f;
                        nr.q1 = b1;
                        const float tail = fmaf(nLf, nr.q1, nr.q2);
                        const float v = fmaf(-ncT, tail, nc * nr.p2);
                        vq[i] = v;
                    }
                    // the 8 samples' spike bytes (+1 / -1 / 0): bit k of a nibble -> byte k by one multiply (0x00204081
                    // copies the nibble to bit offsets 0, 7, 14, 21), -1 = 0xff = byte x 255; one byte store per sample
                    {
                        const unsigned int p0 = ((P & 0xfu) * 0x00204081u) & 0x01010101u, p1 = (((P >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
                        const unsigned int n0 = ((Nn & 0xfu) * 0x00204081u) & 0x01010101u, n1 = (((Nn >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
                        const unsigned int b0 = p0 | (n0 * 255u), b1 = p1 | (n1 * 255u);
                        int8_t *s8 = sseg + 8 * o * C2;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            s8[i * C2] = (int8_t)(b0 >> (8 * i));
                            s8[(4 + i) * C2] = (int8_t)(b1 >> (8 * i));
                        }
                    }
                    // v = hi + lo with hi = fp16(v), lo = fp16(v - hi): 22 significant bits for the tensor-core Gram
                    uint4 h4, l4;
                    unsigned int *hp = &h4.x, *lp = &l4.x;
#pragma unroll
                    for (int i2 = 0; i2 < 4; ++i2) {
                        const __half2 hh = __floats2half2_rn(vq[2 * i2], vq[2 * i2 + 1]);
                        const float2 hf = __half22float2(hh);
                        const __half2 ll = __floats2half2_rn(vq[2 * i2] - hf.x, vq[2 * i2 + 1] - hf.y);
                        hp[i2] = *reinterpret_cast<const unsigned int *>(&hh);
                        lp[i2] = *reinterpret_cast<const unsigned int *>(&ll);
                    }
                    *reinterpret_cast<uint4 *>(vseg + 8 * o) = h4;
                    *reinterpret_cast<uint4 *>(vseg + 2 * kVmRows * kVmPitch + 8 * o) = l4;
                    P >>= 8; Nn >>= 8; PD >>= 8; ND >>= 8;
                }
                // the membrane potential behind the clip end does not count (ragged last segment)
                for (int i = nvalid; i < kSeg; ++i) {
                    vseg[i] = __float2half(0.f);
                    vseg[2 * kVmRows * kVmPitch + i] = __float2half(0.f);
                }
            }
        }
        ROLE_BARRIER();
    }
    ROLE_TIMER_FLUSH(kRoleNeuron);
}

__device__ __forceinline__ void mma_f16_16x8x16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ==== Gram warp: C += V V^T of the membrane tile k - dtile - 1 on the tensor cores, then that tile's int8 spike
// raster -> HBM.  The neuron warp leaves every membrane value (x 2^14) as an fp16 pair v = hi + lo (22 significant
// bits).  Per clip slot and 16 time samples one ldmatrix.x4 each fetches the m16n8k16 fragments of hi and lo of
// V^T (16 channels x 16 samples; the same registers serve as the "col" operand, the matrix is V V^T), and
// C += hi hi^T + hi lo^T + lo hi^T runs as three fp16 MMAs per 8-channel column block: exact products (the dropped
// lo lo^T is below 2^-22 relative) accumulated in float32 over kGramFlush tiles -- the tensor cores add with
// truncation, a long chain would bias the sum -- and then folded into float64 registers.
template <typename SMEM, typename GEOM, typename HOOK = NoStepHook>
__device__ __forceinline__ void gram_role(const SMEM &sm, const GEOM &g, int8_t *__restrict__ spikes,
                                          long long clip0, long long B, long long T64, int MMv, int lane, int k_last,
                                          HOOK hook = HOOK()) {
    const int C2 = 2 * MMv;
    const int T = (int)T64;
    // ldmatrix row of this lane: matrix lane / 8 = (channels 0-7 | 8-15) x (samples 0-7 | 8-15) of a k-step
    const int lm_row = (lane & 7) + 8 * ((lane >> 3) & 1), lm_t = 8 * (lane >> 4);
    float accf[kSlots][2][4];           // float32 partial sums: [slot][column block][m16n8 accumulator fragment]
    double accd[kSlots][2][4];
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int i = 0; i < 4; ++i) { accf[s][nb][i] = 0.f; accd[s][nb][i] = 0.0; }
    ROLE_TIMER_DECL;

    for (int k = sm.k_first; k <= k_last; ++k) {
        hook(k);
        const int j = k - g.dtile - 1;
        const int u0 = j * kTile;
        const bool live = j >= 0 && u0 < T;
        if (live) {
            const __half *vm = sm.vms + ((j & 1) * kVmRows + lm_row) * kVmPitch + lm_t;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const unsigned addr = (unsigned)__cvta_generic_to_shared(vm + s * 16 * kVmPitch);
                const unsigned lo_off = 2 * kVmRows * kVmPitch * (unsigned)sizeof(__half);
#pragma unroll 1     // (a rolled loop: the fused kernels stall on instruction fetch before anything else)
                for (int ks = 0; ks < kTile / 16; ++ks) {
                    unsigned hi[4], lo[4];
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(hi[0]), "=r"(hi[1]), "=r"(hi[2]), "=r"(hi[3]) : "r"(addr + 32u * ks));
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(lo[0]), "=r"(lo[1]), "=r"(lo[2]), "=r"(lo[3]) : "r"(addr + lo_off + 32u * ks));
                    // column block 0 = channels 0-7: its k x n fragment is (a0, a2); block 1 = channels 8-15: (a1, a3)
                    mma_f16_16x8x16(accf[s][0], hi, hi[0], hi[2]);
                    mma_f16_16x8x16(accf[s][1], hi, hi[1], hi[3]);
                    mma_f16_16x8x16(accf[s][0], hi, lo[0], lo[2]);
                    mma_f16_16x8x16(accf[s][1], hi, lo[1], lo[3]);
                    mma_f16_16x8x16(accf[s][0], lo, hi[0], hi[2]);
                    mma_f16_16x8x16(accf[s][1], lo, hi[1], hi[3]);
                }
            }
            if ((j % kGramFlush) == kGramFlush - 1 || (j + 1) * kTile >= T) {
#pragma unroll
                for (int s = 0; s < kSlots; ++s)
#pragma unroll
                    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                        for (int i = 0; i < 4; ++i) { accd[s][nb][i] += (double)accf[s][nb][i]; accf[s][nb][i] = 0.f; }
            }
        }
        // int8 spike raster of the tile -> HBM (contiguous [kTile][C2] in both places)
        if (live && spikes) {
            const int nrow = T - u0 < kTile ? T - u0 : kTile;
            for (int s = 0; s < kSlots; ++s) {
                if (clip0 + s >= B) continue;
                const int8_t *src = sm.stage + ((j & 1) * kSlots + s) * kTile * C2;
                int8_t *dst = spikes + ((clip0 + s) * T64 + u0) * C2;
                const int nbytes = nrow * C2;
                if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (nbytes & 15) == 0 &&
                    ((kTile * C2) & 15) == 0) {
#pragma unroll 1
                    for (int v = lane; v < nbytes / 16; v += 32)
                        reinterpret_cast<int4 *>(dst)[v] = reinterpret_cast<const int4 *>(src)[v];
                } else {
#pragma unroll 1
                    for (int e = lane; e < nbytes; e += 32) dst[e] = src[e];
                }
            }
        }
        ROLE_BARRIER();
    }
    ROLE_TIMER_FLUSH(kRoleGram);
    // the Gram matrices of the two clips -> shared memory for the clip epilogue (the audio rings are dead now):
    // accumulator fragment (m16n8): c0, c1 = row lane/4, columns 2 (lane%4) + {0, 1}; c2, c3 = row lane/4 + 8
#pragma unroll
    for (int s = 0; s < kSlots; ++s)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = (lane >> 2) + 8 * (i >> 1), col = 8 * nb + 2 * (lane & 3) + (i & 1);
                sm.gram[s * 256 + row * 16 + col] = accd[s][nb][i] * (1.0 / ((double)kVmScale * (double)kVmScale));
            }
}


}  // namespace micloc
