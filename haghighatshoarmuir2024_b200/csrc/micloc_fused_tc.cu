// micloc_fused_tc.cu -- the fused hot-path kernel with the STHT FIR on the 5th-generation tensor cores.
//
// Reference sites: micloc/snn_beamformer.py:283-370 (STHT :325-327) and the callers' power/argmax
// paper_plots/target_snn_localization.py:462-464.
//
// Same chain, same serial roles (band-pass, RZCC, neuron, Gram: micloc_fused_common.cuh) and the same pipeline of
// 64-sample steps as micloc_fused.cu, but the 240-tap quadrature FIR -- 84 % of the arithmetic -- no longer occupies
// issue slots of the SM sub-partitions: it runs as a Toeplitz GEMM on tcgen05.mma with the accumulators in tensor
// memory, issued by one thread.
//
//   polyphase   The Hilbert kernel has a tap at every other lag, so Q at odd times is a DENSE FIR of the even
//               samples E[n] = x[2n] and Q at even times the same FIR of the odd samples O[n] = x[2n-1]:
//               Q[2n+1] = sum_j g[j] E[n-j],  Q[2n] = sum_j g[j] O[n-j]   (g[j] = h[k0 + 2j], k0 odd).
//   operands    Every stream (microphone x parity) lives time-contiguous in a shared-memory ring as two fp16
//               pieces u = hi + lo (22 significant bits; the clip is scaled by a power of two so that it fits the
//               fp16 range -- band-pass and RZCC are scale invariant).  The A operand of an M=128 x N=16 x K=16
//               instruction is the HANKEL matrix A[m][e] = u[8m - lag + e] read straight from that ring: a K-major
//               no-swizzle descriptor whose 8-row core matrices overlap (row r starts 16 bytes = 8 samples after
//               row r-1, leading byte offset 16), 16 row groups = the 2M streams of one clip at the ring pitch.
//               B is the constant Toeplitz block of the taps for the 8 output phases, split hi | lo (x 2^14):
//               B[a][e] = g[a + lag - e].  2 x 16 instructions per clip and 128 frames give
//               D[m][a] + D[m][8+a] = y[8m + a] with all four hi/lo cross terms, accumulated in float32.
//   in          The producer warp streams 128-frame audio tiles with 1-D TMA bulk copies (cp.async.bulk +
//               mbarrier) into a staging buffer, converts them into the hi/lo rings and issues the MMAs.
//   out         The four serial warps each own one 32-lane quarter of tensor memory: at the top of a step they move
//               their share of the finished quadrature tile into the step's q rows (tcgen05.ld) and rebuild the
//               in-phase samples x[t - K/2] (np.roll: the first K/2 come from the clip tail) from the rings.
//
// One CTA holds two independent clip-pair groups of five warps (four serial roles + producer).
#include <cstdlib>

#include "micloc_fused_common.cuh"

namespace micloc {
namespace tc {

constexpr int kGWarps = 5;                // warps of a clip-pair group
constexpr int kGThreads = kGWarps * 32;
constexpr int kRoleProducer = 4;          // roles 0..3: band-pass, RZCC, neuron, Gram = tensor-memory lane quarter
constexpr int kRingWords = 16;            // spike-bit ring: 16 words of 32 samples per channel and polarity
constexpr int kMac = 2 * kTile;           // frames per MMA tile (8 rows x 8 samples per stream and parity)
constexpr int kMirror = 56;               // ring positions repeated behind the ring end (a row group reads 72 in a row)
constexpr int kQRow = kTile + 4;          // floats per channel row of a q sub-tile: [even times: 32][odd times: 32] + pad
constexpr int kRows = 8;                  // most microphones (16 row groups = 8 microphones x 2 parities)
constexpr float kTapScale = 16384.f;      // taps are stored x 2^14 as fp16 hi + lo
constexpr int kColsPerTile = 32;          // tensor-memory columns of one tile: hi-piece and lo-piece accumulators of N = 16

struct TcGeom {
    int R;          // ring length in stream samples (multiple of 64)
    int pitch_b;    // bytes per stream row: 2 (R + kMirror)
    int lag;        // a row's window starts `lag` samples before its first output (multiple of 16)
    int ksteps;     // MMA K steps of 16 per piece
    int H;          // K/2 / 2: in-phase delay in stream samples
    int d0;         // leading zero taps of the polyphase filter
    int dtile;      // the neuron warp runs dtile steps behind (RZCC decision latency)
    int tiles_is;   // sub-tiles whose in-phase input comes from the clip tail
    int off_ring, off_stin, off_q, off_vm, off_cs, off_seg, off_clus, off_bits, off_stage, off_misc;
    int smem_group; // bytes of one group
    int smem_bytes; // bytes of the CTA (groups + tap matrix)
};

struct TcSmem {
    // what the shared roles use
    float *cs;
    unsigned int *seg;
    int *clus;
    unsigned int *bits;
    __half *vms;
    int8_t *stage;
    double *gram;
    unsigned int *dbg;
    int bar_id, bar_threads, rec;
    // tensor-core side
    unsigned char *ring;      // [kSlots][2 pieces][2M streams][pitch_b] (+ 2 rows of padding)
    unsigned char *stin;      // TMA staging of one audio tile [kMac][M]
    float *q;                 // [2 sub-tiles][kSlots][2M][kQRow]
    float *scale;             // [kSlots] power-of-two clip scale
    unsigned int *amax;       // [kSlots]
    unsigned int *amax_next;  // [kSlots] of the clip pair this group takes next (scanned by the producer warp meanwhile)
    float *carry;             // [kSlots][2][8] last frame of the previous audio tile (first odd-stream sample of the next)
    unsigned long long *mbar; // [0..3] MMA done (slot, buffer), [4] audio tile staged, [5] tile converted into the rings
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_inval(uint32_t a) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_mma_f16(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t accum) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// K-major, no swizzle: 8-row x 16-byte core matrices; `lbo` = bytes between core matrices along K, `sbo` = bytes
// between 8-row groups along M / N (cute::UMMA::SmemDescriptor: start [0,14), LBO [16,30), SBO [32,46), version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// kind::f16, A and B fp16 K-major, D float32, M = 128, N = 16 (cute::UMMA::InstrDescriptor)
constexpr int kMmaN = 16;
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(kMmaN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

template <typename IN_T> __device__ __forceinline__ float clip_scale(unsigned int amax_bits);
template <> __device__ __forceinline__ float clip_scale<int16_t>(unsigned int) { return 1.f; }      // |x| <= 32768 fits fp16; hi + lo is exact
template <> __device__ __forceinline__ float clip_scale<float>(unsigned int amax_bits) {
    const float a = __uint_as_float(amax_bits);
    if (!(a > 0.f) || !(a < 3.0e38f)) return 1.f;
    int e = (int)((amax_bits >> 23) & 0xffu) - 127;      // floor(log2 a) for normal numbers
    int s = 13 - e;                                      // |x| 2^s < 2^14
    s = s > 120 ? 120 : (s < -120 ? -120 : s);
    return __uint_as_float((unsigned int)(s + 127) << 23);
}

// ======================= producer warp: TMA -> staging, MMAs; amax scan of the next clip pair =======================
// Step k belongs to audio tile J = (k + 1) / 2 of clip slot (k odd ? 0 : 1): the four serial warps convert the staged
// tile into the hi / lo rings at the top of the step (StepHook), the producer waits for them, sends the next tile
// on its way and issues the tile's MMAs.
__device__ __forceinline__ bool tile_of_step(int k, int NJ, bool ok1, int &slot, int &J) {
    slot = (k & 1) ? 0 : 1;
    J = (k + 1) >> 1;
    return J < NJ && (slot == 0 || ok1);
}

template <typename IN_T, int MM>
__device__ __forceinline__ void producer_role(const TcSmem &sm, const ChainParams &p, const TcGeom &g,
                                              const IN_T *__restrict__ audio, long long clip0, long long B,
                                              long long T64, int lane, int NJ, int k_last, uint32_t tmem_cols,
                                              uint32_t tapsB, uint32_t &ring_phase, long long next_clip0) {
    const int M = MM ? MM : p.M;
    const int T = (int)T64;
    const bool ok1 = clip0 + 1 < B;
    const IN_T *src[kSlots] = {audio + clip0 * T64 * M, audio + (ok1 ? clip0 + 1 : clip0) * T64 * M};
    const bool al[kSlots] = {(reinterpret_cast<uintptr_t>(src[0]) & 15) == 0, (reinterpret_cast<uintptr_t>(src[1]) & 15) == 0};
    const uint32_t tile_bytes = (uint32_t)(kMac * M * sizeof(IN_T));
    const uint32_t bar_tma = smem_u32(sm.mbar + 4), bar_ring = smem_u32(sm.mbar + 5);
    const uint32_t stin_a = smem_u32(sm.stin);
    IN_T *stin = reinterpret_cast<IN_T *>(sm.stin);
    ROLE_TIMER_DECL;
    PH_DECL;

    // audio tile J of clip slot `slot` -> staging: one 1-D TMA bulk copy when the tile is whole and 16-byte aligned,
    // else plain loads with zero fill behind the clip end; either way the tile's arrival completes a phase of bar_tma
    auto issue_load = [&](int slot, int J) {
        const int f0 = kMac * J;
        const IN_T *s = src[slot] + (long long)f0 * M;
        if (al[slot] && f0 + kMac <= T) {
            if (lane == 0) {
                fence_proxy_async();
                mbar_expect_tx(bar_tma, tile_bytes);
                tma_bulk_g2s(stin_a, s, tile_bytes, bar_tma);
            }
        } else {
            const int nvalid = (T - f0 < kMac ? T - f0 : kMac) * M;
            for (int e = lane; e < kMac * M; e += 32) stin[e] = e < nvalid ? s[e] : (IN_T)0;
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tma);
        }
        __syncwarp();
    };

    // largest magnitude of the NEXT clip pair (float32 input), a slice per step: its clip scale is ready when that pair starts
    const bool scan = sizeof(IN_T) == 4 && next_clip0 >= 0 && ((T64 * M) & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(audio + next_clip0 * T64 * M) & 15) == 0;
    const long long n4 = (T64 * M) >> 2;
    const int per_step = (int)((n4 + k_last + 1) / (k_last + 2));
    const float4 *scan_src[kSlots] = {reinterpret_cast<const float4 *>(audio + (next_clip0 < 0 ? 0 : next_clip0) * T64 * M),
                                      reinterpret_cast<const float4 *>(audio + (next_clip0 < 0 ? 0 : next_clip0 + 1) * T64 * M)};
    const bool scan_ok1 = next_clip0 + 1 < B;
    float mx[kSlots] = {0.f, 0.f};
    long long scan_i = 0;

    issue_load(0, 0);
    for (int k = -1; k <= k_last; ++k) {
        int slot, J;
        if (tile_of_step(k, NJ, ok1, slot, J)) {
            PH_START();
            mbar_wait(bar_ring, ring_phase);          // the serial warps have turned the staged tile into ring samples
            ring_phase ^= 1u;
            PH_END(0);
            // ---- next tile on its way while the tensor cores work on this one ----
            {
                int s2, J2;
                if (tile_of_step(k + 1, NJ, ok1, s2, J2)) issue_load(s2, J2);
                else if (tile_of_step(k + 2, NJ, ok1, s2, J2)) issue_load(s2, J2);
            }
            PH_END(1);
            // ---- STHT of the tile: the hi-piece chain into accumulator 0, the lo-piece chain into accumulator 1
            //      (independent chains pipeline in the tensor core; the epilogue adds them) ----
            tc_fence_after();
            if (lane == 0) {
                const uint32_t d0 = tmem_cols + (uint32_t)((slot * 2 + (J & 1)) * kColsPerTile);
                int pos = (kTile * J) % g.R - g.lag;
                if (pos < 0) pos += g.R;
                const uint32_t ra_hi = smem_u32(sm.ring + (size_t)((slot * 2 + 0) * 2 * M) * g.pitch_b);
                const uint32_t ra_lo = smem_u32(sm.ring + (size_t)((slot * 2 + 1) * 2 * M) * g.pitch_b);
                const uint64_t a_fix = make_desc(0u, 16u, (uint32_t)g.pitch_b);
                uint64_t bd = make_desc(tapsB, 128u, (uint32_t)(2 * g.ksteps) * 128u);
#pragma unroll 2
                for (int ks = 0; ks < g.ksteps; ++ks) {
                    const uint32_t off = (2u * (uint32_t)pos) >> 4;
                    tc_mma_f16(d0, a_fix | (uint64_t)(((ra_hi >> 4) + off) & 0x3FFFu), bd, kIdesc, ks ? 1u : 0u);
                    tc_mma_f16(d0 + 16u, a_fix | (uint64_t)(((ra_lo >> 4) + off) & 0x3FFFu), bd, kIdesc, ks ? 1u : 0u);
                    bd += 16u;                          // next K step of the tap matrix: 256 bytes on
                    pos += 16;
                    if (pos >= g.R) pos -= g.R;
                }
                tc_commit(smem_u32(sm.mbar + slot * 2 + (J & 1)));
            }
            __syncwarp();
            PH_END(2);
        }
        if (scan) {
            PH_START();
            const long long i_end = scan_i + per_step < n4 ? scan_i + per_step : n4;
#pragma unroll
            for (int c = 0; c < kSlots; ++c) {
                if (c == 1 && !scan_ok1) break;
                for (long long i = scan_i + lane; i < i_end; i += 32) {
                    const float4 v = __ldg(scan_src[c] + i);
                    mx[c] = fmaxf(fmaxf(mx[c], fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
                    if (i + per_step < n4) asm volatile("prefetch.global.L2 [%0];" ::"l"(scan_src[c] + i + per_step));
                }
            }
            scan_i = i_end;
            PH_END(3);
        }
        ROLE_BARRIER();
    }
    // amax of the next pair (bit pattern order = magnitude order); ~0 = "not scanned"
#pragma unroll
    for (int c = 0; c < kSlots; ++c) {
        unsigned int mb = __float_as_uint(mx[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned int ot = __shfl_xor_sync(0xffffffffu, mb, o); mb = ot > mb ? ot : mb; }
        if (lane == 0) sm.amax_next[c] = scan ? mb : 0xffffffffu;
    }
    ROLE_TIMER_FLUSH(0);
    PH_FLUSH(sm.dbg, 0, 4);
}

// ===== top of a step, serial warps: quadrature tile out of tensor memory, in-phase samples out of the rings, the
// ===== next audio tile from the staging buffer into the rings =====
template <typename IN_T, int MM>
struct StepHook {
    const TcSmem &sm;
    const ChainParams &p;
    const TcGeom &g;
    const IN_T *__restrict__ audio;
    long long clip0, B, T64;
    int role, lane, NT, NJ;
    uint32_t tmem_cols;
    uint32_t &tma_phase;

    __device__ __forceinline__ void operator()(int k) const {
        const int M = MM ? MM : p.M, C2 = 2 * M;
        const int T = (int)T64;
        const int s = k - 1;                  // sub-tile the band-pass warp takes at the next step
        if (s >= 0 && s < NT) {
            const int J = s >> 1, h = s & 1, buf = J & 1;
            // ---- Q: row m = 8 (2 mic + parity) + r holds y[8 r + a] of that stream: columns a / 8 + a (taps hi / lo)
            //      of the hi-piece accumulator and the same of the lo-piece accumulator ----
            {
                const int m = 32 * role + lane;
                const int mic = m >> 4, par = (m >> 3) & 1, r8 = m & 7;
#pragma unroll
                for (int slot = 0; slot < kSlots; ++slot) {
                    if (clip0 + slot >= B) continue;
                    if (h == 0) mbar_wait(smem_u32(sm.mbar + slot * 2 + buf), (uint32_t)((J >> 1) & 1));
                    tc_fence_after();
                    uint32_t r[32];
                    tc_ld32(tmem_cols + (uint32_t)((slot * 2 + buf) * kColsPerTile) + ((uint32_t)(32 * role) << 16), r);
                    if ((r8 >> 2) == h && mic < M) {
                        // even stream -> Q at odd times, odd stream -> Q at even times
                        float *dst = sm.q + (((s & 1) * kSlots + slot) * C2 + M + mic) * kQRow + (par ? 0 : 32) + 8 * (r8 & 3);
                        float v[8];
#pragma unroll
                        for (int a = 0; a < 8; ++a)
                            v[a] = (__uint_as_float(r[a]) + __uint_as_float(r[16 + a])) + (__uint_as_float(r[8 + a]) + __uint_as_float(r[24 + a]));
                        reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
                        reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
                    }
                }
                tc_fence_before();
            }
            // ---- I[t] = x[(t - K/2) mod T] (x 2^14 like Q): warps 0, 1 serve clip slot 0, warps 2, 3 slot 1;
            //      a lane rebuilds 8 consecutive samples of one (microphone, parity) ----
            {
                const int gid = (role & 1) * 32 + lane, slot = role >> 1;
                if (gid < 8 * M && clip0 + slot < B) {
                    const int mic = gid >> 3, ipar = (gid >> 2) & 1, chunk = gid & 3;
                    float v[8];
                    if (s < g.tiles_is) {
                        // np.roll: the first K/2 in-phase samples are the clip's last ones (snn_beamformer.py:325)
                        const IN_T *clip = audio + (clip0 + slot) * T64 * M;
                        const float sc = sm.scale[slot] * kTapScale;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = kTile * s + 2 * (8 * chunk + i) + ipar;
                            float x = 0.f;
                            if (t < T) {
                                int srci = (t - p.half) % T;
                                if (srci < 0) srci += T;
                                x = to_f32<IN_T>(clip[(long long)srci * M + mic]) * sc;
                            }
                            v[i] = x;
                        }
                    } else {
                        // I[2n] = E[n - H], I[2n+1] = O[n - H + 1]
                        const int q0 = (kTile / 2) * s + 8 * chunk - g.H + ipar;
                        const int pos = q0 % g.R;
                        const __half *hi = reinterpret_cast<const __half *>(sm.ring + (size_t)((slot * 2 + 0) * C2 + 2 * mic + ipar) * g.pitch_b) + pos;
                        const __half *lo = reinterpret_cast<const __half *>(sm.ring + (size_t)((slot * 2 + 1) * C2 + 2 * mic + ipar) * g.pitch_b) + pos;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = (__half2float(hi[i]) + __half2float(lo[i])) * kTapScale;
                    }
                    float *dst = sm.q + (((s & 1) * kSlots + slot) * C2 + mic) * kQRow + ipar * 32 + 8 * chunk;
                    reinterpret_cast<float4 *>(dst)[0] = make_float4(v[0], v[1], v[2], v[3]);
                    reinterpret_cast<float4 *>(dst)[1] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
        }
        // ---- staging -> rings for the tile of this step: a thread owns two consecutive samples of one stream
        //      (one 32-bit store per piece).  The positions written were last read by the MMAs of tile J - 1, whose
        //      completion every serial warp has observed in its epilogue above (this step or the one before). ----
        {
            int slot, J;
            if (tile_of_step(k, NJ, clip0 + 1 < B, slot, J)) {
                mbar_wait(smem_u32(sm.mbar + 4), tma_phase);
                tma_phase ^= 1u;
                const IN_T *stin = reinterpret_cast<const IN_T *>(sm.stin);
                const int f0 = kMac * J;
                const int pb = (kTile * J) % g.R;
                unsigned char *rhi = sm.ring + (size_t)((slot * 2 + 0) * C2) * g.pitch_b;
                unsigned char *rlo = sm.ring + (size_t)((slot * 2 + 1) * C2) * g.pitch_b;
                const float s_x = sm.scale[slot];
                const float *cin = sm.carry + (slot * 2 + ((J + 1) & 1)) * 8;       // x[128 J - 1]: last frame of tile J - 1
                const int t128 = role * 32 + lane;
                for (int q = t128; q < M * kTile; q += 128) {
                    const int mic = q % M, rest = q / M;
                    const int par = rest & 1, pi = rest >> 1;
                    const int fa = 4 * pi - par, fb = fa + 2;          // frames of the pair inside the tile (O[n] = x[2n-1])
                    float u0, u1;
                    if (fa >= 0) u0 = (f0 + fa < T) ? to_f32<IN_T>(stin[fa * M + mic]) * s_x : 0.f;
                    else u0 = J > 0 ? cin[mic] : 0.f;
                    u1 = (f0 + fb < T) ? to_f32<IN_T>(stin[fb * M + mic]) * s_x : 0.f;
                    const __half2 h2 = __floats2half2_rn(u0, u1);
                    const float2 hf = __half22float2(h2);
                    const __half2 l2 = __floats2half2_rn(u0 - hf.x, u1 - hf.y);
                    const int pos = pb + 2 * pi;
                    const size_t off = (size_t)(2 * mic + par) * g.pitch_b + 2 * pos;
                    *reinterpret_cast<__half2 *>(rhi + off) = h2;
                    *reinterpret_cast<__half2 *>(rlo + off) = l2;
                    if (pos < kMirror) {
                        *reinterpret_cast<__half2 *>(rhi + off + 2 * g.R) = h2;
                        *reinterpret_cast<__half2 *>(rlo + off + 2 * g.R) = l2;
                    }
                }
                if (t128 < M) sm.carry[(slot * 2 + (J & 1)) * 8 + t128] = (f0 + kMac - 1 < T) ? to_f32<IN_T>(stin[(kMac - 1) * M + t128]) * s_x : 0.f;
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(sm.mbar + 5));
            }
        }
    }
};

// ============ band-pass warp: SOS cascade + running sum + sign / zero masks, lane = slot*16 + channel ============
// Inputs of sub-tile k-2 come from its q rows (even-time samples, then odd-time samples), in-phase and quadrature alike.
template <int MM, typename HOOK>
__device__ __forceinline__ void bandpass_role(const TcSmem &sm, const ChainParams &p, long long clip0, long long B,
                                              long long T64, int lane, int k_last, HOOK hook) {
    const int M = MM ? MM : p.M, C2 = 2 * M;
    const int T = (int)T64;
    const int c_slot = lane >> 4, c_ch = lane & 15;
    const bool c_valid = c_ch < C2 && clip0 + c_slot < B;
    Sos2 sos;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        sos.b0[k] = p.sos[k][0]; sos.b1[k] = p.sos[k][1]; sos.b2[k] = p.sos[k][2];
        sos.a1[k] = p.sos[k][3]; sos.a2[k] = p.sos[k][4];
    }
    BiquadState bq; biquad_reset(bq);
    float csum = 0.f;
    ROLE_TIMER_DECL;

    for (int k = -1; k <= k_last; ++k) {
        hook(k);
        const int kc = k - 2;
        const int t0 = kc * kTile;
        if (kc >= 0 && t0 < T && c_valid) {
#pragma unroll 1
            for (int sg = 0; sg < kSegsPerTile; ++sg) {
                const int ts = t0 + sg * kSeg;            // first sample of this segment
                if (ts >= T) break;
                float *cs = sm.cs + ((kc & 1) * kSegsPerTile + sg) * kSeg * 32 + lane;
                unsigned int *sgm = sm.seg + ((kc & 1) * kSegsPerTile + sg) * 3 * 32 + lane;
                const float *xe = sm.q + (((kc & 1) * kSlots + c_slot) * C2 + c_ch) * kQRow + (kSeg / 2) * sg;
                const float *xo = xe + kTile / 2;
                const float carry = csum;
                unsigned int neg = 0u, zero = 0u;
                // sample by sample with explicit sign / zero masks (ragged segments, exact zeros)
                auto slow_segment = [&](int nvalid) {
#pragma unroll 1
                    for (int i = 0; i < nvalid; ++i) {
                        const float z = biquad2_step(sos, bq, (i & 1) ? xo[i >> 1] : xe[i >> 1]);
                        csum += z;
                        cs[i * 32] = csum;
                        neg |= (__float_as_uint(z) >> 31) << (31 - i);
                        zero |= (z == 0.f ? 1u : 0u) << (31 - i);
                    }
                };
                if (ts + kSeg <= T) {
                    const BiquadState bq0 = bq;
                    float zmin = 1.f;                   // smallest |z| of the segment: exact zeros are rare (silence)
                    float4 en = *reinterpret_cast<const float4 *>(xe), on = *reinterpret_cast<const float4 *>(xo);
#pragma unroll 1
                    for (int o = 0; o < kSeg / 8; ++o) {
                        const float xc[8] = {en.x, on.x, en.y, on.y, en.z, on.z, en.w, on.w};
                        if (o + 1 < kSeg / 8) {         // inputs of the next group: their latency hides behind this one
                            en = *reinterpret_cast<const float4 *>(xe + 4 * (o + 1));
                            on = *reinterpret_cast<const float4 *>(xo + 4 * (o + 1));
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float z = biquad2_step(sos, bq, xc[i]);
                            csum += z;
                            cs[(8 * o + i) * 32] = csum;
                            neg = __funnelshift_l(__float_as_uint(z), neg, 1);
                            zmin = fminf(zmin, fabsf(z));
                        }
                    }
                    if (zmin == 0.f) {                  // redo this lane's segment for its zero mask (same arithmetic)
                        bq = bq0; csum = carry; neg = 0u;
                        slow_segment(kSeg);
                    }
                } else {
                    slow_segment(T - ts);
                }
                sgm[0] = neg; sgm[32] = zero; sgm[64] = __float_as_uint(carry);
            }
        }
        ROLE_BARRIER();
    }
    ROLE_TIMER_FLUSH(kRoleBandpass);
}

// GROUPS = 2: warps 0..3 / 4..7 are the serial roles of group 0 / 1 (warp id mod 4 = role = tensor-memory lane
// quarter), warps 8, 9 the producers.  GROUPS = 1: warps 0..3 + producer warp 4.
template <typename IN_T, int MM, int GROUPS>
__global__ void __launch_bounds__(kGThreads * GROUPS, 1)
k_fused_tc(const IN_T *__restrict__ audio, const float *__restrict__ taps, const double *__restrict__ Wd,
           int8_t *__restrict__ spikes, float *__restrict__ power, int32_t *__restrict__ doa,
           int32_t *__restrict__ flags, unsigned int *__restrict__ sm_slots,
           const __grid_constant__ ChainParams p, const __grid_constant__ TcGeom g, long long B, long long T) {
    extern __shared__ __align__(128) unsigned char smem_all[];
    __shared__ long long s_pair[GROUPS];
    __shared__ uint32_t s_tmem;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = MM ? MM : p.M, C2 = 2 * M;
    const int group = warp < 4 * GROUPS ? warp >> 2 : warp - 4 * GROUPS;
    const int role = warp < 4 * GROUPS ? warp & 3 : kRoleProducer;
    const int tid = role * 32 + lane;       // thread index inside the group
    const int bar_id = 1 + group;
    auto group_sync = [&]() { tile_barrier(bar_id, kGThreads); };

    unsigned char *smem_raw = smem_all + (size_t)group * g.smem_group;
    TcSmem sm;
    sm.ring = smem_raw + g.off_ring;
    sm.stin = smem_raw + g.off_stin;
    sm.q = reinterpret_cast<float *>(smem_raw + g.off_q);
    sm.vms = reinterpret_cast<__half *>(smem_raw + g.off_vm);
    sm.cs = reinterpret_cast<float *>(smem_raw + g.off_cs);
    sm.seg = reinterpret_cast<unsigned int *>(smem_raw + g.off_seg);
    sm.clus = reinterpret_cast<int *>(smem_raw + g.off_clus);
    sm.bits = reinterpret_cast<unsigned int *>(smem_raw + g.off_bits);
    sm.stage = reinterpret_cast<int8_t *>(smem_raw + g.off_stage);
    sm.mbar = reinterpret_cast<unsigned long long *>(smem_raw + g.off_misc);
    sm.scale = reinterpret_cast<float *>(smem_raw + g.off_misc + 64);
    sm.amax = reinterpret_cast<unsigned int *>(smem_raw + g.off_misc + 72);
    sm.amax_next = reinterpret_cast<unsigned int *>(smem_raw + g.off_misc + 80);
    sm.carry = reinterpret_cast<float *>(smem_raw + g.off_misc + 128);
    sm.gram = reinterpret_cast<double *>(smem_raw + g.off_ring);     // [kSlots][16][16], clip epilogue only (the rings are dead then)
    sm.dbg = sm_slots;
    sm.bar_id = bar_id;
    sm.bar_threads = kGThreads;
    sm.rec = GROUPS * (int)blockIdx.x + group;
    double *red_v = reinterpret_cast<double *>(smem_raw + g.off_cs);         // [kGThreads], clip epilogue only
    int *red_i = reinterpret_cast<int *>(smem_raw + g.off_cs + kGThreads * sizeof(double));
    __half *tapsB = reinterpret_cast<__half *>(smem_all + (size_t)GROUPS * g.smem_group);

    // ---- once per CTA: tap matrix, barriers, tensor memory ----
    {
        // B[n][e] = g'[a + lag - e] x 2^14, n = 8 piece + a; g'[d0 + j] = taps[j].  Canonical K-major no-swizzle
        // layout: element (n, e) at core matrix ((n / 8) (2 ksteps) + e / 8), row n % 8, column e % 8
        const int Kp = 16 * g.ksteps;
        for (int e = threadIdx.x; e < kMmaN * Kp; e += blockDim.x) {
            const int n = e / Kp, kk = e - n * Kp;
            const int j = (n & 7) + g.lag - kk - g.d0;
            float v = 0.f;
            if (j >= 0 && j < p.n_taps) {
                const float gs = taps[j] * kTapScale;
                const float hi = __half2float(__float2half_rn(gs));
                v = (n >> 3) == 0 ? hi : gs - hi;
            }
            tapsB[((n >> 3) * 2 * g.ksteps + (kk >> 3)) * 64 + (n & 7) * 8 + (kk & 7)] = __float2half_rn(v);
        }
        if (tid == 0) {
            mbar_init(smem_u32(sm.mbar + 4), 1);
            mbar_init(smem_u32(sm.mbar + 5), 4);        // one arrival per serial warp
            for (int i = 0; i < 4; ++i) mbar_init(smem_u32(sm.mbar + i), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(4 * kColsPerTile * GROUPS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    const uint32_t tmem_cols = s_tmem + (uint32_t)(group * 4 * kColsPerTile);
    uint32_t tma_phase = 0, ring_phase = 0;

    const int NT = (int)((T + kTile - 1) / kTile);
    const int NJ = (int)((T + kMac - 1) / kMac);
    const int k_last = NT + g.dtile;    // the Gram warp runs dtile + 1 steps behind
    const long long npairs = (B + kSlots - 1) / kSlots;
    const int ring_words = (4 * C2 + 2) * g.pitch_b / 4;

    // Clip pairs are handed out dynamically, one ahead: while a group works on a pair its producer warp scans the
    // next one for its largest magnitude (the fp16 scale of float32 clips)
    if (tid == 0) s_pair[group] = (long long)atomicAdd(sm_slots + kSlotPair, 1u);
    if (tid < kSlots) sm.amax_next[tid] = 0xffffffffu;
    group_sync();
    long long pair = s_pair[group];
    while (pair < npairs) {
        group_sync();
        if (tid == 0) s_pair[group] = (long long)atomicAdd(sm_slots + kSlotPair, 1u);
        if (tid < kSlots) sm.amax[tid] = 0u;
        group_sync();
        const long long next_pair = s_pair[group];
        const long long clip0 = pair * kSlots;
        {   // zero the rings (samples before the clip start are zeros: lfilter's zero state), the spike bits
            // and the membrane tiles (columns of unused lanes stay zero)
            unsigned int *r4 = reinterpret_cast<unsigned int *>(sm.ring);
            for (int i = tid; i < ring_words; i += kGThreads) r4[i] = 0u;
            for (int i = tid; i < 2 * kRingWords * 32; i += kGThreads) sm.bits[i] = 0u;
            for (int i = tid; i < 2 * 2 * kVmRows * kVmPitch / 2; i += kGThreads) reinterpret_cast<unsigned int *>(sm.vms)[i] = 0u;
            if (tid == 0) {
                for (int i = 0; i < 4; ++i) { mbar_inval(smem_u32(sm.mbar + i)); mbar_init(smem_u32(sm.mbar + i), 1); }
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
        }
        const bool scanned = sm.amax_next[0] != 0xffffffffu;        // (written before the barriers above)
        if (sizeof(IN_T) == 4 && !scanned) {
            // first pair of this group (or an unaligned batch): largest magnitude of each clip, all threads
            for (int s = 0; s < kSlots; ++s) {
                if (clip0 + s >= B) break;
                const float *c = reinterpret_cast<const float *>(audio) + (clip0 + s) * T * M;
                const long long n = T * M;
                float mx = 0.f;
                long long i0 = 0;
                if ((reinterpret_cast<uintptr_t>(c) & 15) == 0) {
                    const float4 *c4 = reinterpret_cast<const float4 *>(c);
                    const long long n4 = n >> 2;
#pragma unroll 8
                    for (long long i = tid; i < n4; i += kGThreads) {
                        const float4 v = __ldg(c4 + i);
                        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
                    }
                    i0 = n4 << 2;
                }
                for (long long i = i0 + tid; i < n; i += kGThreads) mx = fmaxf(mx, fabsf(c[i]));
                unsigned int mb = __float_as_uint(mx);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned int ot = __shfl_xor_sync(0xffffffffu, mb, o); mb = ot > mb ? ot : mb; }
                if (lane == 0) atomicMax(sm.amax + s, mb);
            }
        }
        fence_proxy_async();
        group_sync();
        if (tid < kSlots) sm.scale[tid] = clip_scale<IN_T>(scanned ? sm.amax_next[tid] : sm.amax[tid]);
        group_sync();

        const StepHook<IN_T, MM> hook{sm, p, g, audio, clip0, B, T, role, lane, NT, NJ, tmem_cols, tma_phase};
        if (role == kRoleProducer)
            producer_role<IN_T, MM>(sm, p, g, audio, clip0, B, T, lane, NJ, k_last, tmem_cols, smem_u32(tapsB), ring_phase,
                                    next_pair < npairs ? next_pair * kSlots : -1);
        else if (role == 0) bandpass_role<MM>(sm, p, clip0, B, T, lane, k_last, hook);
        else if (role == 1) rzcc_role<TcSmem, kRingWords>(sm, p, flags, clip0, B, T, M, lane, k_last, hook);
        else if (role == 2) neuron_role<TcSmem, TcGeom, kRingWords>(sm, p, g, clip0, B, T, M, lane, k_last, hook);
        else gram_role<TcSmem, TcGeom>(sm, g, spikes, clip0, B, T, M, lane, k_last, hook);
        group_sync();
        // ---- clip epilogue: power[g] = w_g^T C w_g / T (float64), DoA = first argmax ----
        const double inv_T = 1.0 / (double)T;
        for (int s = 0; s < kSlots; ++s) {
            const long long clip = clip0 + s;
            if (clip >= B) break;
            const double *Cd = sm.gram + s * 256;
            double best = -1.0; int besti = 0x7fffffff;
            for (int gg = tid; gg < p.G; gg += kGThreads) {
                double w[2 * kRows];
#pragma unroll
                for (int c = 0; c < 2 * kRows; ++c) w[c] = c < C2 ? Wd[(long long)c * p.G + gg] : 0.0;
                double accp = 0.0;
#pragma unroll 2
                for (int r = 0; r < C2; ++r) {
                    double rr = 0.0;
#pragma unroll
                    for (int c = 0; c < 2 * kRows; ++c)
                        if (c < C2) rr = fma(Cd[r * 16 + c], w[c], rr);
                    double wr = 0.0;
#pragma unroll
                    for (int c = 0; c < 2 * kRows; ++c) wr = c == r ? w[c] : wr;
                    accp = fma(wr, rr, accp);
                }
                accp *= inv_T;
                if (power) power[clip * p.G + gg] = (float)accp;
                if (accp > best) { best = accp; besti = gg; }
            }
            red_v[tid] = best; red_i[tid] = besti;
            group_sync();
            if (tid < 32) {
                for (int i = tid + 32; i < kGThreads; i += 32) {
                    const double ov = red_v[i]; const int oi = red_i[i];
                    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
                }
                if (tid == 0 && doa) doa[clip] = besti;
            }
            group_sync();
        }
        pair = next_pair;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(4 * kColsPerTile * GROUPS) : "memory");
}

static bool make_geom(const ChainParams &p, int dtype, int groups, TcGeom &g) {
    g = TcGeom{};
    g.d0 = (p.tap_first - 1) / 2;
    g.H = p.half / 2;
    const int ntd = g.d0 + p.n_taps;                    // dense polyphase taps incl. leading zeros
    g.lag = (ntd - 1 + 15) / 16 * 16;
    g.ksteps = (g.lag + 8 + 15) / 16;
    int need = g.lag + kTile;                           // MMA window of the tile being filled
    if (g.H + 2 * kTile > need) need = g.H + 2 * kTile; // in-phase read-back of the serial warps
    g.R = (need + kTile - 1) / kTile * kTile;
    g.pitch_b = 2 * (g.R + kMirror);
    g.dtile = 4 + (rzcc_lag(p.w) - 1 + kTile - 1) / kTile;
    g.tiles_is = (p.half + kTile - 1) / kTile;
    const int esz = dtype == MICLOC_I16 ? 2 : 4;
    int off = 0;
    g.off_ring = off; off += (4 * p.C2 + 2) * g.pitch_b;
    off = (off + 127) & ~127;
    g.off_stin = off; off += kMac * p.M * esz;
    off = (off + 15) & ~15;
    g.off_q = off; off += 2 * kSlots * p.C2 * kQRow * (int)sizeof(float);
    g.off_vm = off; off += 2 * 2 * kVmRows * kVmPitch * (int)sizeof(__half);
    g.off_cs = off; off += 2 * kSegsPerTile * kSeg * 32 * (int)sizeof(float);
    g.off_seg = off; off += 2 * kSegsPerTile * 3 * 32 * (int)sizeof(int);
    g.off_clus = off; off += 4 * kClusterMax * 32 * (int)sizeof(int);
    g.off_bits = off; off += 2 * kRingWords * 32 * (int)sizeof(int);
    g.off_stage = off; off += (2 * kSlots * kTile * p.C2 + 15) & ~15;
    g.off_misc = off; off += 256;
    g.smem_group = (off + 127) & ~127;
    g.smem_bytes = groups * g.smem_group + kMmaN * 16 * g.ksteps * (int)sizeof(__half);
    return true;
}

bool fused_supported(const ChainParams &p) {
    // Hilbert-type kernel (every other tap zero) whose taps sit at odd lags and whose in-phase delay K/2 is even,
    // 2-section band-pass, up to 8 microphones
    return p.tap_stride == 2 && (p.tap_first & 1) == 1 && (p.half & 1) == 0 && p.M <= kRows && p.nsec == 2 &&
           (p.n_taps % 8) == 0;
}

template <typename IN_T, int MM, int GROUPS>
static int launch_t(const ChainParams &p, const TcGeom &g, const float *d_taps, const double *d_Wd, const IN_T *audio,
                    long long B, long long T, int8_t *spikes, float *power, int32_t *doa, int32_t *flags,
                    unsigned int *sm_slots, int sm_count, cudaStream_t st) {
    auto kern = k_fused_tc<IN_T, MM, GROUPS>;
    MICLOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes));
    MICLOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    long long grid = sm_count;
    const long long npairs = (B + kSlots - 1) / kSlots;
    const long long want = (npairs + GROUPS - 1) / GROUPS;
    if (grid > want) grid = want;
    MICLOC_CUDA(cudaMemsetAsync(sm_slots, 0, kSlotResetWords * sizeof(unsigned int), st));   // clip-pair counter restarts per launch
    kern<<<(unsigned)grid, kGThreads * GROUPS, g.smem_bytes, st>>>(audio, d_taps, d_Wd, spikes, power, doa, flags, sm_slots, p, g, B, T);
    count_launch(1);
    MICLOC_CUDA(cudaGetLastError());
    return MICLOC_OK;
}

int launch_fused(const ChainParams &p, const float *d_taps, const double *d_Wd, const void *audio, int dtype,
                 long long B, long long T, int8_t *spikes, float *power, int32_t *doa, int32_t *flags,
                 unsigned int *sm_slots, int sm_count, cudaStream_t st) {
    if (!tc::fused_supported(p)) return set_error(MICLOC_ERR_UNSUPPORTED, "tensor-core fused kernel: unsupported chain geometry");
    TcGeom g;
    int groups = 2;
    if (const char *e = getenv("MICLOC_FUSED_GROUPS")) { if (atoi(e) == 1) groups = 1; }
    make_geom(p, dtype, groups, g);
    if (groups == 2 && g.smem_bytes > 227 * 1024) { groups = 1; make_geom(p, dtype, groups, g); }
    if (g.smem_bytes > 227 * 1024)
        return set_error(MICLOC_ERR_UNSUPPORTED, "tensor-core fused kernel needs %d B of shared memory", g.smem_bytes);
    // the spike-bit ring must hold the back warp's oldest read and the front warp's newest write
    if (kTile * (g.dtile - 2) + p.nL + kSeg > kRingWords * 32)
        return set_error(MICLOC_ERR_UNSUPPORTED, "robust_width %d / neuron length %d exceed the tensor-core kernel's spike ring", p.w, p.nL);
    if (kSlots * 256 * (int)sizeof(double) > 4 * p.C2 * g.pitch_b)
        return set_error(MICLOC_ERR_UNSUPPORTED, "shared-memory rings too small for the epilogue");
    if (T + 16 * kTile >= (1ll << 31)) return set_error(MICLOC_ERR_SHAPE, "T too large for the fused kernel");
#define MICLOC_TC_CASE(IN, MMV)                                                                                   \
    do {                                                                                                          \
        if (groups == 2)                                                                                          \
            return launch_t<IN, MMV, 2>(p, g, d_taps, d_Wd, (const IN *)audio, B, T, spikes, power, doa, flags,   \
                                        sm_slots, sm_count, st);                                                  \
        return launch_t<IN, MMV, 1>(p, g, d_taps, d_Wd, (const IN *)audio, B, T, spikes, power, doa, flags,       \
                                    sm_slots, sm_count, st);                                                      \
    } while (0)
    const bool i16 = dtype == MICLOC_I16;
    if (p.M == 7) { if (i16) MICLOC_TC_CASE(int16_t, 7); else MICLOC_TC_CASE(float, 7); }
    if (i16) MICLOC_TC_CASE(int16_t, 0); else MICLOC_TC_CASE(float, 0);
#undef MICLOC_TC_CASE
}

}  // namespace tc
}  // namespace micloc
