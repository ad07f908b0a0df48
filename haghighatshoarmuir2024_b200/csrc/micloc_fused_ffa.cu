// micloc_fused_ffa.cu -- fast-FIR variant of the fused hot-path kernel (MICLOC_FUSED_FIR=ffa; micloc_fused.cu is the default): raw audio in, spikes + per-DoA
// power + DoA index out; nothing else touches HBM.
//
// Reference sites: micloc/snn_beamformer.py:283-370 and the callers' power/argmax
// paper_plots/target_snn_localization.py:462-464.
//
// A clip-pair GROUP of eight warps owns kSlots = 2 clips at a time and walks them in time tiles of
// kTile = 64 samples.  The warps are specialised BY FUNCTION (every role serves both clips with all its
// lanes, so that the serial depth of each role per tile is short) and run as a software pipeline, one
// named barrier per tile (iteration k):
//
//   FIR warps x4 tile k-1   recombine the sub-filter results of the previous step into the quadrature tile
//                tile k     STHT quadrature FIR as three half-length sub-filters, each split into three tap
//                           thirds ("pieces" of 5 tap blocks); every lane runs two pieces of 8 consecutive output
//                           pairs with a sliding register window (packed FFMA2) and trades partial sums with
//                           two partner lanes by shuffles; the finished vectors go to shared memory
//   band-pass    tile k-2   one lane per (clip, channel): SOS band-pass recurrence, running sum, sign masks of
//                           every 32-sample segment -> shared memory; the in-phase input x[(t - K/2) mod T]
//                           (np.roll) comes from the R0 / R1 rings (clip tail: global memory)
//   RZCC         tile k-3   one lane per (clip, channel): the masks are turned into RZCC
//                           candidates and resolved (find_peaks distance rule) into a bit-packed
//                           spike ring
//   neuron       tile k-d   (d = the latency of the exact find_peaks decision) one lane per (clip,
//                           channel): alpha-kernel neuron recurrences driven by the final spike
//                           bits -> membrane tile + int8 spike tile in shared memory
//   Gram+loader  tile k+1   audio (HBM) -> the three sub-sequence rings of every microphone (loads issued at
//                           the start of the step, stored at its end)
//                tile k-d-1 C += V V^T of the membrane tile on the tensor cores (fp16 hi / lo split), int8 spike
//                           raster of the tile -> HBM
//   clip end                power[g] = w_g^T C w_g / T (float64), DoA = first argmax.
//
// Fast FIR.  The Hilbert kernel has taps only at every other lag, h[k0 + 2j] = c_j: on pairs of
// consecutive samples r[i] = (x[2i-k0], x[2i+1-k0]) and output pairs Y[p] = (Q[2p], Q[2p+1]) it is the dense
// FIR Y[p] = sum_j c_j r[p-j] (both samples of a pair take the same tap).  Splitting
// taps and pairs by parity (A_n = c_2n, B_n = c_2n+1, R0[m] = r[2m], R1[m] = r[2m+1]) gives
//      Y[2m]   = (A*R0)[m] + (B*R1)[m-1]
//      Y[2m+1] = (A*R1)[m] + (B*R0)[m]  =  ((A+B)*(R0+R1))[m] - (A*R0)[m] - (B*R1)[m]
// i.e. THREE convolutions of half length U = A*R0, V = B*R1, W = (A+B)*(R0+R1) instead of four: a quarter
// of the multiply-adds of the direct form is never executed (a 2-parallel fast FIR algorithm; the float32
// error against the float64 reference stays at the direct form's ~1e-6 relative).  The loader keeps R0, R1
// and S = R0 + R1 of every microphone in three rings, so that each sub-filter is the same register-blocked
// sliding-window loop on its own ring with its own tap array.  The 12 M tasks (sub-filter x row x half tile) of
// a group-tile do not fill a whole number of warps; cut into tap thirds they fill four warps to 98 %, and a
// warp issues 10 tap blocks per tile where the direct form issues 15 (see fir_role).
//
// Issue slots, not the FMA pipe alone, bound this kernel: an FFMA2 occupies two issue slots of its SM
// sub-partition (tools/sched_probe.py), so every instruction of the serial roles displaces half an FFMA2.
// One CTA holds two groups (GROUPS = 2, sixteen warps, one CTA per SM), one FIR warp of each group per
// sub-partition (see k_fused).
//
// Status (round 1, B200): parity-green; the FIR warps alone sustain 315k clips/s (270k for the direct form), the
// whole kernel 153-158k against the default kernel's 172k: its serial roles run slower here (band-pass 9500
// cycles per tile against 5400) -- larger code (5400 against 4700 instructions: "no instruction" stalls 0.75
// against 0.13 per issue), more band-pass instructions (ring address arithmetic), conflicted shared-memory loads.
#include <cstdlib>

#include "micloc_fused_common.cuh"

namespace micloc {
namespace ffa {

constexpr int kRows = 7;       // most microphones per clip the lane maps cover
constexpr int kRingWords = 16; // spike-bit ring: 16 words of 32 samples per channel and polarity
constexpr int kFirWarps = 4;
constexpr int kPieceBlocks = 5;    // tap blocks of 8 per piece: a sub-filter of 15 blocks is three pieces
constexpr int kTileM = kTile / 4;      // groups of four samples (= one pair of each sub-sequence) per tile
constexpr int kShiftP = 7;     // ring coordinate of pair m is (m + kShiftP) mod ring_p: window chunks start 8-aligned
constexpr int kUvPitch = 2 * kTileM + 4;   // floats per row of the U / V / W hand-over scratch (+ one pair of the previous tile)

struct FusedGeom {
    int ring_p;      // pairs per sub-sequence ring (multiple of 8)
    int pitch_x;     // floats per ring; pitch_x / 4 is odd (LDS.128 of consecutive rings hit distinct bank groups)
    int nblk;        // tap blocks of 8 per sub-filter (multiple of 3: walked in groups of three)
    int tap_pitch;   // floats per tap array (A, B, A+B)
    int dtile;       // the neuron warp runs dtile tiles behind the pipeline step (RZCC decision latency)
    int fir_blocks;  // debug (MICLOC_FUSED_FIRBLOCKS): tap blocks each FIR warp really computes (0 = all; results are garbage)
    int skip;        // debug (MICLOC_FUSED_SKIP): bit r set = role r only attends the tile barriers (results are garbage)
    unsigned char role_map[16];   // GROUPS = 2: (group << 3 | role) of warp w (sub-partition w % 4), see k_fused
    int off_x, off_q, off_uv, off_vm, off_cs, off_seg, off_clus, off_bits, off_stage;   // byte offsets in dynamic smem
    int smem_bytes;
};

// Packed FFMA2 (fma.rn.f32x2): one instruction per tap and output pair.  With three operands from non-uniform
// registers (accumulator pair, window pair, tap) an FFMA2 issues every ~2.6 cycles per sub-partition in this
// loop, not every 2 as with a uniform-register multiplier (the form the FP32 peak is measured with); the scalar
// FFMA form of the same loop was measured 2 % slower end to end (tools/sched_probe.py roles 9-11).
struct Chunk { unsigned long long p[8]; };   // 8 consecutive pairs of one sub-sequence ring

__device__ __forceinline__ void load_chunk(Chunk &c, const float *row, int coord) {
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(row + coord);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const ulonglong2 u = src[v];
        c.p[2 * v] = u.x; c.p[2 * v + 1] = u.y;
    }
}

// 8 taps x 8 output pairs: acc[ip] += g[jj] * W[ip + 7 - jj], W = lo pairs 0..7 | hi pairs 0..6
struct Taps8 { float4 a, b; };
__device__ __forceinline__ void load_taps(Taps8 &t, const float *__restrict__ taps8) {
    t.a = *reinterpret_cast<const float4 *>(taps8);
    t.b = *reinterpret_cast<const float4 *>(taps8 + 4);
}
__device__ __forceinline__ void fir_block(unsigned long long (&acc)[8], const Chunk &lo, const Chunk &hi,
                                          const Taps8 &t) {
    const float g[8] = {t.a.x, t.a.y, t.a.z, t.a.w, t.b.x, t.b.y, t.b.z, t.b.w};
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

struct FusedSmem {
    float *taps;            // [3: A, B, A+B][tap_pitch]
    float *xs;              // [3: R0, R1, R0+R1][kSlots*M rows][pitch_x] sub-sequence rings
    float *qs;              // [2 tiles][kSlots*M][kQPitch]: finished quadrature tiles
    float *uv;              // [3: U, V, W][kSlots*M][kUvPitch]: sub-filter results of one tile, recombined one step later
    __half *vms;            // [2: hi, lo][2 tiles][kVmRows][kVmPitch] membrane tiles (x 2^14, fp16 split), channel-major
    float *cs;
    unsigned int *seg;      // [2 tiles][kTile/kSeg][3: neg mask, zero mask, carry][32 lanes]: band-pass -> RZCC hand-over
    int *clus;
    unsigned int *bits;     // [2 polarities][kRingWords][32 lanes]
    int8_t *stage;          // [2 tiles][kSlots][kTile][C2]
    double *gram;           // [kSlots][16][16], clip epilogue only (reuses the audio rings)
    unsigned int *dbg;      // sm_slots (debug counters behind the first 256 entries)
    int bar_id;             // named barrier of this clip-pair group
    int rec;                // index of this group's debug record (MICLOC_ROLE_TIMING builds)
};

// ======= loader (part of the Gram warp): audio tile k+1 (HBM) -> sub-sequence rings R0, R1, S = R0 + R1 =======
// lane = slot * 16 + group of four consecutive samples u = 4m - k0 + e, e < 4 (one pair of R0, one of R1);
// fetch() issues the global loads at the start of a pipeline step, store() puts them into the rings at its end
template <typename IN_T, int MM>
struct Loader {
    const IN_T *src;
    float *rows;
    int m, coord, M, T, k0, ring_p, pitch_x, f_stride;
    bool clip_ok;
    __device__ __forceinline__ void init(const FusedSmem &sm, const ChainParams &p, const FusedGeom &g,
                                         const IN_T *__restrict__ audio, long long clip0, long long B, long long T64, int lane) {
        M = MM ? MM : p.M; T = (int)T64; k0 = p.tap_first; ring_p = g.ring_p; pitch_x = g.pitch_x;
        const int slot = lane >> 4, gi = lane & 15;
        clip_ok = clip0 + slot < B;
        src = audio + (clip_ok ? clip0 + slot : clip0) * T64 * M;
        rows = sm.xs + slot * M * g.pitch_x;
        f_stride = kSlots * M * g.pitch_x;              // rings are [3: R0, R1, S][clip row][pitch_x]
        m = gi;                                      // pair index of the tile being filled (tile 0 at k = -1)
        coord = (gi + kShiftP) % g.ring_p;           // its ring coordinate, kept incrementally
    }
    __device__ __forceinline__ void fetch(float (&v)[4][kRows]) const {
        const int u0 = 4 * m - k0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int u = u0 + e;
            const bool ok = u >= 0 && u < T;
            const IN_T *fr = src + (ok ? (long long)u * M : 0);
#pragma unroll
            for (int mm = 0; mm < kRows; ++mm) v[e][mm] = (ok && mm < M) ? to_f32<IN_T>(fr[mm]) : 0.f;
        }
    }
    __device__ __forceinline__ void store(const float (&v)[4][kRows]) {
#pragma unroll
        for (int mm = 0; mm < kRows; ++mm)
            if (mm < M) {
                float *r = rows + mm * pitch_x + 2 * coord;
                *reinterpret_cast<float2 *>(r) = make_float2(v[0][mm], v[1][mm]);
                *reinterpret_cast<float2 *>(r + f_stride) = make_float2(v[2][mm], v[3][mm]);
                *reinterpret_cast<float2 *>(r + 2 * f_stride) = make_float2(v[0][mm] + v[2][mm], v[1][mm] + v[3][mm]);
            }
        m += kTileM;
        coord += kTileM; if (coord >= ring_p) coord -= ring_p;
    }
};

// ======= FIR warps: STHT quadrature FIR of tile k as the three half-length sub-filters U, V, W, in tap thirds =======
// A task = 8 output pairs of one (chunk, sub-filter, row); its 15 tap blocks are three PIECES of 5 (a, b, c).  The 12 M
// tasks of a tile are split evenly over the four FIR warps (n = 3 M each, 21 for M = 7) and every lane runs two
// pieces per tile, so that a warp issues 10 tap blocks instead of 15.  With h = n / 2 the lanes form triples
//     lane j      (j < h): pieces a, b of task j        (one running sum)  + piece c from lane h + j
//     lane h + j         : piece a of task h + j, then piece c of task j   (handed on by shuffles)
//     lane 2h + j        : pieces b, c of task h + j    (one running sum)  + piece a from lane h + j
// (odd n: lanes 3h, 3h + 1 share the last task).  Consecutive lanes work on consecutive rows of the same sub-filter
// and mostly the same tap third: their window loads spread over the shared-memory banks (6.4 wavefronts per
// LDS.128 against 10.9 for consecutive pieces per lane).  The finished U, V, W vectors go to shared memory; the FIR
// warps recombine them into the quadrature tile at the start of the next pipeline step
// (Y[2m] = U[m] + V[m-1], Y[2m+1] = W[m] - U[m] - V[m]).
struct Piece { const float *row; const float *tp; int c0; bool on; };

__device__ __forceinline__ void run_piece(unsigned long long (&acc)[8], const Piece &q, int ringf) {
    Chunk A, Bq, Cq;
    Taps8 t0, t1;
    int cn = q.c0 + 16; if (cn >= ringf) cn -= ringf;
    load_chunk(Bq, q.row, cn);
    load_chunk(A, q.row, q.c0);
    cn = q.c0 - 16; if (cn < 0) cn += ringf;
    load_taps(t0, q.tp);
    load_chunk(Cq, q.row, cn); cn -= 16; if (cn < 0) cn += ringf;
    load_taps(t1, q.tp + 8);
    fir_block(acc, A, Bq, t0);
    load_chunk(Bq, q.row, cn); cn -= 16; if (cn < 0) cn += ringf;
    load_taps(t0, q.tp + 16);
    fir_block(acc, Cq, A, t1);
    load_chunk(A, q.row, cn); cn -= 16; if (cn < 0) cn += ringf;
    load_taps(t1, q.tp + 24);
    fir_block(acc, Bq, Cq, t0);
    load_chunk(Cq, q.row, cn);
    load_taps(t0, q.tp + 32);
    fir_block(acc, A, Bq, t1);
    fir_block(acc, Cq, A, t0);
}

template <int MM>
__device__ __forceinline__ void fir_role(const FusedSmem &sm, const ChainParams &p, const FusedGeom &g,
                                         long long clip0, long long B, int warp_f, int lane, int NT, int k_last, int fir_bar) {
    const int M = MM ? MM : p.M;
    const int rows = kSlots * M;
    const int ringf = 2 * g.ring_p;
    const int n_t = 3 * M, h = n_t / 2;
    // pieces of this lane (task within the warp, tap third) and the lanes it trades partial sums with
    int p_task[2] = {-1, -1}, p_third[2] = {0, 0};
    int role3 = 1;                      // 0: owns pieces a, b (+ c after round 1); 2: owns b, c (+ a after round 0); 1: hands on
    int src0 = lane, src1 = lane;       // source lanes of the shuffles after round 0 / round 1
    if (lane < h) { p_task[0] = p_task[1] = lane; p_third[0] = 0; p_third[1] = 1; role3 = 0; src1 = lane + h; }
    else if (lane < 2 * h) { p_task[0] = lane; p_third[0] = 0; p_task[1] = lane - h; p_third[1] = 2; }
    else if (lane < 3 * h) { p_task[0] = p_task[1] = lane - h; p_third[0] = 1; p_third[1] = 2; role3 = 2; src0 = lane - h; }
    else if ((n_t & 1) && lane == 3 * h) { p_task[0] = p_task[1] = 2 * h; p_third[0] = 0; p_third[1] = 1; role3 = 3; src0 = lane + 1; }
    else if ((n_t & 1) && lane == 3 * h + 1) { p_task[0] = 2 * h; p_third[0] = 2; }
    Piece pc[2];
    int t_f = 0, t_row = 0, t_chunk = 0;
    bool t_on = false;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const bool in = p_task[r] >= 0;
        const int tg = warp_f * n_t + (in ? p_task[r] : 0);    // task of the group-tile: (chunk, sub-filter) major, row minor
        const int cf = tg / rows, rowi = tg - cf * rows, chunk = cf / 3, f = cf - 3 * chunk;
        pc[r].on = in && clip0 + rowi / M < B;
        pc[r].row = sm.xs + (f * rows + rowi) * g.pitch_x;
        pc[r].tp = sm.taps + f * g.tap_pitch + 8 * kPieceBlocks * p_third[r];
        int c = (8 * chunk - 8 * kPieceBlocks * p_third[r]) % g.ring_p;
        if (c < 0) c += g.ring_p;
        pc[r].c0 = 2 * c;                                      // ring coordinate (floats) of block 0's window at tile 0
        // the task whose finished vector this lane holds after the exchanges
        if (((role3 == 0 || role3 == 3) && r == 0) || (role3 == 2 && r == 1)) { t_f = f; t_row = rowi; t_chunk = chunk; t_on = pc[r].on; }
    }
    float *t_dst = sm.uv + (t_f * rows + t_row) * kUvPitch + 4 + 16 * t_chunk;
    // recombination of this warp: 2M of the 8M (row, chunk, Y[2m] | Y[2m+1]) vectors, two lanes (four pairs each) per vector
    const int ct = warp_f * 2 * M + (lane >> 1), c_half = lane & 1;
    const int c_cg = ct >> 1, c_odd = ct & 1, c_row = (c_cg >> 1) < rows ? (c_cg >> 1) : 0, c_chunk = c_cg & 1;
    const bool c_on = lane < 4 * M && clip0 + c_row / M < B;
    const float *c_u = sm.uv + c_row * kUvPitch + 4 + 16 * c_chunk + 8 * c_half;
    const float *c_v = c_u + rows * kUvPitch, *c_w = c_v + rows * kUvPitch;
    const unsigned long long one2 = pack2(1.f, 1.f);
    ROLE_TIMER_DECL;

    for (int k = -1; k <= k_last; ++k) {
        // (a) quadrature tile k-1 from the U, V, W vectors the FIR warps left in shared memory one step ago
        if (k >= 1 && k - 1 < NT && c_on) {
            float *qrow = sm.qs + (((k - 1) & 1) * rows + c_row) * kQPitch + 32 * c_chunk + 16 * c_half;
            if (!c_odd) {
                float2 uu[4], vv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uu[i] = *reinterpret_cast<const float2 *>(c_u + 2 * i);
                    vv[i] = *reinterpret_cast<const float2 *>(c_v + 2 * i - 2);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float2 *>(qrow + 4 * i) = make_float2(uu[i].x + vv[i].x, uu[i].y + vv[i].y);
            } else {
                float4 uu[2], vv[2], ww[2];
#pragma unroll
                for (int v4 = 0; v4 < 2; ++v4) {
                    uu[v4] = reinterpret_cast<const float4 *>(c_u)[v4];
                    vv[v4] = reinterpret_cast<const float4 *>(c_v)[v4];
                    ww[v4] = reinterpret_cast<const float4 *>(c_w)[v4];
                }
#pragma unroll
                for (int v4 = 0; v4 < 2; ++v4) {
                    *reinterpret_cast<float2 *>(qrow + 8 * v4 + 2) =
                        make_float2((ww[v4].x - uu[v4].x) - vv[v4].x, (ww[v4].y - uu[v4].y) - vv[v4].y);
                    *reinterpret_cast<float2 *>(qrow + 8 * v4 + 6) =
                        make_float2((ww[v4].z - uu[v4].z) - vv[v4].z, (ww[v4].w - uu[v4].w) - vv[v4].w);
                }
            }
        }
        // every FIR warp has read the old vectors before any of them writes new ones
        asm volatile("bar.sync %0, %1;" ::"r"(fir_bar), "n"(32 * kFirWarps) : "memory");
        // (b) the pieces of tile k
        if (k >= 0 && k < NT) {
            unsigned long long acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0ull;
#pragma unroll 1
            for (int r = 0; r < 2; ++r) {
                // (one copy of the tap loop: the kernel's code footprint matters, its eight roles share the instruction cache)
                Piece q;
                q.row = r ? pc[1].row : pc[0].row; q.tp = r ? pc[1].tp : pc[0].tp;
                q.c0 = r ? pc[1].c0 : pc[0].c0; q.on = r ? pc[1].on : pc[0].on;
                if (q.on) run_piece(acc, q, ringf);
                const int src = r ? src1 : src0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned long long got = __shfl_sync(0xffffffffu, acc[i], src);
                    if (r == 0 ? role3 >= 2 : role3 == 0) ffma2(acc[i], got, one2);
                    if (r == 0 && role3 == 1) acc[i] = 0ull;
                }
            }
            if (t_on && role3 != 1) {
                if (t_f == 1 && t_chunk == 1)           // V[16(k-1) + 15] becomes the V[m-1] of the next tile's first pair
                    *reinterpret_cast<float2 *>(t_dst - 16 - 2) = *reinterpret_cast<const float2 *>(t_dst + 14);
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    float4 o;
                    unpack2(acc[2 * v4], o.x, o.y);
                    unpack2(acc[2 * v4 + 1], o.z, o.w);
                    reinterpret_cast<float4 *>(t_dst)[v4] = o;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) { pc[r].c0 += 2 * kTileM; if (pc[r].c0 >= ringf) pc[r].c0 -= ringf; }
        }
        ROLE_BARRIER();
    }
    ROLE_TIMER_FLUSH(warp_f);
}

// ============ band-pass warp: SOS cascade + running sum + sign / zero masks, lane = slot*16 + channel ============
template <typename IN_T, int MM>
__device__ __forceinline__ void bandpass_role(const FusedSmem &sm, const ChainParams &p, const FusedGeom &g,
                                              const IN_T *__restrict__ audio, long long clip0, long long B,
                                              long long T64, int lane, int k_last) {
    const int M = MM ? MM : p.M, C2 = 2 * M;
    const int T = (int)T64;
    const int c_slot = lane >> 4, c_ch = lane & 15;
    const bool slot_ok = clip0 + c_slot < B;
    const bool c_valid = c_ch < C2 && slot_ok;
    const bool c_inphase = c_ch < M;
    const IN_T *clip_audio = audio + (slot_ok ? clip0 + c_slot : clip0) * T64 * M;
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
        const int kc = k - 2;
        const int t0 = kc * kTile;
        if (kc >= 0 && t0 < T) {
#pragma unroll 1
            for (int sg = 0; sg < kSegsPerTile; ++sg) {
                const int ts = t0 + sg * kSeg;            // first sample of this segment
                float *cs = sm.cs + ((kc & 1) * kSegsPerTile + sg) * kSeg * 32 + lane;
                unsigned int *sgm = sm.seg + ((kc & 1) * kSegsPerTile + sg) * 3 * 32 + lane;
                if (ts >= T || !c_valid) continue;
                // in-phase input x[(t - K/2) mod T] (np.roll, snn_beamformer.py:325): from the R0 / R1 rings where the
                // segment lies inside their history, from global memory for the clip tail in front of the clip
                // (t < K/2) and for ragged segments; quadrature input: the FIR warps' tile
                int src0 = (ts - p.half) % T;             // (warp-uniform) source sample of the segment's first sample
                if (src0 < 0) src0 += T;
                const float *qp = sm.qs + (((kc & 1) * kSlots + c_slot) * M + (c_inphase ? 0 : c_ch - M)) * kQPitch + sg * kSeg;
                // sample u sits in pair m = (u + k0) >> 2 of ring (u + k0) >> 1 & 1, half (u + k0) & 1; a group of 8
                // samples that starts with (u + k0) & 3 == 1 is R0[m].y, R1[m], R0[m+1], R1[m+1], R0[m+2].x
                const int ua = ts - p.half + p.tap_first;
                const bool fast = ts + kSeg <= T && ts >= p.half && (ua & 3) == 1;      // (warp-uniform)
                int cseg = ((ua >> 2) + kShiftP) % g.ring_p;                              // ring coordinate of the segment's first pair
                const float *r0 = sm.xs + (c_slot * M + (c_inphase ? c_ch : 0)) * g.pitch_x;
                const float *r1 = r0 + kSlots * M * g.pitch_x;
                const float carry = csum;
                unsigned int neg = 0u, zero = 0u;
                // sample by sample with explicit sign / zero masks (ragged segments, clip-tail in-phase source, exact zeros)
                auto slow_segment = [&](int nvalid) {
                    int src = src0;
#pragma unroll 1
                    for (int i = 0; i < nvalid; ++i) {
                        const float x = c_inphase ? to_f32<IN_T>(clip_audio[(long long)src * M + c_ch]) : qp[i];
                        if (++src >= T) src = 0;
                        const float z = biquad2_step(sos, bq, x);
                        csum += z;
                        cs[i * 32] = csum;
                        neg |= (__float_as_uint(z) >> 31) << (31 - i);
                        zero |= (z == 0.f ? 1u : 0u) << (31 - i);
                    }
                };
                auto load_group = [&](int o, float (&x)[8]) {
                    if (c_inphase) {
                        int c = cseg + 2 * o; if (c >= g.ring_p) c -= g.ring_p;
                        const int c1 = c + 1 == g.ring_p ? 0 : c + 1, c2 = c1 + 1 == g.ring_p ? 0 : c1 + 1;
                        x[0] = r0[2 * c + 1];
                        const float2 a = *reinterpret_cast<const float2 *>(r1 + 2 * c);
                        const float2 b = *reinterpret_cast<const float2 *>(r0 + 2 * c1);
                        const float2 d = *reinterpret_cast<const float2 *>(r1 + 2 * c1);
                        x[1] = a.x; x[2] = a.y; x[3] = b.x; x[4] = b.y; x[5] = d.x; x[6] = d.y;
                        x[7] = r0[2 * c2];
                    } else {
                        const float4 a = reinterpret_cast<const float4 *>(qp)[2 * o], b = reinterpret_cast<const float4 *>(qp)[2 * o + 1];
                        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                    }
                };
                if (fast) {
                    const BiquadState bq0 = bq;
                    float zmin = 1.f;                   // smallest |z| of the segment: exact zeros are rare (silence)
                    float xn[8];
                    load_group(0, xn);
#pragma unroll 1
                    for (int o = 0; o < kSeg / 8; ++o) {
                        float xc[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) xc[i] = xn[i];
                        if (o + 1 < kSeg / 8) load_group(o + 1, xn);     // inputs of the next group: their latency hides behind this one
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
                    slow_segment(T - ts < kSeg ? T - ts : kSeg);
                }
                sgm[0] = neg; sgm[32] = zero; sgm[64] = __float_as_uint(carry);
            }
        }
        ROLE_BARRIER();
    }
    ROLE_TIMER_FLUSH(kRoleBandpass);
}

// ==== Gram warp: C += V V^T of the membrane tile k - dtile - 1 on the tensor cores, then that tile's int8 spike
// raster -> HBM.  The neuron warp leaves every membrane value (x 2^14) as an fp16 pair v = hi + lo (22 significant
// bits).  Per clip slot and 16 time samples one ldmatrix.x4 each fetches the m16n8k16 fragments of hi and lo of
// V^T (16 channels x 16 samples; the same registers serve as the "col" operand, the matrix is V V^T), and
// C += hi hi^T + hi lo^T + lo hi^T runs as three fp16 MMAs per 8-channel column block: exact products (the dropped
// lo lo^T is below 2^-22 relative) accumulated in float32 over kGramFlush tiles -- the tensor cores add with
// truncation, a long chain would bias the sum -- and then folded into float64 registers.
template <typename IN_T, int MM>
__device__ __forceinline__ void gram_role(const FusedSmem &sm, const ChainParams &p, const FusedGeom &g,
                                          const IN_T *__restrict__ audio, int8_t *__restrict__ spikes,
                                          long long clip0, long long B, long long T64, int lane, int NT, int k_last) {
    const int C2 = 2 * (MM ? MM : p.M);
    const int T = (int)T64;
    Loader<IN_T, MM> ld;
    ld.init(sm, p, g, audio, clip0, B, T64, lane);
    // ldmatrix row of this lane: matrix lane / 8 = (channels 0-7 | 8-15) x (samples 0-3 | 4-7) of a k-step
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

    for (int k = -1; k <= k_last; ++k) {
        // audio tile k+1: the loads are issued here and stored after the Gram work, their latency hides behind it
        const bool filling = k + 1 < NT && ld.clip_ok;
        float av[4][kRows];
        if (filling) ld.fetch(av);
        const int j = k - g.dtile - 1;
        const int u0 = j * kTile;
        const bool live = j >= 0 && u0 < T;
        if (live) {
            const __half *vm = sm.vms + ((j & 1) * kVmRows + lm_row) * kVmPitch + lm_t;
#pragma unroll
            for (int s = 0; s < kSlots; ++s) {
                const unsigned addr = (unsigned)__cvta_generic_to_shared(vm + s * 16 * kVmPitch);
                const unsigned lo_off = 2 * kVmRows * kVmPitch * (unsigned)sizeof(__half);
#pragma unroll
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
                    for (int v = lane; v < nbytes / 16; v += 32)
                        reinterpret_cast<int4 *>(dst)[v] = reinterpret_cast<const int4 *>(src)[v];
                } else {
                    for (int e = lane; e < nbytes; e += 32) dst[e] = src[e];
                }
            }
        }
        if (filling) ld.store(av);
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

// GROUPS = 2: ONE CTA of sixteen warps per SM holding two independent clip-pair groups (own named barriers,
//             own shared-memory region, own clip pairs).  Hardware warp slot w belongs to sub-partition
//             w % 4, whose single issue port all its warps share: one FIR warp of each group per
//             sub-partition, the serial roles of the two groups in opposite order.
// GROUPS = 1: a CTA is one group of eight warps (two CTAs per SM), role = warp.
static const unsigned char kRoleMaps[1][16] = {
    // (group << 3 | role) of warp 4 i + sub-partition: one FIR warp of each group per sub-partition, the serial
    // roles of group 1 in reverse order so that band-pass + Gram/loader and RZCC + neuron share a sub-partition
    {0 << 3 | 0, 0 << 3 | 1, 0 << 3 | 2, 0 << 3 | 3,
     1 << 3 | 0, 1 << 3 | 1, 1 << 3 | 2, 1 << 3 | 3,
     0 << 3 | kRoleBandpass, 0 << 3 | kRoleRzcc, 0 << 3 | kRoleNeuron, 0 << 3 | kRoleGram,
     1 << 3 | kRoleGram, 1 << 3 | kRoleNeuron, 1 << 3 | kRoleRzcc, 1 << 3 | kRoleBandpass}};

template <typename IN_T, int MM, int GROUPS>
__global__ void __launch_bounds__(kThreads * GROUPS, GROUPS == 1 ? 2 : 1)
k_fused(const IN_T *__restrict__ audio, const float *__restrict__ taps, const double *__restrict__ Wd,
        int8_t *__restrict__ spikes, float *__restrict__ power, int32_t *__restrict__ doa,
        int32_t *__restrict__ flags, unsigned int *__restrict__ sm_slots,
        const __grid_constant__ ChainParams p, const __grid_constant__ FusedGeom g, long long B, long long T) {
    extern __shared__ __align__(16) unsigned char smem_all[];
    __shared__ int s_smsp[kWarps * GROUPS], s_role[kWarps * GROUPS];
    __shared__ long long s_pair[GROUPS];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int M = MM ? MM : p.M, C2 = 2 * M;
#ifdef MICLOC_ROLE_TIMING
    long long dbg_c0 = rt_clock();
    unsigned long long dbg_g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_g0));
#endif

    // ---- which group and role this warp serves ----
    int group = 0, role = warp;
    if (GROUPS == 2) {
        const int gr = g.role_map[warp];
        group = gr >> 3;
        role = gr & 7;
    }
    if (lane == 0) {
        unsigned int wid;
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        s_smsp[warp] = (int)(wid & 3u);
        s_role[warp] = role | (group << 3);
    }
    // roles 0..2: FIR warps; 3: loader; 4: band-pass; 5: RZCC; 6: neuron; 7: Gram
    const int tid = role * 32 + lane;       // thread index inside the group
    const int bar_id = 1 + group;
    auto group_sync = [&]() { tile_barrier(bar_id); };

    unsigned char *smem_raw = smem_all + (size_t)group * g.smem_bytes;
    FusedSmem sm;
    sm.taps = reinterpret_cast<float *>(smem_raw);
    sm.xs = reinterpret_cast<float *>(smem_raw + g.off_x);
    sm.qs = reinterpret_cast<float *>(smem_raw + g.off_q);
    sm.uv = reinterpret_cast<float *>(smem_raw + g.off_uv);
    sm.vms = reinterpret_cast<__half *>(smem_raw + g.off_vm);
    sm.cs = reinterpret_cast<float *>(smem_raw + g.off_cs);      // [2][kSegsPerTile][kSeg][32] running sums
    sm.seg = reinterpret_cast<unsigned int *>(smem_raw + g.off_seg);     // [2][kSegsPerTile][3][32]
    sm.clus = reinterpret_cast<int *>(smem_raw + g.off_clus);    // RZCC cluster buffers, interleaved over 32 lanes
    sm.bits = reinterpret_cast<unsigned int *>(smem_raw + g.off_bits);   // [2][kRingWords][32]
    sm.stage = reinterpret_cast<int8_t *>(smem_raw + g.off_stage);       // [2][kSlots][kTile][C2]
    sm.gram = reinterpret_cast<double *>(smem_raw + g.off_x);    // [kSlots][16][16], clip epilogue only
    sm.dbg = sm_slots;
    sm.bar_id = bar_id;
    sm.rec = GROUPS * (int)blockIdx.x + group;
    double *red_v = reinterpret_cast<double *>(smem_raw + g.off_cs);     // [kThreads], clip epilogue only (reuses the running sums)
    int *red_i = reinterpret_cast<int *>(smem_raw + g.off_cs + kThreads * sizeof(double));

    // tap arrays of the three sub-filters: A_n = c_2n, B_n = c_2n+1, A + B (zero padded)
    for (int i = tid; i < g.tap_pitch; i += kThreads) {
        const float ta = 2 * i < p.n_taps ? taps[2 * i] : 0.f;
        const float tb = 2 * i + 1 < p.n_taps ? taps[2 * i + 1] : 0.f;
        sm.taps[i] = ta;
        sm.taps[g.tap_pitch + i] = tb;
        sm.taps[2 * g.tap_pitch + i] = ta + tb;
    }

    const int NT = (int)((T + kTile - 1) / kTile);
    const int k_last = NT + g.dtile;    // the Gram warp runs dtile + 1 tiles behind
    const long long npairs = (B + kSlots - 1) / kSlots;

    // Clip pairs are handed out dynamically: co-resident groups do not run at the same speed, so a static
    // split would wait for the slowest one.
    for (;;) {
        group_sync();
        if (tid == 0) s_pair[group] = (long long)atomicAdd(sm_slots + kSlotPair, 1u);
        group_sync();
        const long long pair = s_pair[group];
        if (pair >= npairs) break;
        const long long clip0 = pair * kSlots;
        {   // zero the audio rings: samples before the clip start are zeros (lfilter's zero state)
            float4 *x4 = reinterpret_cast<float4 *>(sm.xs);
            const int n4 = kSlots * M * 3 * g.pitch_x / 4;
            for (int i = tid; i < n4; i += kThreads) x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = tid; i < 3 * kSlots * M * kUvPitch; i += kThreads) sm.uv[i] = 0.f;    // V[-1] = 0
            // no spikes before the clip start; membrane columns of unused lanes stay zero
            for (int i = tid; i < 2 * kRingWords * 32; i += kThreads) sm.bits[i] = 0u;
            for (int i = tid; i < 2 * 2 * kVmRows * kVmPitch / 2; i += kThreads) reinterpret_cast<unsigned int *>(sm.vms)[i] = 0u;
        }
        group_sync();

        if ((g.skip >> role) & 1) {
            for (int k = -1; k <= k_last; ++k) {
                if (role < kFirWarps) asm volatile("bar.sync %0, %1;" ::"r"(3 + group), "n"(32 * kFirWarps) : "memory");
                tile_barrier(bar_id);
            }
        } else if (role < kFirWarps)
            fir_role<MM>(sm, p, g, clip0, B, role, lane, NT, k_last, 3 + group);
        else if (role == kRoleBandpass) bandpass_role<IN_T, MM>(sm, p, g, audio, clip0, B, T, lane, k_last);
        else if (role == kRoleRzcc) rzcc_role<FusedSmem, kRingWords>(sm, p, flags, clip0, B, T, M, lane, k_last);
        else if (role == kRoleNeuron) neuron_role<FusedSmem, FusedGeom, kRingWords>(sm, p, g, clip0, B, T, M, lane, k_last);
        else gram_role<IN_T, MM>(sm, p, g, audio, spikes, clip0, B, T, lane, NT, k_last);
        group_sync();
        // ---- clip epilogue: power[g] = w_g^T C w_g / T (float64), DoA = first argmax ----
        // one DoA column per thread and pass: its 2M weights are fetched first (independent loads, one
        // memory latency), then the quadratic form runs from registers and shared memory
        const double inv_T = 1.0 / (double)T;
        for (int s = 0; s < kSlots; ++s) {
            const long long clip = clip0 + s;
            if (clip >= B) break;
            const double *Cd = sm.gram + s * 256;
            double best = -1.0; int besti = 0x7fffffff;
            for (int gg = tid; gg < p.G; gg += kThreads) {
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
            for (int st = kThreads / 2; st > 0; st >>= 1) {
                if (tid < st) {
                    const double ov = red_v[tid + st]; const int oi = red_i[tid + st];
                    if (ov > red_v[tid] || (ov == red_v[tid] && oi < red_i[tid])) { red_v[tid] = ov; red_i[tid] = oi; }
                }
                group_sync();
            }
            if (tid == 0 && doa) doa[clip] = red_i[0];
            group_sync();
        }
    }
#ifdef MICLOC_ROLE_TIMING
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long g1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        unsigned long long *d = reinterpret_cast<unsigned long long *>(sm_slots + kSlotDbg);
        d[16] = (unsigned long long)(rt_clock() - dbg_c0);
        d[17] = g1 - dbg_g0;
    }
    if (tid == 0 && GROUPS * blockIdx.x + group < 512) {
        unsigned long long g1;
        unsigned int smid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        unsigned long long *d = reinterpret_cast<unsigned long long *>(sm_slots + kSlotCta) + 16 * (GROUPS * blockIdx.x + group);
        unsigned long long map = 0;
        for (int w = 0; w < kWarps * GROUPS; ++w)
            if ((s_role[w] >> 3) == group) {
                const int r = s_role[w] & 7;
                map |= (unsigned long long)(r | ((s_smsp[w] & 3) << 3)) << (8 * r);
            }
        d[0] = dbg_g0; d[1] = g1; d[2] = smid; d[3] = map;     // d[4..11]: busy cycles per role
    }
#endif
}

bool fused_supported(const ChainParams &p) {
    // sub-filters of exactly three pieces of kPieceBlocks tap blocks (STHT kernels of 385...480 samples: 10 ms at 48 kHz)
    const int sub_blocks = ((p.n_taps + 1) / 2 + 7) / 8;
    return p.tap_stride == 2 && p.M <= kRows && p.nsec == 2 && (p.n_taps % 8) == 0 &&
           (sub_blocks + 2) / 3 * 3 == 3 * kPieceBlocks;
}

template <typename IN_T, int MM, int GROUPS>
static int launch_fused_t(const ChainParams &p, const FusedGeom &g, const float *d_taps, const double *d_Wd,
                          const IN_T *audio, long long B, long long T, int8_t *spikes, float *power, int32_t *doa,
                          int32_t *flags, unsigned int *sm_slots, int sm_count, cudaStream_t st) {
    auto kern = k_fused<IN_T, MM, GROUPS>;
    const int smem = GROUPS * g.smem_bytes;
    MICLOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // all of the SM's L1/shared array as shared memory: two clip-pair groups of ~100 KB must be resident together
    MICLOC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int per_sm = 1;
    MICLOC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads * GROUPS, smem));
    if (per_sm > 2 / GROUPS) per_sm = 2 / GROUPS;     // two groups per SM: the role placement balances exactly two
    if (per_sm < 1) return set_error(MICLOC_ERR_UNSUPPORTED, "fused kernel does not fit (smem %d B)", smem);
    long long grid = (long long)sm_count * per_sm;
    const long long npairs = (B + kSlots - 1) / kSlots;
    const long long want = (npairs + GROUPS - 1) / GROUPS;
    if (grid > want) grid = want;
    MICLOC_CUDA(cudaMemsetAsync(sm_slots, 0, kSlotResetWords * sizeof(unsigned int), st));   // FIR placement counters + pair counter restart per launch
    kern<<<(unsigned)grid, kThreads * GROUPS, smem, st>>>(audio, d_taps, d_Wd, spikes, power, doa, flags, sm_slots, p, g, B, T);
    count_launch(1);
    MICLOC_CUDA(cudaGetLastError());
    return MICLOC_OK;
}

int launch_fused(const ChainParams &p, const float *d_taps, const double *d_Wd, const void *audio, int dtype,
                 long long B, long long T, int8_t *spikes, float *power, int32_t *doa, int32_t *flags,
                 unsigned int *sm_slots, int sm_count, cudaStream_t st) {
    if (!ffa::fused_supported(p))
        return set_error(MICLOC_ERR_UNSUPPORTED,
                         "fused kernel covers Hilbert-type STHT kernels (every other tap zero), a 2-section band-pass "
                         "and up to %d microphones; use the staged path", kRows);
    FusedGeom g{};
    if (const char *e = getenv("MICLOC_FUSED_SKIP")) g.skip = (int)strtol(e, nullptr, 0);   // role ablation, debugging only
    if (const char *e = getenv("MICLOC_FUSED_FIRBLOCKS")) g.fir_blocks = (int)strtol(e, nullptr, 0);
    {
        for (int w = 0; w < 16; ++w) g.role_map[w] = kRoleMaps[0][w];
    }
    // sub-filters of n_taps / 2 taps in blocks of 8, walked in groups of three (zero taps appended)
    const int sub_taps = (p.n_taps + 1) / 2;
    g.nblk = ((sub_taps + 7) / 8 + 2) / 3 * 3;
    g.tap_pitch = 8 * g.nblk + 12;                       // + padding: the three tap arrays start in different bank groups
    // a ring holds the oldest pair a tile's windows load (8 nblk - 1 back), the tile in the FIR and the tile being filled
    g.ring_p = (8 * g.nblk + 2 * kTileM + 7) / 8 * 8;
    g.pitch_x = 2 * g.ring_p + 4;
    if ((g.pitch_x / 4) % 2 == 0) g.pitch_x += 4;
    // a spike at p is final once the RZCC warp passed p + rzcc_lag(w) - 1; the neuron warp works on
    // tile k - dtile while the RZCC warp has completed tile k - 4
    g.dtile = 4 + (rzcc_lag(p.w) - 1 + kTile - 1) / kTile;
    const int rows = kSlots * p.M;
    int off = (3 * g.tap_pitch * (int)sizeof(float) + 15) & ~15;
    g.off_x = off; off += rows * 3 * g.pitch_x * (int)sizeof(float);
    g.off_q = off; off += 2 * rows * kQPitch * (int)sizeof(float);
    g.off_uv = off; off += 3 * rows * kUvPitch * (int)sizeof(float);
    g.off_vm = off; off += 2 * 2 * kVmRows * kVmPitch * (int)sizeof(__half);
    g.off_cs = off; off += 2 * kSegsPerTile * kSeg * 32 * (int)sizeof(float);
    g.off_seg = off; off += 2 * kSegsPerTile * 3 * 32 * (int)sizeof(int);
    g.off_clus = off; off += 4 * kClusterMax * 32 * (int)sizeof(int);
    g.off_bits = off; off += 2 * kRingWords * 32 * (int)sizeof(int);
    g.off_stage = off; off += (2 * kSlots * kTile * p.C2 + 15) & ~15;
    g.smem_bytes = (off + 15) & ~15;
    if (g.nblk != 3 * kPieceBlocks)
        return set_error(MICLOC_ERR_UNSUPPORTED, "fast-FIR variant: sub-filters of %d tap blocks (it covers %d)", g.nblk, 3 * kPieceBlocks);
    // the spike-bit ring must hold the neuron warp's oldest read and the RZCC warp's newest write
    // (the neuron warp reads back to (k - dtile) * kTile - nL while the RZCC warp clears the words of tile k - 3)
    if (kTile * (g.dtile - 2) + p.nL + kSeg > kRingWords * 32)
        return set_error(MICLOC_ERR_UNSUPPORTED, "robust_width %d / neuron length %d exceed the fused kernel's spike ring; "
                         "use the staged path", p.w, p.nL);
    if (kSlots * 256 * (int)sizeof(double) > rows * 3 * g.pitch_x * (int)sizeof(float))
        return set_error(MICLOC_ERR_UNSUPPORTED, "shared-memory tiles too small for the epilogue");
    if (g.smem_bytes > 227 * 1024)
        return set_error(MICLOC_ERR_UNSUPPORTED, "fused kernel needs %d B of shared memory; use the staged path", g.smem_bytes);
    if (T + 16 * kTile >= (1ll << 29)) return set_error(MICLOC_ERR_SHAPE, "T too large for the fused kernel");
    // one CTA of two clip-pair groups per SM when both fit its shared memory (the usual case), else the
    // single-group CTA twice per SM; MICLOC_FUSED_GROUPS=1 forces the latter (A/B measurements)
    int groups = 2 * g.smem_bytes <= 227 * 1024 ? 2 : 1;
    if (const char *e = getenv("MICLOC_FUSED_GROUPS")) { if (atoi(e) == 1) groups = 1; }
#define MICLOC_FUSED_CASE(IN, MMV)                                                                                \
    do {                                                                                                          \
        if (groups == 2)                                                                                          \
            return launch_fused_t<IN, MMV, 2>(p, g, d_taps, d_Wd, (const IN *)audio, B, T, spikes, power, doa,    \
                                              flags, sm_slots, sm_count, st);                                     \
        return launch_fused_t<IN, MMV, 1>(p, g, d_taps, d_Wd, (const IN *)audio, B, T, spikes, power, doa, flags, \
                                          sm_slots, sm_count, st);                                                \
    } while (0)
    const bool i16 = dtype == MICLOC_I16;
    if (p.M == 7) { if (i16) MICLOC_FUSED_CASE(int16_t, 7); else MICLOC_FUSED_CASE(float, 7); }
    if (i16) MICLOC_FUSED_CASE(int16_t, 0); else MICLOC_FUSED_CASE(float, 0);
#undef MICLOC_FUSED_CASE
}

}  // namespace ffa
}  // namespace micloc
