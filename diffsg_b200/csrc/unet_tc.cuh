// Tensor-core engine: one CTA carries a 128-row tile through the whole UNet1D stage program;
// two CTAs are co-resident per SM (fp16x2) so one tile's MMAs overlap the other's epilogue.
//
//   warps 0-3     epilogue / operand producers: thread == row == TMEM lane, one warp per SM sub-partition.
//                 Every epilogue op STREAMS its source (the TMEM accumulator, a skip vector, the input row)
//                 in groups of 16 columns held as 8 packed f32x2 register pairs: pass 1 accumulates the
//                 shifted LayerNorm moments, pass 2 re-reads the accumulator, normalises, applies Swish,
//                 splits into fp16 (hi, lo) and writes the next GEMM's A operand as core-matrix K-chunks.
//                 All per-element arithmetic is packed (FFMA2 / FADD2 / FMUL2); the only scalar ops are the
//                 two MUFU of Swish, the fp16 unpack and the operand stores.
//   warp 4        bulk-TMA producer: streams fp16 weight K-chunk images (and the bias chunk images: static
//                 ones from the weight blob, the per-step time-bias ones from the step image table) into the
//                 W ring, and the per-stage LayerNorm gamma / beta package into the package ring
//   warp 5        MMA issuer: one thread issues tcgen05.mma (A, W from shared memory, fp32 accumulators in
//                 TMEM) and commits completion to mbarriers
//
// Biases never touch the CUDA cores: every GEMM group ends with one K = 16 "bias MMA" whose A operand is a
// constant tile of ones (columns 0..2) and whose W rows hold the bias as three fp16 terms (hi, mid, lo), so
// the accumulator already contains `x = W a + b` when the epilogue reads it.  The bias image travels with the
// group's last weight chunk (same W-ring stage, one barrier).
//
// Program format: diffsg_b200/tc_packer.py.  Reference semantics: ddpm_opt/UNetCF.py:83-95,
// :318-356; sampler: ddpm_opt/classifier_free_MSR.py:124-137.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace diffsg {
namespace tc {

constexpr int kRows = 128;
constexpr int kChunkK = 64;                          // K columns per operand chunk
constexpr int kSlotBytes = kRows * kChunkK * 2;      // one fp16 A chunk (16 KB)
constexpr int kASlots = 2;
constexpr int kWStages = 2;
constexpr int kRegionCols = 128;                     // widest vector the engine carries
constexpr int kWStageBytes = kRegionCols * kChunkK * 2;   // one fp16 W chunk (N <= kRegionCols)
constexpr int kBiasK = 16;                           // K of the bias MMA
constexpr int kBiasBytes = kRegionCols * kBiasK * 2; // bias image of one GEMM group (rides behind its last weight chunk)
constexpr int kOnesBytes = kRows * kBiasK * 2;       // the constant A tile of the bias MMAs
constexpr int kCtasPerSm = 2;
constexpr int kEpiWarps = 4, kEpiThreads = 32 * kEpiWarps;
constexpr int kThreads = kEpiThreads + 64;           // + TMA warp + MMA warp
constexpr int kTmemCols = 2 * kRegionCols;           // two accumulator regions
constexpr int kMaxStages = 256, kMaxChunks = 640, kMaxEpi = 512;   // program lives in __constant__ memory
constexpr int kPkgFloats = 512, kPSlots = 2;         // gamma | beta (x2 for a cat LayerNorm) of one stage

// streaming epilogue ops (diffsg_b200/tc_packer.py)
enum : int { OP_LN = 1, OP_CATLN, OP_RAW_T, OP_RAW_S, OP_RAW_IN, OP_OUT };
constexpr int kChunkCond = 1;
constexpr int kFTime = 1, kFCond = 2, kFPush = 4, kFDefer = 8;
constexpr int kStatusOverflow = 1;                   // a raw fp16 operand exceeded the fp16 range
// Every epilogue thread arrives on the operand / package barriers itself and waits for its operand slot itself
// (a_empty) before writing: the form compute-sanitizer's racecheck / synccheck can follow.  (An elected lane per warp
// + skipping the slot wait where the accumulator wait already proves the slot free measured 0.3 % faster and is not
// worth tools that can no longer verify the protocol.)
constexpr int kArrivals = 128;                       // arrivals per phase on a_full / p_empty: one per epilogue thread

struct __align__(8) Epi { uint8_t kind, np, dt, misc, slot, off1; uint16_t tt_src4; };   // misc: region | flags << 1
struct __align__(8) Chunk { uint16_t kw, flags; uint32_t w_off16; };
struct __align__(16) Stage {
    uint16_t chunk_begin, epi_begin;
    uint8_t n_chunks, n_epi, n16, bits;          // bits: region | accumulate << 1 | has_gemm << 2 | time_bias << 3
    uint16_t pkg_off16;                          // LayerNorm package: offset (x 16 floats) and size (x 4 floats)
    uint8_t pkg_f4, pad_;
    uint32_t bias_off16;                         // bias image (x 16 bytes): in the weight blob, or in a step-image row
};
static_assert(sizeof(Epi) == 8 && sizeof(Chunk) == 8 && sizeof(Stage) == 16, "program record layout");

// The active stage program (uploaded by the host before a launch whenever the plan changes).
__constant__ Stage c_stages[kMaxStages];
__constant__ Chunk c_chunks[kMaxChunks];
__constant__ Epi c_epis[kMaxEpi];

struct TcDev {
    int n_stages, n_chunks, n_epi;
    const uint8_t* w_hi; const uint8_t* w_lo;     // fp16 weight images (lo: nterms == 3 only)
    const float* params;                           // LayerNorm gamma / beta packages
    const float* tt;                               // fp32 time table [rows][tt_stride] (forward mode)
    const uint8_t* tt_img;                         // fp16 bias-chunk images of the time table, one row per step
    int tt_stride, tt_img_stride, nterms;          // tt_img_stride in bytes
    int M, Mp, C, Cp;
    long long* debug;
    int* status;                                   // kStatus* bits (atomicOr)
    float* scratch;                                // per CTA: skip stack + eps stash + cond image + skip moments
    size_t scratch_floats;                         // per CTA
    int skip_off[kMaxSkip];                        // float offset of each skip slot inside the CTA scratch
    int stash_off, cond_off, stats_off;            // float offsets (cond image: hi then lo, fp16)
};

struct SmemLayout {
    uint8_t a_hi[kASlots][kSlotBytes];
    uint8_t a_lo[kASlots][kSlotBytes];
    uint8_t ones[kOnesBytes];
    float pkg[kPSlots][kPkgFloats];
    uint64_t a_full[kASlots], a_empty[kASlots], w_full[kWStages], w_empty[kWStages], p_full[kPSlots],
        p_empty[kPSlots], acc_full;
    uint32_t tmem_base, pad_;
#ifdef DIFFSG_TC_TIMING
    uint32_t t_wait[96], t_work[96];        // per-stage cycles of thread 0 (accumulator wait / everything else)
#endif
    // followed by the W ring: kWStages x [(nterms == 3 ? 2 : 1) weight images | bias image] (dynamic)
};
#ifndef DIFFSG_TC_TIMING
static_assert((((sizeof(SmemLayout) + 127) & ~size_t(127)) + kWStages * (kWStageBytes + kBiasBytes) + 128 + 1024) * kCtasPerSm <= 233472,
              "kCtasPerSm CTAs (fp16x2) must fit the SM's 228 KB of shared memory");
#endif

// What one launch does: kSampler -> steps step_hi..step_lo, two passes each; else one forward.
struct RunArgs {
    const float* x; const int32_t* t_idx; const float* cond; const float* mask; float* eps;      // forward
    float* y; const float* noise; float* rec_y; float* rec_eps; double* stats;                  // sampler
    int64_t B;
    int T, step_hi, step_lo, norm_steps;
    float omega;
    uint64_t seed, offset;
    float c_eps[64], c_rs[64], c_noise[64];
};

// ------------------------------------------------------------------------------------------ packed fp32
typedef unsigned long long f2;      // two fp32 in one 64-bit register pair (FFMA2 / FADD2 / FMUL2 operands)
__device__ __forceinline__ f2 pk(float a, float b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ f2 pku(uint32_t a, uint32_t b) { f2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void upk(f2 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// (lo, hi) -> packed fp16x2, round to nearest, saturating at +-65504 instead of producing inf
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ f2 h2_to_f2(uint32_t h) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
    return pk(f.x, f.y);
}
// v (two fp32) -> fp16 hi pair and fp16 residual pair: v ~= hi + lo to 22 significant bits
__device__ __forceinline__ void split2(f2 v, uint32_t& hi, uint32_t& lo) {
    float a, b;
    upk(v, a, b);
    hi = cvt_h2(a, b);
    upk(sub2(v, h2_to_f2(hi)), a, b);
    lo = cvt_h2(a, b);
}
// x * sigmoid(x), two lanes.
//   kFast = false: one MUFU.EX2 and one MUFU.RCP per lane (absolute error <= 1.4e-6, measured on B200)
//   kFast = true:  h + h * tanh(h), h = x / 2: ONE MUFU per lane.  MUFU.TANH on B200 is good to 8e-6 absolute
//                  (tools/ubench/tanh_err.cu), the Swish built on it to 1.03e-5 absolute over the whole real line --
//                  far below the fp16 weight rounding of the fp16x2 mode that uses it; fp16x3 keeps the exact form.
template <bool kFast>
__device__ __forceinline__ f2 swish2(f2 u) {
    if (kFast) {
        const f2 h = mul2(u, pk(0.5f, 0.5f));
        float h0, h1, t0, t1;
        upk(h, h0, h1);
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
        return fma2(h, pk(t0, t1), h);
    }
    float e0, e1;
    upk(mul2(u, pk(-1.4426950408889634f, -1.4426950408889634f)), e0, e1);
    upk(add2(pk(ex2_approx(e0), ex2_approx(e1)), pk(1.0f, 1.0f)), e0, e1);
    return mul2(u, pk(rcp_approx(e0), rcp_approx(e1)));
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}

// ------------------------------------------------------------------------------------------ TMEM
// 16 consecutive columns of this thread's lane; results are valid after tmem_wait16 on the same registers
// (the wait takes them as in/out operands so no consumer can be scheduled ahead of it).
__device__ __forceinline__ void tmem_ld16u(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}

// 4 consecutive columns, complete on return
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ld32u(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait32(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]));
}

#ifdef DIFFSG_TC_TIMING
#define TCT_BEGIN(v) const long long v = clock64()
#define TCT_END(v, slot) E.tacc[slot] += clock64() - v
#else
#define TCT_BEGIN(v)
#define TCT_END(v, slot)
#endif

// Per-thread epilogue context
struct EpiCtx {
#ifdef DIFFSG_TC_TIMING
    long long tacc[12];   // 0 acc wait, 1 pkg wait, 2 LN pass 1, 3 LN pass 2, 4 cat-LN, 5 raw ops, 6 cond, 7 out, 8 total
#endif
    int row, lane;
    uint32_t tmem_row;      // TMEM address of this thread's lane, column 0
    uint32_t aseq;          // A-ring sequence number (chunks published so far by the tile)
    uint32_t a_row;         // byte offset of this row inside an 8-row core-matrix group: (row & 7) * 16
    uint32_t a_hi0;         // shared-memory address of a_hi[0]
    float amax;             // largest |raw operand| seen (fp16 range check)
    float* scr;             // this CTA's global scratch
    int64_t grow;           // global row
    bool valid;
};

// ---- A-operand ring (producer side): one vector of `np` 8-column pieces -> ceil(np / 8) K-chunks of up to four
// 16-column groups; a slot is written after waiting for the MMA that read its previous chunk (a_empty).
struct Emitter {
    uint32_t seq0, base;
    int np, ng;
    bool defer;
};
__device__ __forceinline__ void emit_begin(Emitter& em, const EpiCtx& E, int np, bool defer) {
    em.seq0 = E.aseq; em.np = np; em.ng = np >> 1; em.defer = defer; em.base = 0;
}
__device__ __forceinline__ void emit_publish(SmemLayout& S, const EpiCtx& E, uint32_t sq) {
    fence_proxy_async_smem();
    tcgen05_fence_before();
    mbar_arrive(&S.a_full[sq % kASlots]);
}
// shared-memory address (a_hi image) where group g of the vector goes; opens the chunk when g is its first group
__device__ __forceinline__ uint32_t emit_addr(SmemLayout& S, EpiCtx& E, Emitter& em, int g) {
    if ((g & 3) == 0) {
        const int c = g >> 2;
        const uint32_t sq = em.seq0 + c, sl = sq % kASlots;
        mbar_wait(&S.a_empty[sl], ((sq / kASlots) & 1) ^ 1);
        const int cnt = min(8, em.np - 8 * c);                      // 8-column pieces in chunk c
        em.base = E.a_hi0 + sl * kSlotBytes + (uint32_t)(E.row >> 3) * (uint32_t)(cnt * 128) + E.a_row;
    }
    return em.base + (g & 3) * 256;
}
__device__ __forceinline__ void emit_done(SmemLayout& S, const EpiCtx& E, const Emitter& em, int g) {
    if (!em.defer && ((g & 3) == 3 || g == em.ng - 1)) emit_publish(S, E, em.seq0 + (g >> 2));
}
__device__ __forceinline__ void emit_end(SmemLayout& S, EpiCtx& E, const Emitter& em) {
    const int nch = (em.np + 7) >> 3;
    if (em.defer)
        for (int c = 0; c < nch; ++c) emit_publish(S, E, em.seq0 + c);
    E.aseq += nch;
}
constexpr uint32_t kLoOff = kASlots * kSlotBytes;     // a_lo[s] = a_hi[s] + kLoOff

// one group: raw values -> fp16 (hi, lo) operand pieces
__device__ __forceinline__ void store_split(uint32_t addr, const f2 (&v)[8]) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split2(v[j], h[j], l[j]);
    sts128(addr, h[0], h[1], h[2], h[3]);
    sts128(addr + 128, h[4], h[5], h[6], h[7]);
    sts128(addr + kLoOff, l[0], l[1], l[2], l[3]);
    sts128(addr + kLoOff + 128, l[4], l[5], l[6], l[7]);
}
// t = x * sc + sh for one group straight out of the TMEM registers (which are then free for the next load)
__device__ __forceinline__ void normalise16(const uint32_t (&r)[16], f2 sc, f2 sh, f2 (&t)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = fma2(pku(r[2 * j], r[2 * j + 1]), sc, sh);
}
// one group: swish(t * gamma + beta) -> operand pieces.  Pad columns have gamma = beta = 0 -> exact zeros.
template <bool kFast>
__device__ __forceinline__ void store_ln_swish(uint32_t addr, const f2 (&t)[8], const float4* gamma, const float4* beta) {
    f2 v[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 g = gamma[q], b = beta[q];
        v[2 * q] = swish2<kFast>(fma2(t[2 * q], pk(g.x, g.y), pk(b.x, b.y)));
        v[2 * q + 1] = swish2<kFast>(fma2(t[2 * q + 1], pk(g.z, g.w), pk(b.z, b.w)));
    }
    store_split(addr, v);
}
// shifted one-pass moments of one group: d = x - shift (computed straight out of the TMEM registers, which are
// then free for the next load), s1 += d, s2 += d^2 (two packed accumulators each)
__device__ __forceinline__ void centre16(const uint32_t (&r)[16], f2 shift, f2 (&d)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) d[j] = sub2(pku(r[2 * j], r[2 * j + 1]), shift);
}
__device__ __forceinline__ void moments16(const f2 (&d)[8], f2 (&s1)[2], f2 (&s2)[2]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s1[j & 1] = add2(s1[j & 1], d[j]);
        s2[j & 1] = fma2(d[j], d[j], s2[j & 1]);
    }
}
__device__ __forceinline__ void pack16(const uint32_t (&r)[16], f2 (&x)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = pku(r[2 * j], r[2 * j + 1]);
}
__device__ __forceinline__ void fold_moments(const f2 (&s1)[2], const f2 (&s2)[2], float& a, float& q) {
    float a0, a1, q0, q1;
    upk(add2(s1[0], s1[1]), a0, a1);
    upk(add2(s2[0], s2[1]), q0, q1);
    a = a0 + a1;
    q = q0 + q1;
}
__device__ __forceinline__ void track_amax(EpiCtx& E, const uint32_t (&r)[16]) {
#pragma unroll
    for (int j = 0; j < 16; ++j) E.amax = fmaxf(E.amax, fabsf(__uint_as_float(r[j])));
}
__device__ __forceinline__ void store_group_skip(const uint32_t (&r)[16], uint4* sk) {
#pragma unroll
    for (int q = 0; q < 4; ++q) sk[q * kRows] = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
}
// forward mode only: add this row's slice of the fp32 time table (lin1.bias + time embedding) to a group
__device__ __forceinline__ void add_time16(uint32_t (&r)[16], const float* tt) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(tt + q * 4);
        r[q * 4 + 0] = __float_as_uint(__uint_as_float(r[q * 4 + 0]) + b.x);
        r[q * 4 + 1] = __float_as_uint(__uint_as_float(r[q * 4 + 1]) + b.y);
        r[q * 4 + 2] = __float_as_uint(__uint_as_float(r[q * 4 + 2]) + b.z);
        r[q * 4 + 3] = __float_as_uint(__uint_as_float(r[q * 4 + 3]) + b.w);
    }
}

// Walk a skip vector (global scratch slab, L2 / HBM resident) in batches of two groups: all 8 16-byte loads of a
// batch are issued back to back, so the latency is paid once per batch.  (Loads are NOT kept in flight across
// body(): the operand-publish fence would wait for them.)
template <typename F>
__device__ __forceinline__ void skip_walk2(const uint4* sk, int ng, F body) {
    for (int gb = 0; gb < ng; gb += 2) {
        const bool two = gb + 1 < ng;
        const uint4* p0 = sk + (size_t)gb * 4 * kRows;
        const uint4* p1 = p0 + (two ? 4 * kRows : 0);
        const uint4 a0 = p0[0], a1 = p0[kRows], a2 = p0[2 * kRows], a3 = p0[3 * kRows];
        const uint4 c0 = p1[0], c1 = p1[kRows], c2 = p1[2 * kRows], c3 = p1[3 * kRows];
        {
            const uint32_t r[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
            body(gb, r);
        }
        if (two) {
            const uint32_t r[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
            body(gb + 1, r);
        }
    }
}

// Shifted moments of a TMEM vector (pass 1 of a LayerNorm), 32 columns per TMEM round trip; optionally spills the
// vector to a skip slot.  Returns shift, s1 = sum(x - shift), s2 = sum((x - shift)^2) over the first dt columns.
// (An odd group count reads 16 columns past the vector: still inside the 128-column region, and ignored.)
template <bool kSampler>
__device__ __forceinline__ void tmem_moments(uint32_t ta, int ng, int dt, const float* tt_row, uint4* sk_push,
                                             float& shift, float& s1o, float& s2o) {
    uint32_t r[32];
    f2 s1[2] = {0ull, 0ull}, s2[2] = {0ull, 0ull}, sh2 = 0ull;
    shift = 0.f;
    tmem_ld32u(ta, r);
    for (int g = 0; g < ng; g += 2) {
        tmem_wait32(r);
        uint32_t (&ra)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[0]);
        uint32_t (&rb)[16] = *reinterpret_cast<uint32_t(*)[16]>(&r[16]);
        const bool two = g + 1 < ng;
        if (!kSampler && tt_row) { add_time16(ra, tt_row + g * 16); if (two) add_time16(rb, tt_row + g * 16 + 16); }
        if (g == 0) { shift = __uint_as_float(r[0]); sh2 = pk(shift, shift); }
        if (sk_push) {
            store_group_skip(ra, sk_push + (size_t)g * 4 * kRows);
            if (two) store_group_skip(rb, sk_push + (size_t)(g + 1) * 4 * kRows);
        }
        // pad columns (exact zeros; or, for an odd group count, the 16 columns past the vector) must not enter the moments
        const int nval = dt - g * 16;
        if (nval < 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j >= nval) r[j] = __float_as_uint(shift);
        }
        f2 da[8], db[8];
        centre16(ra, sh2, da);
        centre16(rb, sh2, db);
        if (g + 2 < ng) tmem_ld32u(ta + (g + 2) * 16, r);
        moments16(da, s1, s2);
        moments16(db, s1, s2);
    }
    fold_moments(s1, s2, s1o, s2o);
}

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Swish(cond) operand chunks: the tile's pre-split image lives in global scratch already in operand-chunk layout,
// so a chunk is two bulk-TMA copies (hi, lo) straight into the A-ring slot, issued by one thread; the slot's barrier
// counts the bytes on top of the usual one arrival per epilogue warp.
__device__ __forceinline__ void emit_cond(SmemLayout& S, EpiCtx& E, const TcDev& P) {
    const uint8_t* img = reinterpret_cast<const uint8_t*>(E.scr + P.cond_off);
    const int nkc = P.Cp / 8;                      // 16-byte K pieces per row
    uint32_t done = 0;                             // bytes of the chunks before this one (per term)
    for (int c0 = 0; c0 < nkc; c0 += 8) {
        const int nk = min(8, nkc - c0);
        const uint32_t sq = E.aseq, sl = sq % kASlots;
        const uint32_t bytes = (uint32_t)nk * 128u * (kRows / 8);
        // every thread waits for the slot: an arrival must not land in the barrier phase of the slot's previous chunk
        mbar_wait(&S.a_empty[sl], ((sq / kASlots) & 1) ^ 1);
        if (E.row == 0) {
            mbar_expect_tx(&S.a_full[sl], 2 * bytes);
            tma_load_1d(S.a_hi[sl], img + done, bytes, &S.a_full[sl]);
            tma_load_1d(S.a_lo[sl], img + (size_t)P.Cp * kRows * 2 + done, bytes, &S.a_full[sl]);
        }
        mbar_arrive(&S.a_full[sl]);
        done += bytes;
        ++E.aseq;
    }
}

template <bool kSampler, bool kFast>
__device__ __forceinline__ void run_epilogue(SmemLayout& S, const TcDev& P, const RunArgs& R, EpiCtx& E, int trow,
                                             bool use_cond, int pass, int step, uint32_t& acc_phase,
                                             uint32_t& pseq, double& st_s, double& st_q) {
    const int row = E.row;
    for (int si = 0; si < P.n_stages; ++si) {
        const Stage sg = c_stages[si];
        const bool has_pkg = sg.pkg_f4 != 0;
        const uint32_t psl = pseq % kPSlots;
#ifdef DIFFSG_TC_TIMING
        const long long _s0 = clock64();
        long long _s1 = _s0, _s2 = _s0;
#endif
        // the package was requested long ago: check it first, in the shadow of the accumulator wait
        if (has_pkg) { TCT_BEGIN(_tk); mbar_wait(&S.p_full[psl], (pseq / kPSlots) & 1); TCT_END(_tk, 1); }
        const float* pk_ = S.pkg[psl];
        if (sg.bits & 4) {
            TCT_BEGIN(_ta);
#ifdef DIFFSG_TC_TIMING
            _s1 = clock64();
#endif
            mbar_wait(&S.acc_full, acc_phase);
            acc_phase ^= 1;
            tcgen05_fence_after();
            TCT_END(_ta, 0);
#ifdef DIFFSG_TC_TIMING
            _s2 = clock64();
#endif
        }
        for (int ei = sg.epi_begin; ei < sg.epi_begin + sg.n_epi; ++ei) {
            const Epi op = c_epis[ei];
            const int np = op.np, ng = np >> 1, dt = op.dt, dp = np * 8;
            const int region = op.misc & 1, flags = op.misc >> 1;
            const uint32_t ta = E.tmem_row + region * kRegionCols;
            const float* tt_row = (!kSampler && (flags & kFTime)) ? P.tt + (size_t)trow * P.tt_stride + op.tt_src4 * 4 : nullptr;
            uint32_t r[16];
            f2 x[8];
            switch (op.kind) {
                case OP_LN: {
                    // ---- pass 1: moments (and the skip push)
                    TCT_BEGIN(_t1);
                    uint4* sk = (flags & kFPush) ? reinterpret_cast<uint4*>(E.scr + P.skip_off[op.slot]) + row : nullptr;
                    float shift, s1, s2;
                    tmem_moments<kSampler>(ta, ng, dt, tt_row, sk, shift, s1, s2);
                    if (flags & kFPush)
                        reinterpret_cast<float4*>(E.scr + P.stats_off)[op.slot * kRows + row] = make_float4(s1, s2, shift, 0.f);
                    const float inv_n = 1.0f / (float)dt;
                    const float md = s1 * inv_n;
                    const float rstd = rsqrtf(fmaxf(s2 * inv_n - md * md, 0.f) + kLnEps);
                    const f2 sc = pk(rstd, rstd), sh = pk(-(shift + md) * rstd, -(shift + md) * rstd);
                    TCT_END(_t1, 2);
                    // ---- pass 2: normalise, Swish, split, publish
                    TCT_BEGIN(_t2);
                    const float4* pg = reinterpret_cast<const float4*>(pk_ + op.off1 * 4);
                    const float4* pb = pg + np * 2;
                    Emitter em;
                    emit_begin(em, E, np, (flags & kFDefer) != 0);
                    tmem_ld16u(ta, r);
                    for (int g = 0; g < ng; ++g) {
                        const uint32_t addr = emit_addr(S, E, em, g);
                        tmem_wait16(r);
                        if (!kSampler && tt_row) add_time16(r, tt_row + g * 16);
                        normalise16(r, sc, sh, x);
                        if (g + 1 < ng) tmem_ld16u(ta + (g + 1) * 16, r);
                        store_ln_swish<kFast>(addr, x, pg + g * 4, pb + g * 4);
                        emit_done(S, E, em, g);
                    }
                    emit_end(S, E, em);
                    TCT_END(_t2, 3);
                    if ((flags & kFCond) && use_cond) { TCT_BEGIN(_t4); emit_cond(S, E, P); TCT_END(_t4, 6); }
                    break;
                }
                case OP_CATLN: {
                    TCT_BEGIN(_t5);
                    // LayerNorm over cat(x, skip): the skip part's moments were stored when it was pushed;
                    // operands: skip part, then x part
                    const uint4* sk = reinterpret_cast<const uint4*>(E.scr + P.skip_off[op.slot]) + row;
                    const float4 ss = reinterpret_cast<const float4*>(E.scr + P.stats_off)[op.slot * kRows + row];
                    float shift, s1, s2;
                    tmem_moments<kSampler>(ta, ng, dt, nullptr, nullptr, shift, s1, s2);
                    const float n = (float)dt, inv_n = 1.0f / n;
                    const float mx = shift + s1 * inv_n, ms = ss.z + ss.x * inv_n;
                    const float m2 = (s2 - s1 * s1 * inv_n) + (ss.y - ss.x * ss.x * inv_n) + (mx - ms) * (mx - ms) * (0.5f * n);
                    const float mean = 0.5f * (mx + ms);
                    const float rstd = rsqrtf(fmaxf(m2 * (0.5f * inv_n), 0.f) + kLnEps);
                    const f2 sc = pk(rstd, rstd), sh = pk(-mean * rstd, -mean * rstd);
                    const float4* pgx = reinterpret_cast<const float4*>(pk_ + op.off1 * 4);   // gamma_x | beta_x | gamma_s | beta_s
                    const int d4 = np * 2;
                    Emitter em;
                    emit_begin(em, E, np, false);
                    skip_walk2(sk, ng, [&](int g, const uint32_t (&rs)[16]) {
                        const uint32_t addr = emit_addr(S, E, em, g);
                        f2 xs[8];
                        normalise16(rs, sc, sh, xs);
                        store_ln_swish<kFast>(addr, xs, pgx + 2 * d4 + g * 4, pgx + 3 * d4 + g * 4);
                        emit_done(S, E, em, g);
                    });
                    emit_end(S, E, em);
                    emit_begin(em, E, np, false);
                    tmem_ld16u(ta, r);
                    for (int g = 0; g < ng; ++g) {
                        const uint32_t addr = emit_addr(S, E, em, g);
                        tmem_wait16(r);
                        normalise16(r, sc, sh, x);
                        if (g + 1 < ng) tmem_ld16u(ta + (g + 1) * 16, r);
                        store_ln_swish<kFast>(addr, x, pgx + g * 4, pgx + d4 + g * 4);
                        emit_done(S, E, em, g);
                    }
                    emit_end(S, E, em);
                    TCT_END(_t5, 4);
                    break;
                }
                case OP_RAW_T: {
                    TCT_BEGIN(_t6);
                    // raw TMEM vector as an operand (Down/Upsample, shortcut, attention); when pushed, its moments
                    // are stored for the cat LayerNorm that pops it
                    const bool push = (flags & kFPush) != 0;
                    uint4* sk = reinterpret_cast<uint4*>(E.scr + P.skip_off[push ? op.slot : 0]) + row;
                    f2 s1[2] = {0ull, 0ull}, s2[2] = {0ull, 0ull}, sh2 = 0ull;
                    float shift = 0.f;
                    Emitter em;
                    emit_begin(em, E, np, (flags & kFDefer) != 0);     // deferred when the next GEMM accumulates into this region
                    tmem_ld16u(ta, r);
                    for (int g = 0; g < ng; ++g) {
                        const uint32_t addr = emit_addr(S, E, em, g);
                        tmem_wait16(r);
                        track_amax(E, r);
                        pack16(r, x);                                // pad columns are exact zeros (zero W rows, zero bias)
                        store_split(addr, x);
                        if (push) {
                            if (g == 0) { shift = __uint_as_float(r[0]); sh2 = pk(shift, shift); }
                            store_group_skip(r, sk + (size_t)g * 4 * kRows);
                            if (g == ng - 1 && dt < ng * 16) {
                                const int nval = dt - g * 16;
#pragma unroll
                                for (int j = 0; j < 16; ++j) if (j >= nval) r[j] = __float_as_uint(shift);
                            }
                            centre16(r, sh2, x);
                            moments16(x, s1, s2);
                        }
                        if (g + 1 < ng) tmem_ld16u(ta + (g + 1) * 16, r);
                        emit_done(S, E, em, g);
                    }
                    emit_end(S, E, em);
                    if (push) {
                        float a, q;
                        fold_moments(s1, s2, a, q);
                        reinterpret_cast<float4*>(E.scr + P.stats_off)[op.slot * kRows + row] = make_float4(a, q, shift, 0.f);
                    }
                    TCT_END(_t6, 5);
                    break;
                }
                case OP_RAW_S: {
                    TCT_BEGIN(_t7);
                    const uint4* sk = reinterpret_cast<const uint4*>(E.scr + P.skip_off[op.slot]) + row;
                    Emitter em;
                    emit_begin(em, E, np, false);
                    skip_walk2(sk, ng, [&](int g, const uint32_t (&rs)[16]) {
                        const uint32_t addr = emit_addr(S, E, em, g);
                        track_amax(E, rs);
                        f2 xs[8];
                        pack16(rs, xs);
                        store_split(addr, xs);
                        emit_done(S, E, em, g);
                    });
                    emit_end(S, E, em);
                    TCT_END(_t7, 5);
                    break;
                }
                case OP_RAW_IN: {
                    TCT_BEGIN(_t8);
                    const float* src = (kSampler ? R.y : R.x) + E.grow * P.M;
                    const bool vec4 = (P.M & 3) == 0;
                    Emitter em;
                    emit_begin(em, E, np, false);
                    for (int g = 0; g < ng; ++g) {
                        const uint32_t addr = emit_addr(S, E, em, g);
                        if (vec4) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 t = make_uint4(0u, 0u, 0u, 0u);
                                if (E.valid && g * 16 + q * 4 < dt) t = *reinterpret_cast<const uint4*>(src + g * 16 + q * 4);
                                r[q * 4] = t.x; r[q * 4 + 1] = t.y; r[q * 4 + 2] = t.z; r[q * 4 + 3] = t.w;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j) r[j] = (E.valid && g * 16 + j < dt) ? __float_as_uint(src[g * 16 + j]) : 0u;
                        }
                        track_amax(E, r);
                        pack16(r, x);
                        store_split(addr, x);
                        emit_done(S, E, em, g);
                    }
                    emit_end(S, E, em);
                    TCT_END(_t8, 5);
                    break;
                }
                case OP_OUT: {
                    TCT_BEGIN(_t9);
                    // One column quad (4 columns) per iteration of a ROLLED loop: this op runs once per pass over at most
                    // 32 quads, and unrolled (with the Philox rounds and Box-Muller inlined per quad) it was a third of the
                    // kernel's instructions -- instruction-cache footprint the hot LayerNorm loops pay for.
                    float4* stash = reinterpret_cast<float4*>(E.scr + P.stash_off) + row;
                    const float w1 = 1.0f + R.omega, w0 = R.omega;
                    const float ce = R.c_eps[step], crs = R.c_rs[step], cn = R.c_noise[step];
                    const bool add_noise = step > 1;
                    const bool want_stats = step > R.T - 1 - R.norm_steps;
                    const int64_t plane = R.B * (int64_t)P.M;
                    const int64_t pidx = (int64_t)(R.T - 1 - step) * plane;
                    const bool vec4 = (P.M & 3) == 0;           // rows are 16-byte aligned: float4 traffic
                    const int nq = (dt + 3) >> 2;
#pragma unroll 1
                    for (int q = 0; q < nq; ++q) {
                        const int c0 = q * 4;
                        float xv[4];
                        tmem_ld4(ta + c0, xv);
                        if (!kSampler) {
                            if (E.valid) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (c0 + j < dt) R.eps[E.grow * P.M + c0 + j] = xv[j];
                            }
                            continue;
                        }
                        if (pass == 0) {           // unconditional pass: park eps_0
                            stash[(size_t)q * kRows] = make_float4(xv[0], xv[1], xv[2], xv[3]);
                            continue;
                        }
                        if (!E.valid) continue;
                        // conditional pass: guidance mix + posterior update (classifier_free_MSR.py:132-134)
                        const int64_t idx = E.grow * P.M + c0;
                        const float4 e0 = stash[(size_t)q * kRows];
                        const float e0a[4] = {e0.x, e0.y, e0.z, e0.w};
                        float z[4] = {0.f, 0.f, 0.f, 0.f}, yo[4], yn[4], ev[4];
                        if (add_noise) {
                            if (R.noise == nullptr) {
                                philox_normal4((uint64_t)E.grow + R.offset, (uint32_t)step, (uint32_t)q, R.seed, z);
                            } else if (vec4) {
                                const float4 t = *reinterpret_cast<const float4*>(R.noise + pidx + idx);
                                z[0] = t.x; z[1] = t.y; z[2] = t.z; z[3] = t.w;
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j) if (c0 + j < dt) z[j] = R.noise[pidx + idx + j];
                            }
                        }
                        if (vec4) {
                            const float4 t = *reinterpret_cast<const float4*>(R.y + idx);
                            yo[0] = t.x; yo[1] = t.y; yo[2] = t.z; yo[3] = t.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) yo[j] = (c0 + j < dt) ? R.y[idx + j] : 0.f;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ev[j] = w1 * xv[j] - w0 * e0a[j];
                            yn[j] = (yo[j] - ce * ev[j]) * crs;
                            if (add_noise) yn[j] += cn * z[j];
                            if (want_stats && c0 + j < dt) { st_s += (double)yn[j]; st_q += (double)yn[j] * (double)yn[j]; }
                        }
                        if (vec4) {
                            *reinterpret_cast<float4*>(R.y + idx) = make_float4(yn[0], yn[1], yn[2], yn[3]);
                            if (R.rec_eps) *reinterpret_cast<float4*>(R.rec_eps + pidx + idx) = make_float4(ev[0], ev[1], ev[2], ev[3]);
                            if (R.rec_y && !want_stats) *reinterpret_cast<float4*>(R.rec_y + pidx + idx) = make_float4(yn[0], yn[1], yn[2], yn[3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (c0 + j < dt) {
                                    R.y[idx + j] = yn[j];
                                    if (R.rec_eps) R.rec_eps[pidx + idx + j] = ev[j];
                                    if (R.rec_y && !want_stats) R.rec_y[pidx + idx + j] = yn[j];
                                }
                        }
                    }
                    TCT_END(_t9, 7);
                    break;
                }
                default:
                    break;
            }
        }
        if (has_pkg) {
            mbar_arrive(&S.p_empty[psl]);
            ++pseq;
        }
#ifdef DIFFSG_TC_TIMING
        if (E.row == 0 && si < 96) {
            S.t_wait[si] += (uint32_t)(_s2 - _s1);
            S.t_work[si] += (uint32_t)(clock64() - _s0 - (_s2 - _s1));
        }
#endif
    }
}

// Which chunks of a stage are issued in this pass: cond chunks only in a conditional pass.
__device__ __forceinline__ bool chunk_active(const Chunk& ch, bool use_cond) {
    return !(ch.flags & kChunkCond) || use_cond;
}
// index of the last chunk of the stage issued in this pass (its W-ring stage also carries the bias image)
__device__ __forceinline__ int last_active_chunk(const Stage& sg, bool use_cond) {
    int last = sg.chunk_begin;
    for (int ci = sg.chunk_begin; ci < sg.chunk_begin + sg.n_chunks; ++ci)
        if (chunk_active(c_chunks[ci], use_cond)) last = ci;
    return last;
}
// The bias MMA is skipped only for time biases in forward mode (rows carry their own time index: the epilogue
// adds the row's fp32 table slice instead).
template <bool kSampler>
__device__ __forceinline__ bool stage_has_bias_mma(const Stage& sg) { return kSampler || !(sg.bits & 8); }

template <bool kSampler, bool kFast>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) tc_unet_kernel(TcDev P, RunArgs R) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    SmemLayout& S = *reinterpret_cast<SmemLayout*>(sm);
    const int w_terms = P.nterms == 3 ? 2 : 1;
    const int w_stage_bytes = w_terms * kWStageBytes + kBiasBytes;
    uint8_t* w_ring = sm + ((sizeof(SmemLayout) + 127) & ~size_t(127));   // [kWStages] x [w_terms weight images | bias image]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < kASlots; ++i) { mbar_init(&S.a_full[i], kArrivals); mbar_init(&S.a_empty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (int i = 0; i < kPSlots; ++i) { mbar_init(&S.p_full[i], 1); mbar_init(&S.p_empty[i], kArrivals); }
        mbar_init(&S.acc_full, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < kRows) {
        // constant A tile of the bias chunks: K columns 0..2 = 1 (the three fp16 terms of the bias), rest 0;
        // core-matrix layout [row / 8][k / 8][row % 8][8]
        const uint32_t rr = threadIdx.x;
        uint4* o = reinterpret_cast<uint4*>(S.ones + (rr >> 3) * 256 + (rr & 7) * 16);
        o[0] = make_uint4(0x3C003C00u, 0x00003C00u, 0u, 0u);
        o[8] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
    }
    if (warp == kEpiWarps) { tmem_alloc(&S.tmem_base, kTmemCols); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();

    const int64_t n_tiles = (R.B + kRows - 1) / kRows;
    const int n_pass = kSampler ? 2 : 1;
    const int step_hi = kSampler ? R.step_hi : 0, step_lo = kSampler ? R.step_lo : 0;

    if (warp == kEpiWarps) {
        // =========================== TMA producer: parameter packages + weight / bias chunks
        if (lane == 0) {
            uint32_t wseq = 0, pseq = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int step = step_hi; step >= step_lo; --step)
                    for (int pass = 0; pass < n_pass; ++pass) {
                        const bool use_cond = kSampler ? (pass == 1) : true;
                        for (int si = 0; si < P.n_stages; ++si) {
                            const Stage sg = c_stages[si];
                            if (sg.pkg_f4) {
                                const uint32_t sl = pseq % kPSlots;
                                const uint32_t bytes = (uint32_t)sg.pkg_f4 * 16u;
                                mbar_wait_parked(&S.p_empty[sl], ((pseq / kPSlots) & 1) ^ 1);
                                mbar_arrive_expect_tx(&S.p_full[sl], bytes);
                                tma_load_1d(S.pkg[sl], P.params + (size_t)sg.pkg_off16 * 16, bytes, &S.p_full[sl]);
                                ++pseq;
                            }
                            if (!(sg.bits & 4)) continue;
                            const int last = last_active_chunk(sg, use_cond);
                            for (int ci = sg.chunk_begin; ci <= last; ++ci) {
                                const Chunk ch = c_chunks[ci];
                                if (!chunk_active(ch, use_cond)) continue;
                                const uint32_t st = wseq % kWStages, ph = (wseq / kWStages) & 1;
                                const uint32_t bytes = (uint32_t)sg.n16 * 16u * ch.kw * 2u;
                                const bool bias = ci == last && stage_has_bias_mma<kSampler>(sg);
                                const uint32_t bias_bytes = bias ? (uint32_t)sg.n16 * 16u * kBiasK * 2u : 0u;
                                mbar_wait_parked(&S.w_empty[st], ph ^ 1);
                                mbar_arrive_expect_tx(&S.w_full[st], bytes * w_terms + bias_bytes);
                                uint8_t* dst = w_ring + (size_t)st * w_stage_bytes;
                                tma_load_1d(dst, P.w_hi + (size_t)ch.w_off16 * 16, bytes, &S.w_full[st]);
                                if (w_terms == 2)
                                    tma_load_1d(dst + kWStageBytes, P.w_lo + (size_t)ch.w_off16 * 16, bytes, &S.w_full[st]);
                                if (bias) {
                                    const uint8_t* src = (sg.bits & 8) ? P.tt_img + (size_t)step * P.tt_img_stride : P.w_hi;
                                    tma_load_1d(dst + w_terms * kWStageBytes, src + (size_t)sg.bias_off16 * 16, bias_bytes, &S.w_full[st]);
                                }
                                ++wseq;
                            }
                        }
                    }
        }
    } else if (warp == kEpiWarps + 1) {
        // =========================== MMA issuer
        if (lane == 0) {
            uint32_t wseq = 0, aseq = 0;
#ifdef DIFFSG_TC_TIMING
            long long mma_t_w = 0, mma_t_a = 0;
#endif
            const uint64_t d_ones = make_smem_desc(smem_u32(S.ones), 128, 256, 0);
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int step = step_hi; step >= step_lo; --step)
                    for (int pass = 0; pass < n_pass; ++pass) {
                        const bool use_cond = kSampler ? (pass == 1) : true;
                        for (int si = 0; si < P.n_stages; ++si) {
                            const Stage sg = c_stages[si];
                            if (!(sg.bits & 4)) continue;
                            const uint32_t idesc = make_idesc_f16(128, (uint32_t)sg.n16 * 16u);
                            const uint32_t d_tmem = S.tmem_base + (sg.bits & 1) * kRegionCols;
                            uint32_t acc = (sg.bits >> 1) & 1;
                            const int last = last_active_chunk(sg, use_cond);
                            for (int ci = sg.chunk_begin; ci <= last; ++ci) {
                                const Chunk ch = c_chunks[ci];
                                if (!chunk_active(ch, use_cond)) continue;
                                const uint32_t st = wseq % kWStages, wph = (wseq / kWStages) & 1;
                                const uint32_t sl = aseq % kASlots, aph = (aseq / kASlots) & 1;
                                const uint32_t sbo = (uint32_t)ch.kw * 16u;
                                const uint8_t* wst = w_ring + (size_t)st * w_stage_bytes;
                                const uint64_t dw_hi = make_smem_desc(smem_u32(wst), 128, sbo, 0);
                                const uint64_t dw_lo = make_smem_desc(smem_u32(wst + kWStageBytes), 128, sbo, 0);
                                const uint64_t da_hi = make_smem_desc(smem_u32(S.a_hi[sl]), 128, sbo, 0);
                                const uint64_t da_lo = make_smem_desc(smem_u32(S.a_lo[sl]), 128, sbo, 0);
#ifdef DIFFSG_TC_TIMING
                                const long long _m0 = clock64();
#endif
                                mbar_wait_parked(&S.w_full[st], wph);       // weights were requested long ago: off the critical path
#ifdef DIFFSG_TC_TIMING
                                const long long _m1 = clock64();
#endif
                                mbar_wait_spin(&S.a_full[sl], aph);         // the operand chunk is what the tile is waiting for
#ifdef DIFFSG_TC_TIMING
                                // how long the weights were still missing AFTER the operand chunk was ready = exposed W latency
                                mma_t_w += _m1 - _m0; mma_t_a += clock64() - _m1;
#endif
                                tcgen05_fence_after();
                                for (uint32_t ks = 0; ks < ch.kw / 16u; ++ks) {
                                    const uint64_t adv = (uint64_t)(ks * 16u);      // 256 bytes >> 4
                                    umma_f16(d_tmem, da_hi + adv, dw_hi + adv, idesc, acc);
                                    acc = 1;
                                    if (P.nterms >= 2) umma_f16(d_tmem, da_lo + adv, dw_hi + adv, idesc, 1);
                                    if (P.nterms >= 3) umma_f16(d_tmem, da_hi + adv, dw_lo + adv, idesc, 1);
                                }
                                umma_commit(&S.a_empty[sl]);
                                if (ci == last && stage_has_bias_mma<kSampler>(sg))
                                    // accumulator += 1 . [b_hi, b_mid, b_lo]: the constant ones tile is the A operand
                                    umma_f16(d_tmem, d_ones, make_smem_desc(smem_u32(wst + w_terms * kWStageBytes), 128, 256, 0), idesc, 1);
                                umma_commit(&S.w_empty[st]);
                                ++wseq; ++aseq;
                            }
                            umma_commit(&S.acc_full);
                        }
                    }
#ifdef DIFFSG_TC_TIMING
            if (blockIdx.x == 0 && P.debug) { P.debug[9] = mma_t_w; P.debug[10] = mma_t_a; }
#endif
        }
    } else {
        // =========================== epilogue / operand producers (thread == row == TMEM lane)
        EpiCtx E;
        E.lane = lane;
        E.row = threadIdx.x;
        E.tmem_row = S.tmem_base + ((uint32_t)(32 * warp) << 16);
        E.aseq = 0;
        E.a_row = (uint32_t)(E.row & 7) * 16u;
        E.a_hi0 = smem_u32(S.a_hi[0]);
        E.amax = 0.f;
#ifdef DIFFSG_TC_TIMING
        for (int i = 0; i < 12; ++i) E.tacc[i] = 0;
        if (threadIdx.x == 0) for (int i = 0; i < 96; ++i) { S.t_wait[i] = 0; S.t_work[i] = 0; }
#endif
        E.scr = P.scratch + (size_t)blockIdx.x * P.scratch_floats;
        uint32_t acc_phase = 0, pseq = 0;
        double st_s = 0.0, st_q = 0.0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            E.grow = tile * kRows + E.row;
            E.valid = E.grow < R.B;
            // cond image: swish(cond * mask) as fp16 (hi, lo), written in operand-chunk layout (per <= 64-column chunk:
            // [row / 8][piece][row % 8][8 halves]; all hi chunks, then all lo chunks) so that emit_cond can bulk-copy it
            {
                uint8_t* img = reinterpret_cast<uint8_t*>(E.scr + P.cond_off);
                const int nkc = P.Cp / 8;
                const float mk = (!kSampler && R.mask && E.valid) ? R.mask[E.grow] : 1.0f;
                for (int kc = 0; kc < nkc; ++kc) {
                    float xc[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = kc * 8 + j;
                        xc[j] = (E.valid && c < P.C) ? swish_exact(R.cond[E.grow * P.C + c] * mk) : 0.f;
                    }
                    uint4 hi, lo;
                    split2(pk(xc[0], xc[1]), hi.x, lo.x);
                    split2(pk(xc[2], xc[3]), hi.y, lo.y);
                    split2(pk(xc[4], xc[5]), hi.z, lo.z);
                    split2(pk(xc[6], xc[7]), hi.w, lo.w);
                    const int c0 = kc & ~7, nk = min(8, nkc - c0);     // chunk of this piece and its piece count
                    const size_t off = (size_t)c0 * 128 * (kRows / 8) + (size_t)(E.row >> 3) * (nk * 128) + (kc - c0) * 128 + E.a_row;
                    *reinterpret_cast<uint4*>(img + off) = hi;
                    *reinterpret_cast<uint4*>(img + (size_t)P.Cp * kRows * 2 + off) = lo;
                }
                fence_proxy_async_all();           // generic-proxy global writes -> visible to the bulk-TMA reads of emit_cond
            }
            const int trow_fwd = (!kSampler && E.valid) ? R.t_idx[E.grow] : 0;
            for (int step = step_hi; step >= step_lo; --step)
                for (int pass = 0; pass < n_pass; ++pass) {
                    const bool use_cond = kSampler ? (pass == 1) : true;
                    TCT_BEGIN(_tt);
                    run_epilogue<kSampler, kFast>(S, P, R, E, kSampler ? step : trow_fwd, use_cond, pass, step, acc_phase,
                                           pseq, st_s, st_q);
                    TCT_END(_tt, 8);
                }
        }
#ifdef DIFFSG_TC_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 0 && P.debug) {
            for (int i = 0; i < 9; ++i) P.debug[i] = E.tacc[i];
            for (int i = 0; i < 96; ++i) { P.debug[12 + i] = S.t_wait[i]; P.debug[108 + i] = S.t_work[i]; }
        }
#endif
        if (E.amax > 65504.0f && P.status) atomicOr(P.status, kStatusOverflow);
        if (kSampler && R.step_hi > R.T - 1 - R.norm_steps) {
            st_s = warp_sum(st_s);
            st_q = warp_sum(st_q);
            if (lane == 0) {
                atomicAdd(R.stats + 2 * R.step_hi, st_s);
                atomicAdd(R.stats + 2 * R.step_hi + 1, st_q);
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == kEpiWarps) tmem_dealloc(S.tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace diffsg
