// Tensor-core engine: one CTA carries a 128-row tile through the whole UNet1D stage program;
// two CTAs are co-resident per SM (fp16x2) so one tile's MMAs overlap the other's epilogue.
//
//   warp 0        bulk-TMA producer: streams fp16 weight K-chunk images into the W ring and the
//                 per-stage fp32 parameter package (bias / LayerNorm gamma, beta / time-bias slice)
//                 into the package ring
//   warp 1        MMA issuer: one thread issues tcgen05.mma (A, W from shared memory, fp32
//                 accumulators in TMEM) and commits completion to mbarriers
//   warps 2-5     epilogue / operand producers: thread == row == TMEM lane.  Every epilogue op STREAMS
//                 its source (the TMEM accumulator, a skip vector, the input row) in groups of 16
//                 columns: pass 1 accumulates the LayerNorm statistics, pass 2 re-reads the
//                 accumulator, normalises, applies Swish in fp32, splits into fp16 (hi, lo) and writes
//                 the next GEMM's A operand as core-matrix K-chunks.  Only one 16-column group lives
//                 in registers, so the epilogue is a handful of small loops that stay in the
//                 instruction cache (the unrolled row-in-registers version was instruction-fetch bound).
//
// Program format: diffsg_b200/tc_packer.py.  Reference semantics: ddpm_opt/UNetCF.py:83-95,
// :318-356; sampler: ddpm_opt/classifier_free_MSR.py:124-137.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace diffsg {
namespace tc {

// Build variant (diffsg_b200/_lib.py TC_VARIANT): K columns per operand chunk, A-ring depth, TMEM columns per
// accumulator region and the number of co-resident CTAs (tiles) per SM the resources are budgeted for.
#ifndef DIFFSG_TC_CHUNK
#define DIFFSG_TC_CHUNK 64
#endif
#ifndef DIFFSG_TC_ASLOTS
#define DIFFSG_TC_ASLOTS 2
#endif
#ifndef DIFFSG_TC_REGION
#define DIFFSG_TC_REGION 128
#endif
#ifndef DIFFSG_TC_CTAS
#define DIFFSG_TC_CTAS 2
#endif
constexpr int kRows = 128;
constexpr int kChunkK = DIFFSG_TC_CHUNK;
constexpr int kGroupsPerChunk = kChunkK / 16, kPiecesPerChunk = kChunkK / 8;
constexpr int kSlotBytes = kRows * kChunkK * 2;      // one fp16 A chunk (16 KB at 64 columns)
constexpr int kASlots = DIFFSG_TC_ASLOTS;
constexpr int kWStages = 2;
constexpr int kRegionCols = DIFFSG_TC_REGION;        // widest vector the engine carries
constexpr int kWStageBytes = kRegionCols * kChunkK * 2;   // one fp16 W chunk (N <= kRegionCols)
static_assert(kChunkK == 32 || kChunkK == 64, "chunk width");
static_assert(kRegionCols == 64 || kRegionCols == 128, "region width");
#ifndef DIFFSG_TC_SETS
#define DIFFSG_TC_SETS 1
#endif
// Epilogue warp sets: set s owns the 16-column groups g == s (mod kEpiSets) of every vector, so with
// two sets eight warps (two per TMEM lane quarter) share one tile's epilogue.
constexpr int kEpiSets = DIFFSG_TC_SETS;
constexpr int kEpiThreads = 128 * kEpiSets;
constexpr int kCtasPerSm = DIFFSG_TC_CTAS;
// Register split (setmaxnreg): needed for > 2 CTAs per SM or two epilogue sets.  setmaxnreg is a WARPGROUP
// instruction (4 aligned warps execute the same one), so the split build pads the producer side to a full
// warpgroup: warps 0-3 = producers (TMA lane, MMA lane, two idle warps), epilogue from warp 4.
#ifndef DIFFSG_TC_SPLITREGS
#define DIFFSG_TC_SPLITREGS (kEpiSets == 2 || kCtasPerSm > 2)
#endif
constexpr bool kSplitRegs = DIFFSG_TC_SPLITREGS;
constexpr int kEpiWarp0 = kSplitRegs ? 4 : 2;        // no split: warps 2..5 are the epilogue
constexpr int kThreads = kEpiWarp0 * 32 + kEpiThreads;   // 192 (default) / 256 / 384
// Launch budget L per thread: the register file is 4 x 16384 (one bank per scheduler), a scheduler hosts
// ceil(CTAs x warps / 4) warps, L is that share rounded down to 8 (what ptxas derives from __launch_bounds__).
// The producer warpgroup keeps kRegsProducer, the epilogue threads get everything that frees.
constexpr int kWarpsPerScheduler = (kCtasPerSm * (kThreads / 32) + 3) / 4;
constexpr int kRegsLaunch = (16384 / (kWarpsPerScheduler * 32)) / 8 * 8;
constexpr int kRegsProducer = 24;
constexpr int kRegsEpilogue = (kRegsLaunch + (kRegsLaunch - kRegsProducer) * (kEpiWarp0 * 32) / kEpiThreads) / 8 * 8;
static_assert(!kSplitRegs || (kRegsProducer * kEpiWarp0 * 32 + kRegsEpilogue * kEpiThreads <= kRegsLaunch * kThreads && kRegsEpilogue <= 232),
              "setmaxnreg split exceeds the CTA's register allocation");
constexpr int kTmemCols = 2 * kRegionCols;           // two accumulator regions
constexpr int kMaxStages = 256, kMaxChunks = 512, kMaxEpi = 1024;   // program lives in __constant__ memory (16 KB)
constexpr int kPkgFloats = 640, kPSlots = 2;

// streaming epilogue ops (diffsg_b200/tc_packer.py)
enum : int { OP_LN = 1, OP_CATLN, OP_RAW_T, OP_RAW_S, OP_RAW_IN, OP_OUT };
constexpr int kChunkCond = 1;
constexpr int kFTime = 1, kFCond = 2, kFPush = 4, kFDefer = 8;

struct __align__(8) Epi { uint8_t kind, np, dt, misc, slot, off0, off1, off2; };   // misc: region | flags << 1
struct __align__(8) Chunk { uint16_t kw, flags; uint32_t w_off16; };
struct __align__(16) Stage {
    uint16_t chunk_begin, epi_begin;
    uint8_t n_chunks, n_epi, n16, bits;          // bits: region | accumulate << 1 | has_gemm << 2 | has_time << 3
    uint32_t pkg_off4;
    uint16_t tt_src4;
    uint8_t pkg_f4, tt_f4;
};
static_assert(sizeof(Epi) == 8 && sizeof(Chunk) == 8 && sizeof(Stage) == 16, "program record layout");

// The active stage program (uploaded by the host before a launch whenever the plan changes).
__constant__ Stage c_stages[kMaxStages];
__constant__ Chunk c_chunks[kMaxChunks];
__constant__ Epi c_epis[kMaxEpi];

struct TcDev {
    int n_stages, n_chunks, n_epi;
    const uint8_t* w_hi; const uint8_t* w_lo;     // fp16 weight images (lo: nterms == 3 only)
    const float* params; const float* tt;
    int tt_stride, nterms;
    int M, Mp, C, Cp;
    long long* debug;
    float* scratch;                                 // per CTA: skip stack + eps stash + cond image
    size_t scratch_floats;                          // per CTA
    int skip_off[kMaxSkip];                         // float offset of each skip slot inside the CTA scratch
    int stash_off, cond_off;                        // float offsets (cond image: hi then lo, fp16)
};

struct SmemLayout {
    uint8_t a_hi[kASlots][kSlotBytes];
    uint8_t a_lo[kASlots][kSlotBytes];
    float pkg[kPSlots][kPkgFloats];
    float2 xch[2][kEpiSets == 2 ? 256 : 1];         // per-row moment exchange between the two sets
    uint64_t a_full[kASlots], a_empty[kASlots], w_full[kWStages], w_empty[kWStages], p_full[kPSlots],
        p_empty[kPSlots], acc_full;
    uint32_t tmem_base, pad_;
    // followed by the W ring: kWStages * (nterms == 3 ? 2 : 1) * kWStageBytes (dynamic)
};
static_assert((((sizeof(SmemLayout) + 127) & ~size_t(127)) + kWStages * kWStageBytes + 128 + 1024) * kCtasPerSm <= 233472,
              "kCtasPerSm CTAs (fp16x2) must fit the SM's 228 KB of shared memory");

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

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// x * sigmoid(x) with one MUFU.EX2 and one MUFU.RCP (relative error ~3e-7; correct limits at +-inf)
__device__ __forceinline__ float swish_f(float x) {
    return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer)); }
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue)); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __uint_as_float(r);
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
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
    long long tacc[12];   // 0 acc wait, 1 pkg wait, 2 LN pass 1, 3 exchange, 4 LN pass 2 (incl 5, 6), 5 a_empty wait, 6 publish, 7 cond, 8 catln, 9 out, 10 total, 11 raw ops
#endif
    int row, set, et;
    uint32_t xpar;          // parity of the moment-exchange buffer
    uint32_t tmem_row;      // TMEM address of this thread's lane, column 0
    uint32_t aseq;          // A-ring sequence number (chunks published so far by the tile)
    float* scr;             // this CTA's global scratch
    int64_t grow;           // global row
    bool valid;
};

// ---- A-operand ring (producer side): one vector of `np` 8-column pieces -> ceil(np / 8) K-chunks.
// Group g (pieces 2g, 2g+1) is written by set g % kEpiSets; EVERY epilogue thread arrives once per
// chunk (after waiting for the slot), whether or not it wrote into it.
struct Emitter {
    uint32_t seq0;
    int np, ng;
    bool defer;
};
__device__ __forceinline__ void emit_begin(Emitter& em, const EpiCtx& E, int np, bool defer) {
    em.seq0 = E.aseq; em.np = np; em.ng = np >> 1; em.defer = defer;
}
#ifdef DIFFSG_TC_TIMING
#define emit_wait_slot(S, sq) { TCT_BEGIN(_tw); mbar_wait(&(S).a_empty[(sq) % kASlots], ((((sq)) / kASlots) & 1) ^ 1); TCT_END(_tw, 5); }
#define emit_publish(S, sq) { TCT_BEGIN(_tp); fence_proxy_async_smem(); tcgen05_fence_before(); mbar_arrive(&(S).a_full[(sq) % kASlots]); TCT_END(_tp, 6); }
#else
__device__ __forceinline__ void emit_wait_slot(SmemLayout& S, uint32_t sq) {
    mbar_wait(&S.a_empty[sq % kASlots], ((sq / kASlots) & 1) ^ 1);
}
__device__ __forceinline__ void emit_publish(SmemLayout& S, uint32_t sq) {
    fence_proxy_async_smem();
    tcgen05_fence_before();
    mbar_arrive(&S.a_full[sq % kASlots]);
}
#endif
// first / last group of chunk c owned by set `set` (first > last: none)
__device__ __forceinline__ void my_groups(const Emitter& em, int c, int set, int& first, int& last) {
    const int lo = kGroupsPerChunk * c, hi = min(kGroupsPerChunk * c + kGroupsPerChunk - 1, em.ng - 1);
    first = lo + ((set - lo) & (kEpiSets - 1));
    last = hi - ((hi - set) & (kEpiSets - 1));
}
__device__ __forceinline__ void emit_group(SmemLayout& S, EpiCtx& E, const Emitter& em, int g, const float (&x)[16]) {
    const int c = g / kGroupsPerChunk;
    const uint32_t sq = em.seq0 + c;
    const uint32_t sl = sq % kASlots;
    int first, last;
    my_groups(em, c, E.set, first, last);
    if (g == first) emit_wait_slot(S, sq);
    const int cnt = min(kPiecesPerChunk, em.np - kPiecesPerChunk * c);         // pieces in chunk c
    const uint32_t off = (uint32_t)(E.row >> 3) * (cnt * 128) + ((g % kGroupsPerChunk) * 2) * 128 + (E.row & 7) * 16;
    uint4 hi, lo;
    split_pack8(*reinterpret_cast<const float(*)[8]>(&x[0]), hi, lo);
    *reinterpret_cast<uint4*>(S.a_hi[sl] + off) = hi;
    *reinterpret_cast<uint4*>(S.a_lo[sl] + off) = lo;
    split_pack8(*reinterpret_cast<const float(*)[8]>(&x[8]), hi, lo);
    *reinterpret_cast<uint4*>(S.a_hi[sl] + off + 128) = hi;
    *reinterpret_cast<uint4*>(S.a_lo[sl] + off + 128) = lo;
    if (!em.defer && g == last) emit_publish(S, sq);
}
__device__ __forceinline__ void emit_end(SmemLayout& S, EpiCtx& E, const Emitter& em) {
    const int nch = (em.np + kPiecesPerChunk - 1) / kPiecesPerChunk;
    for (int c = 0; c < nch; ++c) {
        int first, last;
        my_groups(em, c, E.set, first, last);
        const bool mine = first <= last;
        if (!mine) emit_wait_slot(S, em.seq0 + c);          // no group of this chunk is mine: still take part
        if (!mine || em.defer) emit_publish(S, em.seq0 + c);
    }
    E.aseq += nch;
}

// ---- group sources --------------------------------------------------------------------------
// accumulator columns [16g, 16g+16) of `region` + bias (bias: shared-memory package, or the row's
// time-table slice in global memory for forward mode)
__device__ __forceinline__ void load_group_tmem(float (&x)[16], uint32_t ta, const float* bias) {
    tmem_ld16(ta, x);
    float4 b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) b[q] = *reinterpret_cast<const float4*>(bias + q * 4);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        x[q * 4 + 0] += b[q].x; x[q * 4 + 1] += b[q].y; x[q * 4 + 2] += b[q].z; x[q * 4 + 3] += b[q].w;
    }
}
// Software-pipelined walk over the accumulator: the raw TMEM load of group g+1 is in flight while
// group g is processed.  `raw` holds the prefetched group (no bias yet).
struct TmemWalk {
    float raw[16];
};
__device__ __forceinline__ void walk_begin(TmemWalk& w, uint32_t ta) {
    tmem_ld16(ta, w.raw);
    tmem_ld_wait();
}
// x = prefetched group + bias; start fetching the next group (if any); caller must call walk_sync()
// before the next walk_next()
__device__ __forceinline__ void walk_next(TmemWalk& w, float (&x)[16], const float* bias, bool more, uint32_t ta_next) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(bias + q * 4);
        x[q * 4 + 0] = w.raw[q * 4 + 0] + b.x; x[q * 4 + 1] = w.raw[q * 4 + 1] + b.y;
        x[q * 4 + 2] = w.raw[q * 4 + 2] + b.z; x[q * 4 + 3] = w.raw[q * 4 + 3] + b.w;
    }
    if (more) tmem_ld16(ta_next, w.raw);
}
__device__ __forceinline__ void walk_sync() { tmem_ld_wait(); }
__device__ __forceinline__ void load_group_skip(float (&x)[16], const float4* sk) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 t = sk[q * kRows];
        x[q * 4 + 0] = t.x; x[q * 4 + 1] = t.y; x[q * 4 + 2] = t.z; x[q * 4 + 3] = t.w;
    }
}
__device__ __forceinline__ void store_group_skip(const float (&x)[16], float4* sk) {
#pragma unroll
    for (int q = 0; q < 4; ++q) sk[q * kRows] = make_float4(x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
}

// Walk a skip vector (global scratch slab, L2 / HBM resident, ~1 us latency) in batches of two groups:
// all 8 16-byte loads of a batch are issued back to back, so the latency is paid once per batch
// instead of once per group.  (Loads are NOT kept in flight across body(): the operand-publish fence
// inside emit_group would wait for them.)
template <typename F>
__device__ __forceinline__ void skip_walk2(const float4* sk, int g0, int G, int ng, F body) {
    for (int gb = g0; gb < ng; gb += 2 * G) {
        const bool two = gb + G < ng;
        const float4* p0 = sk + (size_t)gb * 4 * kRows;
        const float4* p1 = sk + (size_t)(two ? gb + G : gb) * 4 * kRows;
        const float4 a0 = p0[0], a1 = p0[kRows], a2 = p0[2 * kRows], a3 = p0[3 * kRows];
        const float4 c0 = p1[0], c1 = p1[kRows], c2 = p1[2 * kRows], c3 = p1[3 * kRows];
        float x[16];
        x[0] = a0.x; x[1] = a0.y; x[2] = a0.z; x[3] = a0.w; x[4] = a1.x; x[5] = a1.y; x[6] = a1.z; x[7] = a1.w;
        x[8] = a2.x; x[9] = a2.y; x[10] = a2.z; x[11] = a2.w; x[12] = a3.x; x[13] = a3.y; x[14] = a3.z; x[15] = a3.w;
        body(gb, x);
        if (two) {
            x[0] = c0.x; x[1] = c0.y; x[2] = c0.z; x[3] = c0.w; x[4] = c1.x; x[5] = c1.y; x[6] = c1.z; x[7] = c1.w;
            x[8] = c2.x; x[9] = c2.y; x[10] = c2.z; x[11] = c2.w; x[12] = c3.x; x[13] = c3.y; x[14] = c3.z; x[15] = c3.w;
            body(gb + G, x);
        }
    }
}

// shifted one-pass moments: s1 += (x - shift), s2 += (x - shift)^2 over the first `nval` columns
__device__ __forceinline__ void moments_group(const float (&x)[16], int nval, float shift, float& s1, float& s2) {
    float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
    if (nval >= 16) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
            const float d0 = x[j] - shift, d1 = x[j + 1] - shift;
            a0 += d0; a1 += d1;
            q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float d = j < nval ? x[j] - shift : 0.f;
            a0 += d;
            q0 = fmaf(d, d, q0);
        }
    }
    s1 += a0 + a1;
    s2 += q0 + q1;
}
// x <- swish((x * a_scale + a_shift) * gamma + beta), zero beyond nval
__device__ __forceinline__ void ln_swish_group(float (&x)[16], int nval, float a_scale, float a_shift,
                                               const float* gamma, const float* beta) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + q * 4);
        const float4 b = *reinterpret_cast<const float4*>(beta + q * 4);
        x[q * 4 + 0] = swish_f(fmaf(fmaf(x[q * 4 + 0], a_scale, a_shift), g.x, b.x));
        x[q * 4 + 1] = swish_f(fmaf(fmaf(x[q * 4 + 1], a_scale, a_shift), g.y, b.y));
        x[q * 4 + 2] = swish_f(fmaf(fmaf(x[q * 4 + 2], a_scale, a_shift), g.z, b.z));
        x[q * 4 + 3] = swish_f(fmaf(fmaf(x[q * 4 + 3], a_scale, a_shift), g.w, b.w));
    }
    if (nval < 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (j >= nval) x[j] = 0.f;
    }
}
__device__ __forceinline__ void finish_moments(float s1, float s2, float shift, float n, float& a_scale, float& a_shift) {
    const float md = s1 / n;
    const float var = fmaxf(s2 / n - md * md, 0.f);
    const float rstd = rsqrtf(var + kLnEps);
    a_scale = rstd;
    a_shift = -(shift + md) * rstd;
}

// sum the partial moments of the two sets of a row
__device__ __forceinline__ void exchange_moments(SmemLayout& S, EpiCtx& E, float& s1, float& s2) {
    if (kEpiSets == 2) {
        S.xch[E.xpar][E.et] = make_float2(s1, s2);
        epi_bar_sync();
        const float2 o = S.xch[E.xpar][E.et ^ 128];
        E.xpar ^= 1;
        s1 += o.x;
        s2 += o.y;
    }
}

__device__ __forceinline__ void emit_cond(SmemLayout& S, EpiCtx& E, const TcDev& P) {
    const uint4* img = reinterpret_cast<const uint4*>(E.scr + P.cond_off);
    const int nkc = P.Cp / 8;                      // 16-byte K pieces per row
    for (int c0 = 0; c0 < nkc; c0 += kPiecesPerChunk) {
        const int nk = min(kPiecesPerChunk, nkc - c0);
        const uint32_t sq = E.aseq, sl = sq % kASlots;
        mbar_wait(&S.a_empty[sl], ((sq / kASlots) & 1) ^ 1);
        const uint32_t base = (uint32_t)(E.row >> 3) * (nk * 128) + (E.row & 7) * 16;
        for (int k0 = E.set; k0 < nk; k0 += 4 * kEpiSets) {   // 8 independent 16-byte loads in flight; piece k -> set k % kEpiSets
            uint4 h[4], l[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k0 + k * kEpiSets < nk) {
                    h[k] = img[(size_t)(c0 + k0 + k * kEpiSets) * kRows + E.row];
                    l[k] = img[(size_t)(nkc + c0 + k0 + k * kEpiSets) * kRows + E.row];
                }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k0 + k * kEpiSets < nk) {
                    *reinterpret_cast<uint4*>(S.a_hi[sl] + base + (k0 + k * kEpiSets) * 128) = h[k];
                    *reinterpret_cast<uint4*>(S.a_lo[sl] + base + (k0 + k * kEpiSets) * 128) = l[k];
                }
        }
        fence_proxy_async_smem();
        mbar_arrive(&S.a_full[sl]);
        ++E.aseq;
    }
}

template <bool kSampler>
__device__ __forceinline__ void run_epilogue(SmemLayout& S, const TcDev& P, const RunArgs& R, EpiCtx& E, int trow,
                                             bool use_cond, int pass, int step, uint32_t& acc_phase,
                                             uint32_t& pseq, double& st_s, double& st_q) {
    const int row = E.row;
    for (int si = 0; si < P.n_stages; ++si) {
        const Stage sg = c_stages[si];
        if (sg.bits & 4) {
            TCT_BEGIN(_ta);
            mbar_wait(&S.acc_full, acc_phase);
            acc_phase ^= 1;
            tcgen05_fence_after();
            TCT_END(_ta, 0);
        }
        const bool has_pkg = (sg.pkg_f4 | sg.tt_f4) != 0;
        const uint32_t psl = pseq % kPSlots;
        if (has_pkg) { TCT_BEGIN(_tk); mbar_wait(&S.p_full[psl], (pseq / kPSlots) & 1); TCT_END(_tk, 1); }
        const float* pk = S.pkg[psl];
        for (int ei = sg.epi_begin; ei < sg.epi_begin + sg.n_epi; ++ei) {
            const Epi op = c_epis[ei];
            const int np = op.np, ng = np >> 1, dt = op.dt;
            const int region = op.misc & 1, flags = op.misc >> 1;
            const uint32_t ta = E.tmem_row + region * kRegionCols;
            const float* bias = pk + op.off0 * 4;
            if (!kSampler && (flags & kFTime)) bias = P.tt + (size_t)trow * P.tt_stride + sg.tt_src4 * 4;
            float x[16];
            constexpr int G = kEpiSets;       // group stride: this thread owns groups set, set + G, ...
            const int g0 = E.set;
            switch (op.kind) {
                case OP_LN: {
                    float4* sk = reinterpret_cast<float4*>(E.scr + P.skip_off[op.slot]) + row;
                    // shift for the one-pass moments: column 0 of the (bias-added) vector
                    float shift = 0.f, s1 = 0.f, s2 = 0.f;
                    if (G == 2 && g0 != 0) shift = tmem_ld1(ta) + bias[0];
                    TmemWalk w;
                    TCT_BEGIN(_t1);
                    if (g0 < ng) walk_begin(w, ta + g0 * 16);
                    for (int g = g0; g < ng; g += G) {
                        walk_next(w, x, bias + g * 16, g + G < ng, ta + (g + G) * 16);
                        if (flags & kFPush) store_group_skip(x, sk + (size_t)g * 4 * kRows);
                        if (g == 0) shift = x[0];
                        moments_group(x, dt - g * 16, shift, s1, s2);
                        walk_sync();
                    }
                    TCT_END(_t1, 2);
                    TCT_BEGIN(_t2);
                    exchange_moments(S, E, s1, s2);
                    TCT_END(_t2, 3);
                    TCT_BEGIN(_t3);
                    float a_scale, a_shift;
                    finish_moments(s1, s2, shift, (float)dt, a_scale, a_shift);
                    Emitter em;
                    emit_begin(em, E, np, (flags & kFDefer) != 0);
                    const float* pg = pk + op.off1 * 4;
                    const float* pb = pk + op.off2 * 4;
                    if (g0 < ng) walk_begin(w, ta + g0 * 16);
                    for (int g = g0; g < ng; g += G) {
                        walk_next(w, x, bias + g * 16, g + G < ng, ta + (g + G) * 16);
                        ln_swish_group(x, dt - g * 16, a_scale, a_shift, pg + g * 16, pb + g * 16);
                        emit_group(S, E, em, g, x);
                        walk_sync();
                    }
                    emit_end(S, E, em);
                    TCT_END(_t3, 4);
                    if ((flags & kFCond) && use_cond) { TCT_BEGIN(_t4); emit_cond(S, E, P); TCT_END(_t4, 7); }
                    break;
                }
                case OP_CATLN: {
                    TCT_BEGIN(_t5);
                    // LayerNorm over cat(x, skip): statistics over both, operands: skip part, then x part
                    const float4* sk = reinterpret_cast<const float4*>(E.scr + P.skip_off[op.slot]) + row;
                    float shift = 0.f, s1 = 0.f, s2 = 0.f;
                    if (G == 2 && g0 != 0) shift = tmem_ld1(ta) + bias[0];
                    for (int g = g0; g < ng; g += G) {
                        load_group_tmem(x, ta + g * 16, bias + g * 16);
                        if (g == 0) shift = x[0];
                        moments_group(x, dt - g * 16, shift, s1, s2);
                    }
                    skip_walk2(sk, g0, G, ng, [&](int g, const float (&xg)[16]) { moments_group(xg, dt - g * 16, shift, s1, s2); });
                    exchange_moments(S, E, s1, s2);
                    float a_scale, a_shift;
                    finish_moments(s1, s2, shift, (float)(2 * dt), a_scale, a_shift);
                    const float* pgx = pk + op.off1 * 4;            // gamma_x | beta_x | gamma_s | beta_s
                    const int dp = np * 8;
                    Emitter em;
                    emit_begin(em, E, np, false);
                    skip_walk2(sk, g0, G, ng, [&](int g, float (&xg)[16]) {
                        ln_swish_group(xg, dt - g * 16, a_scale, a_shift, pgx + 2 * dp + g * 16, pgx + 3 * dp + g * 16);
                        emit_group(S, E, em, g, xg);
                    });
                    emit_end(S, E, em);
                    emit_begin(em, E, np, false);
                    for (int g = g0; g < ng; g += G) {
                        load_group_tmem(x, ta + g * 16, bias + g * 16);
                        ln_swish_group(x, dt - g * 16, a_scale, a_shift, pgx + g * 16, pgx + dp + g * 16);
                        emit_group(S, E, em, g, x);
                    }
                    emit_end(S, E, em);
                    TCT_END(_t5, 8);
                    break;
                }
                case OP_RAW_T: {
                    float4* sk = reinterpret_cast<float4*>(E.scr + P.skip_off[op.slot]) + row;
                    Emitter em;
                    emit_begin(em, E, np, (flags & kFDefer) != 0);     // deferred when the next GEMM accumulates into this region (attention)
                    for (int g = g0; g < ng; g += G) {
                        load_group_tmem(x, ta + g * 16, bias + g * 16);
                        if (flags & kFPush) store_group_skip(x, sk + (size_t)g * 4 * kRows);
                        emit_group(S, E, em, g, x);                  // pad columns are exact zeros (zero W rows, zero bias)
                    }
                    emit_end(S, E, em);
                    break;
                }
                case OP_RAW_S: {
                    const float4* sk = reinterpret_cast<const float4*>(E.scr + P.skip_off[op.slot]) + row;
                    Emitter em;
                    emit_begin(em, E, np, false);
                    skip_walk2(sk, g0, G, ng, [&](int g, const float (&xg)[16]) { emit_group(S, E, em, g, xg); });
                    emit_end(S, E, em);
                    break;
                }
                case OP_RAW_IN: {
                    const float* src = (kSampler ? R.y : R.x) + E.grow * P.M;
                    Emitter em;
                    emit_begin(em, E, np, false);
                    for (int g = g0; g < ng; g += G) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) x[j] = (E.valid && g * 16 + j < dt) ? src[g * 16 + j] : 0.f;
                        emit_group(S, E, em, g, x);
                    }
                    emit_end(S, E, em);
                    break;
                }
                case OP_OUT: {
                    float4* stash = reinterpret_cast<float4*>(E.scr + P.stash_off) + row;
                    const float w1 = 1.0f + R.omega, w0 = R.omega;
                    const float ce = R.c_eps[step], crs = R.c_rs[step], cn = R.c_noise[step];
                    const bool add_noise = step > 1;
                    const bool want_stats = step > R.T - 1 - R.norm_steps;
                    const int64_t plane = R.B * (int64_t)P.M;
                    const int64_t pidx = (int64_t)(R.T - 1 - step) * plane;
                    for (int g = g0; g < ng; g += G) {
                        load_group_tmem(x, ta + g * 16, bias + g * 16);
                        if (!kSampler) {
                            if (E.valid) {
#pragma unroll
                                for (int j = 0; j < 16; ++j)
                                    if (g * 16 + j < dt) R.eps[E.grow * P.M + g * 16 + j] = x[j];
                            }
                            continue;
                        }
                        if (pass == 0) {           // unconditional pass: park eps_0
                            store_group_skip(x, stash + (size_t)g * 4 * kRows);
                            continue;
                        }
                        // conditional pass: guidance mix + posterior update (classifier_free_MSR.py:132-134)
                        const bool vec4 = (P.M & 3) == 0;           // rows are 16-byte aligned: float4 traffic
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int c0 = g * 16 + q * 4;
                            if (c0 >= dt || !E.valid) continue;
                            const int64_t idx = E.grow * P.M + c0;
                            const float4 e0 = stash[(size_t)(g * 4 + q) * kRows];
                            const float e0a[4] = {e0.x, e0.y, e0.z, e0.w};
                            float z[4] = {0.f, 0.f, 0.f, 0.f}, yo[4], yn[4], ev[4];
                            if (add_noise) {
                                if (R.noise == nullptr) {
                                    philox_normal4((uint64_t)E.grow + R.offset, (uint32_t)step, (uint32_t)(g * 4 + q), R.seed, z);
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
                                ev[j] = w1 * x[q * 4 + j] - w0 * e0a[j];
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
                    }
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
    }
}

template <bool kSampler>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) tc_unet_kernel(TcDev P, RunArgs R) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    SmemLayout& S = *reinterpret_cast<SmemLayout*>(sm);
    const int w_terms = P.nterms == 3 ? 2 : 1;
    uint8_t* w_ring = sm + ((sizeof(SmemLayout) + 127) & ~size_t(127));   // [kWStages][w_terms][kWStageBytes]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup
    if (threadIdx.x == 0) {
        for (int i = 0; i < kASlots; ++i) { mbar_init(&S.a_full[i], kEpiThreads); mbar_init(&S.a_empty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        for (int i = 0; i < kPSlots; ++i) { mbar_init(&S.p_full[i], 1); mbar_init(&S.p_empty[i], kEpiThreads); }
        mbar_init(&S.acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(&S.tmem_base, kTmemCols); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();

    const int64_t n_tiles = (R.B + kRows - 1) / kRows;
    const int n_pass = kSampler ? 2 : 1;
    const int step_hi = kSampler ? R.step_hi : 0, step_lo = kSampler ? R.step_lo : 0;

    if (warp < kEpiWarp0) {
      if (kSplitRegs) setmaxnreg_dec();
      if (warp == 0) {
        // =========================== TMA producer: parameter packages + weight chunks
        if (lane == 0) {
            uint32_t wseq = 0, pseq = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int step = step_hi; step >= step_lo; --step)
                    for (int pass = 0; pass < n_pass; ++pass) {
                        const bool use_cond = kSampler ? (pass == 1) : true;
                        for (int si = 0; si < P.n_stages; ++si) {
                            const Stage sg = c_stages[si];
                            if (sg.pkg_f4 | sg.tt_f4) {
                                const uint32_t sl = pseq % kPSlots;
                                const bool tma_time = kSampler && sg.tt_f4;
                                const uint32_t tt_bytes = (uint32_t)sg.tt_f4 * 16u, st_bytes = (uint32_t)sg.pkg_f4 * 16u;
                                mbar_wait_parked(&S.p_empty[sl], ((pseq / kPSlots) & 1) ^ 1);
                                mbar_arrive_expect_tx(&S.p_full[sl], st_bytes + (tma_time ? tt_bytes : 0u));
                                if (tma_time)
                                    tma_load_1d(S.pkg[sl], P.tt + (size_t)step * P.tt_stride + (size_t)sg.tt_src4 * 4, tt_bytes, &S.p_full[sl]);
                                if (st_bytes)
                                    tma_load_1d(S.pkg[sl] + sg.tt_f4 * 4, P.params + (size_t)sg.pkg_off4 * 4, st_bytes, &S.p_full[sl]);
                                ++pseq;
                            }
                            if (!(sg.bits & 4)) continue;
                            for (int ci = sg.chunk_begin; ci < sg.chunk_begin + sg.n_chunks; ++ci) {
                                const Chunk ch = c_chunks[ci];
                                if ((ch.flags & kChunkCond) && !use_cond) continue;
                                const uint32_t st = wseq % kWStages, ph = (wseq / kWStages) & 1;
                                const uint32_t bytes = (uint32_t)sg.n16 * 16u * ch.kw * 2u;
                                mbar_wait_parked(&S.w_empty[st], ph ^ 1);
                                mbar_arrive_expect_tx(&S.w_full[st], bytes * w_terms);
                                uint8_t* dst = w_ring + (size_t)st * w_terms * kWStageBytes;
                                tma_load_1d(dst, P.w_hi + (size_t)ch.w_off16 * 16, bytes, &S.w_full[st]);
                                if (w_terms == 2)
                                    tma_load_1d(dst + kWStageBytes, P.w_lo + (size_t)ch.w_off16 * 16, bytes, &S.w_full[st]);
                                ++wseq;
                            }
                        }
                    }
        }
      } else if (warp == 1) {
        // =========================== MMA issuer
        if (lane == 0) {
            uint32_t wseq = 0, aseq = 0;
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
                            for (int ci = sg.chunk_begin; ci < sg.chunk_begin + sg.n_chunks; ++ci) {
                                const Chunk ch = c_chunks[ci];
                                if ((ch.flags & kChunkCond) && !use_cond) continue;
                                const uint32_t st = wseq % kWStages, wph = (wseq / kWStages) & 1;
                                const uint32_t sl = aseq % kASlots, aph = (aseq / kASlots) & 1;
                                mbar_wait_parked(&S.a_full[sl], aph);
                                mbar_wait_parked(&S.w_full[st], wph);
                                tcgen05_fence_after();
                                const uint32_t sbo = (uint32_t)ch.kw * 16u;
                                const uint8_t* wst = w_ring + (size_t)st * w_terms * kWStageBytes;
                                const uint64_t da_hi = make_smem_desc(smem_u32(S.a_hi[sl]), 128, sbo, 0);
                                const uint64_t da_lo = make_smem_desc(smem_u32(S.a_lo[sl]), 128, sbo, 0);
                                const uint64_t dw_hi = make_smem_desc(smem_u32(wst), 128, sbo, 0);
                                const uint64_t dw_lo = make_smem_desc(smem_u32(wst + kWStageBytes), 128, sbo, 0);
                                for (uint32_t ks = 0; ks < ch.kw / 16u; ++ks) {
                                    const uint64_t adv = (uint64_t)(ks * 16u);      // 256 bytes >> 4
                                    umma_f16(d_tmem, da_hi + adv, dw_hi + adv, idesc, acc);
                                    acc = 1;
                                    if (P.nterms >= 2) umma_f16(d_tmem, da_lo + adv, dw_hi + adv, idesc, 1);
                                    if (P.nterms >= 3) umma_f16(d_tmem, da_hi + adv, dw_lo + adv, idesc, 1);
                                }
                                umma_commit(&S.a_empty[sl]);
                                umma_commit(&S.w_empty[st]);
                                ++wseq; ++aseq;
                            }
                            umma_commit(&S.acc_full);
                        }
                    }
        }
      }
    } else {
        // =========================== epilogue / operand producers (thread == row == TMEM lane)
        if (kSplitRegs) setmaxnreg_inc();
        EpiCtx E;
        E.et = threadIdx.x - kEpiWarp0 * 32;
        E.set = E.et >> 7;
        E.row = 32 * (warp & 3) + lane;
        E.tmem_row = S.tmem_base + ((uint32_t)(32 * (warp & 3)) << 16);
        E.aseq = 0;
        E.xpar = 0;
#ifdef DIFFSG_TC_TIMING
        for (int i = 0; i < 12; ++i) E.tacc[i] = 0;
#endif
        E.scr = P.scratch + (size_t)blockIdx.x * P.scratch_floats;
        uint32_t acc_phase = 0, pseq = 0;
        double st_s = 0.0, st_q = 0.0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            E.grow = tile * kRows + E.row;
            E.valid = E.grow < R.B;
            // cond image: swish(cond * mask) as fp16 (hi, lo) 16-byte K pieces (thread-private scratch)
            {
                uint4* img = reinterpret_cast<uint4*>(E.scr + P.cond_off);
                const int nkc = P.Cp / 8;
                const float mk = (!kSampler && R.mask && E.valid) ? R.mask[E.grow] : 1.0f;
                for (int kc = E.set; kc < nkc; kc += kEpiSets) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = kc * 8 + j;
                        x[j] = (E.valid && c < P.C) ? swish_exact(R.cond[E.grow * P.C + c] * mk) : 0.f;
                    }
                    uint4 hi, lo;
                    split_pack8(x, hi, lo);
                    img[(size_t)kc * kRows + E.row] = hi;
                    img[(size_t)(nkc + kc) * kRows + E.row] = lo;
                }
            }
            const int trow_fwd = (!kSampler && E.valid) ? R.t_idx[E.grow] : 0;
            for (int step = step_hi; step >= step_lo; --step)
                for (int pass = 0; pass < n_pass; ++pass) {
                    const bool use_cond = kSampler ? (pass == 1) : true;
                    TCT_BEGIN(_tt);
                    run_epilogue<kSampler>(S, P, R, E, kSampler ? step : trow_fwd, use_cond, pass, step, acc_phase,
                                           pseq, st_s, st_q);
                    TCT_END(_tt, 10);
                }
        }
#ifdef DIFFSG_TC_TIMING
        if (blockIdx.x == 0 && E.et == 0 && P.debug)
            for (int i = 0; i < 12; ++i) P.debug[i] = E.tacc[i];
#endif
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
    if (warp == 0) tmem_dealloc(S.tmem_base, kTmemCols);
}

}  // namespace tc
}  // namespace diffsg
