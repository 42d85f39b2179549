// Tensor-core engine: one CTA carries a 128-row tile through the whole UNet1D stage program;
// two CTAs are co-resident per SM (fp16x2) so one tile's MMAs overlap the other's epilogue.
//
//   warp 0        bulk-TMA producer: streams fp16 weight K-chunk images into the W ring and the
//                 per-stage fp32 parameter package (bias / LayerNorm gamma, beta / time-bias slice)
//                 into the package ring
//   warp 1        MMA issuer: one thread issues tcgen05.mma (A, W from shared memory, fp32
//                 accumulators in TMEM) and commits completion to mbarriers
//   warps 4-11    epilogue / operand producers.  TMEM lane == row; the row's vector is split across
//                 TWO threads (warps 4-7: low half of the columns, warps 8-11: high half), so eight
//                 warps hide each other's latencies.  They load the accumulator into registers, add
//                 the bias, LayerNorm (statistics exchanged between the two halves through shared
//                 memory) + Swish in fp32, split into fp16 (hi, lo) and write the next GEMM's A
//                 operand as core-matrix K-chunks.
//
// Program format: diffsg_b200/tc_packer.py.  Reference semantics: ddpm_opt/UNetCF.py:83-95,
// :318-356; sampler: ddpm_opt/classifier_free_MSR.py:124-137.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace diffsg {
namespace tc {

constexpr int kRows = 128;
constexpr int kChunkK = 64;
constexpr int kSlotBytes = kRows * kChunkK * 2;      // 16 KB: one fp16 A chunk
constexpr int kASlots = 2;
constexpr int kWStages = 2;
constexpr int kWStageBytes = 128 * kChunkK * 2;      // 16 KB: one fp16 W chunk (N <= 128)
#ifndef DIFFSG_TC_SPLIT
#define DIFFSG_TC_SPLIT 1
#endif
constexpr int kSplit = DIFFSG_TC_SPLIT;              // threads per row in the epilogue (1 or 2)
constexpr int kEpiThreads = 128 * kSplit;
constexpr int kThreads = 128 + kEpiThreads;
constexpr int kEpiWarp0 = 4;                         // epilogue warps 4.. (aligned warpgroups)
constexpr int kVecRegs = 128 / kSplit;               // columns a thread keeps in registers
constexpr int kTmemCols = 256;                       // two 128-column regions
constexpr int kMaxStages = 256, kMaxChunks = 512, kMaxEpi = 1024;   // program lives in __constant__ memory (16 KB)
constexpr int kPkgFloats = 640, kPSlots = 2;
constexpr int kCtasPerSm = 2;
// setmaxnreg split of the per-thread launch budget (split 1: 128 -> 32 / 224; split 2: 80 -> 24 / 104)
constexpr int kRegsProducer = kSplit == 1 ? 32 : 24, kRegsEpilogue = kSplit == 1 ? 224 : 104;

enum : int { TE_LOAD = 1, TE_LOAD_SKIP, TE_LOAD_INPUT, TE_STORE_SKIP, TE_STORE_OUT, TE_STATS, TE_EMIT_LN,
             TE_EMIT_RAW, TE_EMIT_COND, TE_LN_BLOCK };
constexpr int kStatsReset = 1, kStatsFinish = 2, kChunkCond = 1;
constexpr int kFTime = 1, kFCond = 2, kFPush = 4;

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
    long long* debug;                               // DIFFSG_TC_TIMING builds: 12 clock64 accumulators
    float* scratch;                                 // per CTA: skip stack + eps stash + cond image
    size_t scratch_floats;                          // per CTA
    int skip_off[kMaxSkip];                         // float offset of each skip slot inside the CTA scratch
    int stash_off, cond_off;                        // float offsets (cond image: hi then lo, fp16)
};

struct SmemLayout {
    uint8_t a_hi[kASlots][kSlotBytes];
    uint8_t a_lo[kASlots][kSlotBytes];
    float pkg[kPSlots][kPkgFloats];
    float2 xchg[2][kSplit == 2 ? kEpiThreads : 1];
    uint64_t a_full[kASlots], a_empty[kASlots], w_full[kWStages], w_empty[kWStages], p_full[kPSlots],
        p_empty[kPSlots], acc_full;
    uint32_t tmem_base, pad_;
    // followed by the W ring: kWStages * (nterms == 3 ? 2 : 1) * kWStageBytes (dynamic)
};
static_assert(((sizeof(SmemLayout) + 127) & ~size_t(127)) + kWStages * kWStageBytes + 128 <= 114688, "two CTAs per SM: measured limit 112 KB each");

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
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// Per-thread epilogue state.  A vector of `np` 8-column pieces is split between the two threads of
// a row: half 0 owns pieces [0, np/2), half 1 owns [np/2, np); each thread keeps its pieces in v[].
#ifdef DIFFSG_TC_TIMING
#define TCT_BEGIN() const long long _t0 = clock64()
#define TCT_END(slot) E.tacc[slot] += clock64() - _t0
#else
#define TCT_BEGIN()
#define TCT_END(slot)
#endif
struct EpiCtx {
#ifdef DIFFSG_TC_TIMING
    long long tacc[12];   // 0 acc wait, 1 pkg wait, 2 load, 3 stats, 4 emit total, 5 a_empty wait, 6 publish, 7 cond, 8 skip ld/st, 9 out, 10 whole
#endif
    float v[kVecRegs];
    float mean, rstd, m2, cnt, cnt_all;
    uint32_t aseq;          // A-ring sequence number (chunks published so far by the whole tile)
    uint32_t xpar;          // parity of the statistics exchange buffer
    int half, row, et;
};

// ---- A-operand ring (producer side) --------------------------------------------------------
// Write this thread's NPC pieces of an np-piece vector into the chunk slots aseq, aseq+1 and publish.
// `piece(i, hi, lo)` yields the packed fp16 (hi, lo) of local piece i.
template <int NPC, typename F>
__device__ __forceinline__ void emit_pieces(SmemLayout& S, EpiCtx& E, int np, F piece) {
    const int pbeg = E.half * NPC;
    const int nch = (np + 7) >> 3;
    // chunks touched by this thread: first = pbeg >> 3, last = (pbeg + NPC - 1) >> 3  (at most two)
    const int c_first = pbeg >> 3, c_last = (pbeg + NPC - 1) >> 3;
    {
        TCT_BEGIN();
        for (int c = c_first; c <= c_last; ++c) {
            const uint32_t sq = E.aseq + c;
            mbar_wait(&S.a_empty[sq % kASlots], ((sq / kASlots) & 1) ^ 1);
        }
        TCT_END(5);
    }
#pragma unroll
    for (int i = 0; i < NPC; ++i) {
        const int p = pbeg + i, c = p >> 3, kc = p & 7;
        const int cnt = min(8, np - 8 * c);                     // pieces in chunk c
        const uint32_t sl = (E.aseq + c) % kASlots;
        const uint32_t off = (uint32_t)(E.row >> 3) * (cnt * 128) + kc * 128 + (E.row & 7) * 16;
        uint4 hi, lo;
        piece(i, hi, lo);
        *reinterpret_cast<uint4*>(S.a_hi[sl] + off) = hi;
        *reinterpret_cast<uint4*>(S.a_lo[sl] + off) = lo;
    }
    {
        TCT_BEGIN();
        fence_proxy_async_smem();
        tcgen05_fence_before();
        const int n0 = np >> 1;
        for (int c = c_first; c <= c_last; ++c) {
            // split rows: both halves contribute to chunk c iff it straddles the split point n0
            const bool both = kSplit == 2 && (8 * c < n0) && (min(8 * c + 8, np) > n0);
            mbar_arrive_n(&S.a_full[(E.aseq + c) % kASlots], (kSplit == 1 || both) ? 1u : 2u);
        }
        TCT_END(6);
    }
    E.aseq += nch;
}

template <int MODE, int NPC, bool FULL>   // MODE 0: raw, 1: swish(LN(v) * gamma + beta)
__device__ __forceinline__ void emit_vec(SmemLayout& S, EpiCtx& E, int np, int nv, const float* pk_g, const float* pk_b) {
    const float a_scale = E.rstd, a_shift = -E.mean * E.rstd;
    const int cb = E.half * NPC * 8;
    emit_pieces<NPC>(S, E, np, [&](int i, uint4& hi, uint4& lo) {
        float x[8];
        if (MODE) {
            const float4 ga = *reinterpret_cast<const float4*>(pk_g + cb + i * 8), gb = *reinterpret_cast<const float4*>(pk_g + cb + i * 8 + 4);
            const float4 ba = *reinterpret_cast<const float4*>(pk_b + cb + i * 8), bb = *reinterpret_cast<const float4*>(pk_b + cb + i * 8 + 4);
            const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
            const float bet[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float t = fmaf(fmaf(E.v[i * 8 + j], a_scale, a_shift), gam[j], bet[j]);
                x[j] = (FULL || i * 8 + j < nv) ? swish_f(t) : 0.f;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = (FULL || i * 8 + j < nv) ? E.v[i * 8 + j] : 0.f;
        }
        split_pack8(x, hi, lo);
    });
}

// LayerNorm statistics of this thread's valid columns merged into the running (cnt, mean, m2);
// on `finish` the two halves of the row exchange their partial results (Chan et al. merge).
template <int NPC, bool FULL>
__device__ __forceinline__ void stats_vec(SmemLayout& S, EpiCtx& E, int nv, int dt, int flags) {
    if (flags & kStatsReset) { E.cnt = 0.f; E.mean = 0.f; E.m2 = 0.f; E.cnt_all = 0.f; }
    constexpr int W = NPC * 8;
    if (FULL || nv > 0) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < W; ++j) s4[j & 3] += (FULL || j < nv) ? E.v[j] : 0.f;
        const float n = FULL ? (float)W : (float)nv;
        const float m = ((s4[0] + s4[1]) + (s4[2] + s4[3])) / n;
        float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < W; ++j) {
            const float d = (FULL || j < nv) ? E.v[j] - m : 0.f;
            q4[j & 3] = fmaf(d, d, q4[j & 3]);
        }
        const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        const float tot = E.cnt + n, delta = m - E.mean;
        E.mean += delta * (n / tot);
        E.m2 += q + delta * delta * (E.cnt * n / tot);
        E.cnt = tot;
    }
    E.cnt_all += (float)dt;
    if (flags & kStatsFinish) {
        if (kSplit == 2) {
            S.xchg[E.xpar][E.et] = make_float2(E.mean, E.m2);
            epi_bar_sync();
            const float2 o = S.xchg[E.xpar][E.et ^ 128];
            E.xpar ^= 1;
            const float on = E.cnt_all - E.cnt;                    // the partner's column count
            if (on > 0.f) {
                const float delta = o.x - E.mean;
                E.mean += delta * (on / E.cnt_all);
                E.m2 += o.y + delta * delta * (E.cnt * on / E.cnt_all);
            }
        }
        E.rstd = rsqrtf(E.m2 / E.cnt_all + kLnEps);
    }
}

// v = accumulator row slice + bias (bias from the shared-memory package, or gathered per row from
// the time table in forward mode)
template <int NPC>
__device__ __forceinline__ void load_vec(EpiCtx& E, uint32_t taddr, const float* bias) {
    static_assert(NPC * 8 <= kVecRegs, "vector does not fit the per-thread register slice");
    if constexpr (NPC % 2 == 0) {
#pragma unroll
        for (int i = 0; i < NPC / 2; ++i) tmem_ld16(taddr + i * 16, *reinterpret_cast<float(*)[16]>(&E.v[i * 16]));
    } else {
#pragma unroll
        for (int i = 0; i < NPC; ++i) tmem_ld8(taddr + i * 8, &E.v[i * 8]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < NPC * 2; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(bias + q * 4);
        E.v[q * 4 + 0] += b.x; E.v[q * 4 + 1] += b.y; E.v[q * 4 + 2] += b.z; E.v[q * 4 + 3] += b.w;
    }
}
template <int NPC>
__device__ __forceinline__ void load_skip_vec(EpiCtx& E, const float4* sk) {
#pragma unroll
    for (int q = 0; q < NPC * 2; ++q) {
        const float4 t = sk[q * kRows];
        E.v[q * 4 + 0] = t.x; E.v[q * 4 + 1] = t.y; E.v[q * 4 + 2] = t.z; E.v[q * 4 + 3] = t.w;
    }
}
template <int NPC>
__device__ __forceinline__ void store_skip_vec(const EpiCtx& E, float4* sk) {
#pragma unroll
    for (int q = 0; q < NPC * 2; ++q)
        sk[q * kRows] = make_float4(E.v[q * 4], E.v[q * 4 + 1], E.v[q * 4 + 2], E.v[q * 4 + 3]);
}

// npc = pieces (8 columns each) owned by one thread: np / kSplit.  Vectors are padded to 16 columns,
// so with split 1 npc is even (2..16), with split 2 it is 1..8.
#if DIFFSG_TC_SPLIT == 2
#define DIFFSG_TC_NPC_SWITCH(npc, CALL)        \
    switch (npc) {                             \
        case 1: CALL(1); break;                \
        case 2: CALL(2); break;                \
        case 3: CALL(3); break;                \
        case 4: CALL(4); break;                \
        case 5: CALL(5); break;                \
        case 6: CALL(6); break;                \
        case 7: CALL(7); break;                \
        default: CALL(8); break;               \
    }
// LayerNorm'd vectors are internal widths: powers of two (8 columns are padded to 16)
#define DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)   \
    switch (npc) {                             \
        case 1: CALL(1); break;                \
        case 2: CALL(2); break;                \
        case 4: CALL(4); break;                \
        default: CALL(8); break;               \
    }
#else
#define DIFFSG_TC_NPC_SWITCH(npc, CALL)        \
    switch (npc) {                             \
        case 2: CALL(2); break;                \
        case 4: CALL(4); break;                \
        case 6: CALL(6); break;                \
        case 8: CALL(8); break;                \
        case 10: CALL(10); break;              \
        case 12: CALL(12); break;              \
        case 14: CALL(14); break;              \
        default: CALL(16); break;              \
    }
#define DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)   \
    switch (npc) {                             \
        case 2: CALL(2); break;                \
        case 4: CALL(4); break;                \
        case 8: CALL(8); break;                \
        default: CALL(16); break;              \
    }
#endif
constexpr int kNpcPartial = kSplit == 2 ? 1 : 2;   // the only vector that may be partially valid: 16 padded columns

__device__ __forceinline__ void emit_cond(SmemLayout& S, EpiCtx& E, const TcDev& P, const float* scr) {
    const uint4* img = reinterpret_cast<const uint4*>(scr + P.cond_off);
    const int nkc = P.Cp / 8;                      // 16-byte K pieces per row; piece k belongs to half k & 1
    for (int c0 = 0; c0 < nkc; c0 += 8) {
        const int nk = min(8, nkc - c0);
        const uint32_t sq = E.aseq;
        const uint32_t sl = sq % kASlots;
        mbar_wait(&S.a_empty[sl], ((sq / kASlots) & 1) ^ 1);
        const uint32_t base = (uint32_t)(E.row >> 3) * (nk * 128) + (E.row & 7) * 16;
        for (int k = E.half; k < nk; k += kSplit) {
            *reinterpret_cast<uint4*>(S.a_hi[sl] + base + k * 128) = img[(size_t)(c0 + k) * kRows + E.row];
            *reinterpret_cast<uint4*>(S.a_lo[sl] + base + k * 128) = img[(size_t)(nkc + c0 + k) * kRows + E.row];
        }
        fence_proxy_async_smem();
        mbar_arrive_n(&S.a_full[sl], 1u);
        ++E.aseq;
    }
}

template <bool kSampler>
__device__ __forceinline__ void run_epilogue(SmemLayout& S, const TcDev& P, const RunArgs& R, EpiCtx& E,
                                             float* scr, int64_t grow, bool valid, int trow, bool use_cond,
                                             int pass, int step, uint32_t& acc_phase, uint32_t& pseq,
                                             double& st_s, double& st_q) {
    const int row = E.row;
    const uint32_t tmem_row = S.tmem_base + ((uint32_t)(row & ~31) << 16);
    for (int si = 0; si < P.n_stages; ++si) {
        const Stage sg = c_stages[si];
        if (sg.bits & 4) {
            TCT_BEGIN();
            mbar_wait(&S.acc_full, acc_phase);
            acc_phase ^= 1;
            tcgen05_fence_after();
            TCT_END(0);
        }
        const bool has_pkg = (sg.pkg_f4 | sg.tt_f4) != 0;
        const uint32_t psl = pseq % kPSlots;
        if (has_pkg) { TCT_BEGIN(); mbar_wait(&S.p_full[psl], (pseq / kPSlots) & 1); TCT_END(1); }
        const float* pk = S.pkg[psl];
        for (int ei = sg.epi_begin; ei < sg.epi_begin + sg.n_epi; ++ei) {
            const Epi op = c_epis[ei];
            const int np = op.np, npc = np / kSplit, dt = op.dt;
            const int cb = E.half * npc * 8;                              // first column owned by this thread
            const int nv = max(0, min(dt - cb, npc * 8));                 // valid (un-padded) columns owned
            const bool full = nv == npc * 8;
            const int region = op.misc & 1, flags = op.misc >> 1;
            switch (op.kind) {
                case TE_LOAD:
                case TE_LN_BLOCK: {
                    const uint32_t ta = tmem_row + region * 128 + cb;
                    const float* bias = pk + op.off0 * 4 + cb;
                    if (!kSampler && (flags & kFTime)) bias = P.tt + (size_t)trow * P.tt_stride + sg.tt_src4 * 4 + cb;
                    {
                        TCT_BEGIN();
#define CALL(W) load_vec<W>(E, ta, bias)
                        DIFFSG_TC_NPC_SWITCH(npc, CALL)
#undef CALL
                        TCT_END(2);
                    }
                    if (op.kind == TE_LOAD) break;
                    if (flags & kFPush) {
                        float4* sk = reinterpret_cast<float4*>(scr + P.skip_off[op.slot]) + (size_t)(cb / 4) * kRows + row;
#define CALL(W) store_skip_vec<W>(E, sk)
                        DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)
#undef CALL
                    }
                    const float* pg = pk + op.off1 * 4;
                    const float* pb = pk + op.off2 * 4;
                    if (full) {
                        {
                            TCT_BEGIN();
#define CALL(W) stats_vec<W, true>(S, E, nv, dt, kStatsReset | kStatsFinish)
                            DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)
#undef CALL
                            TCT_END(3);
                        }
                        TCT_BEGIN();
#define CALL(W) emit_vec<1, W, true>(S, E, np, nv, pg, pb)
                        DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)
#undef CALL
                        TCT_END(4);
                    } else {
                        stats_vec<kNpcPartial, false>(S, E, nv, dt, kStatsReset | kStatsFinish);
                        emit_vec<1, kNpcPartial, false>(S, E, np, nv, pg, pb);
                    }
                    if ((flags & kFCond) && use_cond) { TCT_BEGIN(); emit_cond(S, E, P, scr); TCT_END(7); }
                    break;
                }
                case TE_LOAD_SKIP: {
                    TCT_BEGIN();
                    const float4* sk = reinterpret_cast<const float4*>(scr + P.skip_off[op.slot]) + (size_t)(cb / 4) * kRows + row;
#define CALL(W) load_skip_vec<W>(E, sk)
                    DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)
#undef CALL
                    E.v[0] += 0.f * E.v[npc * 8 - 1];   // (timing builds only) make the loads complete inside the timed region
                    TCT_END(8);
                    break;
                }
                case TE_STORE_SKIP: {
                    float4* sk = reinterpret_cast<float4*>(scr + P.skip_off[op.slot]) + (size_t)(cb / 4) * kRows + row;
#define CALL(W) store_skip_vec<W>(E, sk)
                    DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)
#undef CALL
                    break;
                }
                case TE_LOAD_INPUT: {
                    const float* src = (kSampler ? R.y : R.x) + grow * P.M + cb;
#pragma unroll
                    for (int j = 0; j < kVecRegs; ++j)
                        if (j < npc * 8) E.v[j] = (valid && j < nv) ? src[j] : 0.f;
                    break;
                }
                case TE_STATS: {
                    TCT_BEGIN();
                    if (full) {
#define CALL(W) stats_vec<W, true>(S, E, nv, dt, flags)
                        DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)
#undef CALL
                    } else {
                        stats_vec<kNpcPartial, false>(S, E, nv, dt, flags);
                    }
                    TCT_END(3);
                    break;
                }
                case TE_EMIT_LN: {
                    TCT_BEGIN();
                    const float* pg = pk + op.off0 * 4;
                    const float* pb = pk + op.off1 * 4;
                    if (full) {
#define CALL(W) emit_vec<1, W, true>(S, E, np, nv, pg, pb)
                        DIFFSG_TC_NPC_SWITCH_POW2(npc, CALL)
#undef CALL
                    } else {
                        emit_vec<1, kNpcPartial, false>(S, E, np, nv, pg, pb);
                    }
                    TCT_END(4);
                    break;
                }
                case TE_EMIT_RAW: {
                    TCT_BEGIN();
                    if (full) {
#define CALL(W) emit_vec<0, W, true>(S, E, np, nv, nullptr, nullptr)
                        DIFFSG_TC_NPC_SWITCH(npc, CALL)
#undef CALL
                    } else {
#define CALL(W) emit_vec<0, W, false>(S, E, np, nv, nullptr, nullptr)
                        DIFFSG_TC_NPC_SWITCH(npc, CALL)
#undef CALL
                    }
                    TCT_END(11);
                    break;
                }
                case TE_EMIT_COND:
                    if (use_cond) emit_cond(S, E, P, scr);
                    break;
                case TE_STORE_OUT: {
                    TCT_BEGIN();
                    if (!kSampler) {
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < kVecRegs; ++j)
                                if (j < nv) R.eps[grow * P.M + cb + j] = E.v[j];
                        }
                        break;
                    }
                    float4* stash = reinterpret_cast<float4*>(scr + P.stash_off) + (size_t)(cb / 4) * kRows + row;
                    if (pass == 0) {           // unconditional pass: park eps_0
#pragma unroll
                        for (int q = 0; q < kVecRegs / 4; ++q)
                            if (q < npc * 2)
                                stash[q * kRows] = make_float4(E.v[q * 4], E.v[q * 4 + 1], E.v[q * 4 + 2], E.v[q * 4 + 3]);
                        break;
                    }
                    // conditional pass: guidance mix + posterior update (classifier_free_MSR.py:132-134)
                    const float w1 = 1.0f + R.omega, w0 = R.omega;
                    const float ce = R.c_eps[step], crs = R.c_rs[step], cn = R.c_noise[step];
                    const bool add_noise = step > 1;
                    const bool want_stats = step > R.T - 1 - R.norm_steps;
                    const int64_t plane = R.B * (int64_t)P.M;
                    const int64_t pidx = (int64_t)(R.T - 1 - step) * plane;
#pragma unroll
                    for (int q = 0; q < kVecRegs / 4; ++q)
                        if (q < npc * 2 && q * 4 < nv) {
                            const float4 e0 = stash[q * kRows];
                            const float e0a[4] = {e0.x, e0.y, e0.z, e0.w};
                            float z[4] = {0.f, 0.f, 0.f, 0.f};
                            if (valid && add_noise && R.noise == nullptr)
                                philox_normal4((uint64_t)grow + R.offset, (uint32_t)step, (uint32_t)(cb / 4 + q), R.seed, z);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int c = q * 4 + j;
                                if (valid && c < nv) {
                                    const int64_t idx = grow * P.M + cb + c;
                                    if (add_noise && R.noise != nullptr) z[j] = R.noise[pidx + idx];
                                    const float e = w1 * E.v[c] - w0 * e0a[j];
                                    float yn = (R.y[idx] - ce * e) * crs;
                                    if (add_noise) yn += cn * z[j];
                                    R.y[idx] = yn;
                                    if (R.rec_eps) R.rec_eps[pidx + idx] = e;
                                    if (R.rec_y && !want_stats) R.rec_y[pidx + idx] = yn;
                                    if (want_stats) { st_s += (double)yn; st_q += (double)yn * (double)yn; }
                                }
                            }
                        }
                    TCT_END(9);
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
      setmaxnreg_dec();
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
                                mbar_wait(&S.p_empty[sl], ((pseq / kPSlots) & 1) ^ 1);
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
                                mbar_wait(&S.w_empty[st], ph ^ 1);
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
                            const uint32_t d_tmem = S.tmem_base + (sg.bits & 1) * 128;
                            uint32_t acc = (sg.bits >> 1) & 1;
                            for (int ci = sg.chunk_begin; ci < sg.chunk_begin + sg.n_chunks; ++ci) {
                                const Chunk ch = c_chunks[ci];
                                if ((ch.flags & kChunkCond) && !use_cond) continue;
                                const uint32_t st = wseq % kWStages, wph = (wseq / kWStages) & 1;
                                const uint32_t sl = aseq % kASlots, aph = (aseq / kASlots) & 1;
                                mbar_wait(&S.a_full[sl], aph);
                                mbar_wait(&S.w_full[st], wph);
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
        // =========================== epilogue / operand producers (two threads per row)
        setmaxnreg_inc();
        EpiCtx E;
        E.et = threadIdx.x - kEpiWarp0 * 32;
        E.half = kSplit == 2 ? (E.et >> 7) : 0;
        E.row = E.et & 127;
#ifdef DIFFSG_TC_TIMING
        for (int i = 0; i < 12; ++i) E.tacc[i] = 0;
#endif
        E.aseq = 0; E.xpar = 0; E.mean = 0.f; E.rstd = 1.f; E.m2 = 0.f; E.cnt = 0.f; E.cnt_all = 0.f;
#pragma unroll
        for (int j = 0; j < kVecRegs; ++j) E.v[j] = 0.f;
        uint32_t acc_phase = 0, pseq = 0;
        double st_s = 0.0, st_q = 0.0;
        float* scr = P.scratch + (size_t)blockIdx.x * P.scratch_floats;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t grow = tile * kRows + E.row;
            const bool valid = grow < R.B;
            // cond image: swish(cond * mask) as fp16 (hi, lo) 16-byte K pieces; piece k is built (and
            // later copied into the operand ring) by the thread of half k & 1 -> thread-private scratch
            {
                uint4* img = reinterpret_cast<uint4*>(scr + P.cond_off);
                const int nkc = P.Cp / 8;
                const float mk = (!kSampler && R.mask && valid) ? R.mask[grow] : 1.0f;
                for (int kc = E.half; kc < nkc; kc += kSplit) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = kc * 8 + j;
                        x[j] = (valid && c < P.C) ? swish_exact(R.cond[grow * P.C + c] * mk) : 0.f;
                    }
                    uint4 hi, lo;
                    split_pack8(x, hi, lo);
                    img[(size_t)kc * kRows + E.row] = hi;
                    img[(size_t)(nkc + kc) * kRows + E.row] = lo;
                }
            }
            const int trow_fwd = (!kSampler && valid) ? R.t_idx[grow] : 0;
            for (int step = step_hi; step >= step_lo; --step)
                for (int pass = 0; pass < n_pass; ++pass) {
                    const bool use_cond = kSampler ? (pass == 1) : true;
                    TCT_BEGIN();
                    run_epilogue<kSampler>(S, P, R, E, scr, grow, valid, kSampler ? step : trow_fwd, use_cond, pass,
                                           step, acc_phase, pseq, st_s, st_q);
                    TCT_END(10);
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
