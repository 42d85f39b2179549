// Tensor-core engine: one CTA carries a 128-row tile through the whole UNet1D stage program.
//
//   warp 0      bulk-TMA producer: streams fp16 weight K-chunk images into the W ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma (A, W from shared memory, fp32
//               accumulators in TMEM), commits completion to mbarriers
//   warps 4-7   epilogue / operand producers: thread == row == TMEM lane.  Load the accumulator
//               row into registers, add biases, LayerNorm + Swish in fp32, split into fp16
//               (hi, lo) and write the next GEMM's A operand as core-matrix K-chunks
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
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;                         // epilogue warps 4..7 (one aligned warpgroup)
constexpr int kTmemCols = 256;                       // two 128-column regions
constexpr int kMaxStages = 128, kMaxChunks = 256, kMaxEpi = 640;
constexpr int kCtasPerSm = 2;                        // fp16x2: two co-resident tiles per SM
constexpr int kRegsProducer = 32, kRegsEpilogue = 224;  // setmaxnreg split of the 128-per-thread launch budget

enum : int { TE_LOAD_TMEM = 1, TE_LOAD_SKIP, TE_LOAD_INPUT, TE_STORE_SKIP, TE_STORE_OUT, TE_STATS, TE_EMIT_LN,
             TE_EMIT_RAW, TE_EMIT_COND };
constexpr int kStatsReset = 1, kStatsFinish = 2, kChunkCond = 1;

struct __align__(16) Epi { uint8_t kind, region, dp16, flags; uint16_t dt, slot; int32_t off0, off1; };
struct __align__(8) Chunk { uint16_t kw, flags; uint32_t w_off16; };
struct __align__(16) Stage { uint16_t chunk_begin, n_chunks, epi_begin, n_epi; uint8_t n16, region, accumulate, has_gemm; uint32_t pad; };
static_assert(sizeof(Epi) == 16 && sizeof(Chunk) == 8 && sizeof(Stage) == 16, "program record layout");

struct TcDev {
    const Stage* stages; const Chunk* chunks; const Epi* epis;
    int n_stages, n_chunks, n_epi;
    const uint8_t* w_hi; const uint8_t* w_lo;     // fp16 weight images (lo: nterms == 3 only)
    const float* params; const float* tt;
    int tt_stride, nterms;
    int M, Mp, C, Cp;
    float* scratch;                                 // per CTA: skip stack + eps stash + cond image
    size_t scratch_floats;                          // per CTA
    int skip_off[kMaxSkip];                         // float offset of each skip slot inside the CTA scratch
    int stash_off, cond_off;                        // float offsets (cond image: hi then lo, fp16)
};

struct SmemLayout {
    uint8_t a_hi[kASlots][kSlotBytes];
    uint8_t a_lo[kASlots][kSlotBytes];
    uint64_t a_full[kASlots], a_empty[kASlots], w_full[kWStages], w_empty[kWStages], acc_full;
    uint32_t tmem_base, pad_;
    Stage stages[kMaxStages];
    Chunk chunks[kMaxChunks];
    Epi epis[kMaxEpi];
    // followed by the W ring: kWStages * (nterms == 3 ? 2 : 1) * kWStageBytes (dynamic)
};

// What one launch does: kSampler -> steps step_hi..step_lo, two passes each; else one forward.
struct RunArgs {
    // forward
    const float* x; const int32_t* t_idx; const float* cond; const float* mask; float* eps;
    // sampler
    float* y; const float* noise; float* rec_y; float* rec_eps; double* stats;
    int64_t B;
    int T, step_hi, step_lo, norm_steps;
    float omega;
    uint64_t seed, offset;
    float c_eps[64], c_rs[64], c_noise[64];
};

// ------------------------------------------------------------------------------------------
struct EpiCtx {
    float v[128];
    float mean, rstd, m2, cnt;
    uint32_t aseq;          // A-ring sequence number (chunks produced so far)
};

// x * sigmoid(x) with MUFU ex2 + rcp (relative error ~3e-7; exact limits at +-inf)
__device__ __forceinline__ float swish_f(float x) {
    const float e = exp2f(-1.4426950408889634f * x);
    return __fdividef(x, 1.0f + e);
}
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsProducer)); }
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue)); }

// producer side of one A chunk: wait for the slot, return its base offset for this row
__device__ __forceinline__ void a_slot_acquire(SmemLayout& S, uint32_t aseq) {
    const uint32_t sl = aseq % kASlots, ph = (aseq / kASlots) & 1;
    mbar_wait(&S.a_empty[sl], ph ^ 1);
}
__device__ __forceinline__ void a_slot_publish(SmemLayout& S, uint32_t aseq) {
    fence_proxy_async_smem();
    tcgen05_fence_before();
    mbar_arrive(&S.a_full[aseq % kASlots]);
}

// Emit v[0:16*DP16) as K-chunks of <= 64.  MODE 0: raw, 1: swish(LN(v) * gamma + beta).
// FULL: every column is real (dt == 16*DP16) -> no per-element predicates.
template <int MODE, int DP16, bool FULL>
__device__ __forceinline__ void emit_vec(SmemLayout& S, EpiCtx& E, const TcDev& P, int row, int dt,
                                         int off_g, int off_b) {
    const float4* g4 = reinterpret_cast<const float4*>(P.params + (MODE ? off_g : 0));
    const float4* b4 = reinterpret_cast<const float4*>(P.params + (MODE ? off_b : 0));
    const float a_scale = E.rstd, a_shift = -E.mean * E.rstd;
#pragma unroll
    for (int c = 0; c * 4 < DP16; ++c) {
        constexpr int kDummy = 0; (void)kDummy;
        const int ngr = (DP16 - c * 4) < 4 ? (DP16 - c * 4) : 4;
        const uint32_t sbo = (uint32_t)ngr * 256;                 // kw * 16
        const uint32_t sl = E.aseq % kASlots;
        a_slot_acquire(S, E.aseq);
        uint8_t* hi_base = S.a_hi[sl] + (row >> 3) * sbo + (row & 7) * 16;
        uint8_t* lo_base = S.a_lo[sl] + (row >> 3) * sbo + (row & 7) * 16;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
            if (gg < ngr) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j0 = (c * 4 + gg) * 16 + h * 8;
                    float x[8];
                    if (MODE) {
                        const float4 ga = __ldg(g4 + j0 / 4), gb = __ldg(g4 + j0 / 4 + 1);
                        const float4 ba = __ldg(b4 + j0 / 4), bb = __ldg(b4 + j0 / 4 + 1);
                        const float gam[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
                        const float bet[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float t = fmaf(fmaf(E.v[j0 + j], a_scale, a_shift), gam[j], bet[j]);
                            x[j] = (FULL || j0 + j < dt) ? swish_f(t) : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) x[j] = (FULL || j0 + j < dt) ? E.v[j0 + j] : 0.f;
                    }
                    uint4 hi, lo;
                    split_pack8(x, hi, lo);
                    *reinterpret_cast<uint4*>(hi_base + (gg * 2 + h) * 128) = hi;
                    *reinterpret_cast<uint4*>(lo_base + (gg * 2 + h) * 128) = lo;
                }
            }
        }
        a_slot_publish(S, E.aseq);
        ++E.aseq;
    }
}

// Per-row LayerNorm statistics of v[0:dt), merged into the running (cnt, mean, m2) (Chan et al.)
template <int DP16, bool FULL>
__device__ __forceinline__ void stats_vec(EpiCtx& E, int dt, int flags) {
    if (flags & kStatsReset) { E.cnt = 0.f; E.mean = 0.f; E.m2 = 0.f; }
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < DP16 * 16; ++j) s4[j & 3] += (FULL || j < dt) ? E.v[j] : 0.f;
    const float n = FULL ? (float)(DP16 * 16) : (float)dt;
    const float m = ((s4[0] + s4[1]) + (s4[2] + s4[3])) / n;
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < DP16 * 16; ++j) {
        const float d = (FULL || j < dt) ? E.v[j] - m : 0.f;
        q4[j & 3] = fmaf(d, d, q4[j & 3]);
    }
    const float q = (q4[0] + q4[1]) + (q4[2] + q4[3]);
    const float tot = E.cnt + n, delta = m - E.mean;
    E.mean += delta * (n / tot);
    E.m2 += q + delta * delta * (E.cnt * n / tot);
    E.cnt = tot;
    if (flags & kStatsFinish) E.rstd = rsqrtf(E.m2 / E.cnt + kLnEps);
}

template <int DP16>
__device__ __forceinline__ void load_tmem_vec(EpiCtx& E, const TcDev& P, uint32_t ta, int off0, int off1, int trow) {
#pragma unroll
    for (int g = 0; g < DP16; ++g) tmem_ld16(ta + g * 16, *reinterpret_cast<float(*)[16]>(&E.v[g * 16]));
    const float4* b4 = reinterpret_cast<const float4*>(P.params + off0);
    const float4* t4 = reinterpret_cast<const float4*>(P.tt + (size_t)trow * P.tt_stride + (off1 >= 0 ? off1 : 0));
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < DP16; ++g) {           // one 16-column group (4 x float4) of bias in flight at a time
        float4 b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) b[q] = __ldg(b4 + g * 4 + q);
        if (off1 >= 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 t = __ldg(t4 + g * 4 + q);
                b[q].x += t.x; b[q].y += t.y; b[q].z += t.z; b[q].w += t.w;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            E.v[g * 16 + q * 4 + 0] += b[q].x; E.v[g * 16 + q * 4 + 1] += b[q].y;
            E.v[g * 16 + q * 4 + 2] += b[q].z; E.v[g * 16 + q * 4 + 3] += b[q].w;
        }
    }
}
template <int DP16>
__device__ __forceinline__ void load_skip_vec(EpiCtx& E, const float4* sk) {
#pragma unroll
    for (int q = 0; q < DP16 * 4; ++q) {
        const float4 t = sk[q * kRows];
        E.v[q * 4 + 0] = t.x; E.v[q * 4 + 1] = t.y; E.v[q * 4 + 2] = t.z; E.v[q * 4 + 3] = t.w;
    }
}
template <int DP16>
__device__ __forceinline__ void store_skip_vec(const EpiCtx& E, float4* sk) {
#pragma unroll
    for (int q = 0; q < DP16 * 4; ++q)
        sk[q * kRows] = make_float4(E.v[q * 4], E.v[q * 4 + 1], E.v[q * 4 + 2], E.v[q * 4 + 3]);
}

// widths the tensor-core engine accepts: 16, 32, 64, 80, 96, 112, 128 columns (dp16 = 1,2,4,5,6,7,8;
// 3 is folded into 4 by the packer never emitting it -> treated as invalid at attach time)
#define DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)     \
    switch (dp16) {                            \
        case 1: CALL(1); break;                \
        case 2: CALL(2); break;                \
        case 3: CALL(3); break;                \
        case 4: CALL(4); break;                \
        case 5: CALL(5); break;                \
        case 6: CALL(6); break;                \
        case 7: CALL(7); break;                \
        default: CALL(8); break;               \
    }

template <bool kSampler>
__device__ __forceinline__ void run_epilogue(SmemLayout& S, const TcDev& P, const RunArgs& R, EpiCtx& E,
                                             float* scr, int row, int64_t grow, bool valid, int trow,
                                             bool use_cond, int pass, int step, uint32_t& acc_phase,
                                             double& st_s, double& st_q) {
    const uint32_t lane_base = (uint32_t)(row & ~31);
    const uint32_t tmem_row = S.tmem_base + (lane_base << 16);
    for (int si = 0; si < P.n_stages; ++si) {
        const Stage sg = S.stages[si];
        if (sg.has_gemm) {
            mbar_wait(&S.acc_full, acc_phase);
            acc_phase ^= 1;
            tcgen05_fence_after();
        }
        for (int ei = sg.epi_begin; ei < sg.epi_begin + sg.n_epi; ++ei) {
            const Epi op = S.epis[ei];
            const int dp16 = op.dp16, dt = op.dt;
            switch (op.kind) {
                case TE_LOAD_TMEM: {
                    const uint32_t ta = tmem_row + op.region * 128;
#define CALL(W) load_tmem_vec<W>(E, P, ta, op.off0, op.off1, trow)
                    DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)
#undef CALL
                    break;
                }
                case TE_LOAD_SKIP: {
                    const float4* sk = reinterpret_cast<const float4*>(scr + P.skip_off[op.slot]) + row;
#define CALL(W) load_skip_vec<W>(E, sk)
                    DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)
#undef CALL
                    break;
                }
                case TE_STORE_SKIP: {
                    float4* sk = reinterpret_cast<float4*>(scr + P.skip_off[op.slot]) + row;
#define CALL(W) store_skip_vec<W>(E, sk)
                    DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)
#undef CALL
                    break;
                }
                case TE_LOAD_INPUT: {
                    const float* src = (kSampler ? R.y : R.x) + grow * P.M;
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < dp16) {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                E.v[g * 16 + j] = (valid && g * 16 + j < dt) ? src[g * 16 + j] : 0.f;
                        }
                    break;
                }
                case TE_STATS:
                    if (dt == dp16 * 16) {
#define CALL(W) stats_vec<W, true>(E, dt, op.flags)
                        DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)
#undef CALL
                    } else {
                        stats_vec<1, false>(E, dt, op.flags);      // only 16-wide vectors may be partial
                    }
                    break;
                case TE_EMIT_LN:
                    if (dt == dp16 * 16) {
#define CALL(W) emit_vec<1, W, true>(S, E, P, row, dt, op.off0, op.off1)
                        DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)
#undef CALL
                    } else {
                        emit_vec<1, 1, false>(S, E, P, row, dt, op.off0, op.off1);
                    }
                    break;
                case TE_EMIT_RAW:
                    if (dt == dp16 * 16) {
#define CALL(W) emit_vec<0, W, true>(S, E, P, row, dt, 0, 0)
                        DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)
#undef CALL
                    } else if (dp16 == 1) {
                        emit_vec<0, 1, false>(S, E, P, row, dt, 0, 0);
                    } else {
#define CALL(W) emit_vec<0, W, false>(S, E, P, row, dt, 0, 0)
                        DIFFSG_TC_WIDTH_SWITCH(dp16, CALL)
#undef CALL
                    }
                    break;
                case TE_EMIT_COND: {
                    if (!use_cond) break;
                    const uint4* img = reinterpret_cast<const uint4*>(scr + P.cond_off);
                    const int nkc = P.Cp / 8;                      // 16-byte K pieces per row
                    for (int c0 = 0; c0 < nkc; c0 += 8) {
                        const int nk = min(8, nkc - c0);
                        const uint32_t sbo = (uint32_t)nk * 128;
                        const uint32_t sl = E.aseq % kASlots;
                        a_slot_acquire(S, E.aseq);
                        uint8_t* hi_base = S.a_hi[sl] + (row >> 3) * sbo + (row & 7) * 16;
                        uint8_t* lo_base = S.a_lo[sl] + (row >> 3) * sbo + (row & 7) * 16;
                        for (int k = 0; k < nk; ++k) {
                            *reinterpret_cast<uint4*>(hi_base + k * 128) = img[(size_t)(c0 + k) * kRows + row];
                            *reinterpret_cast<uint4*>(lo_base + k * 128) = img[(size_t)(nkc + c0 + k) * kRows + row];
                        }
                        a_slot_publish(S, E.aseq);
                        ++E.aseq;
                    }
                    break;
                }
                case TE_STORE_OUT: {
                    if (!kSampler) {
                        if (valid) {
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                if (g < dp16) {
#pragma unroll
                                    for (int j = 0; j < 16; ++j)
                                        if (g * 16 + j < dt) R.eps[grow * P.M + g * 16 + j] = E.v[g * 16 + j];
                                }
                        }
                        break;
                    }
                    float4* stash = reinterpret_cast<float4*>(scr + P.stash_off) + row;
                    if (pass == 0) {           // unconditional pass: park eps_0
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            if (g < dp16) {
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    stash[(g * 4 + q) * kRows] = make_float4(E.v[g * 16 + q * 4], E.v[g * 16 + q * 4 + 1],
                                                                             E.v[g * 16 + q * 4 + 2], E.v[g * 16 + q * 4 + 3]);
                            }
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
                    for (int g = 0; g < 8; ++g)
                        if (g < dp16) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 e0 = stash[(g * 4 + q) * kRows];
                                const float e0a[4] = {e0.x, e0.y, e0.z, e0.w};
                                float z[4] = {0.f, 0.f, 0.f, 0.f};
                                if (valid && add_noise && R.noise == nullptr && (g * 16 + q * 4) < dt)
                                    philox_normal4((uint64_t)grow + R.offset, (uint32_t)step, (uint32_t)(g * 4 + q), R.seed, z);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int c = g * 16 + q * 4 + j;
                                    if (valid && c < dt) {
                                        const int64_t idx = grow * P.M + c;
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
                        }
                    break;
                }
                default:
                    break;
            }
        }
    }
}

template <bool kSampler>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) tc_unet_kernel(TcDev P, RunArgs R) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    SmemLayout& S = *reinterpret_cast<SmemLayout*>(sm);
    const int w_terms = P.nterms == 3 ? 2 : 1;
    uint8_t* w_ring = sm + ((sizeof(SmemLayout) + 1023) & ~size_t(1023));   // [kWStages][w_terms][kWStageBytes]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup
    for (int i = threadIdx.x; i < P.n_stages; i += kThreads) S.stages[i] = P.stages[i];
    for (int i = threadIdx.x; i < P.n_chunks; i += kThreads) S.chunks[i] = P.chunks[i];
    for (int i = threadIdx.x; i < P.n_epi; i += kThreads) S.epis[i] = P.epis[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < kASlots; ++i) { mbar_init(&S.a_full[i], 128); mbar_init(&S.a_empty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
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
        // =========================== weight producer
        if (lane == 0) {
            uint32_t wseq = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int step = step_hi; step >= step_lo; --step)
                    for (int pass = 0; pass < n_pass; ++pass) {
                        const bool use_cond = kSampler ? (pass == 1) : true;
                        for (int si = 0; si < P.n_stages; ++si) {
                            const Stage sg = S.stages[si];
                            if (!sg.has_gemm) continue;
                            for (int ci = sg.chunk_begin; ci < sg.chunk_begin + sg.n_chunks; ++ci) {
                                const Chunk ch = S.chunks[ci];
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
                            const Stage sg = S.stages[si];
                            if (!sg.has_gemm) continue;
                            const uint32_t idesc = make_idesc_f16(128, (uint32_t)sg.n16 * 16u);
                            const uint32_t d_tmem = S.tmem_base + sg.region * 128;
                            uint32_t acc = sg.accumulate;
                            for (int ci = sg.chunk_begin; ci < sg.chunk_begin + sg.n_chunks; ++ci) {
                                const Chunk ch = S.chunks[ci];
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
        // =========================== epilogue / operand producers (thread == row)
        setmaxnreg_inc();
        const int row = 32 * (warp & 3) + lane;
        EpiCtx E;
        E.aseq = 0; E.mean = 0.f; E.rstd = 1.f; E.m2 = 0.f; E.cnt = 0.f;
#pragma unroll
        for (int j = 0; j < 128; ++j) E.v[j] = 0.f;
        uint32_t acc_phase = 0;
        double st_s = 0.0, st_q = 0.0;
        float* scr = P.scratch + (size_t)blockIdx.x * P.scratch_floats;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t grow = tile * kRows + row;
            const bool valid = grow < R.B;
            // cond image: swish(cond * mask) as fp16 (hi, lo), K pieces of 8, thread-private scratch
            {
                uint4* img = reinterpret_cast<uint4*>(scr + P.cond_off);
                const int nkc = P.Cp / 8;
                const float mk = (!kSampler && R.mask && valid) ? R.mask[grow] : 1.0f;
                for (int kc = 0; kc < nkc; ++kc) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = kc * 8 + j;
                        x[j] = (valid && c < P.C) ? swish_exact(R.cond[grow * P.C + c] * mk) : 0.f;
                    }
                    uint4 hi, lo;
                    split_pack8(x, hi, lo);
                    img[(size_t)kc * kRows + row] = hi;
                    img[(size_t)(nkc + kc) * kRows + row] = lo;
                }
            }
            const int trow_fwd = (!kSampler && valid) ? R.t_idx[grow] : 0;
            for (int step = step_hi; step >= step_lo; --step)
                for (int pass = 0; pass < n_pass; ++pass) {
                    const bool use_cond = kSampler ? (pass == 1) : true;
                    run_epilogue<kSampler>(S, P, R, E, scr, row, grow, valid, kSampler ? step : trow_fwd, use_cond,
                                           pass, step, acc_phase, st_s, st_q);
                }
        }
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
