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
constexpr int kWStages = 3;
constexpr int kWStageBytes = 128 * kChunkK * 2;      // 16 KB: one fp16 W chunk (N <= 128)
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;                         // epilogue warps 4..7 (one aligned warpgroup)
constexpr int kTmemCols = 256;                       // two 128-column regions
constexpr int kMaxStages = 160, kMaxChunks = 512, kMaxEpi = 768;

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

__device__ __forceinline__ float swish_f(float x) { return x / (1.0f + __expf(-x)); }

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

// Emit v[0:dp) as K-chunks.  MODE 0: raw, 1: swish(LN(v) * gamma + beta)
template <int MODE>
__device__ __forceinline__ void emit_vec(SmemLayout& S, EpiCtx& E, const TcDev& P, int row, int dp16, int dt,
                                         int off_g, int off_b) {
    const float4* g4 = reinterpret_cast<const float4*>(P.params + (MODE ? off_g : 0));
    const float4* b4 = reinterpret_cast<const float4*>(P.params + (MODE ? off_b : 0));
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c * 4 < dp16) {
            const int ngr = min(4, dp16 - c * 4);
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
                                const float t = (E.v[j0 + j] - E.mean) * E.rstd * gam[j] + bet[j];
                                x[j] = (j0 + j < dt) ? swish_f(t) : 0.f;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) x[j] = (j0 + j < dt) ? E.v[j0 + j] : 0.f;
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
}

// Per-row LayerNorm statistics of v[0:dt), merged into the running (cnt, mean, m2) (Chan et al.)
__device__ __forceinline__ void stats_vec(EpiCtx& E, int dp16, int dt, int flags) {
    if (flags & kStatsReset) { E.cnt = 0.f; E.mean = 0.f; E.m2 = 0.f; }
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g)
        if (g < dp16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) s += (g * 16 + j < dt) ? E.v[g * 16 + j] : 0.f;
        }
    const float n = (float)dt, m = s / n;
    float q = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g)
        if (g < dp16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float d = (g * 16 + j < dt) ? E.v[g * 16 + j] - m : 0.f;
                q = fmaf(d, d, q);
            }
        }
    const float tot = E.cnt + n, delta = m - E.mean;
    E.mean += delta * (n / tot);
    E.m2 += q + delta * delta * (E.cnt * n / tot);
    E.cnt = tot;
    if (flags & kStatsFinish) E.rstd = 1.0f / sqrtf(E.m2 / E.cnt + kLnEps);
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
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < dp16) tmem_ld16(ta + g * 16, *reinterpret_cast<float(*)[16]>(&E.v[g * 16]));
                    tmem_ld_wait();
                    const float4* b4 = reinterpret_cast<const float4*>(P.params + op.off0);
                    const float4* t4 = reinterpret_cast<const float4*>(P.tt + (size_t)trow * P.tt_stride + (op.off1 >= 0 ? op.off1 : 0));
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < dp16) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float4 b = __ldg(b4 + g * 4 + q);
                                if (op.off1 >= 0) {
                                    const float4 t = __ldg(t4 + g * 4 + q);
                                    b.x += t.x; b.y += t.y; b.z += t.z; b.w += t.w;
                                }
                                E.v[g * 16 + q * 4 + 0] += b.x; E.v[g * 16 + q * 4 + 1] += b.y;
                                E.v[g * 16 + q * 4 + 2] += b.z; E.v[g * 16 + q * 4 + 3] += b.w;
                            }
                        }
                    break;
                }
                case TE_LOAD_SKIP: {
                    const float4* sk = reinterpret_cast<const float4*>(scr + P.skip_off[op.slot]) + row;
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < dp16) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 t = sk[(g * 4 + q) * kRows];
                                E.v[g * 16 + q * 4 + 0] = t.x; E.v[g * 16 + q * 4 + 1] = t.y;
                                E.v[g * 16 + q * 4 + 2] = t.z; E.v[g * 16 + q * 4 + 3] = t.w;
                            }
                        }
                    break;
                }
                case TE_STORE_SKIP: {
                    float4* sk = reinterpret_cast<float4*>(scr + P.skip_off[op.slot]) + row;
#pragma unroll
                    for (int g = 0; g < 8; ++g)
                        if (g < dp16) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                sk[(g * 4 + q) * kRows] = make_float4(E.v[g * 16 + q * 4], E.v[g * 16 + q * 4 + 1],
                                                                      E.v[g * 16 + q * 4 + 2], E.v[g * 16 + q * 4 + 3]);
                        }
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
                    stats_vec(E, dp16, dt, op.flags);
                    break;
                case TE_EMIT_LN:
                    emit_vec<1>(S, E, P, row, dp16, dt, op.off0, op.off1);
                    break;
                case TE_EMIT_RAW:
                    emit_vec<0>(S, E, P, row, dp16, dt, 0, 0);
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
__global__ void __launch_bounds__(kThreads, 1) tc_unet_kernel(TcDev P, RunArgs R) {
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
    } else if (warp >= kEpiWarp0) {
        // =========================== epilogue / operand producers (thread == row)
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
