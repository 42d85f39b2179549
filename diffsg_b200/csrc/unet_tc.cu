// Tensor-core (tcgen05 / TMEM / bulk-TMA) engine of the UNet1D op program, sm_100a.
//
// Building block validated in isolation by diffsg_debug_tc_gemm (tests/test_gpu_tc.py):
//   C[128, N] = A[128, K] . W[N, K]^T,  A split on the fly into fp16 (hi, lo) by the
//   "epilogue" threads (thread == row == TMEM lane), W pre-packed on the host as fp16 core-matrix
//   images and streamed with 1-D bulk TMA, fp32 accumulation in TMEM.
#include <cstdio>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"
#include "unet_tc.cuh"
#include "plan.h"

namespace diffsg {
namespace tc {

constexpr int kMaxN = 128;
constexpr int kStageBytes = kWStageBytes;

struct GemmTestArgs {
    const float* A;        // [128][K] fp32
    const __half* W_hi;    // chunk images, see pack_w_image() in tests / packer
    const __half* W_lo;    // or null
    float* C;              // [128][N]
    int K, N;
    int nterms;            // 1: hi.W  2: + lo.W  3: + hi.W_lo
    uint32_t layout;       // UMMA layout type (0 = interleave)
    uint32_t lbo;          // K-direction core-matrix stride (bytes)
    int swap_lbo_sbo;      // experiment switch
};

struct __align__(128) GemmTestSmem {
    uint8_t a_hi[kASlots][kSlotBytes];
    uint8_t a_lo[kASlots][kSlotBytes];
    uint8_t w_hi[kWStages][kStageBytes];
    uint8_t w_lo[kWStages][kStageBytes];
    uint64_t a_full[kASlots], a_empty[kASlots], w_full[kWStages], w_empty[kWStages], acc_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 1) tc_gemm_test_kernel(GemmTestArgs g) {
    extern __shared__ uint8_t smem_raw[];
    GemmTestSmem& S = *reinterpret_cast<GemmTestSmem*>(
        smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (g.K + kChunkK - 1) / kChunkK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kASlots; ++i) { mbar_init(&S.a_full[i], 128); mbar_init(&S.a_empty[i], 1); }
        for (int i = 0; i < kWStages; ++i) { mbar_init(&S.w_full[i], 1); mbar_init(&S.w_empty[i], 1); }
        mbar_init(&S.acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(&S.tmem_base, 128); tmem_relinquish(); }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = S.tmem_base;

    if (warp == 0) {
        // ---------------- weight producer (bulk TMA)
        if (lane == 0) {
            for (int c = 0; c < n_chunks; ++c) {
                const int st = c % kWStages, ph = (c / kWStages) & 1;
                const int kw = min(kChunkK, g.K - c * kChunkK);
                const uint32_t bytes = (uint32_t)g.N * kw * 2;
                mbar_wait(&S.w_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&S.w_full[st], g.nterms == 3 ? 2 * bytes : bytes);
                const size_t off = (size_t)g.N * c * kChunkK;   // halves
                tma_load_1d(S.w_hi[st], g.W_hi + off, bytes, &S.w_full[st]);
                if (g.nterms == 3) tma_load_1d(S.w_lo[st], g.W_lo + off, bytes, &S.w_full[st]);
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(128, g.N);
            for (int c = 0; c < n_chunks; ++c) {
                const int st = c % kWStages, wph = (c / kWStages) & 1;
                const int sl = c % kASlots, aph = (c / kASlots) & 1;
                const int kw = min(kChunkK, g.K - c * kChunkK);
                const uint32_t sbo = (uint32_t)kw * 16;        // (kw/8) core matrices of 128 B per 8-row group
                mbar_wait(&S.a_full[sl], aph);
                mbar_wait(&S.w_full[st], wph);
                tcgen05_fence_after();
                const uint32_t l = g.swap_lbo_sbo ? sbo : g.lbo, s = g.swap_lbo_sbo ? g.lbo : sbo;
                for (int ks = 0; ks < kw / 16; ++ks) {
                    const uint32_t koff = ks * 2 * g.lbo;
                    const uint64_t a_hi = make_smem_desc(smem_u32(S.a_hi[sl]) + koff, l, s, g.layout);
                    const uint64_t a_lo = make_smem_desc(smem_u32(S.a_lo[sl]) + koff, l, s, g.layout);
                    const uint64_t w_hi = make_smem_desc(smem_u32(S.w_hi[st]) + koff, l, s, g.layout);
                    const uint64_t w_lo = make_smem_desc(smem_u32(S.w_lo[st]) + koff, l, s, g.layout);
                    umma_f16(tmem, a_hi, w_hi, idesc, (c | ks) != 0);
                    if (g.nterms >= 2) umma_f16(tmem, a_lo, w_hi, idesc, 1);
                    if (g.nterms >= 3) umma_f16(tmem, a_hi, w_lo, idesc, 1);
                }
                umma_commit(&S.a_empty[sl]);
                umma_commit(&S.w_empty[st]);
            }
            umma_commit(&S.acc_full);
        }
    } else {
        // ---------------- operand producer + epilogue: thread == row
        const int row = 32 * (warp & 3) + lane;
        for (int c = 0; c < n_chunks; ++c) {
            const int sl = c % kASlots, aph = (c / kASlots) & 1;
            const int kw = min(kChunkK, g.K - c * kChunkK);
            const uint32_t sbo = (uint32_t)kw * 16;
            mbar_wait(&S.a_empty[sl], aph ^ 1);
            const float* src = g.A + (size_t)row * g.K + c * kChunkK;
            const uint32_t base = (row >> 3) * sbo + (row & 7) * 16;
            for (int kc = 0; kc < kw / 8; ++kc) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = src[kc * 8 + j];
                uint4 hi, lo;
                split_pack8(x, hi, lo);
                *reinterpret_cast<uint4*>(S.a_hi[sl] + base + kc * g.lbo) = hi;
                *reinterpret_cast<uint4*>(S.a_lo[sl] + base + kc * g.lbo) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(&S.a_full[sl]);
        }
        mbar_wait(&S.acc_full, 0);
        tcgen05_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        for (int n0 = 0; n0 < g.N; n0 += 16) {
            float v[16];
            tmem_ld16(taddr + n0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) g.C[(size_t)row * g.N + n0 + j] = v[j];
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

}  // namespace tc
}  // namespace diffsg

// ---------------------------------------------------------------------------- host side
namespace diffsg {
namespace tc {

static std::mutex g_prog_mutex;
static uint64_t g_resident[64] = {0};
// Recorded after every tensor-core launch: the kernels read the stage program from __constant__ memory, so
// before ANOTHER plan's program is uploaded (possibly on a different stream) that stream must wait for the
// last kernel of the program it replaces.
static cudaEvent_t g_last_launch[64] = {nullptr};

struct TcHost {
    TcDev dev{};
    std::vector<Stage> stages;
    std::vector<Chunk> chunks;
    std::vector<Epi> epis;
    uint64_t serial = 0;
    float* d_scratch = nullptr;
    int* d_status = nullptr;
    int grid_max = 0;
    int occupancy = 0;
    size_t smem_bytes = 0;
    int tt_rows = 0, img_rows = 0;
    bool have_weights = false;
};

void tc_destroy(diffsg_plan* p) {
    if (!p || !p->tc) return;
    TcHost* h = p->tc;
    {
        std::lock_guard<std::mutex> lock(g_prog_mutex);
        if (g_resident[p->cfg.device & 63] == h->serial) g_resident[p->cfg.device & 63] = 0;
    }
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->d_status) cudaFree(h->d_status);
    delete h;
    p->tc = nullptr;
}

int tc_attach(diffsg_plan* p, const diffsg_tc_program* g) {
    if (!p || !g || !g->stages || !g->chunks || !g->epis) { set_error("attach_tc: null argument"); return DIFFSG_E_INVALID; }
    if (g->n_stages <= 0 || g->n_stages > kMaxStages || g->n_chunks <= 0 || g->n_chunks > kMaxChunks || g->n_epi <= 0 ||
        g->n_epi > kMaxEpi || g->n_skip < 0 || g->n_skip > kMaxSkip) {
        set_error("attach_tc: program too large (%d stages, %d chunks, %d epilogue ops, %d skips)", g->n_stages, g->n_chunks, g->n_epi, g->n_skip);
        return DIFFSG_E_UNSUPPORTED;
    }
    if (g->nterms < 1 || g->nterms > 3) { set_error("attach_tc: nterms must be 1..3"); return DIFFSG_E_INVALID; }
    const int M = p->cfg.input_dim, C = p->cfg.cond_dim;
    if (M > kRegionCols || C > 128) { set_error("attach_tc: input_dim > %d or cond_dim > 128", kRegionCols); return DIFFSG_E_UNSUPPORTED; }
    // validate records
    const Stage* st = (const Stage*)g->stages;
    const Chunk* ch = (const Chunk*)g->chunks;
    const Epi* ep = (const Epi*)g->epis;
    for (int i = 0; i < g->n_stages; ++i) {
        if (st[i].chunk_begin + st[i].n_chunks > g->n_chunks || st[i].epi_begin + st[i].n_epi > g->n_epi ||
            st[i].n16 < 1 || st[i].n16 * 16 > kRegionCols || st[i].pkg_f4 * 4 > kPkgFloats) {
            set_error("attach_tc: stage %d malformed", i); return DIFFSG_E_INVALID;
        }
    }
    for (int i = 0; i < g->n_chunks; ++i)
        if (ch[i].kw == 0 || ch[i].kw > kChunkK || ch[i].kw % 16) { set_error("attach_tc: chunk %d kw=%d", i, ch[i].kw); return DIFFSG_E_INVALID; }
    for (int i = 0; i < g->n_stages; ++i) {      // a GEMM group needs at least one chunk that is issued in every pass
        if (!(st[i].bits & 4)) continue;
        bool any = false;
        for (int c = st[i].chunk_begin; c < st[i].chunk_begin + st[i].n_chunks; ++c) any = any || !(ch[c].flags & kChunkCond);
        if (!any) { set_error("attach_tc: stage %d has no unconditional chunk", i); return DIFFSG_E_INVALID; }
    }
    for (int i = 0; i < g->n_epi; ++i) {
        const Epi& e = ep[i];
        const bool skip = e.kind == OP_CATLN || e.kind == OP_RAW_S || ((e.misc >> 1) & kFPush);
        const bool ln = e.kind == OP_LN || e.kind == OP_CATLN;
        if (e.kind < OP_LN || e.kind > OP_OUT || e.np * 8 > kRegionCols || (e.np & 1) || e.np == 0 || e.dt > e.np * 8 || e.dt == 0 ||
            (skip && e.slot >= g->n_skip) || (ln && (e.off1 * 4 + e.np * 8 * (e.kind == OP_CATLN ? 4 : 2)) > kPkgFloats)) {
            set_error("attach_tc: epilogue op %d malformed", i); return DIFFSG_E_INVALID;
        }
    }
    tc_destroy(p);
    TcHost* h = new (std::nothrow) TcHost();
    if (!h) { set_error("out of host memory"); return DIFFSG_E_INVALID; }
    p->tc = h;
    DIFFSG_CUDA_OK(cudaSetDevice(p->cfg.device));
    TcDev& D = h->dev;
    memset(&D, 0, sizeof(D));
    D.n_stages = g->n_stages; D.n_chunks = g->n_chunks; D.n_epi = g->n_epi;
    D.nterms = g->nterms; D.tt_stride = g->tt_stride;
    D.M = M; D.Mp = ((M + 15) / 16) * 16; D.C = C; D.Cp = ((C + 15) / 16) * 16;
    size_t off = 0;
    for (int s = 0; s < g->n_skip; ++s) {
        if (g->skip_widths[s] <= 0 || g->skip_widths[s] > 128 || g->skip_widths[s] % 16) { set_error("attach_tc: skip width"); return DIFFSG_E_INVALID; }
        D.skip_off[s] = (int)off;
        off += (size_t)g->skip_widths[s] * kRows;
    }
    D.stash_off = (int)off; off += (size_t)D.Mp * kRows;
    D.cond_off = (int)off;  off += (size_t)D.Cp * kRows;      // 2 images x Cp x 128 fp16 = Cp*128 floats
    D.stats_off = (int)off; off += (size_t)(g->n_skip > 0 ? g->n_skip : 1) * kRows * 4;   // (s1, s2, shift, -) per pushed row
    D.scratch_floats = (off + 31) & ~size_t(31);
    const int w_terms = g->nterms == 3 ? 2 : 1;
    h->smem_bytes = 128 + ((sizeof(SmemLayout) + 127) & ~size_t(127)) + (size_t)kWStages * (w_terms * kWStageBytes + kBiasBytes);
    if (getenv("DIFFSG_TC_ONE_CTA")) h->smem_bytes = 180 * 1024;   // experiment: one tile per SM (no co-resident CTA)
    // fp16x2 fits two CTAs (tiles) per SM: 2 x (<= 113 KB smem, 256 TMEM columns, 30 K registers)
    h->grid_max = p->sm_count;
    if ((int)h->smem_bytes > p->max_smem) { set_error("attach_tc: needs %zu B shared memory", h->smem_bytes); return DIFFSG_E_UNSUPPORTED; }
    static std::atomic<uint64_t> next_serial{1};
    h->serial = next_serial.fetch_add(1);
    h->stages.assign(st, st + g->n_stages);
    h->chunks.assign(ch, ch + g->n_chunks);
    h->epis.assign(ep, ep + g->n_epi);
    if (cudaMalloc(&h->d_scratch, sizeof(float) * D.scratch_floats * p->sm_count * kCtasPerSm) != cudaSuccess) {
        set_error("attach_tc: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
        tc_destroy(p);
        return DIFFSG_E_CUDA;
    }
    DIFFSG_CUDA_OK(cudaMemset(h->d_scratch, 0, sizeof(float) * D.scratch_floats * p->sm_count * kCtasPerSm));
    D.scratch = h->d_scratch;
    DIFFSG_CUDA_OK(cudaMalloc(&h->d_status, sizeof(int)));
    DIFFSG_CUDA_OK(cudaMemset(h->d_status, 0, sizeof(int)));
    D.status = h->d_status;
    for (const void* fn : {(const void*)tc_unet_kernel<true, true>, (const void*)tc_unet_kernel<true, false>,
                           (const void*)tc_unet_kernel<false, true>, (const void*)tc_unet_kernel<false, false>}) {
        DIFFSG_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
        DIFFSG_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    int occ = 0;
    DIFFSG_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tc_unet_kernel<true, true>, kThreads, h->smem_bytes));
    h->occupancy = occ;
    if (getenv("DIFFSG_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, tc_unet_kernel<true, true>);
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, p->cfg.device);
        fprintf(stderr, "[diffsg] tc kernel: regs %d, static smem %zu, dyn smem %zu (max %d), maxThreads %d, occ %d | SM: regs %d, smem %zu, "
                "reserved/block %zu, max blocks %d, max threads %d\n", fa.numRegs, fa.sharedSizeBytes, h->smem_bytes,
                fa.maxDynamicSharedSizeBytes, fa.maxThreadsPerBlock, occ, prop.regsPerMultiprocessor, prop.sharedMemPerMultiprocessor,
                prop.reservedSharedMemPerBlock, prop.maxBlocksPerMultiProcessor, prop.maxThreadsPerMultiProcessor);
        for (int thr = 128; thr <= 512; thr += 64) {
            int o = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, tc_unet_kernel<true, true>, thr, h->smem_bytes);
            fprintf(stderr, "[diffsg]   occupancy at %d threads: %d\n", thr, o);
        }
        for (size_t sm = 32768; sm <= 131072; sm += 16384) {
            int o = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, tc_unet_kernel<true, true>, kThreads, sm);
            fprintf(stderr, "[diffsg]   occupancy at %zu B smem: %d\n", sm, o);
        }
    }
    if (occ < 1) { set_error("attach_tc: kernel does not fit on an SM (smem %zu B)", h->smem_bytes); return DIFFSG_E_UNSUPPORTED; }
    // The occupancy API reports 1 for every kernel that contains tcgen05.alloc (it cannot see how many
    // TMEM columns are requested); two CTAs x (256 columns, <= 113 KB smem, 30 K registers) do fit.
    int fit = (int)(233472 / (h->smem_bytes + 1024));
    h->occupancy = fit < 1 ? 1 : (fit > kCtasPerSm ? kCtasPerSm : fit);
    h->grid_max = p->sm_count * h->occupancy;
    return DIFFSG_OK;
}

int tc_set_weights(diffsg_plan* p, const void* w_hi, const void* w_lo, size_t w_bytes, const float* params,
                   size_t n_params, const float* tt, int tt_rows, const void* tt_img, int img_rows, int64_t img_stride) {
    if (!p || !p->tc) { set_error("set_tc_weights: no tensor-core program attached"); return DIFFSG_E_STATE; }
    if (!w_hi || !params || (p->tc->dev.nterms == 3 && !w_lo)) { set_error("set_tc_weights: null argument"); return DIFFSG_E_INVALID; }
    if ((!tt || tt_rows <= 0) && (!tt_img || img_rows <= 0)) { set_error("set_tc_weights: neither a time table nor step images given"); return DIFFSG_E_INVALID; }
    if (((uintptr_t)w_hi | (uintptr_t)w_lo | (uintptr_t)params | (uintptr_t)tt | (uintptr_t)tt_img | (uintptr_t)img_stride) & 15) {
        set_error("set_tc_weights: blobs must be 16-byte aligned"); return DIFFSG_E_INVALID;
    }
    if (img_stride < 0 || img_stride > 0x7fffffff) { set_error("set_tc_weights: image stride"); return DIFFSG_E_INVALID; }
    (void)w_bytes; (void)n_params;
    TcDev& D = p->tc->dev;
    D.w_hi = (const uint8_t*)w_hi; D.w_lo = (const uint8_t*)w_lo; D.params = params;
    D.tt = tt; D.tt_img = (const uint8_t*)tt_img; D.tt_img_stride = (int)img_stride;
    p->tc->tt_rows = tt ? tt_rows : 0;
    p->tc->img_rows = tt_img ? img_rows : 0;
    p->tc->have_weights = true;
    return DIFFSG_OK;
}

// kStatus* bits raised by the kernels of this plan since the last reset.  Synchronises `st`.
int tc_status(diffsg_plan* p, int32_t* flags, int reset, cudaStream_t st) {
    if (!p || !p->tc || !flags) { set_error("plan_status: no tensor-core program attached"); return DIFFSG_E_STATE; }
    DIFFSG_CUDA_OK(cudaSetDevice(p->cfg.device));
    int h = 0;
    DIFFSG_CUDA_OK(cudaMemcpyAsync(&h, p->tc->d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (reset) DIFFSG_CUDA_OK(cudaMemsetAsync(p->tc->d_status, 0, sizeof(int), st));
    DIFFSG_CUDA_OK(cudaStreamSynchronize(st));
    *flags = h;
    return DIFFSG_OK;
}

int tc_query(const diffsg_plan* p, int what) {
    if (!p->tc) return -1;
    switch (what) {
        case 1: return p->tc->occupancy;
        case 2: return (int)p->tc->smem_bytes;
        case 3: return p->tc->grid_max;
        case 4: return p->tc->dev.nterms;
        case 6: return kChunkK;
        case 7: return kRegionCols;
        case 8: return (int)(p->tc->dev.scratch_floats * sizeof(float));
        default: return -1;
    }
}

// The stage program is read from __constant__ memory: upload it (stream-ordered) whenever another
// plan's program is resident on this device.  One tensor-core program is active per device at a time.
static int tc_activate(diffsg_plan* p, cudaStream_t st) {
    TcHost* h = p->tc;
    std::lock_guard<std::mutex> lock(g_prog_mutex);
    const int dev = p->cfg.device & 63;
    if (g_resident[dev] == h->serial) return DIFFSG_OK;
    if (g_last_launch[dev]) DIFFSG_CUDA_OK(cudaStreamWaitEvent(st, g_last_launch[dev], 0));
    DIFFSG_CUDA_OK(cudaMemcpyToSymbolAsync(c_stages, h->stages.data(), sizeof(Stage) * h->stages.size(), 0, cudaMemcpyHostToDevice, st));
    DIFFSG_CUDA_OK(cudaMemcpyToSymbolAsync(c_chunks, h->chunks.data(), sizeof(Chunk) * h->chunks.size(), 0, cudaMemcpyHostToDevice, st));
    DIFFSG_CUDA_OK(cudaMemcpyToSymbolAsync(c_epis, h->epis.data(), sizeof(Epi) * h->epis.size(), 0, cudaMemcpyHostToDevice, st));
    g_resident[dev] = h->serial;
    return DIFFSG_OK;
}

static int tc_mark_launch(const diffsg_plan* p, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_prog_mutex);
    const int dev = p->cfg.device & 63;
    if (!g_last_launch[dev]) DIFFSG_CUDA_OK(cudaEventCreateWithFlags(&g_last_launch[dev], cudaEventDisableTiming));
    DIFFSG_CUDA_OK(cudaEventRecord(g_last_launch[dev], st));
    return DIFFSG_OK;
}

static int tc_grid(const diffsg_plan* p, int64_t B) {
    const int64_t tiles = (B + kRows - 1) / kRows;
    return (int)(tiles < p->tc->grid_max ? tiles : p->tc->grid_max);
}

int tc_forward(diffsg_plan* p, const float* x, const int32_t* t_idx, const float* cond, const float* mask,
               float* eps, int64_t B, cudaStream_t st) {
    if (!p->tc || !p->tc->have_weights || !p->tc->tt_rows) { set_error("tensor-core engine selected but not initialised (forward needs the fp32 time table)"); return DIFFSG_E_STATE; }
    if (int rc = tc_activate(p, st)) return rc;
    RunArgs R;
    memset(&R, 0, sizeof(R));
    R.x = x; R.t_idx = t_idx; R.cond = cond; R.mask = mask; R.eps = eps; R.B = B;
    // fp16x2 (nterms <= 2) runs the one-MUFU Swish; fp16x3 (the ~fp32 mode) the exact two-MUFU form
    if (p->tc->dev.nterms <= 2) tc_unet_kernel<false, true><<<tc_grid(p, B), kThreads, p->tc->smem_bytes, st>>>(p->tc->dev, R);
    else tc_unet_kernel<false, false><<<tc_grid(p, B), kThreads, p->tc->smem_bytes, st>>>(p->tc->dev, R);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return tc_mark_launch(p, st);
}

int tc_sample_check(diffsg_plan* p, const diffsg_sample_args* a) {
    if (!p->tc || !p->tc->have_weights) { set_error("tensor-core engine selected but not initialised"); return DIFFSG_E_STATE; }
    if (a->T > p->tc->img_rows) { set_error("tensor-core step images cover %d steps, T=%d", p->tc->img_rows, a->T); return DIFFSG_E_INVALID; }
    return DIFFSG_OK;
}

// one launch: reverse steps step_hi .. step_lo of every tile (a re-normalised step is launched alone)
int tc_sample_launch(diffsg_plan* p, const diffsg_sample_args* a, int norm_steps, int step_hi, int step_lo, cudaStream_t st) {
    const int T = a->T;
    if (int rc = tc_activate(p, st)) return rc;
    RunArgs R;
    memset(&R, 0, sizeof(R));
    R.cond = a->cond_dev; R.y = a->y_dev; R.noise = a->noise_dev; R.rec_y = a->rec_y_dev; R.rec_eps = a->rec_eps_dev;
    R.stats = a->stat_ws_dev; R.B = a->B; R.T = T; R.norm_steps = norm_steps; R.omega = a->omega;
    R.seed = a->philox_seed; R.offset = a->philox_offset;
    for (int i = 0; i < T; ++i) { R.c_eps[i] = a->coef_host[i]; R.c_rs[i] = a->coef_host[T + i]; R.c_noise[i] = a->coef_host[2 * T + i]; }
    R.step_hi = step_hi; R.step_lo = step_lo;
    if (p->tc->dev.nterms <= 2) tc_unet_kernel<true, true><<<tc_grid(p, a->B), kThreads, p->tc->smem_bytes, st>>>(p->tc->dev, R);
    else tc_unet_kernel<true, false><<<tc_grid(p, a->B), kThreads, p->tc->smem_bytes, st>>>(p->tc->dev, R);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    if (int rc = tc_mark_launch(p, st)) return rc;
#ifdef DIFFSG_TC_TIMING
    {
        static long long* dbg = nullptr;
        if (!dbg) { cudaMalloc(&dbg, 204 * sizeof(long long)); }
        p->tc->dev.debug = dbg;
        cudaStreamSynchronize(st);
        long long h[204];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        const char* names[12] = {"acc_wait", "pkg_wait", "ln_pass1", "ln_pass2", "catln", "raw", "cond", "out", "total", "mma_wait_w", "mma_wait_a", "-"};
        fprintf(stderr, "[tc timing, cycles of thread 0 / CTA 0, previous launch]");
        for (int i = 0; i < 12; ++i) fprintf(stderr, " %s=%lld", names[i], h[i]);
        fprintf(stderr, "\n[tc stage wait]");       // per stage: cycles thread 0 waited for the accumulator ...
        for (int i = 0; i < 96; ++i) fprintf(stderr, " %lld", h[12 + i]);
        fprintf(stderr, "\n[tc stage work]");       // ... and spent on everything else
        for (int i = 0; i < 96; ++i) fprintf(stderr, " %lld", h[108 + i]);
        fprintf(stderr, "\n");
    }
#endif
    return DIFFSG_OK;
}

}  // namespace tc
}  // namespace diffsg

using namespace diffsg;

extern "C" int diffsg_debug_tc_gemm(const float* A_dev, const void* W_hi_dev, const void* W_lo_dev, float* C_dev,
                                    int32_t K, int32_t N, int32_t nterms, uint32_t layout, uint32_t lbo,
                                    int32_t swap_lbo_sbo, void* stream) {
    if (!A_dev || !W_hi_dev || !C_dev || K <= 0 || K % 16 || N < 16 || N % 16 || N > tc::kMaxN || nterms < 1 || nterms > 3 ||
        (nterms == 3 && !W_lo_dev)) {
        set_error("debug_tc_gemm: bad argument (K %% 16 == 0, 16 <= N <= %d, N %% 16 == 0)", tc::kMaxN);
        return DIFFSG_E_INVALID;
    }
    tc::GemmTestArgs g{A_dev, (const __half*)W_hi_dev, (const __half*)W_lo_dev, C_dev, K, N, nterms, layout, lbo, swap_lbo_sbo};
    const int smem = (int)sizeof(tc::GemmTestSmem) + 1024;
    DIFFSG_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    tc::tc_gemm_test_kernel<<<1, 192, smem, (cudaStream_t)stream>>>(g);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}
