// Tensor-core (tcgen05 / TMEM / bulk-TMA) engine of the UNet1D op program, sm_100a.
//
// Building block validated in isolation by diffsg_debug_tc_gemm (tests/test_gpu_tc.py):
//   C[128, N] = A[128, K] . W[N, K]^T,  A split on the fly into fp16 (hi, lo) by the
//   "epilogue" threads (thread == row == TMEM lane), W pre-packed on the host as fp16 core-matrix
//   images and streamed with 1-D bulk TMA, fp32 accumulation in TMEM.
#include <cstdio>

#include "common.cuh"
#include "tc_common.cuh"

namespace diffsg {
namespace tc {

constexpr int kTileRows = 128;      // rows per CTA tile == UMMA M == TMEM lanes
constexpr int kChunkK = 64;         // K elements per operand chunk (A slot / W stage)
constexpr int kSlotBytes = kTileRows * kChunkK * 2;   // one fp16 A chunk (16 KB)
constexpr int kASlots = 2;
constexpr int kWStages = 2;
constexpr int kMaxN = 128;
constexpr int kStageBytes = kMaxN * kChunkK * 2;      // one fp16 W chunk (16 KB)

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
