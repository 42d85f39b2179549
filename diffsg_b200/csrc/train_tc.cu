// Training GEMMs on the 5th-generation tensor cores (sm_100a): the Linear layers of the eps-MSE training step
// (reference ddpm_opt/UNetCF.py:83-95 ResidualBlock, :318-356 UNet1D.forward; loss classifier_free_MSR.py:100-112)
// as hand-written tcgen05 kernels with the LayerNorm -> Swish pairs fused in:
//
//   tlin_fwd_kernel           y  = [swish(LN(a)) | a] . W^T (+ a2 . W2^T) + bias (+ bias2) (+ add) (+ gadd[gidx])
//   tlin_bwd_kernel, dgrad    dx = dy . W, optionally pushed through the LayerNorm -> Swish backward in the epilogue
//                             (dgamma / dbeta column sums by a shuffle butterfly), plus an optional addend
//   tlin_bwd_kernel, wgrad    dW += dy^T . act(a), db += column sums of dy, d(gadd) += scatter of dy by gidx — the last
//                             two on the tensor cores as well (a ones column and T one-hot columns appended to act(a))
//   (one backward launch runs up to two dgrad and two wgrad problems of a node; CTAs pick their role by block index)
//
// Arithmetic: operands are fp32 in HBM; every operand element is split into bf16 (hi, lo) while it is staged into
// shared memory and each product runs as three kind::f16 (bf16) MMAs (hi.hi + lo.hi + hi.lo) with fp32 accumulation
// in TMEM: 16 significant operand bits, fp32 range (gradients of 1e-8 do not underflow) — "bf16x3".
// One CTA = 256 threads = one 128-row tile (UMMA M = 128 = TMEM lanes; in the epilogues a thread owns one accumulator
// row); operands are UMMA K-major no-swizzle core-matrix chunks of 64 contraction columns (the sampler engine's
// layout, unet_tc.cuh).
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace diffsg {
namespace ttc {

using namespace diffsg::tc;

constexpr int kRows = 128;              // rows per tile = UMMA M = TMEM lanes
constexpr int kKC = 64;                 // contraction columns per staged chunk
constexpr uint32_t kLBO = 128;          // K-adjacent core matrices are contiguous
constexpr uint32_t kSBO = kKC * 16;     // next 8 rows
constexpr int kThreads = 256;           // 8 warps; see the scaffolding notes
constexpr int kWgradKT = 224;           // feature columns per wgrad CTA (+ <= 32 extra columns = 256 = max UMMA N)
constexpr int kMaxExtra = 32;

struct Mat {                            // logical row-major [rows, k0 + k1] = cat(p0[rows, k0], p1[rows, k1])
    const float* p0;
    const float* p1;
    int k0, k1, vec;                    // vec: both halves 8-column / 16-byte aligned
};
struct MatOut {
    float* p0;
    float* p1;
    int k0, k1, vec;
};

__device__ __forceinline__ uint32_t op_off(int mn, int k) {
    return (uint32_t)(mn >> 3) * kSBO + (uint32_t)(k >> 3) * 128u + (uint32_t)(mn & 7) * 16u + (uint32_t)(k & 7) * 2u;
}
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(uint32_t M, uint32_t N) {
    // c = f32 (bit 4), a = b = bf16 (format 1 at bits 7 and 10), both K-major
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
        const float2 hf = __bfloat1622float2(hh);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void store_split(uint8_t* hi_base, uint32_t lo_delta, uint32_t off, const float (&v)[8]) {
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(hi_base + off) = hi;
    *reinterpret_cast<uint4*>(hi_base + lo_delta + off) = lo;
}

// 8 consecutive columns [c, c + 8) of one row; columns beyond the matrix read as 0
template <bool kV>
__device__ __forceinline__ void load8(const Mat& m, int64_t row, int c, float (&v)[8]) {
    const int K = m.k0 + m.k1;
    if (kV || m.vec) {
        if (c >= K) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            return;
        }
        const float* p = c < m.k0 ? m.p0 + row * m.k0 + c : m.p1 + row * m.k1 + (c - m.k0);
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cc = c + j;
            v[j] = cc < m.k0 ? __ldg(m.p0 + row * m.k0 + cc) : (cc < K ? __ldg(m.p1 + row * m.k1 + (cc - m.k0)) : 0.f);
        }
    }
}
__device__ __forceinline__ float load1(const Mat& m, int64_t row, int c) {
    return c < m.k0 ? __ldg(m.p0 + row * m.k0 + c) : __ldg(m.p1 + row * m.k1 + (c - m.k0));
}
template <bool kV>
__device__ __forceinline__ void store8(const MatOut& m, int64_t row, int c, const float (&v)[8]) {
    const int K = m.k0 + m.k1;
    if (kV || m.vec) {
        if (c >= K) return;
        float* p = c < m.k0 ? m.p0 + row * m.k0 + c : m.p1 + row * m.k1 + (c - m.k0);
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int cc = c + j;
            if (cc < m.k0) m.p0[row * m.k0 + cc] = v[j];
            else if (cc < K) m.p1[row * m.k1 + (cc - m.k0)] = v[j];
        }
    }
}
// branch-free (MUFU.EX2 + MUFU.RCP, ~2 ulp): an IEEE division here puts a slow-path branch between the elements of a
// row and serialises their dependent chains
// x / d and x % d for the small item counters of the staging loops (x < 4096, d <= 256): q = (x ceil(2^20 / d)) >> 20
// is exact there (x (d - 1) < 2^20) and fits 32 bits
struct FastDiv {
    uint32_t d, m;
    __device__ __forceinline__ explicit FastDiv(int d_) : d((uint32_t)d_), m(((1u << 20) + (uint32_t)d_ - 1) / (uint32_t)d_) {}
    __device__ __forceinline__ void divmod(int x, int& q, int& r) const {
        q = (int)(((uint32_t)x * m) >> 20);
        r = x - q * (int)d;
    }
};
__device__ __forceinline__ float sigmoidf_(float n) { return __fdividef(1.0f, 1.0f + __expf(-n)); }

__device__ __forceinline__ void red_add(float* p, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// column sums over the 32 lanes of a warp of 16 per-lane values: 16 shuffles (recursive halving).  Every lane returns the
// sum of column 8 b4 + 4 b3 + 2 b2 + b1 (b_i = bit i of the lane id); lanes l and l ^ 1 hold the same column.
__device__ __forceinline__ float colsum16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    const bool u4 = lane & 16, u3 = lane & 8, u2 = lane & 4, u1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = u4 ? v[i] : v[i + 8], keep = u4 ? v[i + 8] : v[i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = u3 ? a[i] : a[i + 4], keep = u3 ? a[i + 4] : a[i];
        b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = u2 ? b[i] : b[i + 2], keep = u2 ? b[i + 2] : b[i];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = u1 ? c[0] : c[1], keep = u1 ? c[1] : c[0];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    return d;
}
__device__ __forceinline__ int colsum16_col(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

// Programmatic dependent launch: the kernels are launched with programmatic stream serialization, so a CTA may start
// while the previous kernel of the stream is still draining.  Everything before pdl_wait() touches only parameters
// (weights, LayerNorm gamma / beta: last written by the optimiser, many full stream barriers ago) and on-chip state
// (barriers, TMEM allocation); activations and gradients are read and written after it.  pdl_wait() returns when the
// previous kernel has completed and flushed, which (every kernel of this file waits before it finishes) orders the
// whole chain.  launch_dependents right after the wait lets the NEXT kernel's prologue overlap this kernel's body.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ CTA scaffolding
// 256 threads (8 warps), 2 CTAs per SM.  Operand staging is done by rolled loops shared by all threads, one 8-element
// item (8 consecutive contraction elements of one operand row) per iteration: load fp32, optional LayerNorm -> Swish,
// split into bf16 (hi, lo), two 16-byte stores into the core-matrix chunk.  The 16 warps of an SM cover the load
// latency; an earlier version that unrolled 8 items per thread and prefetched the next chunk into registers was 3-4x the
// code and bound by instruction fetch.  All eight warps share the epilogues: warp w reads the TMEM
// lanes of sub-partition w % 4, the two warps of a sub-partition ("sides") take alternating 32-column rounds.
struct Smem {
    uint64_t bar;
    uint32_t tmem_base;
    uint32_t pad_;
};

// operand buffers: [A hi | A lo | B hi | B lo], A = 128 x 64 bf16 = 16 KB per term, B = bn_pad x 64 bf16 per term
__device__ __forceinline__ uint32_t a_term_bytes() { return kRows * kKC * 2; }

__device__ __forceinline__ void cta_setup(Smem& S, uint32_t tmem_cols) {
    if (threadIdx.x == 0) {
        mbar_init(&S.bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        __syncwarp();
        tmem_alloc(&S.tmem_base, tmem_cols);
        tmem_relinquish();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
}
__device__ __forceinline__ void cta_teardown(Smem& S, uint32_t tmem_cols) {
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(S.tmem_base, tmem_cols);
}

// publish the staged chunk and issue its MMAs (one thread); completion arrives on S.bar
__device__ __forceinline__ void publish_and_issue(Smem& S, uint8_t* a_hi, uint8_t* b_hi, uint32_t b_term, int kw, uint32_t idesc,
                                                  uint32_t d_tmem, bool first) {
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            tcgen05_fence_after();
            const uint64_t da_hi = make_smem_desc(smem_u32(a_hi), kLBO, kSBO, 0);
            const uint64_t da_lo = make_smem_desc(smem_u32(a_hi + a_term_bytes()), kLBO, kSBO, 0);
            const uint64_t db_hi = make_smem_desc(smem_u32(b_hi), kLBO, kSBO, 0);
            const uint64_t db_lo = make_smem_desc(smem_u32(b_hi + b_term), kLBO, kSBO, 0);
            uint32_t acc = first ? 0u : 1u;
            for (int ks = 0; ks < kw / 16; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 16);        // two core matrices = 256 bytes, >> 4
                umma_f16(d_tmem, da_hi + adv, db_hi + adv, idesc, acc);
                umma_f16(d_tmem, da_lo + adv, db_hi + adv, idesc, 1);
                umma_f16(d_tmem, da_hi + adv, db_lo + adv, idesc, 1);
                acc = 1;
            }
            umma_commit(&S.bar);
        }
        __syncwarp();       // the other lanes park here: a lane spinning on the mbarrier would starve the issuing lane
    }
}
__device__ __forceinline__ void wait_consumed(Smem& S, uint32_t& phase) {
    mbar_wait(&S.bar, phase);
    phase ^= 1;
    tcgen05_fence_after();
}

// ---- LayerNorm moments: lane l of warp w (w < 4) covers rows 32 w + 8 g + (l & 7), g = 0..3, and every fourth
// 8-column piece starting at l >> 3: one warp-wide load touches 8 rows x 128 contiguous bytes (8 cache lines instead of
// the 32 of a thread-per-row walk); the four lanes that share a row meet through two shuffles.
struct RowMap {
    int r_in, pq, wbase;
    __device__ __forceinline__ explicit RowMap(int t) : r_in(t & 7), pq((t & 31) >> 3), wbase(t & ~31) {}
    __device__ __forceinline__ int row(int u) const { return wbase + 8 * (u >> 1) + r_in; }      // row inside the tile
    __device__ __forceinline__ int piece(int u) const { return pq + 4 * (u & 1); }
};
// ---- epilogue through shared memory: the accumulator rows (thread == row == TMEM lane) are parked in a padded row-major
// tile in the operand buffers (free after the last MMA) and leave it through row-contiguous, fully coalesced accesses
constexpr int kTilePad = 4;             // floats; row stride 4 (mod 32) words: conflict-free 16-byte accesses from 8 rows
__device__ __forceinline__ void tmem_rows_to_tile(float* tile, int ld, uint32_t d_tmem, int warp, int side, int t, int col0, int ncols16) {
    const uint32_t t_row = d_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)col0;
    for (int g = 2 * side; g < ncols16; g += 4) {
        const bool two = g + 1 < ncols16;
        float v[2][16];
        tmem_ld16(t_row + g * 16, v[0]);
        if (two) tmem_ld16(t_row + (g + 1) * 16, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) continue;
            float4* dst = reinterpret_cast<float4*>(tile + (size_t)t * ld + (g + u) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[u][4 * j], v[u][4 * j + 1], v[u][4 * j + 2], v[u][4 * j + 3]);
        }
    }
}

// ================================================================================================ forward
struct FwdArgs {
    Mat a;
    const float* w;
    const float* bias;
    const float* gamma;
    const float* beta;
    float* mean;
    float* rstd;
    Mat a2;
    const float* w2;
    const float* bias2;
    const float* add;
    const float* gadd;
    const int64_t* gidx;
    float* y;
    int64_t B;
    int N, n_pad, tmem_cols, wvec, wvec2, yvec, gadd_ld;
};

template <bool kV>
__global__ void __launch_bounds__(kThreads, 2) tlin_fwd_kernel(const FwdArgs P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Smem S;
    __shared__ float s_gamma[256], s_beta[256], s_mu[kRows], s_rs[kRows];
    uint8_t* a_hi = smem_raw;
    uint8_t* b_hi = smem_raw + 2 * a_term_bytes();
    const uint32_t b_term = (uint32_t)P.n_pad * kKC * 2;

    const int tid = threadIdx.x, warp = tid >> 5, side = warp >> 2, t = tid & 127;
    const int n0 = blockIdx.y * 128;
    const int n_valid = min(128, P.N - n0);
    const int n_pad = min(P.n_pad, (n_valid + 15) & ~15);
    const bool ln = P.gamma != nullptr;
    const int K0 = P.a.k0 + P.a.k1, K1 = P.a2.p0 ? P.a2.k0 + P.a2.k1 : 0;
    const int nc0 = (K0 + kKC - 1) / kKC, nchunk = nc0 + (K1 + kKC - 1) / kKC;

    if (ln)
        for (int i = tid; i < 256; i += kThreads) { s_gamma[i] = i < K0 ? P.gamma[i] : 0.f; s_beta[i] = i < K0 ? P.beta[i] : 0.f; }
    cta_setup(S, P.tmem_cols);
    const uint32_t d_tmem = S.tmem_base;
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n_pad);

    const FastDiv fd(n_pad);
    pdl_wait();

    // LayerNorm moments of the thread's four rows (warps 0-3):
    // shifted one-pass sums over the lane's pieces, then over the four lanes that share a row
    const RowMap rm(t);
    const int64_t row0 = (int64_t)blockIdx.x * kRows;
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, rs[4] = {0.f, 0.f, 0.f, 0.f};
    if (side == 0 && ln) {
        float x0[4], sm[4], sq[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int64_t r = row0 + rm.row(2 * g);
            x0[g] = r < P.B ? load1(P.a, r, 0) : 0.f;
            sm[g] = sq[g] = 0.f;
        }
        for (int c = rm.pq * 8; c < K0; c += 32) {
            float v[4][8];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int64_t r = row0 + rm.row(2 * g);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[g][j] = 0.f;
                if (r < P.B) load8<kV>(P.a, r, c, v[g]);                       // columns >= K read as 0
            }
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = (c + j < K0) ? v[g][j] - x0[g] : 0.f;
                    sm[g] += d;
                    sq[g] = fmaf(d, d, sq[g]);
                }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            sm[g] += __shfl_xor_sync(0xffffffffu, sm[g], 8);
            sq[g] += __shfl_xor_sync(0xffffffffu, sq[g], 8);
            sm[g] += __shfl_xor_sync(0xffffffffu, sm[g], 16);
            sq[g] += __shfl_xor_sync(0xffffffffu, sq[g], 16);
            const float md = sm[g] / (float)K0;
            mu[g] = x0[g] + md;
            rs[g] = rsqrtf(fmaxf(sq[g] / (float)K0 - md * md, 0.f) + kLnEps);
            const int64_t r = row0 + rm.row(2 * g);
            if (blockIdx.y == 0 && rm.pq == 0 && r < P.B) { P.mean[r] = mu[g]; P.rstd[r] = rs[g]; }
            if (rm.pq == 0) { s_mu[rm.row(2 * g)] = mu[g]; s_rs[rm.row(2 * g)] = rs[g]; }
        }
    }

    // Staging: rolled loops shared by all 256 threads, one 8-element item per iteration (see dgrad_body)
    uint32_t phase = 0;
    {
        __syncthreads();                                     // row statistics visible to every staging thread
        for (int i = 0; i < nchunk; ++i) {
            const int seg = i >= nc0;
            const int K = seg ? K1 : K0, k0 = (seg ? i - nc0 : i) * kKC;
            const int kw = min(kKC, (K - k0 + 15) & ~15);
            const int pieces = kw >> 3;
            const Mat& A = seg ? P.a2 : P.a;
            const bool act = ln && seg == 0;
            if (i > 0) wait_consumed(S, phase);
#pragma unroll 2
            for (int a = tid; a < kRows * 8; a += kThreads) {                     // A rows: 8 rows x 128 B per warp load
                const int pc = (a >> 3) & 7, rl = ((a >> 6) << 3) | (a & 7);
                if (pc >= pieces) continue;
                const int c = k0 + pc * 8;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
                if (row0 + rl < P.B) load8<kV>(A, row0 + rl, c, v);
                if (act) {
                    const float m = s_mu[rl], r = s_rs[rl];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float n = fmaf((v[q] - m) * r, s_gamma[c + q], s_beta[c + q]);
                        v[q] = n * sigmoidf_(n);
                    }
                }
                store_split(a_hi, a_term_bytes(), op_off(rl, pc * 8), v);
            }
            const float* W = seg ? P.w2 : P.w;
            const bool wv = kV || (seg ? P.wvec2 : P.wvec);
            const int total = n_pad * pieces;
#pragma unroll 2
            for (int it = tid; it < total; it += kThreads) {                       // weights [N, K]
                int n, j;
                fd.divmod(it, j, n);
                const int c = k0 + j * 8;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
                if (n < n_valid && c < K) {
                    const float* src = W + (size_t)(n0 + n) * K + c;
                    if (wv) {
                        const float4 x = __ldg(reinterpret_cast<const float4*>(src)), y = __ldg(reinterpret_cast<const float4*>(src) + 1);
                        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = (c + q < K) ? __ldg(src + q) : 0.f;
                    }
                }
                store_split(b_hi, b_term, op_off(n, j * 8), v);
            }
            publish_and_issue(S, a_hi, b_hi, b_term, kw, idesc, d_tmem, i == 0);
        }
    }
    wait_consumed(S, phase);

    // epilogue: accumulator rows -> shared-memory tile -> (+ bias, + add, + gathered add) -> y, row-contiguous
    float* tile = reinterpret_cast<float*>(smem_raw);
    const int ld = n_pad + kTilePad;
    tmem_rows_to_tile(tile, ld, d_tmem, warp, side, t, 0, n_pad / 16);
    __syncthreads();
    if (kV || P.yvec) {
        const int nc4 = n_valid >> 2;                       // N % 4 == 0
        const FastDiv f4(nc4);
        const int total = kRows * nc4;
        for (int it0 = tid; it0 < total; it0 += 4 * kThreads) {
            float4 ex[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u * kThreads;
                int r, c4;
                f4.divmod(it, r, c4);
                const int64_t grow_ = row0 + r;
                float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                if (it < total && grow_ < P.B) {
                    const int c = n0 + c4 * 4;
                    if (P.bias) { const float4 x = __ldg(reinterpret_cast<const float4*>(P.bias + c)); e.x += x.x; e.y += x.y; e.z += x.z; e.w += x.w; }
                    if (P.bias2) { const float4 x = __ldg(reinterpret_cast<const float4*>(P.bias2 + c)); e.x += x.x; e.y += x.y; e.z += x.z; e.w += x.w; }
                    if (P.add) { const float4 x = __ldg(reinterpret_cast<const float4*>(P.add + grow_ * P.N + c)); e.x += x.x; e.y += x.y; e.z += x.z; e.w += x.w; }
                    if (P.gadd) {
                        const float4 x = __ldg(reinterpret_cast<const float4*>(P.gadd + P.gidx[grow_] * P.gadd_ld + c));
                        e.x += x.x; e.y += x.y; e.z += x.z; e.w += x.w;
                    }
                }
                ex[u] = e;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int it = it0 + u * kThreads;
                int r, c4;
                f4.divmod(it, r, c4);
                const int64_t grow_ = row0 + r;
                if (it < total && grow_ < P.B) {
                    const float4 a = *reinterpret_cast<const float4*>(tile + (size_t)r * ld + c4 * 4);
                    *reinterpret_cast<float4*>(P.y + grow_ * P.N + n0 + c4 * 4) = make_float4(a.x + ex[u].x, a.y + ex[u].y, a.z + ex[u].z, a.w + ex[u].w);
                }
            }
        }
    } else {
        const FastDiv f1(n_valid);
        const int total = kRows * n_valid;
        for (int it = tid; it < total; it += kThreads) {
            int r, cl;
            f1.divmod(it, r, cl);
            const int64_t grow_ = row0 + r;
            if (grow_ >= P.B) continue;
            const int c = n0 + cl;
            float e = tile[(size_t)r * ld + cl];
            if (P.bias) e += __ldg(P.bias + c);
            if (P.bias2) e += __ldg(P.bias2 + c);
            if (P.add) e += __ldg(P.add + grow_ * P.N + c);
            if (P.gadd) e += __ldg(P.gadd + P.gidx[grow_] * P.gadd_ld + c);
            P.y[grow_ * P.N + c] = e;
        }
    }
    cta_teardown(S, P.tmem_cols);
}

// ================================================================================================ dgrad
struct DgradArgs {
    const float* dy;       // [B, N]
    const float* w;        // [N, K]
    Mat x;                 // LN mode: the forward input
    const float* gamma;
    const float* beta;
    const float* mean;
    const float* rstd;
    Mat dres;              // optional addend (p0 == nullptr: none)
    MatOut dx;
    float* dgamma;
    float* dbeta;
    int64_t B;
    int N, K, kt, kt_pad, tmem_cols, dyvec;
};

template <bool kV>
__device__ __forceinline__ void dgrad_body(const DgradArgs& P, uint8_t* smem_raw, Smem& S, int bx, int by) {
    __shared__ float s_gamma[256], s_beta[256], s_dg[256], s_db[256], s_p1[2][kRows], s_p2[2][kRows];
    uint8_t* a_hi = smem_raw;
    uint8_t* b_hi = smem_raw + 2 * a_term_bytes();
    const uint32_t b_term = (uint32_t)P.kt_pad * kKC * 2;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, side = warp >> 2, t = tid & 127;
    const int64_t row = (int64_t)bx * kRows + t;
    const bool valid = row < P.B;
    const int kb = by * P.kt;                       // first output column of this CTA
    const int k_valid = min(P.kt, P.K - kb);
    const int k_pad = (k_valid + 15) & ~15;
    const bool ln = P.gamma != nullptr;

    if (ln)
        for (int i = tid; i < 256; i += kThreads) {
            s_gamma[i] = i < P.K ? P.gamma[i] : 0.f;
            s_beta[i] = i < P.K ? P.beta[i] : 0.f;
            s_dg[i] = 0.f;
            s_db[i] = 0.f;
        }
    cta_setup(S, P.tmem_cols);
    const uint32_t d_tmem = S.tmem_base;
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)k_pad);
    const Mat DY{P.dy, nullptr, P.N, 0, P.dyvec};
    const int nchunk = (P.N + kKC - 1) / kKC;
    const FastDiv fd(k_pad);
    const int64_t row0 = (int64_t)bx * kRows;

    // Staging: rolled loops shared by all 256 threads, one 8-element item per iteration (the two CTAs x 8 warps of an SM
    // cover the load latency; an unrolled register-prefetch version of these loops was ~4x the code and bound by
    // instruction fetch, stall_no_instructions 45-50 %)
    uint32_t phase = 0;
    pdl_wait();
    for (int i = 0; i < nchunk; ++i) {
        const int n0 = i * kKC, kw = min(kKC, (P.N - n0 + 15) & ~15);
        const int pieces = kw >> 3;
        if (i > 0) wait_consumed(S, phase);
#pragma unroll 2
        for (int a = tid; a < kRows * 8; a += kThreads) {                     // A = dy rows: 8 rows x 128 B per warp load
            const int pc = (a >> 3) & 7, rl = ((a >> 6) << 3) | (a & 7);
            if (pc >= pieces) continue;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = 0.f;
            if (row0 + rl < P.B) load8<kV>(DY, row0 + rl, n0 + pc * 8, v);
            store_split(a_hi, a_term_bytes(), op_off(rl, pc * 8), v);
        }
        const int total = k_pad * pieces;
#pragma unroll 2
        for (int it = tid; it < total; it += kThreads) {                       // B = W^T: (mn = k, kk = n), lanes along k
            int k, n8;
            fd.divmod(it, n8, k);
            const int nb = n0 + n8 * 8;
            const int nv = k < k_valid ? min(P.N - nb, 8) : 0;                 // rows of W beyond N / columns beyond K: 0
            const float* base = P.w + (size_t)nb * P.K + kb + k;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = q < nv ? __ldg(base + q * P.K) : 0.f;
            store_split(b_hi, b_term, op_off(k, n8 * 8), v);
        }
        publish_and_issue(S, a_hi, b_hi, b_term, kw, idesc, d_tmem, i == 0);
    }
    wait_consumed(S, phase);

    const uint32_t t_row = d_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int ngroups = k_pad / 16;
    if (!ln) {
        // accumulator rows -> shared-memory tile -> (+ dres) -> dx, row-contiguous
        float* tile = reinterpret_cast<float*>(smem_raw);
        const int ld = k_pad + kTilePad;
        tmem_rows_to_tile(tile, ld, d_tmem, warp, side, t, 0, ngroups);
        __syncthreads();
        if (kV || (P.dx.vec && (!P.dres.p0 || P.dres.vec))) {
            const int nc4 = k_valid >> 2;
            const FastDiv f4(nc4);
            const int total = kRows * nc4;
            for (int it0 = tid; it0 < total; it0 += 4 * kThreads) {
                float4 ex[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int it = it0 + u * kThreads;
                    int r, c4;
                    f4.divmod(it, r, c4);
                    const int64_t grow_ = row0 + r;
                    const int c = kb + c4 * 4;
                    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (it < total && grow_ < P.B && P.dres.p0)
                        e = __ldg(reinterpret_cast<const float4*>(c < P.dres.k0 ? P.dres.p0 + grow_ * P.dres.k0 + c : P.dres.p1 + grow_ * P.dres.k1 + (c - P.dres.k0)));
                    ex[u] = e;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int it = it0 + u * kThreads;
                    int r, c4;
                    f4.divmod(it, r, c4);
                    const int64_t grow_ = row0 + r;
                    const int c = kb + c4 * 4;
                    if (it < total && grow_ < P.B) {
                        const float4 a = *reinterpret_cast<const float4*>(tile + (size_t)r * ld + c4 * 4);
                        float* dst = c < P.dx.k0 ? P.dx.p0 + grow_ * P.dx.k0 + c : P.dx.p1 + grow_ * P.dx.k1 + (c - P.dx.k0);
                        *reinterpret_cast<float4*>(dst) = make_float4(a.x + ex[u].x, a.y + ex[u].y, a.z + ex[u].z, a.w + ex[u].w);
                    }
                }
            }
        } else {
            const FastDiv f1(k_valid);
            const int total = kRows * k_valid;
            for (int it = tid; it < total; it += kThreads) {
                int r, cl;
                f1.divmod(it, r, cl);
                const int64_t grow_ = row0 + r;
                if (grow_ >= P.B) continue;
                const int c = kb + cl;
                float e = tile[(size_t)r * ld + cl];
                if (P.dres.p0) e += load1(P.dres, grow_, c);
                if (c < P.dx.k0) P.dx.p0[grow_ * P.dx.k0 + c] = e;
                else P.dx.p1[grow_ * P.dx.k1 + (c - P.dx.k0)] = e;
            }
        }
    } else {
        // LayerNorm -> Swish backward on the accumulator row g = d(swish(n)), n = gamma xh + beta, xh = (x - mu) rstd:
        //   dn = g swish'(n);  dxh = dn gamma;  dx = rstd (dxh - mean(dxh) - xh mean(dxh xh));  dgamma += dn xh;  dbeta += dn
        // Two passes over the accumulator (row sums first); each side owns every other 32-column round of the row and the
        // two partial row sums meet in shared memory.  x loads are issued before the TMEM loads of a round.
        const float mu = valid ? P.mean[row] : 0.f, rs = valid ? P.rstd[row] : 0.f;
        float s1 = 0.f, s2 = 0.f;
        for (int g = 2 * side; g < ngroups; g += 4) {
            const bool two = g + 1 < ngroups;
            float x[4][8];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
#pragma unroll
                for (int q = 0; q < 8; ++q) x[h][q] = 0.f;
                if (valid && (h < 2 || two)) load8<kV>(P.x, row, g * 16 + h * 8, x[h]);
            }
            float v[2][16];
            tmem_ld16(t_row + g * 16, v[0]);
            if (two) tmem_ld16(t_row + (g + 1) * 16, v[1]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) continue;            // uniform
                float pg[16], pb[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = (g + u) * 16 + j;
                    const float gm = s_gamma[c];
                    const float xh = (x[2 * u + (j >> 3)][j & 7] - mu) * rs;
                    const float n = fmaf(xh, gm, s_beta[c]);
                    const float sg = sigmoidf_(n);
                    const float dn = v[u][j] * (sg * fmaf(n, 1.0f - sg, 1.0f));    // invalid rows / columns >= K: accumulator is 0
                    const float dxh = dn * gm;
                    s1 += dxh;
                    s2 = fmaf(dxh, xh, s2);
                    pg[j] = dn * xh;
                    pb[j] = dn;
                }
                const float cg = colsum16(pg, lane), cb = colsum16(pb, lane);
                if (!(lane & 1)) {
                    const int c = (g + u) * 16 + colsum16_col(lane);
                    atomicAdd(&s_dg[c], cg);
                    atomicAdd(&s_db[c], cb);
                }
            }
        }
        s_p1[side][t] = s1;
        s_p2[side][t] = s2;
        __syncthreads();
        const float m1 = (s_p1[0][t] + s_p1[1][t]) / (float)P.K, m2 = (s_p2[0][t] + s_p2[1][t]) / (float)P.K;
        for (int g = 2 * side; g < ngroups; g += 4) {
            const bool two = g + 1 < ngroups;
            float x[4][8], o[4][8];
#pragma unroll
            for (int h = 0; h < 4; ++h) {
#pragma unroll
                for (int q = 0; q < 8; ++q) { x[h][q] = 0.f; o[h][q] = 0.f; }
                if (valid && (h < 2 || two)) {
                    load8<kV>(P.x, row, g * 16 + h * 8, x[h]);
                    if (P.dres.p0) load8<kV>(P.dres, row, g * 16 + h * 8, o[h]);
                }
            }
            float v[2][16];
            tmem_ld16(t_row + g * 16, v[0]);
            if (two) tmem_ld16(t_row + (g + 1) * 16, v[1]);
            tmem_ld_wait();
            if (!valid) continue;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int c = g * 16 + h * 8;
                if ((h >= 2 && !two) || c >= P.K) continue;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float gm = s_gamma[c + q];
                    const float xh = (x[h][q] - mu) * rs;
                    const float n = fmaf(xh, gm, s_beta[c + q]);
                    const float sg = sigmoidf_(n);
                    const float dxh = v[h >> 1][(h & 1) * 8 + q] * (sg * fmaf(n, 1.0f - sg, 1.0f)) * gm;
                    o[h][q] += rs * (dxh - m1 - xh * m2);
                }
                store8<kV>(P.dx, row, c, o[h]);
            }
        }
        __syncthreads();
        for (int i = tid; i < P.K; i += kThreads) {
            red_add(P.dgamma + i, s_dg[i]);
            red_add(P.dbeta + i, s_db[i]);
        }
    }
    cta_teardown(S, P.tmem_cols);
}

// ================================================================================================ wgrad
struct WgradArgs {
    const float* dy;       // [B, N]
    Mat a;                 // [B, K]: the forward A operand (before LayerNorm -> Swish when gamma != nullptr)
    const float* gamma;
    const float* beta;
    const float* mean;
    const float* rstd;
    const int64_t* gidx;   // [B] row -> gathered-add row (with dgadd)
    float* dw;             // [N, K]  +=
    float* dbias;          // [N]     +=   (nullable)
    float* dgadd;          // [T, N]  +=   (nullable)
    int64_t B;
    int N, K, T, n_chunks, tmem_cols, bcols_pad, dwvec, dgadd_ld;
};

template <bool kV>
__device__ __forceinline__ void wgrad_body(const WgradArgs& P, uint8_t* smem_raw, Smem& S, int bx, int by, int bz, int gx, int gz) {
    __shared__ float s_mu[kKC], s_rs[kKC];
    __shared__ int s_gi[kKC];
    uint8_t* a_hi = smem_raw;
    uint8_t* b_hi = smem_raw + 2 * a_term_bytes();
    const uint32_t b_term = (uint32_t)P.bcols_pad * kKC * 2;

    const int tid = threadIdx.x, warp = tid >> 5, side = warp >> 2;
    const int n0 = by * 128;
    const int n_valid = min(128, P.N - n0);
    const int kb = bz * kWgradKT;
    const int kcols = min(kWgradKT, P.K - kb);
    const bool last_z = bz == gz - 1;
    const int n_extra = last_z ? ((P.dbias ? 1 : 0) + (P.dgadd ? P.T : 0)) : 0;
    const int one_col = (last_z && P.dbias) ? kcols : -1;
    const int hot0 = (last_z && P.dgadd) ? kcols + (P.dbias ? 1 : 0) : -1;
    const int bcols = kcols + n_extra;
    const int bcols_pad = (bcols + 15) & ~15;
    const bool ln = P.gamma != nullptr;
    cta_setup(S, P.tmem_cols);
    const uint32_t d_tmem = S.tmem_base;
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)bcols_pad);

    pdl_wait();                          // both operands are activations / gradients of earlier kernels
    uint32_t phase = 0;
    const FastDiv fd_n(n_valid), fd_k(kcols), fd_x(max(bcols_pad - kcols, 1));

    // Persistent over 64-row chunks; staging by rolled loops shared by all 256 threads (see dgrad_body).  Transposed
    // operands: (mn = feature, kk = row), lanes along the feature (contiguous in memory), 8 rows per item.
    int it_no = 0;
    for (int ch = bx; ch < P.n_chunks; ch += gx, ++it_no) {
        const int64_t r0 = (int64_t)ch * kKC;
        const int kw = (int)min((int64_t)kKC, (P.B - r0 + 15) & ~(int64_t)15);
        const int nr8 = kw >> 3;
        const int rows_left = (int)min(P.B - r0, (int64_t)kKC);
        if (it_no > 0) wait_consumed(S, phase);
        if (tid < kKC) {                         // per-row scalars of the chunk (LayerNorm statistics, gather index)
            const int64_t r = r0 + tid;
            const bool ok = r < P.B;
            s_mu[tid] = (ln && ok) ? __ldg(P.mean + r) : 0.f;
            s_rs[tid] = (ln && ok) ? __ldg(P.rstd + r) : 0.f;
            s_gi[tid] = (hot0 >= 0 && ok) ? (int)P.gidx[r] : -1;
        }
        __syncthreads();
#pragma unroll 2
        for (int it = tid; it < n_valid * nr8; it += kThreads) {             // A operand = dy^T
            int nn, r8;
            fd_n.divmod(it, r8, nn);
            const float* base = P.dy + (r0 + r8 * 8) * P.N + n0 + nn;
            const int nv = rows_left - r8 * 8;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = q < nv ? __ldg(base + q * P.N) : 0.f;
            store_split(a_hi, a_term_bytes(), op_off(nn, r8 * 8), v);
        }
#pragma unroll 2
        for (int it = tid; it < kcols * nr8; it += kThreads) {               // B operand, feature columns = act(a)^T
            int c, r8;
            fd_k.divmod(it, r8, c);
            const int f = kb + c;
            const int ld = f < P.a.k0 ? P.a.k0 : P.a.k1;
            const float* base = (f < P.a.k0 ? P.a.p0 + f : P.a.p1 + (f - P.a.k0)) + (r0 + r8 * 8) * ld;
            const int nv = rows_left - r8 * 8;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = q < nv ? __ldg(base + q * ld) : 0.f;
            if (ln) {
                const float gm = __ldg(P.gamma + f), bt = __ldg(P.beta + f);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int rr = r8 * 8 + q;
                    const float n = fmaf((v[q] - s_mu[rr]) * s_rs[rr], gm, bt);
                    v[q] = n * sigmoidf_(n);              // rows >= B: finite, and the dy operand is 0 there
                }
            }
            store_split(b_hi, b_term, op_off(c, r8 * 8), v);
        }
        const int nx = bcols_pad - kcols;
        for (int it = tid; it < nx * nr8; it += kThreads) {                  // [1 | onehot(gidx)] and the zero padding
            int c, r8;
            fd_x.divmod(it, r8, c);
            c += kcols;
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int rr = r8 * 8 + q;
                v[q] = (rr < rows_left && (c == one_col || (hot0 >= 0 && s_gi[rr] == c - hot0))) ? 1.f : 0.f;
            }
            store_split(b_hi, b_term, op_off(c, r8 * 8), v);
        }
        publish_and_issue(S, a_hi, b_hi, b_term, kw, idesc, d_tmem, it_no == 0);
    }
    if (it_no > 0) wait_consumed(S, phase);

    // epilogue: TMEM lane n = output feature n0 + n; columns = [dW row | db | d(gadd) column]; 16-column groups
    // alternate between the two sides
    const int n_loc = (warp & 3) * 32 + (tid & 31);
    const uint32_t t_row = d_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const bool has_n = n_loc < n_valid && it_no > 0;
    const int n = n0 + n_loc;
    const bool v4 = kV || P.dwvec != 0;
    for (int g = side; g < bcols_pad / 16; g += 2) {
        float v[16];
        tmem_ld16(t_row + g * 16, v);
        tmem_ld_wait();
        if (!has_n) continue;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            const int c = g * 16 + q4 * 4;
            if (v4 && c + 4 <= kcols) {
                red_add4(P.dw + (size_t)n * P.K + kb + c, v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int cc = c + q;
                    if (cc < kcols) red_add(P.dw + (size_t)n * P.K + kb + cc, v[q4 * 4 + q]);
                    else if (cc == one_col) red_add(P.dbias + n, v[q4 * 4 + q]);
                    else if (hot0 >= 0 && cc >= hot0 && cc < hot0 + P.T) red_add(P.dgadd + (size_t)(cc - hot0) * P.dgadd_ld + n, v[q4 * 4 + q]);
                }
            }
        }
    }
    cta_teardown(S, P.tmem_cols);
}

// ================================================================================================ backward launch
// One launch runs up to two wgrad and two dgrad problems of a node side by side (they only share inputs): the persistent
// wgrad CTAs first, then the dgrad tiles.  Each CTA picks its role from its block index.
struct BwdArgs {
    DgradArgs d[2];
    WgradArgs w[2];
    int nd, nw;
    int d_cta[2], d_gy[2];
    int w_cta[2], w_gx[2], w_gy[2], w_gz[2];
};

template <bool kV>
__global__ void __launch_bounds__(kThreads, 2) tlin_bwd_kernel(const __grid_constant__ BwdArgs P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ Smem S;
    int b = blockIdx.x;
    for (int j = 0; j < P.nw; ++j) {
        if (b < P.w_cta[j]) {
            const int gx = P.w_gx[j], gy = P.w_gy[j];
            wgrad_body<kV>(P.w[j], smem_raw, S, b % gx, (b / gx) % gy, b / (gx * gy), gx, P.w_gz[j]);
            return;
        }
        b -= P.w_cta[j];
    }
    for (int j = 0; j < P.nd; ++j) {
        if (b < P.d_cta[j]) {
            dgrad_body<kV>(P.d[j], smem_raw, S, b / P.d_gy[j], b % P.d_gy[j]);
            return;
        }
        b -= P.d_cta[j];
    }
}

static inline uint32_t pow2_cols(int n) {
    uint32_t c = 32;
    while ((int)c < n) c <<= 1;
    return c;
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline Mat to_mat(const diffsg_mat& m) {
    Mat r;
    r.p0 = m.p0;
    r.p1 = m.k1 > 0 ? m.p1 : nullptr;
    r.k0 = m.k0;
    r.k1 = m.k1 > 0 ? m.k1 : 0;
    r.vec = (r.k0 % 8 == 0) && (r.k1 % 8 == 0) && aligned16(r.p0) && (!r.p1 || aligned16(r.p1));
    return r;
}
static inline bool mat_ok(const diffsg_mat& m) { return m.p0 && m.k0 > 0 && m.k1 >= 0 && (m.k1 == 0 || m.p1); }

}  // namespace ttc
}  // namespace diffsg

using namespace diffsg;
using namespace diffsg::ttc;

// launch with programmatic stream serialization (see pdl_wait); DIFFSG_NO_PDL=1 falls back to a plain launch
template <typename Args>
static cudaError_t launch_pdl(void (*kernel)(Args), dim3 grid, size_t smem, cudaStream_t st, const Args& P) {
    static const bool no_pdl = [] { const char* e = getenv("DIFFSG_NO_PDL"); return e && e[0] == '1'; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, P);
}

// opt-in dynamic shared memory, once per (kernel, device)
static int set_smem(const void* fn, size_t bytes, int which) {
    static bool done[4][64] = {};
    int dev = 0;
    DIFFSG_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[which][dev]) return DIFFSG_OK;
    DIFFSG_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    if (dev >= 0 && dev < 64) done[which][dev] = true;
    return DIFFSG_OK;
}

extern "C" {

int diffsg_tlin_forward(const diffsg_tlin_fwd_args* a, void* stream) {
    if (a && a->B == 0) return DIFFSG_OK;
    if (!a || !mat_ok(a->a) || !a->w || !a->y || a->B < 0 || a->N <= 0) { set_error("tlin_forward: bad argument"); return DIFFSG_E_INVALID; }
    const bool ln = a->gamma != nullptr;
    if (ln && (!a->beta || !a->mean || !a->rstd)) { set_error("tlin_forward: LayerNorm mode needs beta, mean, rstd"); return DIFFSG_E_INVALID; }
    if (ln && a->a.k0 + a->a.k1 > 256) { set_error("tlin_forward: LayerNorm width %d > 256", a->a.k0 + a->a.k1); return DIFFSG_E_UNSUPPORTED; }
    if ((a->gadd != nullptr) != (a->gidx != nullptr)) { set_error("tlin_forward: gadd and gidx go together"); return DIFFSG_E_INVALID; }
    const bool seg2 = a->a2.p0 != nullptr;
    if (seg2 && (!mat_ok(a->a2) || !a->w2)) { set_error("tlin_forward: bad second segment"); return DIFFSG_E_INVALID; }
    FwdArgs P{};
    P.a = to_mat(a->a);
    P.w = a->w; P.bias = a->bias; P.gamma = a->gamma; P.beta = a->beta; P.mean = a->mean; P.rstd = a->rstd;
    if (seg2) P.a2 = to_mat(a->a2); else P.a2 = Mat{nullptr, nullptr, 0, 0, 0};
    P.w2 = a->w2; P.bias2 = seg2 ? a->bias2 : nullptr;
    P.add = a->add; P.gadd = a->gadd; P.gidx = a->gidx; P.y = a->y; P.B = a->B; P.N = a->N;
    P.gadd_ld = a->gadd_ld > 0 ? a->gadd_ld : a->N;
    P.n_pad = a->N >= 128 ? 128 : ((a->N + 15) & ~15);
    P.tmem_cols = (int)pow2_cols(P.n_pad);
    const int K = P.a.k0 + P.a.k1;
    P.wvec = (K % 8 == 0) && aligned16(a->w);
    P.wvec2 = seg2 && ((P.a2.k0 + P.a2.k1) % 8 == 0) && aligned16(a->w2);
    P.yvec = (a->N % 4 == 0) && aligned16(a->y) && aligned16(a->add) && aligned16(a->gadd) && (P.gadd_ld % 4 == 0) && aligned16(a->bias) && aligned16(P.bias2);
    size_t smem = 2 * (size_t)kRows * kKC * 2 + 2 * (size_t)P.n_pad * kKC * 2;
    const size_t tile = (size_t)kRows * (P.n_pad + kTilePad) * sizeof(float);        // epilogue tile shares the operand buffers
    if (tile > smem) smem = tile;
    // fast variant: every operand 8-column / 16-byte aligned (the generic paths compiled out: half the instructions)
    const bool all_vec = P.a.vec && P.wvec && P.yvec && (!seg2 || (P.a2.vec && P.wvec2));
    void (*kern)(FwdArgs) = all_vec ? tlin_fwd_kernel<true> : tlin_fwd_kernel<false>;
    if (int rc = set_smem((const void*)kern, (size_t)kRows * (128 + kTilePad) * sizeof(float), all_vec ? 0 : 1)) return rc;
    const dim3 grid((unsigned)((a->B + kRows - 1) / kRows), (unsigned)((a->N + 127) / 128));
    DIFFSG_CUDA_OK(launch_pdl(kern, grid, smem, (cudaStream_t)stream, P));
    count_launch();
    return DIFFSG_OK;
}

// argument checks + kernel-side structs; `grid` receives the CTA geometry of the problem
static int prep_dgrad(const diffsg_tlin_dgrad_args* a, DgradArgs& P, int& n_tiles, int& gy, size_t& smem) {
    if (!a || !a->dy || !a->w || a->N <= 0 || a->K <= 0 || a->B <= 0 || !a->dx.p0) { set_error("tlin_dgrad: bad argument"); return DIFFSG_E_INVALID; }
    if (a->dx.k0 + (a->dx.k1 > 0 ? a->dx.k1 : 0) != a->K) { set_error("tlin_dgrad: dx split does not add up to K"); return DIFFSG_E_INVALID; }
    const bool ln = a->gamma != nullptr;
    if (ln && (!a->beta || !a->mean || !a->rstd || !a->dgamma || !a->dbeta || !mat_ok(a->x) || a->x.k0 + (a->x.k1 > 0 ? a->x.k1 : 0) != a->K)) {
        set_error("tlin_dgrad: LayerNorm mode needs x[B, K], beta, mean, rstd, dgamma, dbeta");
        return DIFFSG_E_INVALID;
    }
    if (ln && a->K > 256) { set_error("tlin_dgrad: LayerNorm width %d > 256", a->K); return DIFFSG_E_UNSUPPORTED; }
    if (a->dres.p0 && a->dres.k0 + (a->dres.k1 > 0 ? a->dres.k1 : 0) != a->K) { set_error("tlin_dgrad: dres split does not add up to K"); return DIFFSG_E_INVALID; }
    P = DgradArgs{};
    P.dy = a->dy; P.w = a->w; P.gamma = a->gamma; P.beta = a->beta; P.mean = a->mean; P.rstd = a->rstd;
    if (ln) P.x = to_mat(a->x);
    if (a->dres.p0) P.dres = to_mat(a->dres); else P.dres = Mat{nullptr, nullptr, 0, 0, 0};
    diffsg_mat dxm{a->dx.p0, a->dx.p1, a->dx.k0, a->dx.k1};
    const Mat t = to_mat(dxm);
    P.dx = MatOut{a->dx.p0, a->dx.k1 > 0 ? a->dx.p1 : nullptr, t.k0, t.k1, t.vec};
    P.dgamma = a->dgamma; P.dbeta = a->dbeta; P.B = a->B; P.N = a->N; P.K = a->K;
    P.kt = ln ? 256 : 128;
    const int kmax = a->K < P.kt ? a->K : P.kt;
    P.kt_pad = (kmax + 15) & ~15;
    P.tmem_cols = (int)pow2_cols(P.kt_pad);
    P.dyvec = (a->N % 8 == 0) && aligned16(a->dy);
    smem = 2 * (size_t)kRows * kKC * 2 + 2 * (size_t)P.kt_pad * kKC * 2;
    const size_t tile = ln ? 0 : (size_t)kRows * (P.kt_pad + kTilePad) * sizeof(float);
    if (tile > smem) smem = tile;
    n_tiles = (int)((a->B + kRows - 1) / kRows);
    gy = (a->K + P.kt - 1) / P.kt;
    return DIFFSG_OK;
}

static int prep_wgrad(const diffsg_tlin_wgrad_args* a, WgradArgs& P, int& gx, int& gy, int& gz, size_t& smem) {
    if (!a || !a->dy || !mat_ok(a->a) || !a->dw || a->N <= 0 || a->B <= 0) { set_error("tlin_wgrad: bad argument"); return DIFFSG_E_INVALID; }
    const bool ln = a->gamma != nullptr;
    if (ln && (!a->beta || !a->mean || !a->rstd)) { set_error("tlin_wgrad: LayerNorm mode needs beta, mean, rstd"); return DIFFSG_E_INVALID; }
    if (a->dgadd && (!a->gidx || a->gadd_rows <= 0)) { set_error("tlin_wgrad: dgadd needs gidx and gadd_rows"); return DIFFSG_E_INVALID; }
    const int n_extra = (a->dbias ? 1 : 0) + (a->dgadd ? a->gadd_rows : 0);
    if (n_extra > kMaxExtra) { set_error("tlin_wgrad: %d gathered rows > %d", a->gadd_rows, kMaxExtra - 1); return DIFFSG_E_UNSUPPORTED; }
    P = WgradArgs{};
    P.dy = a->dy; P.a = to_mat(a->a); P.gamma = a->gamma; P.beta = a->beta; P.mean = a->mean; P.rstd = a->rstd;
    P.gidx = a->gidx; P.dw = a->dw; P.dbias = a->dbias; P.dgadd = a->dgadd; P.B = a->B; P.N = a->N;
    P.K = P.a.k0 + P.a.k1; P.T = a->dgadd ? a->gadd_rows : 0;
    P.dgadd_ld = a->dgadd_ld > 0 ? a->dgadd_ld : a->N;
    P.n_chunks = (int)((a->B + kKC - 1) / kKC);
    gz = (P.K + kWgradKT - 1) / kWgradKT;
    gy = (a->N + 127) / 128;
    const int k_first = P.K < kWgradKT ? P.K : kWgradKT;
    const int k_last = P.K - (gz - 1) * kWgradKT;
    int widest = gz > 1 ? kWgradKT : k_first;
    if (k_last + n_extra > widest) widest = k_last + n_extra;
    P.bcols_pad = (widest + 15) & ~15;
    P.dwvec = (P.K % 4 == 0) && aligned16(a->dw);
    P.tmem_cols = (int)pow2_cols(P.bcols_pad);
    gx = 296 / (gy * gz);
    if (gx < 1) gx = 1;
    if (gx > P.n_chunks) gx = P.n_chunks;
    smem = 2 * (size_t)kRows * kKC * 2 + 2 * (size_t)P.bcols_pad * kKC * 2;
    return DIFFSG_OK;
}

int diffsg_tlin_backward(const diffsg_tlin_dgrad_args* dgrads, int32_t n_dgrad, const diffsg_tlin_wgrad_args* wgrads,
                         int32_t n_wgrad, void* stream) {
    if (n_dgrad < 0 || n_dgrad > 2 || n_wgrad < 0 || n_wgrad > 2 || (n_dgrad && !dgrads) || (n_wgrad && !wgrads)) {
        set_error("tlin_backward: at most two dgrad and two wgrad problems per launch");
        return DIFFSG_E_INVALID;
    }
    BwdArgs P{};
    size_t smem = 0;
    int total = 0;
    for (int j = 0; j < n_wgrad; ++j) {
        if (wgrads[j].B == 0) continue;
        size_t sm = 0;
        int gx = 0, gy = 0, gz = 0;
        if (int rc = prep_wgrad(&wgrads[j], P.w[P.nw], gx, gy, gz, sm)) return rc;
        P.w_gx[P.nw] = gx; P.w_gy[P.nw] = gy; P.w_gz[P.nw] = gz; P.w_cta[P.nw] = gx * gy * gz;
        total += P.w_cta[P.nw];
        if (sm > smem) smem = sm;
        ++P.nw;
    }
    for (int j = 0; j < n_dgrad; ++j) {
        if (dgrads[j].B == 0) continue;
        size_t sm = 0;
        int nt = 0, gy = 0;
        if (int rc = prep_dgrad(&dgrads[j], P.d[P.nd], nt, gy, sm)) return rc;
        P.d_gy[P.nd] = gy; P.d_cta[P.nd] = nt * gy;
        total += P.d_cta[P.nd];
        if (sm > smem) smem = sm;
        ++P.nd;
    }
    if (total == 0) return DIFFSG_OK;
    bool all_vec = true;
    for (int j = 0; j < P.nw; ++j) all_vec = all_vec && P.w[j].dwvec && P.w[j].a.vec && (P.w[j].N % 8 == 0) && aligned16(P.w[j].dy);
    for (int j = 0; j < P.nd; ++j)
        all_vec = all_vec && P.d[j].dyvec && P.d[j].dx.vec && (!P.d[j].dres.p0 || P.d[j].dres.vec) && (!P.d[j].gamma || P.d[j].x.vec);
    if (int rc = set_smem(all_vec ? (const void*)tlin_bwd_kernel<true> : (const void*)tlin_bwd_kernel<false>,
                          2 * (size_t)kRows * kKC * 2 + 2 * (size_t)256 * kKC * 2, all_vec ? 2 : 3)) return rc;
    DIFFSG_CUDA_OK(launch_pdl(all_vec ? tlin_bwd_kernel<true> : tlin_bwd_kernel<false>, dim3((unsigned)total), smem, (cudaStream_t)stream, P));
    count_launch();
    return DIFFSG_OK;
}

int diffsg_tlin_dgrad(const diffsg_tlin_dgrad_args* a, void* stream) { return diffsg_tlin_backward(a, 1, nullptr, 0, stream); }

int diffsg_tlin_wgrad(const diffsg_tlin_wgrad_args* a, void* stream) { return diffsg_tlin_backward(nullptr, 0, a, 1, stream); }

}  // extern "C"
