// fp32 "warp-row" evaluation of the UNet1D op program.
//
// One warp carries kRowsPerWarp rows through the whole network: lanes span output
// columns, every weight fetched from L2/L1 is reused for all 8 rows in registers, the
// per-row vectors live in a private shared-memory slab, the skip stack in an L2-resident
// global scratch slab.  Warps never synchronise with each other.  This is the exact-fp32
// engine (parity mode, generic shapes); the tcgen05 engine (unet_tc.cuh) is the fast one.
//
// Semantics: reference ddpm_opt/UNetCF.py:83-95 (ResidualBlock), :318-356 (UNet1D.forward).
#pragma once
#include "common.cuh"

namespace diffsg {

struct PlanDev {
    const diffsg_op* ops;
    const float* params;
    const float* tt;          // [tt_rows][tt_stride]
    float* scratch;           // [n_warps_total][kRowsPerWarp * scratch_floats]
    int n_ops, tt_stride;
    int M, C;
    int buf_off[DIFFSG_N_BUF + 1];  // float offset of each buffer inside a warp slab
    int buf_ld[DIFFSG_N_BUF + 1];   // row stride (floats, multiple of 4)
    int slab_floats;                // floats per warp slab
    int skip_off[kMaxSkip];         // per-row float offset of each skip slot
    int stash_off;                  // per-row float offset of the eps_0 stash
    int scratch_floats;             // per-row floats of global scratch
    int in_buf, out_buf;
};

template <int VEC>
__device__ __forceinline__ void load_w(const float* p, float (&w)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
    } else if constexpr (VEC == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(p));
        w[0] = t.x; w[1] = t.y;
    } else {
        w[0] = __ldg(p);
    }
}

// One column tile (32*VEC wide) of dst = (acc? dst : 0) + bias + time + src . W
template <int VEC>
__device__ __forceinline__ void gemm_tile(const PlanDev& P, const diffsg_op& op, const float* src,
                                          int lds, float* dst, int ldd, int n0, int trow, int lane) {
    constexpr int R = kRowsPerWarp;
    const int c0 = n0 + lane * VEC;
    const bool active = c0 < op.N;   // VEC > 1 tiles are always full
    const float* W = P.params + op.w_off + c0;
    const int ldw = op.ldw;
    float acc[R][VEC];

    float b[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) b[v] = 0.f;
    if (active && !(op.flags & DIFFSG_F_NOBIAS)) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) b[v] = __ldg(P.params + op.b_off + c0 + v);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int tr = __shfl_sync(0xffffffffu, trow, r);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float a = b[v];
            if (active) {
                if (op.flags & DIFFSG_F_TIME) a += __ldg(P.tt + (size_t)tr * P.tt_stride + op.t_off + c0 + v);
                if (op.flags & DIFFSG_F_ACC) a += dst[r * ldd + c0 + v];
            }
            acc[r][v] = a;
        }
    }

    const int K = op.K;
    int k = 0;
#pragma unroll 2
    for (; k + 4 <= K; k += 4) {
        float4 a[R];
#pragma unroll
        for (int r = 0; r < R; ++r) a[r] = *reinterpret_cast<const float4*>(src + r * lds + k);
        float w[4][VEC];
        if (active) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) load_w<VEC>(W + (size_t)(k + kk) * ldw, w[kk]);
        } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int v = 0; v < VEC; ++v) w[kk][v] = 0.f;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                acc[r][v] = fmaf(a[r].x, w[0][v], acc[r][v]);
                acc[r][v] = fmaf(a[r].y, w[1][v], acc[r][v]);
                acc[r][v] = fmaf(a[r].z, w[2][v], acc[r][v]);
                acc[r][v] = fmaf(a[r].w, w[3][v], acc[r][v]);
            }
    }
    for (; k < K; ++k) {
        float w[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) w[v] = 0.f;
        if (active) load_w<VEC>(W + (size_t)k * ldw, w);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float a = src[r * lds + k];
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[r][v] = fmaf(a, w[v], acc[r][v]);
        }
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int v = 0; v < VEC; ++v) dst[r * ldd + c0 + v] = acc[r][v];
    }
}

__device__ __forceinline__ void gemm_op(const PlanDev& P, const diffsg_op& op, float* slab, int trow,
                                        int lane) {
    const float* src = slab + P.buf_off[op.src];
    const int lds = P.buf_ld[op.src];
    float* dst = slab + P.buf_off[op.dst] + op.dcol;
    const int ldd = P.buf_ld[op.dst];
    int n0 = 0;
    while (n0 < op.N) {
        const int rem = op.N - n0;
        if (rem >= 128) { gemm_tile<4>(P, op, src, lds, dst, ldd, n0, trow, lane); n0 += 128; }
        else if (rem >= 64) { gemm_tile<2>(P, op, src, lds, dst, ldd, n0, trow, lane); n0 += 64; }
        else { gemm_tile<1>(P, op, src, lds, dst, ldd, n0, trow, lane); n0 += 32; }
    }
}

// dst = swish(LayerNorm(src)) row-wise, two-pass moments in fp32 (torch semantics, eps 1e-5)
__device__ __forceinline__ void lnsw_op(const PlanDev& P, const diffsg_op& op, float* slab, int lane) {
    constexpr int R = kRowsPerWarp;
    const float* src = slab + P.buf_off[op.src];
    const int lds = P.buf_ld[op.src];
    float* dst = slab + P.buf_off[op.dst];
    const int ldd = P.buf_ld[op.dst];
    const int D = op.N;
    const float inv_d = 1.0f / (float)D;
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = 0.f;
    for (int c = lane; c < D; c += 32)
#pragma unroll
        for (int r = 0; r < R; ++r) mean[r] += src[r * lds + c];
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = warp_sum(mean[r]) * inv_d;
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = 0.f;
    for (int c = lane; c < D; c += 32)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float d = src[r * lds + c] - mean[r];
            rstd[r] = fmaf(d, d, rstd[r]);
        }
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = 1.0f / sqrtf(warp_sum(rstd[r]) * inv_d + kLnEps);
    const float* gamma = P.params + op.w_off;
    const float* beta = P.params + op.b_off;
    for (int c = lane; c < D; c += 32) {
        const float g = __ldg(gamma + c), b = __ldg(beta + c);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float v = (src[r * lds + c] - mean[r]) * rstd[r] * g + b;
            dst[r * ldd + c] = swish_exact(v);
        }
    }
}

// Run the whole program for the 8 rows of this warp.  `trow`: lane r (< 8) holds the
// time-table row of row r.  `use_cond` = false skips the condition GEMMs (uncond pass:
// swish(0) = 0, so only the cond_emb bias — folded into lin2's bias — remains).
__device__ __forceinline__ void run_program(const PlanDev& P, float* slab, float* gscr, int trow,
                                            bool use_cond, int lane) {
    constexpr int R = kRowsPerWarp;
    for (int i = 0; i < P.n_ops; ++i) {
        const diffsg_op op = P.ops[i];
        switch (op.kind) {
            case DIFFSG_OP_GEMM:
                if (op.src == DIFFSG_BUF_COND && !use_cond) break;
                gemm_op(P, op, slab, trow, lane);
                break;
            case DIFFSG_OP_LNSW:
                lnsw_op(P, op, slab, lane);
                break;
            case DIFFSG_OP_PUSH: {
                const float* src = slab + P.buf_off[op.src];
                const int lds = P.buf_ld[op.src];
                float* g = gscr + (size_t)P.skip_off[op.dcol] * R;
                for (int c = lane; c < op.N; c += 32)
#pragma unroll
                    for (int r = 0; r < R; ++r) g[r * op.N + c] = src[r * lds + c];
                break;
            }
            case DIFFSG_OP_POP: {
                float* dst = slab + P.buf_off[op.dst] + op.dcol;
                const int ldd = P.buf_ld[op.dst];
                const float* g = gscr + (size_t)P.skip_off[op.K] * R;
                for (int c = lane; c < op.N; c += 32)
#pragma unroll
                    for (int r = 0; r < R; ++r) dst[r * ldd + c] = g[r * op.N + c];
                break;
            }
            default:
                break;
        }
        __syncwarp();
    }
}

// Load `n` columns of up to 8 rows (global, row stride ld) into a slab buffer; rows past
// `nrows` are zero-filled.  `scale`-free; optional swish and per-row mask for the cond slab.
__device__ __forceinline__ void load_rows(const float* g, int64_t row0, int nrows, int n, int ldg_,
                                          float* dst, int ldd, int lane) {
    constexpr int R = kRowsPerWarp;
    for (int c = lane; c < n; c += 32)
#pragma unroll
        for (int r = 0; r < R; ++r)
            dst[r * ldd + c] = (r < nrows) ? g[(row0 + r) * (int64_t)ldg_ + c] : 0.f;
}

}  // namespace diffsg
