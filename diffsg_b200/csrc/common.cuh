// Shared helpers for the diffsg_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/diffsg_b200.h"

namespace diffsg {

constexpr int kRowsPerWarp = 8;     // rows a warp carries through the whole network
constexpr int kMaxWarps = 16;       // warps per CTA (upper bound; runtime picks <= this)
constexpr int kMaxWidth = 256;      // widest per-row vector (cat(x, skip) of the top level)
constexpr int kMaxSkip = 64;        // skip slots
constexpr int kMaxOps = 1024;
constexpr float kLnEps = 1e-5f;

// thread-local error string + launch counter (host side, diffsg.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define DIFFSG_CUDA_OK(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            diffsg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                              __FILE__, __LINE__);                                        \
            return DIFFSG_E_CUDA;                                                         \
        }                                                                                 \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// x * sigmoid(x), fp32-accurate (reference UNetCF.py:6-14)
__device__ __forceinline__ float swish_exact(float x) { return x / (1.0f + expf(-x)); }

// ---- Philox4x32-10 (Salmon et al. 2011) ------------------------------------------------
struct Philox4 {
    uint32_t v[4];
};
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

// Four standard normals for (row, step, column quad q) — the sampler's noise stream.
// counter = (row_lo, row_hi, step, q), key = (seed_lo, seed_hi); Box-Muller on 24-bit uniforms.
__device__ __forceinline__ void philox_normal4(uint64_t row, uint32_t step, uint32_t q,
                                               uint64_t seed, float out[4]) {
    const Philox4 r = philox4x32_10((uint32_t)row, (uint32_t)(row >> 32), step, q,
                                    (uint32_t)seed, (uint32_t)(seed >> 32));
    const float k2pi = 6.283185307179586f, s = 5.9604644775390625e-8f;  // 2^-24
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u0 = (float)((r.v[2 * h] >> 8) + 1u) * s;       // (0, 1]
        const float u1 = (float)(r.v[2 * h + 1] >> 8) * s;          // [0, 1)
        // MUFU-based log / sincos (abs error ~1e-6): angle folded to [-pi, pi), where the fast
        // sin/cos are most accurate; cos(t + pi) = -cos(t), sin(t + pi) = -sin(t)
        const float rad = sqrtf(-2.0f * __logf(u0));
        float sn, cs;
        __sincosf(k2pi * (u1 - 0.5f), &sn, &cs);
        out[2 * h] = -rad * cs;
        out[2 * h + 1] = -rad * sn;
    }
}

}  // namespace diffsg
