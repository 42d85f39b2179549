// HBM-bound side kernels: EMA update, decoder statistics, per-problem decode + objective.
// Reference semantics (reference repo paths):
//   EMA            ddpm_opt/ema.py:3-14
//   MSR decode/obj ddpm_opt/classifier_free_MSR.py:239-245, 284-288
//   NU decode/rate ddpm_opt/classifier_free_NU.py:267-303
//   CO decode/cost ddpm_opt/classifier_free_CO.py:255-290
#include <cfloat>
#include <cstring>
#include "common.cuh"

namespace diffsg {

static inline int blocks_for(int64_t n, int per_block, int cap = 148 * 16) {
    int64_t b = (n + per_block - 1) / per_block;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

// ---------------------------------------------------------------------------------- EMA
__global__ void ema_flat_kernel(float* __restrict__ avg, const float* __restrict__ p, int64_t n, float d,
                                float omd, int copy_first) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n4 = (((uintptr_t)avg | (uintptr_t)p) & 15) == 0 ? n / 4 : 0;
    for (int64_t j = i; j < n4; j += stride) {
        const float4 pv = reinterpret_cast<const float4*>(p)[j];
        float4 av = pv;
        if (!copy_first) {
            av = reinterpret_cast<float4*>(avg)[j];
            av.x = d * av.x + omd * pv.x; av.y = d * av.y + omd * pv.y;
            av.z = d * av.z + omd * pv.z; av.w = d * av.w + omd * pv.w;
        }
        reinterpret_cast<float4*>(avg)[j] = av;
    }
    for (int64_t j = n4 * 4 + i; j < n; j += stride) avg[j] = copy_first ? p[j] : d * avg[j] + omd * p[j];
}

__global__ void ema_multi_kernel(float* const* __restrict__ avgs, const float* const* __restrict__ ps,
                                 const int64_t* __restrict__ sizes, float d, float omd, int copy_first) {
    const int t = blockIdx.y;
    float* avg = avgs[t];
    const float* p = ps[t];
    const int64_t n = sizes[t];
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x)
        avg[j] = copy_first ? p[j] : d * avg[j] + omd * p[j];
}

// ------------------------------------------------------------------------------ Adam (+ EMA)
// One launch over the flat parameter / gradient buffers: torch.optim.Adam's update (no weight decay, no amsgrad;
// reference loop classifier_free_MSR.py:213,225) and, when hyper[5] != 0, the EMA of ddpm_opt/ema.py:10-14 on the
// freshly updated parameters in the same pass (12 + 8 B/param instead of two kernels re-reading p).
// Every hyper-parameter and the step counter live in DEVICE memory so the launch can sit in a CUDA graph:
//   hyper = [lr, beta1, beta2, eps, ema_decay, ema_mode (0 off, 1 copy, 2 blend)]
__global__ void adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, float* __restrict__ ema, int64_t n,
                                 const float* __restrict__ hyper, const int64_t* __restrict__ step_dev) {
    __shared__ float sh[2];
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], d = hyper[4];
    const int ema_mode = ema ? (int)hyper[5] : 0;
    if (threadIdx.x == 0) {
        const double s = (double)(step_dev[0] + 1);
        sh[0] = (float)(1.0 - pow((double)b1, s));            // bias_correction1
        sh[1] = (float)sqrt(1.0 - pow((double)b2, s));        // sqrt(bias_correction2)
    }
    __syncthreads();
    const float step_size = lr / sh[0], rs2 = 1.0f / sh[1];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float gi = g[i];
        const float mi = m[i] + (gi - m[i]) * (1.0f - b1);    // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float pi = p[i] - step_size * (mi / (sqrtf(vi) * rs2 + eps));
        p[i] = pi;
        if (ema_mode == 1) ema[i] = pi;
        else if (ema_mode == 2) ema[i] = d * ema[i] + (1.0f - d) * pi;
    }
}
__global__ void adam_bump_kernel(int64_t* step_dev) { step_dev[0] += 1; }

// ------------------------------------------------------------------------------ batched small MLP
// The comparison baselines of the reference are tiny MLPs evaluated row by row (MTFNN: baselines/MTFNN.py:43-52,
// 122-131, 187-211; the PPO actor / critic: baselines/PPO.py:44-62): Linear -> {ReLU, Tanh} chains with a Sigmoid /
// Softmax / split head.  One thread carries one row through all layers; the weights sit in shared memory and are
// read as warp-wide broadcasts, so the kernel moves in + out floats per row over HBM and nothing else.
constexpr int kMlpMaxLayers = 8, kMlpMaxWidth = 128;
struct MlpDesc {
    int n_layers, in_dim;
    int out[kMlpMaxLayers];      // output width of layer l
    int act[kMlpMaxLayers];      // 0 none, 1 ReLU, 2 Tanh, 3 Sigmoid
    int w_off[kMlpMaxLayers];    // float offsets of W [out][in] and b [out] in the parameter blob
    int b_off[kMlpMaxLayers];
    int head, head_split;        // head: 0 none, 1 softmax over all columns, 2 sigmoid on [0, split) + softmax on [split, N)
    int n_params;
};
__global__ void mlp_forward_kernel(const float* __restrict__ x, const float* __restrict__ params, MlpDesc d,
                                   float* __restrict__ out, int64_t B) {
    extern __shared__ float sp[];
    for (int i = threadIdx.x; i < d.n_params; i += blockDim.x) sp[i] = params[i];
    __syncthreads();
    float a[kMlpMaxWidth], h[kMlpMaxWidth];
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < B; row += (int64_t)gridDim.x * blockDim.x) {
        for (int k = 0; k < d.in_dim; ++k) a[k] = x[row * d.in_dim + k];
        int kin = d.in_dim;
        for (int l = 0; l < d.n_layers; ++l) {
            const float* W = sp + d.w_off[l];
            const float* bias = sp + d.b_off[l];
            const int n_out = d.out[l];
            for (int n = 0; n < n_out; ++n) {
                float acc = bias[n];
                for (int k = 0; k < kin; ++k) acc = fmaf(W[n * kin + k], a[k], acc);
                if (d.act[l] == 1) acc = fmaxf(acc, 0.f);
                else if (d.act[l] == 2) acc = tanhf(acc);
                else if (d.act[l] == 3) acc = 1.0f / (1.0f + expf(-acc));
                h[n] = acc;
            }
            for (int n = 0; n < n_out; ++n) a[n] = h[n];
            kin = n_out;
        }
        const int s0 = d.head == 2 ? d.head_split : 0;
        if (d.head == 2)
            for (int n = 0; n < s0 && n < kin; ++n) a[n] = 1.0f / (1.0f + expf(-a[n]));
        if (d.head != 0 && kin > s0) {
            float m = -FLT_MAX, z = 0.f;
            for (int n = s0; n < kin; ++n) m = fmaxf(m, a[n]);
            for (int n = s0; n < kin; ++n) { a[n] = expf(a[n] - m); z += a[n]; }
            for (int n = s0; n < kin; ++n) a[n] /= z;
        }
        for (int n = 0; n < kin; ++n) out[row * kin + n] = a[n];
    }
}

// ------------------------------------------------------------------------------ min/max
__device__ __forceinline__ void atomic_min_f(float* a, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__global__ void minmax_init_kernel(float* mm) { mm[0] = FLT_MAX; mm[1] = -FLT_MAX; }
__global__ void minmax_kernel(const float* __restrict__ y, int64_t B, int ld, int col0, int width, float* mm) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    const int64_t total = B * width;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / width;
        const int c = (int)(t - r * width);
        const float v = y[r * ld + col0 + c];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    __shared__ float slo[32], shi[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { slo[w] = lo; shi[w] = hi; }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        lo = lane < nw ? slo[lane] : FLT_MAX;
        hi = lane < nw ? shi[lane] : -FLT_MAX;
        lo = warp_min(lo); hi = warp_max(hi);
        if (lane == 0) { atomic_min_f(mm, lo); atomic_max_f(mm + 1, hi); }
    }
}

// ------------------------------------------------------------------------------ softmax
// Row softmax over a sub-warp group of G lanes (G = 2^k <= 32); values strided by G.
template <typename F>
__device__ __forceinline__ float group_reduce(float v, int G, F f) {
    for (int o = G >> 1; o > 0; o >>= 1) v = f(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
static inline int group_size(int M) {
    int g = 1;
    while (g < M && g < 32) g <<= 1;
    return g;
}

// MSR: p = W * softmax((y - mn) / (mx - mn)); rate = sum log2(1 + p * g)
__global__ void objective_msr_kernel(const float* __restrict__ y, const float* __restrict__ g,
                                     const float* __restrict__ mm, float W, float* __restrict__ p_out,
                                     float* __restrict__ rate, int64_t B, int M, int G, int decode) {
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, gl = lane % G, per_warp = 32 / G;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float mn = decode ? mm[0] : 0.f, range = decode ? (mm[1] - mm[0]) : 1.f;
    for (int64_t base = warp_global * per_warp; base < B; base += n_warps * per_warp) {
        const int64_t row = base + sub;
        const bool ok = row < B;
        float r = 0.f;
        if (decode) {
            float vmax = -FLT_MAX;
            if (ok) for (int c = gl; c < M; c += G) vmax = fmaxf(vmax, (y[row * M + c] - mn) / range);
            vmax = group_reduce(vmax, G, [](float a, float b) { return fmaxf(a, b); });
            float s = 0.f;
            if (ok) for (int c = gl; c < M; c += G) s += expf((y[row * M + c] - mn) / range - vmax);
            s = group_reduce(s, G, [](float a, float b) { return a + b; });
            if (ok) for (int c = gl; c < M; c += G) {
                const float p = W * (expf((y[row * M + c] - mn) / range - vmax) / s);
                if (p_out) p_out[row * M + c] = p;
                r += log2f(1.0f + p * g[row * M + c]);
            }
        } else if (ok) {
            for (int c = gl; c < M; c += G) r += log2f(1.0f + y[row * M + c] * g[row * M + c]);
        }
        r = group_reduce(r, G, [](float a, float b) { return a + b; });
        if (ok && gl == 0) rate[row] = r;
    }
}

// NU decode: uav = (y[:, :2] - mn) / (mx - mn) * (width, height); p = P_sum * softmax(y[:, 2:])
constexpr int kMaxUsers = 32;
__global__ void decode_nu_kernel(const float* __restrict__ y, const float* __restrict__ mm, float width,
                                 float height, float P_sum, float* __restrict__ dec, int64_t B, int K) {
    const float mn = mm[0], range = mm[1] - mm[0];
    const int ld = K + 2;
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < B; row += (int64_t)gridDim.x * blockDim.x) {
        const float* yr = y + row * ld;
        float* d = dec + row * ld;
        d[0] = (yr[0] - mn) / range * width;
        d[1] = (yr[1] - mn) / range * height;
        float vmax = -FLT_MAX;
        for (int j = 0; j < K; ++j) vmax = fmaxf(vmax, yr[2 + j]);
        float s = 0.f;
        for (int j = 0; j < K; ++j) s += expf(yr[2 + j] - vmax);
        for (int j = 0; j < K; ++j) d[2 + j] = expf(yr[2 + j] - vmax) / s * P_sum;
    }
}

// NOMA-UAV sum rate with SIC ordered by channel gain (descending; ties by index).
__global__ void rate_nu_kernel(const float* __restrict__ dec, const float* __restrict__ xy,
                               float* __restrict__ rate, int64_t B, int K) {
    const float sigma_sq = 110.f, rou_0 = 60.f, H2 = 150.f * 150.f;
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < B; row += (int64_t)gridDim.x * blockDim.x) {
        const float* d = dec + row * (K + 2);
        const float* x = xy + row * 2 * K;
        float h[kMaxUsers];
        for (int j = 0; j < K; ++j) {
            const float dx = x[2 * j] - d[0], dy = x[2 * j + 1] - d[1];
            h[j] = sqrtf(rou_0 / (H2 + dx * dx + dy * dy));
        }
        float r = 0.f;
        for (int j = 0; j < K; ++j) {
            float stronger = 0.f;
            int rank = 0;
            for (int i = 0; i < K; ++i)
                if (h[i] > h[j] || (h[i] == h[j] && i < j)) { stronger += d[2 + i]; ++rank; }
            const float h2 = h[j] * h[j];
            const float sinr = rank == 0 ? d[2 + j] * h2 / sigma_sq : d[2 + j] / (stronger + sigma_sq / h2);
            r += log2f(1.0f + sinr);
        }
        rate[row] = r;
    }
}

// CO decode: softmax(y), zeroed when every logit < -10
constexpr int kMaxNodes = 32;
__global__ void decode_co_kernel(const float* __restrict__ y, float* __restrict__ dec, int64_t B, int n) {
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < B; row += (int64_t)gridDim.x * blockDim.x) {
        const float* yr = y + row * n;
        float vmax = -FLT_MAX;
        bool all_low = true;
        for (int j = 0; j < n; ++j) { vmax = fmaxf(vmax, yr[j]); all_low = all_low && (yr[j] < -10.f); }
        float s = 0.f;
        for (int j = 0; j < n; ++j) s += expf(yr[j] - vmax);
        for (int j = 0; j < n; ++j) dec[row * n + j] = all_low ? 0.f : expf(yr[j] - vmax) / s;
    }
}

// CO cost: threshold decisions at 0.1, renormalise the offloaded shares, local vs offload cost
__global__ void cost_co_kernel(const float* __restrict__ x, const float* __restrict__ alloc,
                               float* __restrict__ cost, int64_t B, int n) {
    for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < B; row += (int64_t)gridDim.x * blockDim.x) {
        const float* xr = x + row * 3 * n;
        const float* a = alloc + row * n;
        float ysum = 0.f;
        int dsum = 0;
        for (int j = 0; j < n; ++j)
            if (a[j] > 0.1f) { ysum += a[j]; ++dsum; }
        const float dden = dsum == 0 ? 0.00001f : (float)dsum;
        const float diff = (1.0f - ysum) / dden;
        float c = 0.f;
        for (int j = 0; j < n; ++j) {
            if (a[j] > 0.1f) c += xr[3 * j + 1] + xr[3 * j + 2] / (a[j] + diff);
            else c += xr[3 * j];
        }
        cost[row] = c;
    }
}

}  // namespace diffsg

using namespace diffsg;

extern "C" {

int diffsg_ema_update(float* avg, const float* p, int64_t n, double decay, int32_t copy_first, void* stream) {
    if (!avg || !p || n < 0) { set_error("ema_update: bad argument"); return DIFFSG_E_INVALID; }
    if (n == 0) return DIFFSG_OK;
    ema_flat_kernel<<<blocks_for(n, 1024), 256, 0, (cudaStream_t)stream>>>(avg, p, n, (float)decay,
                                                                            (float)(1.0 - decay), copy_first);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_ema_update_multi(float* const* avgs, const float* const* ps, const int64_t* sizes,
                            int32_t n_tensors, int64_t max_size, double decay, int32_t copy_first,
                            void* stream) {
    if (!avgs || !ps || !sizes || n_tensors < 0 || n_tensors > 65535) { set_error("ema_update_multi: bad argument"); return DIFFSG_E_INVALID; }
    if (n_tensors == 0) return DIFFSG_OK;
    dim3 grid(blocks_for(max_size, 1024, 64), n_tensors);
    ema_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(avgs, ps, sizes, (float)decay,
                                                             (float)(1.0 - decay), copy_first);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_adam_step(float* p, const float* g, float* m, float* v, float* ema, int64_t n, const float* hyper,
                     int64_t* step_dev, void* stream) {
    if (!p || !g || !m || !v || !hyper || !step_dev || n < 0) { set_error("adam_step: bad argument"); return DIFFSG_E_INVALID; }
    if (n == 0) return DIFFSG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    adam_flat_kernel<<<blocks_for(n, 1024), 256, 0, st>>>(p, g, m, v, ema, n, hyper, step_dev);
    adam_bump_kernel<<<1, 1, 0, st>>>(step_dev);
    count_launch(2);
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_mlp_forward(const float* x, const float* params, int64_t B, int32_t in_dim, int32_t n_layers,
                       const int32_t* out_dims, const int32_t* acts, int32_t head, int32_t head_split, float* out,
                       void* stream) {
    if (!x || !params || !out || !out_dims || !acts || B < 0 || in_dim < 1 || in_dim > kMlpMaxWidth || n_layers < 1 ||
        n_layers > kMlpMaxLayers || head < 0 || head > 2 || head_split < 0) { set_error("mlp_forward: bad argument"); return DIFFSG_E_INVALID; }
    if (B == 0) return DIFFSG_OK;
    MlpDesc d;
    memset(&d, 0, sizeof(d));
    d.n_layers = n_layers; d.in_dim = in_dim; d.head = head; d.head_split = head_split;
    int off = 0, kin = in_dim;
    for (int l = 0; l < n_layers; ++l) {
        if (out_dims[l] < 1 || out_dims[l] > kMlpMaxWidth || acts[l] < 0 || acts[l] > 3) { set_error("mlp_forward: layer %d out of range", l); return DIFFSG_E_UNSUPPORTED; }
        d.out[l] = out_dims[l]; d.act[l] = acts[l];
        d.w_off[l] = off; off += out_dims[l] * kin;      // the blob is W_0 | b_0 | W_1 | b_1 | ... (nn.Linear layouts)
        d.b_off[l] = off; off += out_dims[l];
        kin = out_dims[l];
    }
    d.n_params = off;
    const size_t smem = sizeof(float) * (size_t)off;
    if (smem > 200 * 1024) { set_error("mlp_forward: %zu bytes of weights exceed shared memory", smem); return DIFFSG_E_UNSUPPORTED; }
    DIFFSG_CUDA_OK(cudaFuncSetAttribute(mlp_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mlp_forward_kernel<<<blocks_for(B, 128, 148 * 8), 128, smem, (cudaStream_t)stream>>>(x, params, d, out, B);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_minmax(const float* y, int64_t B, int32_t ld, int32_t col0, int32_t width, float* mm, void* stream) {
    if (!y || !mm || B <= 0 || width <= 0 || col0 < 0 || col0 + width > ld) { set_error("minmax: bad argument"); return DIFFSG_E_INVALID; }
    cudaStream_t st = (cudaStream_t)stream;
    minmax_init_kernel<<<1, 1, 0, st>>>(mm);
    minmax_kernel<<<blocks_for(B * width, 2048, 148 * 4), 256, 0, st>>>(y, B, ld, col0, width, mm);
    count_launch(2);
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_objective_msr(const float* y, const float* g, const float* mm, float W, float* p_out, float* rate,
                         int64_t B, int32_t M, void* stream) {
    if (!y || !g || !mm || !rate || B <= 0 || M <= 0) { set_error("objective_msr: bad argument"); return DIFFSG_E_INVALID; }
    const int G = group_size(M);
    objective_msr_kernel<<<blocks_for(B, 256 / G * 4), 256, 0, (cudaStream_t)stream>>>(y, g, mm, W, p_out, rate, B, M, G, 1);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_rate_msr(const float* p, const float* g, float* rate, int64_t B, int32_t M, void* stream) {
    if (!p || !g || !rate || B <= 0 || M <= 0) { set_error("rate_msr: bad argument"); return DIFFSG_E_INVALID; }
    const int G = group_size(M);
    objective_msr_kernel<<<blocks_for(B, 256 / G * 4), 256, 0, (cudaStream_t)stream>>>(p, g, nullptr, 1.f, nullptr, rate, B, M, G, 0);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_decode_nu(const float* y, const float* mm, float width, float height, float P_sum, float* dec,
                     int64_t B, int32_t K, void* stream) {
    if (!y || !mm || !dec || B <= 0 || K <= 0) { set_error("decode_nu: bad argument"); return DIFFSG_E_INVALID; }
    decode_nu_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(y, mm, width, height, P_sum, dec, B, K);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_rate_nu(const float* dec, const float* xy, float* rate, int64_t B, int32_t K, void* stream) {
    if (!dec || !xy || !rate || B <= 0 || K <= 0 || K > kMaxUsers) { set_error("rate_nu: bad argument (K <= %d)", kMaxUsers); return DIFFSG_E_INVALID; }
    rate_nu_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(dec, xy, rate, B, K);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_decode_co(const float* y, float* dec, int64_t B, int32_t n, void* stream) {
    if (!y || !dec || B <= 0 || n <= 0) { set_error("decode_co: bad argument"); return DIFFSG_E_INVALID; }
    decode_co_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(y, dec, B, n);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_cost_co(const float* x, const float* alloc, float* cost, int64_t B, int32_t n, void* stream) {
    if (!x || !alloc || !cost || B <= 0 || n <= 0) { set_error("cost_co: bad argument"); return DIFFSG_E_INVALID; }
    cost_co_kernel<<<blocks_for(B, 256), 256, 0, (cudaStream_t)stream>>>(x, alloc, cost, B, n);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------- LayerNorm + Swish (training)
// y = swish(LN(x) * gamma + beta) row-wise; one warp per row.  Forward keeps (mean, rstd) for the
// backward pass.  Reference: ddpm_opt/UNetCF.py:90,92,94,356 (nn.LayerNorm -> Swish).
namespace diffsg {

__global__ void lnsw_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean,
                                float* __restrict__ rstd, int64_t B, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < B; r += nw) {
        const float* xr = x + r * D;
        float s = 0.f;
        for (int c = lane; c < D; c += 32) s += xr[c];
        const float m = warp_sum(s) / (float)D;
        float q = 0.f;
        for (int c = lane; c < D; c += 32) { const float d = xr[c] - m; q = fmaf(d, d, q); }
        const float rs = 1.0f / sqrtf(warp_sum(q) / (float)D + kLnEps);
        for (int c = lane; c < D; c += 32) {
            const float t = (xr[c] - m) * rs * gamma[c] + beta[c];
            y[r * D + c] = swish_exact(t);
        }
        if (lane == 0) { mean[r] = m; rstd[r] = rs; }
    }
}

// dx, and per-block partial dgamma / dbeta (reduced by lnsw_bwd_reduce_kernel)
__global__ void lnsw_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                const float* __restrict__ beta, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ dy,
                                float* __restrict__ dx, float* __restrict__ dgamma_part,
                                float* __restrict__ dbeta_part, int64_t B, int D) {
    extern __shared__ float sh[];          // [2][D] per-block accumulators
    float* sg = sh;
    float* sb = sh + D;
    for (int c = threadIdx.x; c < 2 * D; c += blockDim.x) sh[c] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < B; r += nw) {
        const float m = mean[r], rs = rstd[r];
        const float* xr = x + r * D;
        const float* dyr = dy + r * D;
        // dt = dy * swish'(t);  dxhat = dt * gamma;  dx = rs * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < D; c += 32) {
            const float xh = (xr[c] - m) * rs;
            const float t = xh * gamma[c] + beta[c];
            const float sg_ = 1.0f / (1.0f + expf(-t));
            const float dt = dyr[c] * (sg_ * (1.0f + t * (1.0f - sg_)));
            const float dxh = dt * gamma[c];
            s1 += dxh;
            s2 = fmaf(dxh, xh, s2);
            atomicAdd(&sg[c], dt * xh);
            atomicAdd(&sb[c], dt);
        }
        s1 = warp_sum(s1) / (float)D;
        s2 = warp_sum(s2) / (float)D;
        for (int c = lane; c < D; c += 32) {
            const float xh = (xr[c] - m) * rs;
            const float t = xh * gamma[c] + beta[c];
            const float sg_ = 1.0f / (1.0f + expf(-t));
            const float dxh = dyr[c] * (sg_ * (1.0f + t * (1.0f - sg_))) * gamma[c];
            dx[r * D + c] = rs * (dxh - s1 - xh * s2);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        dgamma_part[(size_t)blockIdx.x * D + c] = sg[c];
        dbeta_part[(size_t)blockIdx.x * D + c] = sb[c];
    }
}

__global__ void lnsw_bwd_reduce_kernel(const float* __restrict__ part_g, const float* __restrict__ part_b,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int nblocks, int D) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= D) return;
    float g = 0.f, b = 0.f;
    for (int k = 0; k < nblocks; ++k) { g += part_g[(size_t)k * D + c]; b += part_b[(size_t)k * D + c]; }
    dgamma[c] = g;
    dbeta[c] = b;
}

}  // namespace diffsg

extern "C" {

int diffsg_lnsw_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                        int64_t B, int32_t D, void* stream) {
    if (!x || !gamma || !beta || !y || !mean || !rstd || B <= 0 || D <= 0) { set_error("lnsw_forward: bad argument"); return DIFFSG_E_INVALID; }
    lnsw_fwd_kernel<<<blocks_for(B, 8), 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, y, mean, rstd, B, D);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_lnsw_backward(const float* x, const float* gamma, const float* beta, const float* mean, const float* rstd,
                         const float* dy, float* dx, float* dgamma, float* dbeta, float* workspace,
                         int64_t workspace_floats, int64_t B, int32_t D, void* stream) {
    if (!x || !gamma || !beta || !mean || !rstd || !dy || !dx || !dgamma || !dbeta || !workspace || B <= 0 || D <= 0 || D > 4096) {
        set_error("lnsw_backward: bad argument");
        return DIFFSG_E_INVALID;
    }
    int nb = blocks_for(B, 8, 148 * 2);
    if ((int64_t)nb * 2 * D > workspace_floats) nb = (int)(workspace_floats / (2 * (int64_t)D));
    if (nb < 1) { set_error("lnsw_backward: workspace too small (needs >= %d floats)", 2 * D); return DIFFSG_E_INVALID; }
    float* pg = workspace;
    float* pb = workspace + (size_t)nb * D;
    cudaStream_t st = (cudaStream_t)stream;
    lnsw_bwd_kernel<<<nb, 256, 2 * D * sizeof(float), st>>>(x, gamma, beta, mean, rstd, dy, dx, pg, pb, B, D);
    lnsw_bwd_reduce_kernel<<<(D + 127) / 128, 128, 0, st>>>(pg, pb, dgamma, dbeta, nb, D);
    count_launch(2);
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

}  // extern "C"
