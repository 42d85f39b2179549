// diffsg_b200 C-ABI: plan management, fp32 warp-row UNet forward, CFG sampler driver.
// Interfaces replaced (reference repo): UNet1D.forward ddpm_opt/UNetCF.py:318-356,
// DDPM.sample ddpm_opt/classifier_free_MSR.py:114-155 (== _NU.py:143-180, _CO.py:117-154).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "unet_simt.cuh"
#include "plan.h"

namespace diffsg {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs backward kernels on its own thread

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ----------------------------------------------------------------------------- kernels
extern __shared__ __align__(16) float g_smem[];

// eps = UNet(x, t_idx, cond * mask)       (generic forward; one warp per 8 rows)
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
unet_forward_simt_kernel(PlanDev P, const float* __restrict__ x, const int32_t* __restrict__ t_idx,
                         const float* __restrict__ cond, const float* __restrict__ mask,
                         float* __restrict__ eps, int64_t B) {
    constexpr int R = kRowsPerWarp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float* slab = g_smem + (size_t)warp * P.slab_floats;
    const int64_t wid = (int64_t)warp * gridDim.x + blockIdx.x;
    float* gscr = P.scratch + (size_t)wid * R * P.scratch_floats;
    const int64_t n_groups = (B + R - 1) / R, stride = (int64_t)nwarp * gridDim.x;
    float* cbuf = slab + P.buf_off[DIFFSG_BUF_COND];
    const int ldc = P.buf_ld[DIFFSG_BUF_COND];
    float* inb = slab + P.buf_off[P.in_buf];
    const float* outb = slab + P.buf_off[P.out_buf];
    for (int64_t g = wid; g < n_groups; g += stride) {
        const int64_t row0 = g * R;
        const int nrows = (int)min((int64_t)R, B - row0);
        load_rows(x, row0, nrows, P.M, P.M, inb, P.buf_ld[P.in_buf], lane);
        for (int c = lane; c < P.C; c += 32)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float v = 0.f;
                if (r < nrows) {
                    v = cond[(row0 + r) * (int64_t)P.C + c];
                    if (mask) v *= mask[row0 + r];
                    v = swish_exact(v);
                }
                cbuf[r * ldc + c] = v;
            }
        int trow = 0;
        if (lane < nrows) trow = t_idx[row0 + lane];
        __syncwarp();
        run_program(P, slab, gscr, trow, true, lane);
        for (int c = lane; c < P.M; c += 32)
            for (int r = 0; r < nrows; ++r) eps[(row0 + r) * (int64_t)P.M + c] = outb[r * P.buf_ld[P.out_buf] + c];
        __syncwarp();
    }
}

struct SampleDev {
    const float* cond;
    float* y;
    const float* noise;      // [(T-2)][B][M] or null
    float* rec_y;            // [T][B][M] or null
    float* rec_eps;          // [T][B][M] or null
    double* stats;           // [T][2] (sum, sumsq)
    int64_t B;
    int T, step_hi, step_lo; // this launch runs steps step_hi .. step_lo (descending)
    int norm_steps;
    float omega;
    uint64_t seed, offset;
    float c_eps[64], c_rs[64], c_noise[64];
};

// Steps step_hi..step_lo of the reverse process for every row group (both CFG passes,
// guidance mix, posterior update, noise add, partial sums for the batch re-normalisation).
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
sample_simt_kernel(PlanDev P, SampleDev S) {
    constexpr int R = kRowsPerWarp;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    float* slab = g_smem + (size_t)warp * P.slab_floats;
    const int64_t wid = (int64_t)warp * gridDim.x + blockIdx.x;
    float* gscr = P.scratch + (size_t)wid * R * P.scratch_floats;
    float* stash = gscr + (size_t)P.stash_off * R;
    const int64_t n_groups = (S.B + R - 1) / R, stride = (int64_t)nwarp * gridDim.x;
    float* cbuf = slab + P.buf_off[DIFFSG_BUF_COND];
    const int ldc = P.buf_ld[DIFFSG_BUF_COND];
    float* inb = slab + P.buf_off[P.in_buf];
    const int ldi = P.buf_ld[P.in_buf];
    const float* outb = slab + P.buf_off[P.out_buf];
    const int ldo = P.buf_ld[P.out_buf];
    const int M = P.M, nq = (M + 3) >> 2;
    const int64_t plane = S.B * (int64_t)M;
    const float w1 = 1.0f + S.omega, w0 = S.omega;

    double acc_s = 0.0, acc_q = 0.0;   // this warp's partial sums for the (single) normalised step
    for (int64_t g = wid; g < n_groups; g += stride) {
        const int64_t row0 = g * R;
        const int nrows = (int)min((int64_t)R, S.B - row0);
        for (int c = lane; c < P.C; c += 32)
#pragma unroll
            for (int r = 0; r < R; ++r)
                cbuf[r * ldc + c] = (r < nrows) ? swish_exact(S.cond[(row0 + r) * (int64_t)P.C + c]) : 0.f;
        for (int i = S.step_hi; i >= S.step_lo; --i) {
            // ---- unconditional pass (cond_mask = 0) -> eps_0
            load_rows(S.y, row0, nrows, M, M, inb, ldi, lane);
            __syncwarp();
            run_program(P, slab, gscr, i, false, lane);
            for (int q = lane; q < nq; q += 32)
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int c = 4 * q + v;
                        if (c < M) stash[r * M + c] = outb[r * ldo + c];
                    }
            __syncwarp();
            // ---- conditional pass (cond_mask = 1) -> eps_1
            load_rows(S.y, row0, nrows, M, M, inb, ldi, lane);
            __syncwarp();
            run_program(P, slab, gscr, i, true, lane);
            // ---- guidance mix + posterior update (reference classifier_free_MSR.py:132-134)
            const float ce = S.c_eps[i], crs = S.c_rs[i], cn = S.c_noise[i];
            const bool add_noise = i > 1;
            const bool want_stats = i > S.T - 1 - S.norm_steps;
            for (int q = lane; q < nq; q += 32)
                for (int r = 0; r < nrows; ++r) {
                    const int64_t row = row0 + r;
                    float z[4] = {0.f, 0.f, 0.f, 0.f};
                    if (add_noise && S.noise == nullptr)
                        philox_normal4((uint64_t)row + S.offset, (uint32_t)i, (uint32_t)q, S.seed, z);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int c = 4 * q + v;
                        if (c >= M) continue;
                        const int64_t idx = row * M + c;
                        if (add_noise && S.noise != nullptr)
                            z[v] = S.noise[(int64_t)(S.T - 1 - i) * plane + idx];
                        const float e = w1 * outb[r * ldo + c] - w0 * stash[r * M + c];
                        float yn = (S.y[idx] - ce * e) * crs;
                        if (add_noise) yn += cn * z[v];
                        S.y[idx] = yn;
                        if (S.rec_eps) S.rec_eps[(int64_t)(S.T - 1 - i) * plane + idx] = e;
                        if (S.rec_y && !want_stats) S.rec_y[(int64_t)(S.T - 1 - i) * plane + idx] = yn;
                        if (want_stats) { acc_s += (double)yn; acc_q += (double)yn * (double)yn; }
                    }
                }
            __syncwarp();
        }
    }
    // a launch that contains a normalised step contains only that step (host guarantees)
    if (S.step_hi > S.T - 1 - S.norm_steps) {
        acc_s = warp_sum(acc_s);
        acc_q = warp_sum(acc_q);
        if (lane == 0) {
            atomicAdd(S.stats + 2 * S.step_hi, acc_s);
            atomicAdd(S.stats + 2 * S.step_hi + 1, acc_q);
        }
    }
}

// y = (y - mean) / sqrt(var_unbiased) with the scalars of one step (reference :136-137)
__global__ void renorm_kernel(float* __restrict__ y, float* __restrict__ rec, const double* __restrict__ st,
                              int64_t n, int64_t n_stat) {
    const double s = st[0], q = st[1];
    const double mean = s / (double)n_stat;
    const double var = (q - s * mean) / (double)(n_stat - 1);
    const float mf = (float)mean, sd = sqrtf((float)var);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = (y[i] - mf) / sd;
        y[i] = v;
        if (rec) rec[i] = v;
    }
}

__global__ void philox_fill_kernel(float* __restrict__ out, int64_t B, int M, uint32_t step, uint64_t seed,
                                   uint64_t offset) {
    const int nq = (M + 3) >> 2;
    const int64_t total = B * nq;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / nq;
        const int q = (int)(t - row * nq);
        float z[4];
        philox_normal4((uint64_t)row + offset, step, (uint32_t)q, seed, z);
#pragma unroll
        for (int v = 0; v < 4; ++v)
            if (4 * q + v < M) out[row * M + 4 * q + v] = z[v];
    }
}

// ----------------------------------------------------------------------------- plan
static int validate_program(const diffsg_cfg& c, const diffsg_op* ops, int n_ops) {
    auto bad = [&](int i, const char* why) {
        set_error("op %d: %s", i, why);
        return DIFFSG_E_INVALID;
    };
    for (int i = 0; i < n_ops; ++i) {
        const diffsg_op& o = ops[i];
        switch (o.kind) {
            case DIFFSG_OP_GEMM:
                if (o.src < 0 || o.src > DIFFSG_BUF_COND || o.dst < 0 || o.dst >= DIFFSG_N_BUF) return bad(i, "buffer id");
                if (o.src == o.dst) return bad(i, "GEMM src == dst");
                if (o.K <= 0 || o.N <= 0 || o.K > kMaxWidth || o.dcol < 0 || o.dcol + o.N > kMaxWidth) return bad(i, "GEMM shape");
                if ((o.w_off & 3) || (o.ldw & 3) || o.ldw < o.N) return bad(i, "W alignment");
                if ((o.flags & DIFFSG_F_TIME) && (o.t_off < 0 || o.t_off + o.N > c.tt_stride)) return bad(i, "time offset");
                break;
            case DIFFSG_OP_LNSW:
                if (o.src < 0 || o.src >= DIFFSG_N_BUF || o.dst < 0 || o.dst >= DIFFSG_N_BUF) return bad(i, "buffer id");
                if (o.N <= 0 || o.N > kMaxWidth) return bad(i, "LN width");
                break;
            case DIFFSG_OP_PUSH:
                if (o.dcol < 0 || o.dcol >= c.n_skip || o.src < 0 || o.src >= DIFFSG_N_BUF) return bad(i, "push slot");
                break;
            case DIFFSG_OP_POP:
                if (o.K < 0 || o.K >= c.n_skip || o.dst < 0 || o.dst >= DIFFSG_N_BUF || o.dcol + o.N > kMaxWidth) return bad(i, "pop slot");
                break;
            default:
                return bad(i, "unknown op kind");
        }
    }
    return DIFFSG_OK;
}

}  // namespace diffsg

using namespace diffsg;

extern "C" {

const char* diffsg_last_error(void) { return g_err; }
int diffsg_abi_version(void) { return DIFFSG_ABI_VERSION; }
int64_t diffsg_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int diffsg_plan_create(const diffsg_cfg* cfg, const diffsg_op* ops, int32_t n_ops,
                       const int32_t* skip_widths, diffsg_plan** out) {
    if (!cfg || !ops || !out || n_ops <= 0) { set_error("plan_create: null argument"); return DIFFSG_E_INVALID; }
    if (cfg->abi_version != DIFFSG_ABI_VERSION) { set_error("ABI version mismatch: header %d, caller %d", DIFFSG_ABI_VERSION, cfg->abi_version); return DIFFSG_E_INVALID; }
    if (n_ops > kMaxOps || cfg->n_skip > kMaxSkip || cfg->n_skip < 0) { set_error("program too large (%d ops, %d skips)", n_ops, cfg->n_skip); return DIFFSG_E_UNSUPPORTED; }
    if (cfg->input_dim <= 0 || cfg->input_dim > kMaxWidth || cfg->cond_dim <= 0 || cfg->cond_dim > kMaxWidth) { set_error("input_dim/cond_dim out of range"); return DIFFSG_E_UNSUPPORTED; }
    if (cfg->n_skip > 0 && !skip_widths) { set_error("skip_widths is null"); return DIFFSG_E_INVALID; }
    int rc = validate_program(*cfg, ops, n_ops);
    if (rc) return rc;

    int ndev = 0;
    DIFFSG_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) { set_error("device %d not present", cfg->device); return DIFFSG_E_INVALID; }
    cudaDeviceProp prop;
    DIFFSG_CUDA_OK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) { set_error("diffsg_b200 is built for sm_100a only; device %d is sm_%d%d", cfg->device, prop.major, prop.minor); return DIFFSG_E_UNSUPPORTED; }
    DIFFSG_CUDA_OK(cudaSetDevice(cfg->device));

    diffsg_plan* p = new (std::nothrow) diffsg_plan();
    if (!p) { set_error("out of host memory"); return DIFFSG_E_INVALID; }
    p->cfg = *cfg;
    p->n_ops = n_ops;
    p->sm_count = prop.multiProcessorCount;
    p->max_smem = (int)prop.sharedMemPerBlockOptin;

    // buffer widths: the widest vector each buffer id ever holds
    int bw[DIFFSG_N_BUF + 1] = {0, 0, 0, 0, 0};
    auto grow = [&](int b, int w) { if (b >= 0 && b <= DIFFSG_BUF_COND && w > bw[b]) bw[b] = w; };
    grow(cfg->in_buf, cfg->input_dim);
    grow(DIFFSG_BUF_COND, cfg->cond_dim);
    for (int i = 0; i < n_ops; ++i) {
        const diffsg_op& o = ops[i];
        if (o.kind == DIFFSG_OP_GEMM) { grow(o.src, o.K); grow(o.dst, o.dcol + o.N); }
        else if (o.kind == DIFFSG_OP_LNSW) { grow(o.src, o.N); grow(o.dst, o.N); }
        else if (o.kind == DIFFSG_OP_PUSH) { grow(o.src, o.N); }
        else if (o.kind == DIFFSG_OP_POP) { grow(o.dst, o.dcol + o.N); }
    }
    PlanDev& P = p->dev;
    memset(&P, 0, sizeof(P));
    int off = 0;
    for (int b = 0; b <= DIFFSG_N_BUF; ++b) {
        const int ld = (bw[b] + 3) & ~3;
        if (ld > kMaxWidth) { delete p; set_error("buffer %d width %d exceeds %d", b, ld, kMaxWidth); return DIFFSG_E_UNSUPPORTED; }
        P.buf_off[b] = off;
        P.buf_ld[b] = ld > 0 ? ld : 4;
        off += P.buf_ld[b] * kRowsPerWarp;
    }
    P.slab_floats = off;
    int soff = 0;
    for (int s = 0; s < cfg->n_skip; ++s) { P.skip_off[s] = soff; soff += skip_widths[s]; }
    P.stash_off = soff;
    soff += cfg->input_dim;
    P.scratch_floats = soff;
    P.n_ops = n_ops;
    P.tt_stride = cfg->tt_stride;
    P.M = cfg->input_dim;
    P.C = cfg->cond_dim;
    P.in_buf = cfg->in_buf;
    P.out_buf = cfg->out_buf;

    const size_t slab_bytes = (size_t)P.slab_floats * sizeof(float);
    int warps = (int)((size_t)p->max_smem / slab_bytes);
    if (warps > kMaxWarps) warps = kMaxWarps;
    if (warps < 1) { delete p; set_error("per-warp slab (%zu B) exceeds shared memory", slab_bytes); return DIFFSG_E_UNSUPPORTED; }
    p->warps = warps;
    p->smem_bytes = slab_bytes * warps;

    const size_t n_warps_total = (size_t)p->sm_count * warps;
    const size_t scratch_bytes = n_warps_total * kRowsPerWarp * (size_t)P.scratch_floats * sizeof(float);
    if (cudaMalloc(&p->d_ops, sizeof(diffsg_op) * n_ops) != cudaSuccess ||
        cudaMalloc(&p->d_scratch, scratch_bytes) != cudaSuccess) {
        set_error("cudaMalloc of plan scratch failed: %s", cudaGetErrorString(cudaGetLastError()));
        diffsg_plan_destroy(p);
        return DIFFSG_E_CUDA;
    }
    DIFFSG_CUDA_OK(cudaMemcpy(p->d_ops, ops, sizeof(diffsg_op) * n_ops, cudaMemcpyHostToDevice));
    DIFFSG_CUDA_OK(cudaMemset(p->d_scratch, 0, scratch_bytes));
    P.ops = p->d_ops;
    P.scratch = p->d_scratch;
    DIFFSG_CUDA_OK(cudaFuncSetAttribute(unet_forward_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem));
    DIFFSG_CUDA_OK(cudaFuncSetAttribute(sample_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->max_smem));
    *out = p;
    return DIFFSG_OK;
}

int diffsg_plan_destroy(diffsg_plan* p) {
    if (!p) return DIFFSG_OK;
    tc::tc_destroy(p);
    if (p->d_ops) cudaFree(p->d_ops);
    if (p->d_scratch) cudaFree(p->d_scratch);
    delete p;
    return DIFFSG_OK;
}

int diffsg_plan_set_weights(diffsg_plan* p, const float* params_dev, size_t n_params,
                            const float* time_table_dev, int32_t tt_rows) {
    if (!p || !params_dev || !time_table_dev || tt_rows <= 0) { set_error("set_weights: null argument"); return DIFFSG_E_INVALID; }
    if (((uintptr_t)params_dev & 15) != 0) { set_error("params blob must be 16-byte aligned"); return DIFFSG_E_INVALID; }
    // bounds-check every offset against the blob
    std::vector<diffsg_op> ops(p->n_ops);
    DIFFSG_CUDA_OK(cudaMemcpy(ops.data(), p->d_ops, sizeof(diffsg_op) * p->n_ops, cudaMemcpyDeviceToHost));
    for (int i = 0; i < p->n_ops; ++i) {
        const diffsg_op& o = ops[i];
        size_t hi = 0;
        if (o.kind == DIFFSG_OP_GEMM) {
            hi = (size_t)o.w_off + (size_t)o.K * o.ldw;
            if (!(o.flags & DIFFSG_F_NOBIAS) && (size_t)o.b_off + o.N > hi) hi = (size_t)o.b_off + o.N;
        } else if (o.kind == DIFFSG_OP_LNSW) {
            hi = (size_t)(o.w_off > o.b_off ? o.w_off : o.b_off) + o.N;
        }
        if (hi > n_params) { set_error("op %d reads past the parameter blob (%zu > %zu)", i, hi, n_params); return DIFFSG_E_INVALID; }
    }
    p->dev.params = params_dev;
    p->dev.tt = time_table_dev;
    p->tt_rows = tt_rows;
    p->have_weights = true;
    return DIFFSG_OK;
}

static int grid_for(const diffsg_plan* p, int64_t B) {
    const int64_t groups = (B + kRowsPerWarp - 1) / kRowsPerWarp;
    int64_t g = groups < p->sm_count ? groups : p->sm_count;
    return (int)(g < 1 ? 1 : g);
}

int diffsg_unet_forward(diffsg_plan* p, const float* x, const int32_t* t_idx, const float* cond,
                        const float* mask, float* eps, int64_t B, void* stream) {
    if (!p || !x || !t_idx || !cond || !eps) { set_error("unet_forward: null argument"); return DIFFSG_E_INVALID; }
    if (p->engine != DIFFSG_ENGINE_TC && !p->have_weights) { set_error("unet_forward before set_weights"); return DIFFSG_E_STATE; }
    if (B <= 0) return DIFFSG_OK;
    DIFFSG_CUDA_OK(cudaSetDevice(p->cfg.device));
    if (p->engine == DIFFSG_ENGINE_TC) return tc::tc_forward(p, x, t_idx, cond, mask, eps, B, (cudaStream_t)stream);
    unet_forward_simt_kernel<<<grid_for(p, B), p->warps * 32, p->smem_bytes, (cudaStream_t)stream>>>(
        p->dev, x, t_idx, cond, mask, eps, B);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

// y <- (y - mean) / sd with the two scalars of one step; shared by both engines and diffsg_sample_renorm
static int launch_renorm(float* y, float* rec, const double* stats, int64_t n, int64_t n_stat, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int rb = (int)((n + 1023) / 1024);
    if (rb > sms * 8) rb = sms * 8;
    if (rb < 1) rb = 1;
    renorm_kernel<<<rb, 256, 0, st>>>(y, rec, stats, n, n_stat);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

static int simt_sample_launch(diffsg_plan* p, const diffsg_sample_args* a, int norm_steps, int hi, int lo, cudaStream_t st) {
    SampleDev S;
    memset(&S, 0, sizeof(S));
    const int T = a->T;
    S.cond = a->cond_dev; S.y = a->y_dev; S.noise = a->noise_dev;
    S.rec_y = a->rec_y_dev; S.rec_eps = a->rec_eps_dev; S.stats = a->stat_ws_dev;
    S.B = a->B; S.T = T; S.norm_steps = norm_steps; S.omega = a->omega;
    S.seed = a->philox_seed; S.offset = a->philox_offset;
    for (int i = 0; i < T; ++i) {
        S.c_eps[i] = a->coef_host[i];
        S.c_rs[i] = a->coef_host[T + i];
        S.c_noise[i] = a->coef_host[2 * T + i];
    }
    S.step_hi = hi; S.step_lo = lo;
    sample_simt_kernel<<<grid_for(p, a->B), p->warps * 32, p->smem_bytes, st>>>(p->dev, S);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

int diffsg_sample_steps(diffsg_plan* p, const diffsg_sample_args* a, int32_t step_hi, int32_t step_lo, int32_t renorm,
                        void* stream) {
    if (!p || !a || !a->cond_dev || !a->y_dev || !a->coef_host || !a->stat_ws_dev) { set_error("sample: null argument"); return DIFFSG_E_INVALID; }
    if (p->engine != DIFFSG_ENGINE_TC && !p->have_weights) { set_error("sample before set_weights"); return DIFFSG_E_STATE; }
    if (a->T <= 0 || a->T > 64) { set_error("T=%d outside [1,64]", a->T); return DIFFSG_E_UNSUPPORTED; }
    if (a->norm_steps < 0) { set_error("norm_steps < 0"); return DIFFSG_E_INVALID; }
    const int T = a->T;
    if (step_lo < 0 || step_hi >= T || step_lo > step_hi) { set_error("sample_steps: steps %d..%d outside [0,%d)", step_hi, step_lo, T); return DIFFSG_E_INVALID; }
    const int norm_steps = a->norm_steps > T ? T : a->norm_steps;
    const int first_plain = T - 1 - norm_steps;          // highest step that is NOT re-normalised
    if (!renorm && step_hi > first_plain && step_hi != step_lo) {
        set_error("sample_steps: with renorm == 0 a re-normalised step (%d) must be the only step of the call", step_hi);
        return DIFFSG_E_INVALID;
    }
    if (a->B <= 0) return DIFFSG_OK;
    DIFFSG_CUDA_OK(cudaSetDevice(p->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool tc_engine = p->engine == DIFFSG_ENGINE_TC;
    if (tc_engine) { if (int rc = tc::tc_sample_check(p, a)) return rc; }
    else if (T > p->tt_rows) { set_error("time table has %d rows, T=%d", p->tt_rows, T); return DIFFSG_E_INVALID; }
    const int64_t n = a->B * (int64_t)p->cfg.input_dim;
    int i = step_hi;
    // re-normalised steps: the statistics are over the whole batch -> one launch each, grid-wide sums, then renorm
    for (; i >= step_lo && i > first_plain; --i) {
        DIFFSG_CUDA_OK(cudaMemsetAsync(a->stat_ws_dev + 2 * i, 0, sizeof(double) * 2, st));
        if (int rc = tc_engine ? tc::tc_sample_launch(p, a, norm_steps, i, i, st) : simt_sample_launch(p, a, norm_steps, i, i, st)) return rc;
        if (renorm) {
            if (int rc = launch_renorm(a->y_dev, a->rec_y_dev ? a->rec_y_dev + (int64_t)(T - 1 - i) * n : nullptr,
                                       a->stat_ws_dev + 2 * i, n, n, st)) return rc;
        }
    }
    // remaining steps: rows are independent -> one persistent launch
    if (i >= step_lo) {
        if (int rc = tc_engine ? tc::tc_sample_launch(p, a, norm_steps, i, step_lo, st) : simt_sample_launch(p, a, norm_steps, i, step_lo, st)) return rc;
    }
    return DIFFSG_OK;
}

int diffsg_sample(diffsg_plan* p, const diffsg_sample_args* a, void* stream) {
    if (!a) { set_error("sample: null argument"); return DIFFSG_E_INVALID; }
    return diffsg_sample_steps(p, a, a->T - 1, 0, 1, stream);
}

int diffsg_sample_renorm(float* y, float* rec, const double* stats, int64_t n_local, int64_t n_stat, void* stream) {
    if (!y || !stats || n_local < 0 || n_stat < 2 || n_stat < n_local) { set_error("sample_renorm: bad argument"); return DIFFSG_E_INVALID; }
    if (n_local == 0) return DIFFSG_OK;
    return launch_renorm(y, rec, stats, n_local, n_stat, (cudaStream_t)stream);
}

int diffsg_plan_attach_tc(diffsg_plan* p, const diffsg_tc_program* prog) { return tc::tc_attach(p, prog); }

int diffsg_plan_set_tc_weights(diffsg_plan* p, const void* w_hi, const void* w_lo, size_t w_bytes, const float* params,
                               size_t n_params, const float* tt, int32_t tt_rows, const void* tt_img, int32_t img_rows,
                               int64_t img_stride) {
    return tc::tc_set_weights(p, w_hi, w_lo, w_bytes, params, n_params, tt, tt_rows, tt_img, img_rows, img_stride);
}

int diffsg_plan_status(diffsg_plan* p, int32_t* flags, int32_t reset, void* stream) {
    return tc::tc_status(p, flags, reset, (cudaStream_t)stream);
}

int diffsg_plan_set_engine(diffsg_plan* p, int32_t engine) {
    if (!p || (engine != DIFFSG_ENGINE_SIMT && engine != DIFFSG_ENGINE_TC)) { set_error("set_engine: bad argument"); return DIFFSG_E_INVALID; }
    if (engine == DIFFSG_ENGINE_TC && !p->tc) { set_error("set_engine: no tensor-core program attached"); return DIFFSG_E_STATE; }
    p->engine = engine;
    return DIFFSG_OK;
}

int diffsg_plan_query(const diffsg_plan* p, int32_t what) {
    if (!p) return -1;
    if (what == 0) return p->engine;
    if (what == 5) return p->warps;
    return tc::tc_query(p, what);
}

int diffsg_philox_normal(float* out, int64_t B, int32_t M, int32_t step, uint64_t seed, uint64_t offset,
                         void* stream) {
    if (!out || B < 0 || M <= 0) { set_error("philox_normal: bad argument"); return DIFFSG_E_INVALID; }
    if (B == 0) return DIFFSG_OK;
    const int64_t total = B * ((M + 3) / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    philox_fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, B, M, (uint32_t)step, seed, offset);
    count_launch();
    DIFFSG_CUDA_OK(cudaGetLastError());
    return DIFFSG_OK;
}

}  // extern "C"
