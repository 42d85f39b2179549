/*
 * diffsg_b200 — C-ABI of the B200-native CFG-DDPM solver path.
 *
 * The reference (qiyu3816/DiffSG) is pure Python/PyTorch and has no FFI; its public
 * surface for this path is the nn.Module API.  Every entry point below names the
 * reference interface it replaces (paths relative to the reference repo root).  The
 * Python host side (the modules under diffsg_b200/) keeps the reference's constructor arguments,
 * state_dict layout and sample()/forward() signatures and binds these symbols through
 * ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 on success, a negative DIFFSG_E_* code on failure;
 *     diffsg_last_error() returns a thread-local, NUL-terminated description.
 *   - all `*_dev` pointers are device pointers on the plan's device, row-major, fp32
 *     unless stated.  The caller owns every buffer; the library borrows them for the
 *     duration of the call (weights: until the next diffsg_plan_set_weights / destroy).
 *   - all work is enqueued on the caller's stream (`stream` = cudaStream_t cast to
 *     void*); no call synchronises the device.
 *   - a plan owns per-CTA scratch: calls on ONE plan must be ordered on one stream (or by events);
 *     different plans may run concurrently on different streams (the library orders the swap of the
 *     tensor-core stage program, which lives in __constant__ memory, behind the kernels still using it).
 *   - there is no CPU path: a missing/unsupported GPU is an error, never a fallback.
 */
#ifndef DIFFSG_B200_H
#define DIFFSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIFFSG_ABI_VERSION 3

enum {
    DIFFSG_OK = 0,
    DIFFSG_E_INVALID = -1,   /* bad argument / malformed program */
    DIFFSG_E_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    DIFFSG_E_UNSUPPORTED = -3, /* shape outside the compiled limits, or wrong GPU arch */
    DIFFSG_E_STATE = -4      /* call order (e.g. forward before set_weights) */
};

/* ---- network program ---------------------------------------------------------------
 * The host lowers UNet1D (reference ddpm_opt/UNetCF.py:262-356) to a flat list of ops
 * over four per-row scratch vectors (buffer ids 0..3), a condition vector (id 4) and a
 * skip stack.  Parameters live in one fp32 blob; offsets are in floats.
 */
enum {
    DIFFSG_OP_GEMM = 1,  /* dst[dcol:dcol+N] (+)= src[0:K] . W[K,N] + bias[N]  (nn.Linear) */
    DIFFSG_OP_LNSW = 2,  /* dst[0:N] = swish(layer_norm(src[0:N]; gamma, beta, eps=1e-5)) */
    DIFFSG_OP_PUSH = 3,  /* skip[slot=dcol] = src[0:N]                                     */
    DIFFSG_OP_POP = 4    /* dst[dcol:dcol+N] = skip[slot=K]                                */
};
enum {
    DIFFSG_F_ACC = 1,       /* GEMM: accumulate into dst instead of overwriting            */
    DIFFSG_F_TIME = 2,      /* GEMM: add time_table[t_idx[row]][t_off : t_off+N] to bias   */
    DIFFSG_F_NOBIAS = 4     /* GEMM: no bias vector (b_off ignored)                        */
};
#define DIFFSG_BUF_COND 4   /* src id of the swish(cond * mask) vector                     */
#define DIFFSG_N_BUF 4

typedef struct diffsg_op {
    int32_t kind;   /* DIFFSG_OP_*                                                         */
    int32_t src;    /* source buffer id                                                    */
    int32_t dst;    /* destination buffer id                                               */
    int32_t K;      /* GEMM: input width (unpadded).  POP: skip slot.                      */
    int32_t N;      /* GEMM: output width.  LNSW/PUSH/POP: vector width.                   */
    int32_t flags;  /* DIFFSG_F_*                                                          */
    int32_t w_off;  /* GEMM: W, stored [K][ldw] (W^T of nn.Linear.weight).  LNSW: gamma.   */
    int32_t b_off;  /* GEMM: bias.  LNSW: beta.                                            */
    int32_t t_off;  /* GEMM+TIME: column offset into a time_table row                      */
    int32_t dcol;   /* GEMM/POP: first destination column.  PUSH: skip slot.               */
    int32_t ldw;    /* GEMM: row stride of W in floats (N rounded up to 4)                 */
    int32_t pad_;
} diffsg_op;

typedef struct diffsg_cfg {
    int32_t abi_version;   /* DIFFSG_ABI_VERSION                                           */
    int32_t input_dim;     /* M: width of y / eps (UNet1D input_dim)                       */
    int32_t cond_dim;      /* C: width of the condition vector                             */
    int32_t max_width;     /* widest vector any op touches (<= 256)                        */
    int32_t n_skip;        /* number of skip slots                                         */
    int32_t skip_floats;   /* sum of skip widths (floats per row)                          */
    int32_t tt_stride;     /* floats per time_table row                                    */
    int32_t tt_rows;       /* rows in the time table (T for sampling; U for forward)       */
    int32_t in_buf;        /* buffer id that receives x / y_t                              */
    int32_t out_buf;       /* buffer id holding eps after the last op                      */
    int32_t device;        /* CUDA device ordinal                                          */
    int32_t reserved[5];
} diffsg_cfg;

typedef struct diffsg_plan diffsg_plan;

const char* diffsg_last_error(void);
int diffsg_abi_version(void);

/* Plan = compiled program + plan-lifetime scratch (skip stack, partial sums).
 * Replaces: UNet1D.__init__ (ddpm_opt/UNetCF.py:262-316) as the owner of the topology. */
int diffsg_plan_create(const diffsg_cfg* cfg, const diffsg_op* ops, int32_t n_ops,
                       const int32_t* skip_widths, diffsg_plan** out);
int diffsg_plan_destroy(diffsg_plan* plan);

/* Bind parameters.  `params_dev`: fp32 blob addressed by the ops' offsets;
 * `time_table_dev`: [tt_rows][tt_stride] hoisted per-block time biases
 * (TimeEmbedding + ResidualBlock.time_emb, ddpm_opt/UNetCF.py:30-46,91).
 * Replaces: nn.Module.load_state_dict / .to(device) for the denoiser. */
int diffsg_plan_set_weights(diffsg_plan* plan, const float* params_dev, size_t n_params,
                            const float* time_table_dev, int32_t tt_rows);

/* eps[B,M] = UNet(x[B,M], time row t_idx[B], cond[B,C] * mask[B]).
 * mask_dev may be NULL (all ones).  Replaces: UNet1D.forward (ddpm_opt/UNetCF.py:318-356). */
int diffsg_unet_forward(diffsg_plan* plan, const float* x_dev, const int32_t* t_idx_dev,
                        const float* cond_dev, const float* mask_dev, float* eps_dev,
                        int64_t B, void* stream);

/* Reverse-diffusion CFG sampler.  Replaces: DDPM.sample
 * (ddpm_opt/classifier_free_MSR.py:114-155 == _NU.py:143-180 == _CO.py:117-154).
 *
 *   for i = T-1 .. 0:
 *     eps = (1+omega) * UNet(y, i, cond, 1) - omega * UNet(y, i, cond, 0)
 *     y   = (y - c_eps[i] * eps) * c_rs[i] + c_noise[i] * z_i        (z_i = 0 for i <= 1)
 *     if i > T-5: y = (y - mean(y)) / sqrt(var_unbiased(y))          (scalars over [B,M])
 *
 * coef_host: 3*T floats on the HOST: c_eps[T] | c_rs[T] | c_noise[T].
 * y_dev: in = y_T, out = y_0 (in place).  noise_dev: NULL => counter-based Philox4x32-10
 * normals keyed by (seed, step, row, col); else [(T-2)][B][M] injected draws, plane k used
 * at step i = T-1-k.  rec_y_dev / rec_eps_dev: NULL or [T][B][M] per-step records
 * (record_denoise_path).  stat_ws_dev: >= 2*T doubles of zero-initialisable workspace.
 */
typedef struct diffsg_sample_args {
    const float* cond_dev;
    float* y_dev;
    const float* noise_dev;
    float* rec_y_dev;
    float* rec_eps_dev;
    double* stat_ws_dev;
    const float* coef_host;
    int64_t B;
    int32_t T;
    int32_t norm_steps;   /* number of leading steps that re-normalise (reference: 4)      */
    float omega;
    uint32_t pad_;
    uint64_t philox_seed;
    uint64_t philox_offset;
} diffsg_sample_args;

#ifdef __cplusplus
static_assert(sizeof(diffsg_op) == 48 && sizeof(diffsg_cfg) == 64 && sizeof(diffsg_sample_args) == 96,
              "diffsg_b200 ABI struct layout changed: bump DIFFSG_ABI_VERSION and the bindings");
#endif

int diffsg_sample(diffsg_plan* plan, const diffsg_sample_args* args, void* stream);

/* A slice of diffsg_sample, for callers that shard ONE batch over several GPUs and want the reference's
 * whole-batch statistics in the re-normalised steps (the reference normalises over the batch of the call,
 * classifier_free_MSR.py:136-137): runs reverse steps step_hi .. step_lo (inclusive, descending).
 *   renorm != 0: re-normalised steps use this shard's own statistics (what diffsg_sample does);
 *   renorm == 0: a re-normalised step only ACCUMULATES (sum y, sum y^2) of its un-normalised output into
 *                stat_ws_dev[2*step .. 2*step+1] (zeroed by this call); it must be the only step of the call
 *                (step_hi == step_lo).  The caller sums the two doubles over all ranks (NCCL all-reduce)
 *                and calls diffsg_sample_renorm with the global element count.
 * diffsg_sample == diffsg_sample_steps(T-1, 0, renorm = 1). */
int diffsg_sample_steps(diffsg_plan* plan, const diffsg_sample_args* args, int32_t step_hi, int32_t step_lo,
                        int32_t renorm, void* stream);
/* y = (y - mean) / sqrt(var_unbiased) over n_local elements with mean / var from stats_dev = {sum y, sum y^2}
 * taken over n_stat elements (n_stat = n_local for per-shard statistics); rec_y_plane_dev: NULL or the
 * [B][M] record plane of that step, which receives the normalised values as well. */
int diffsg_sample_renorm(float* y_dev, float* rec_y_plane_dev, const double* stats_dev, int64_t n_local,
                         int64_t n_stat, void* stream);

/* ---- tensor-core program (engine DIFFSG_ENGINE_TC) ------------------------------------
 * Second lowering of the same network for the tcgen05 engine (diffsg_b200/tc_packer.py):
 * stages = GEMM groups accumulating in TMEM, each followed by an epilogue micro-program.
 * Records are packed little-endian structs; see diffsg_b200/csrc/unet_tc.cuh. */
enum { DIFFSG_ENGINE_SIMT = 0, DIFFSG_ENGINE_TC = 1 };

typedef struct diffsg_tc_program {
    const void* stages;        /* n_stages x 16-byte records                                  */
    const void* chunks;        /* n_chunks x  8-byte records                                  */
    const void* epis;          /* n_epi    x  8-byte records                                  */
    const int32_t* skip_widths;/* n_skip padded widths (multiples of 16)                      */
    int32_t n_stages, n_chunks, n_epi, n_skip;
    int32_t nterms;            /* 2: (A_hi + A_lo) . W_fp16;  3: + A_hi . W_lo                */
    int32_t tt_stride;         /* floats per row of the tensor-core time table                */
    int32_t reserved[2];
} diffsg_tc_program;

/* Attach the tensor-core program to a plan (fails with DIFFSG_E_UNSUPPORTED if the topology is
 * outside the engine's limits; the plan then keeps running on the fp32 engine). */
int diffsg_plan_attach_tc(diffsg_plan* plan, const diffsg_tc_program* prog);
/* fp16 weight + static bias-chunk images (w_lo_dev may be NULL when nterms == 2), the fp32 LayerNorm
 * gamma / beta packages, and the hoisted time path of the tensor-core program in two forms:
 *   time_table_dev  fp32 [tt_rows][tt_stride]: lin1.bias + time embedding per block; read per row by
 *                   diffsg_unet_forward (rows carry arbitrary time indices).  May be NULL for sampling only.
 *   time_img_dev    fp16 [img_rows][img_stride_bytes]: the same rows as K = 16 bias-chunk images (three fp16
 *                   terms per value), streamed by the sampler, row = reverse step.  May be NULL for forward only. */
int diffsg_plan_set_tc_weights(diffsg_plan* plan, const void* w_hi_dev, const void* w_lo_dev,
                               size_t w_bytes, const float* params_dev, size_t n_params,
                               const float* time_table_dev, int32_t tt_rows, const void* time_img_dev,
                               int32_t img_rows, int64_t img_stride_bytes);
/* Sticky status bits raised on the device by the tensor-core kernels of this plan since the last reset:
 *   DIFFSG_STATUS_FP16_OVERFLOW  a raw (un-normalised) operand row -- y_t, the residual stream into a Down/Upsample
 *                                or shortcut Linear -- exceeded the fp16 range (|x| > 65504): the fp16-split engines
 *                                saturate it, so results of that call are NOT within tolerance of the fp32 reference;
 *                                re-run with precision "fp32".
 * This call synchronises `stream` (it reads one word back). */
#define DIFFSG_STATUS_FP16_OVERFLOW 1
int diffsg_plan_status(diffsg_plan* plan, int32_t* flags_out, int32_t reset, void* stream);
/* Select the engine used by diffsg_unet_forward / diffsg_sample. */
int diffsg_plan_set_engine(diffsg_plan* plan, int32_t engine);

/* Introspection: what = 0 engine, 1 tensor-core CTAs resident per SM, 2 its dynamic shared memory per CTA
 * (bytes), 3 its maximum grid, 4 nterms, 5 warps per CTA of the fp32 engine, 6 / 7 the K columns per operand
 * chunk and the TMEM columns per accumulator region this library was BUILT for (the program handed to
 * diffsg_plan_attach_tc must be lowered for the same values), 8 the tensor-core engine's global scratch per
 * CTA in bytes.  Returns -1 when not applicable. */
int diffsg_plan_query(const diffsg_plan* plan, int32_t what);

/* Number of kernel launches issued by this library (process-wide) since the last reset
 * (bench.py's `gpu_launches`). */
int64_t diffsg_launch_count(int reset);

/* Fill out[n] with the sampler's Philox normals for (seed, offset, step) — the exact
 * stream diffsg_sample uses when noise_dev == NULL (tests / reproducibility). */
int diffsg_philox_normal(float* out_dev, int64_t B, int32_t M, int32_t step,
                         uint64_t seed, uint64_t offset, void* stream);

/* avg = copy_first ? p : decay*avg + (1-decay)*p over one flat fp32 buffer.
 * Replaces: ExponentialMovingAverage.update_parameters (ddpm_opt/ema.py:3-14). */
int diffsg_ema_update(float* avg_dev, const float* p_dev, int64_t n, double decay,
                      int32_t copy_first, void* stream);

/* Multi-tensor form: ptr tables live on the device. */
int diffsg_ema_update_multi(float* const* avg_ptrs_dev, const float* const* p_ptrs_dev,
                            const int64_t* sizes_dev, int32_t n_tensors, int64_t max_size,
                            double decay, int32_t copy_first, void* stream);

/* Fused Adam (+ optional EMA) over flat fp32 buffers: torch.optim.Adam's update without weight decay / amsgrad
 * (reference training loop: ddpm_opt/classifier_free_MSR.py:213,225 `optimizer.step()`), and in the same pass the
 * EMA of ddpm_opt/ema.py:10-14 on the updated parameters when ema_dev != NULL and hyper_dev[5] != 0.
 *   hyper_dev: 6 floats on the DEVICE = {lr, beta1, beta2, eps, ema_decay, ema_mode (0 off, 1 copy, 2 blend)}
 *   step_dev : int64 on the device, number of steps taken so far (incremented by this call)
 * Device-resident hyper-parameters keep the launch valid inside a CUDA graph while lr / the EMA gate change. */
int diffsg_adam_step(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, float* ema_dev, int64_t n,
                     const float* hyper_dev, int64_t* step_dev, void* stream);

/* Batched forward of a small MLP: x[B, in_dim] -> out[B, out_dims[n_layers-1]].  Replaces the row-by-row torch
 * evaluation of the reference's comparison baselines: MTFNN (baselines/MTFNN.py:43-52, 122-131, 187-211) and the PPO
 * actor / critic (baselines/PPO.py:44-62).  params_dev = W_0 [out_0][in] | b_0 | W_1 | b_1 | ... (nn.Linear layouts);
 * acts[l]: 0 none, 1 ReLU, 2 Tanh, 3 Sigmoid; head: 0 none, 1 row softmax, 2 sigmoid on columns [0, head_split)
 * and softmax on the rest (MTFNN.forward for NU, MTFNN.py:208-210).  Widths <= 128, <= 8 layers. */
int diffsg_mlp_forward(const float* x_dev, const float* params_dev, int64_t B, int32_t in_dim, int32_t n_layers,
                       const int32_t* out_dims, const int32_t* acts, int32_t head, int32_t head_split, float* out_dev,
                       void* stream);

/* Global (min, max) of a strided [B, width] slice -> mm_dev[2] (decoder statistics). */
int diffsg_minmax(const float* y_dev, int64_t B, int32_t ld, int32_t col0, int32_t width,
                  float* mm_dev, void* stream);

/* MSR decode + objective (ddpm_opt/classifier_free_MSR.py:239-245, 284-288):
 *   p = W * softmax_row((y - mm[0]) / (mm[1] - mm[0]));  rate = sum_j log2(1 + p_j * g_j)
 * p_out_dev may be NULL. */
int diffsg_objective_msr(const float* y_dev, const float* g_dev, const float* mm_dev,
                         float W, float* p_out_dev, float* rate_dev, int64_t B, int32_t M,
                         void* stream);

/* MSR rate of a given allocation (labels): rate = sum_j log2(1 + p_j * g_j). */
int diffsg_rate_msr(const float* p_dev, const float* g_dev, float* rate_dev, int64_t B,
                    int32_t M, void* stream);

/* NU decode (ddpm_opt/classifier_free_NU.py:267-276) and NOMA rate (:279-303).
 * y: [B, 2+K]; mm = global (min,max) of y[:, :2]; dec_out: [B, 2+K]. */
int diffsg_decode_nu(const float* y_dev, const float* mm_dev, float width, float height,
                     float P_sum, float* dec_out_dev, int64_t B, int32_t K, void* stream);
int diffsg_rate_nu(const float* dec_dev, const float* xy_dev, float* rate_dev, int64_t B,
                   int32_t K, void* stream);

/* CO decode (ddpm_opt/classifier_free_CO.py:281-290) and cost (:255-278).
 * x: [B, 3*n] (local, transition, ideal-exec per node); y: [B, n]. */
int diffsg_decode_co(const float* y_dev, float* dec_out_dev, int64_t B, int32_t n,
                     void* stream);
int diffsg_cost_co(const float* x_dev, const float* alloc_dev, float* cost_dev, int64_t B,
                   int32_t n, void* stream);

/* ---- training: fused LayerNorm + Swish, forward and backward -----------------------------
 * y = swish(layer_norm(x; gamma, beta, eps=1e-5)) for x[B, D]; forward also returns the per-row
 * (mean, rstd) the backward needs.  Replaces the nn.LayerNorm -> Swish pairs of
 * ddpm_opt/UNetCF.py:90,92,94,356 in the training graph (82 per forward for MSR). */
int diffsg_lnsw_forward(const float* x_dev, const float* gamma_dev, const float* beta_dev, float* y_dev,
                        float* mean_dev, float* rstd_dev, int64_t B, int32_t D, void* stream);
/* workspace: >= 2 * D floats (more lets more CTAs accumulate dgamma/dbeta privately). */
int diffsg_lnsw_backward(const float* x_dev, const float* gamma_dev, const float* beta_dev,
                         const float* mean_dev, const float* rstd_dev, const float* dy_dev, float* dx_dev,
                         float* dgamma_dev, float* dbeta_dev, float* workspace_dev,
                         int64_t workspace_floats, int64_t B, int32_t D, void* stream);

/* ---- training GEMMs on the tensor cores (tcgen05, bf16 hi+lo operands, fp32 accumulate) -------------------
 * The Linear layers of the eps-MSE training step, with the LayerNorm -> Swish in front of them fused in.  Replaces
 * nn.Linear forward / autograd dgrad / wgrad of ddpm_opt/UNetCF.py:83-95 (ResidualBlock), :123-157 (attention at
 * sequence length 1), :318-356 (UNet1D.forward) under the loss of classifier_free_MSR.py:100-112.  All pointers are
 * device pointers to dense row-major fp32 (gidx: int64). */
typedef struct diffsg_mat {       /* logical [rows, k0 + k1] = cat(p0[rows, k0], p1[rows, k1]) along columns; k1 = 0: p0 only */
    const float* p0;
    const float* p1;
    int32_t k0;
    int32_t k1;
} diffsg_mat;
typedef struct diffsg_mat_out {
    float* p0;
    float* p1;
    int32_t k0;
    int32_t k1;
} diffsg_mat_out;

/* y[B, N] = act(a) . w^T + bias (+ a2 . w2^T + bias2) (+ add) (+ gadd[gidx[row]]);
 * act = swish(LayerNorm(.; gamma, beta, eps 1e-5)) over the K = a.k0 + a.k1 columns when gamma != NULL (K <= 256;
 * per-row mean / rstd are written for the backward), identity otherwise. */
typedef struct diffsg_tlin_fwd_args {
    diffsg_mat a;
    const float* w;               /* [N, K] */
    const float* bias;            /* [N] or NULL */
    const float* gamma;           /* [K] or NULL */
    const float* beta;
    float* mean;                  /* [B] out (LayerNorm mode) */
    float* rstd;
    diffsg_mat a2;                /* optional second (identity) segment accumulated into the same tile; a2.p0 = NULL: none */
    const float* w2;              /* [N, a2.k0 + a2.k1] */
    const float* bias2;
    const float* add;             /* [B, N] or NULL */
    const float* gadd;            /* [T, N] (row stride gadd_ld) or NULL: row gidx[row] is added to output row `row` */
    const int64_t* gidx;          /* [B] */
    float* y;                     /* [B, N] */
    int64_t B;
    int32_t N;
    int32_t gadd_ld;              /* row stride of gadd in floats; 0 = N (dense).  A column slice of a wider table is allowed */
} diffsg_tlin_fwd_args;
int diffsg_tlin_forward(const diffsg_tlin_fwd_args* args, void* stream);

/* dx[B, K] = (dy[B, N] . w[N, K]) (+ dres).  With gamma != NULL the product is the gradient w.r.t. swish(LayerNorm(x))
 * and is pushed through the LayerNorm -> Swish backward in the epilogue (K <= 256): dx is then the gradient w.r.t. x and
 * dgamma[K] / dbeta[K] are ACCUMULATED (+=). */
typedef struct diffsg_tlin_dgrad_args {
    const float* dy;
    const float* w;
    diffsg_mat x;                 /* LayerNorm mode: the forward input */
    const float* gamma;
    const float* beta;
    const float* mean;
    const float* rstd;
    diffsg_mat dres;              /* optional addend with K columns; dres.p0 = NULL: none */
    diffsg_mat_out dx;            /* K = dx.k0 + dx.k1 columns */
    float* dgamma;
    float* dbeta;
    int64_t B;
    int32_t N;
    int32_t K;
} diffsg_tlin_dgrad_args;
int diffsg_tlin_dgrad(const diffsg_tlin_dgrad_args* args, void* stream);

/* dw[N, K] += dy^T . act(a);  dbias[N] += column sums of dy;  dgadd[T, N] += rows of dy scattered by gidx
 * (all three from one pass over the rows; dbias / dgadd optional, gadd_rows <= 31; dgadd may be a column slice of a
 * wider [T, dgadd_ld] table). */
typedef struct diffsg_tlin_wgrad_args {
    const float* dy;
    diffsg_mat a;
    const float* gamma;
    const float* beta;
    const float* mean;
    const float* rstd;
    const int64_t* gidx;
    float* dw;
    float* dbias;
    float* dgadd;
    int64_t B;
    int32_t N;
    int32_t gadd_rows;
    int32_t dgadd_ld;             /* row stride of dgadd in floats; 0 = N */
    int32_t reserved;
} diffsg_tlin_wgrad_args;
int diffsg_tlin_wgrad(const diffsg_tlin_wgrad_args* args, void* stream);

#ifdef __cplusplus
static_assert(sizeof(diffsg_mat) == 24 && sizeof(diffsg_tlin_fwd_args) == 160 && sizeof(diffsg_tlin_dgrad_args) == 152 &&
                  sizeof(diffsg_tlin_wgrad_args) == 120,
              "diffsg_b200 ABI struct layout changed: bump DIFFSG_ABI_VERSION and the bindings");
#endif

/* The whole backward of one fused node in ONE launch: up to two dgrad and two wgrad problems (main segment + second
 * segment) that only share inputs run side by side; CTAs pick their role from the block index. */
int diffsg_tlin_backward(const diffsg_tlin_dgrad_args* dgrads, int32_t n_dgrad, const diffsg_tlin_wgrad_args* wgrads,
                         int32_t n_wgrad, void* stream);

/* ---- test hooks (not part of the product surface) ------------------------------------
 * One 128-row tcgen05 GEMM tile: C[128,N] = A[128,K] . W[N,K]^T with A split into fp16
 * (hi, lo) in-kernel and W given as pre-packed fp16 core-matrix images. */
int diffsg_debug_tc_gemm(const float* A_dev, const void* W_hi_dev, const void* W_lo_dev,
                         float* C_dev, int32_t K, int32_t N, int32_t nterms, uint32_t layout,
                         uint32_t lbo, int32_t swap_lbo_sbo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFSG_B200_H */
