// Host-side plan object behind the opaque diffsg_plan handle.
#pragma once
#include "common.cuh"
#include "unet_simt.cuh"

namespace diffsg { namespace tc { struct TcHost; } }

struct diffsg_plan {
    diffsg_cfg cfg{};
    diffsg::PlanDev dev{};
    diffsg_op* d_ops = nullptr;
    float* d_scratch = nullptr;
    int n_ops = 0;
    int sm_count = 0;
    int max_smem = 0;
    int warps = 0;            // warps per CTA of the warp-row kernels
    size_t smem_bytes = 0;
    int tt_rows = 0;
    bool have_weights = false;
    int engine = DIFFSG_ENGINE_SIMT;
    diffsg::tc::TcHost* tc = nullptr;   // tensor-core program + scratch (unet_tc.cu)
};

namespace diffsg {
namespace tc {
// implemented in unet_tc.cu
int tc_attach(diffsg_plan* p, const diffsg_tc_program* prog);
int tc_set_weights(diffsg_plan* p, const void* w_hi, const void* w_lo, size_t w_bytes, const float* params,
                   size_t n_params, const float* tt, int tt_rows, const void* tt_img, int img_rows, int64_t img_stride);
int tc_status(diffsg_plan* p, int32_t* flags, int reset, cudaStream_t st);
int tc_forward(diffsg_plan* p, const float* x, const int32_t* t_idx, const float* cond, const float* mask,
               float* eps, int64_t B, cudaStream_t st);
int tc_sample_check(diffsg_plan* p, const diffsg_sample_args* a);
int tc_sample_launch(diffsg_plan* p, const diffsg_sample_args* a, int norm_steps, int step_hi, int step_lo, cudaStream_t st);
void tc_destroy(diffsg_plan* p);
int tc_query(const diffsg_plan* p, int what);
}  // namespace tc
}  // namespace diffsg
