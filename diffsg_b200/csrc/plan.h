// Host-side plan object behind the opaque diffsg_plan handle.
#pragma once
#include "common.cuh"
#include "unet_simt.cuh"

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
};
