// Issue-rate micro-benchmarks for the epilogue instruction mix (sm_100a): cycles per warp instruction per
// SM sub-partition for FFMA / FFMA2 / FADD2 / FMUL2 / MUFU.{EX2,RCP,TANH} / F2FP / HADD2.F32, at 1 and 2
// warps per sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 512
#define CH 8

__device__ __forceinline__ unsigned long long pk(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
template <int OP>
__global__ void bench(float* out, long long* cyc, float seed) {
    float x[CH];
    unsigned long long p[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { x[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; p[i] = pk(x[i], x[i] + 0.5f); }
    const unsigned long long c2 = pk(seed * 0.999f, seed * 0.998f), d2 = pk(1e-3f, 2e-3f);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(seed), "f"(1e-3f));
            if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
            if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(d2));
            if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
            if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 6) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 7) {   // F2FP pack + unpack
                unsigned h;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x[i]), "f"(seed));
                asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(x[i]) : "r"(h));
            }
            if (OP == 8) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(1e-3f));
            if (OP == 9) {   // swish body, scalar: mul, ex2, add, rcp, mul
                float e;
                asm volatile("mul.rn.f32 %0, %1, 0fBFB8AA3B;" : "=f"(e) : "f"(x[i]));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(e));
                asm volatile("add.rn.f32 %0, %0, 0f3F800000;" : "+f"(e));
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(e));
                asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(e));
            }
            if (OP == 10) {  // swish via tanh: h = 0.5 x; t = tanh(h); y = h * t + h
                float h, t;
                asm volatile("mul.rn.f32 %0, %1, 0f3F000000;" : "=f"(h) : "f"(x[i]));
                asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
                asm volatile("fma.rn.f32 %0, %1, %2, %1;" : "=f"(x[i]) : "f"(h), "f"(t));
            }
            if (OP == 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(*(unsigned*)&x[i]) : "r"(0x1234u), "r"(0x77u));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); s += x[i] + a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int instr_per_body) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    for (int threads : {128, 256, 512}) {
        bench<OP><<<148, threads>>>(out, cyc, 1.0001f);
        bench<OP><<<148, threads>>>(out, cyc, 1.0001f);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        const double per_smsp = (double)ITERS * CH * instr_per_body * (threads / 128);   // warp instrs per sub-partition
        printf("%-28s warps/SMSP %d: %8lld cyc, %.3f cyc per warp-instr per SMSP\n", name, threads / 128, h, h / per_smsp);
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FFMA", 1); run<8>("FADD", 1); run<1>("FFMA2 (f32x2)", 1); run<2>("FADD2", 1); run<3>("FMUL2", 1);
    run<4>("MUFU.EX2", 1); run<5>("MUFU.RCP", 1); run<6>("MUFU.TANH", 1);
    run<7>("F2FP+HADD2.F32", 2); run<11>("LOP3", 1);
    run<9>("swish ex2+rcp (5 instr)", 5); run<10>("swish tanh (3 instr)", 3);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
