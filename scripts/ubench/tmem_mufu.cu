// Micro-benchmarks that size the attention softmax loop: tcgen05.ld / tcgen05.st throughput per SM for 4 and 8 warps,
// MUFU.EX2 throughput, and FFMA-polynomial exp2 throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>
#include "../../sd-webui-text2video_b200/csrc/ptx.cuh"
using namespace t2v;

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// mode 0: LDTM x32 ; 1: STTM x32 ; 2: MUFU.EX2 ; 3: polynomial exp2 on the FMA pipe ; 4: LDTM x32 with 2 in flight
__global__ void k(int mode, int iters, float* out, long long* clk) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    float acc = threadIdx.x * 1e-3f;
    uint32_t r[32];
    for (int i = 0; i < 32; ++i) r[i] = i + threadIdx.x;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
            tmem_ld_32x32(base + (it & 3) * 32, r);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 31]);
        }
    } else if (mode == 4) {
        uint32_t r2[32];
        for (int it = 0; it < iters; it += 2) {
            tmem_ld_32x32(base + (it & 2) * 32, r);
            tmem_ld_32x32(base + (it & 2) * 32 + 32, r2);
            tmem_ld_wait();
            acc += __uint_as_float(r[it & 31]) + __uint_as_float(r2[it & 31]);
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; ++it) {
            r[0] = it;
            tmem_st_32x32(base + (it & 3) * 32, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else if (mode == 2) {
        float x[8];
        for (int i = 0; i < 8; ++i) x[i] = -1e-3f * (i + 1) * (threadIdx.x + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = ex2_approx(x[i]) - 1.5f;
        }
        for (int i = 0; i < 8; ++i) acc += x[i];
    } else if (mode == 3) {
        float x[8];
        for (int i = 0; i < 8; ++i) x[i] = -1e-3f * (i + 1) * (threadIdx.x + 1);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // exp2(x) for x <= 0: split into integer + fraction, cubic on the fraction, exponent add by integer ALU
                float xf = fmaxf(x[i], -126.f);
                float fl = floorf(xf);
                float f = xf - fl;
                float p = fmaf(fmaf(fmaf(0.0555041f, f, 0.2402265f), f, 0.6931472f), f, 1.0f);
                int e = static_cast<int>(fl);
                x[i] = __int_as_float(__float_as_int(p) + (e << 23)) - 1.5f;
            }
        }
        for (int i = 0; i < 8; ++i) acc += x[i];
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

int main() {
    float* out;
    long long* clk;
    cudaMalloc(&out, 148 * 512 * 4);
    cudaMalloc(&clk, 148 * 8);
    const char* names[] = {"LDTM.x32 (ld+wait)", "STTM.x32", "MUFU.EX2", "poly exp2 (FMA pipe)", "LDTM.x32 x2 in flight"};
    for (int mode = 0; mode < 5; ++mode)
        for (int warps = 4; warps <= 16; warps *= 2) {
            const int iters = 4096;
            k<<<148, warps * 32, 0>>>(mode, iters, out, clk);
            cudaDeviceSynchronize();
            k<<<148, warps * 32, 0>>>(mode, iters, out, clk);
            cudaError_t e = cudaDeviceSynchronize();
            long long h;
            cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
            double per_sm;
            const char* unit;
            if (mode == 0 || mode == 1 || mode == 4) { per_sm = double(iters) * warps * 4096 / h; unit = "B/clk/SM"; }
            else { per_sm = double(iters) * 8 * warps * 32 / h; unit = "exp2/clk/SM"; }
            printf("%-24s warps %2d: %8lld clk  %7.1f %s  (%s)\n", names[mode], warps, h, per_sm, unit, cudaGetErrorString(e));
        }
    return 0;
}
