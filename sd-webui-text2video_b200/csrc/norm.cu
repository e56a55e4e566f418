// GroupNorm (32 groups) statistics + apply(+SiLU), and LayerNorm, on channels-last fp16 token matrices.
// HBM-bound kernels: algorithmic bytes = 2 B/element read for stats, 2 B read + 2 B write for apply.
//
// Reduction domains: 4-D GroupNorm = one frame (h*w rows); 5-D GroupNorm = one sample
// (F*h*w rows).  Both are "instances" of `rows_per_inst` consecutive rows here.
// fp32 math throughout (the reference runs group_norm / layer_norm / SiLU-after-norm in fp32 under autocast and
// rounds to fp16 only when the value enters the next conv/linear -- exactly where these kernels round).
#include "common.cuh"
#include "kernels.cuh"

namespace t2v {

namespace {

constexpr int kGroups = 32;
constexpr int kStatsThreads = 256;   // 32 vector lanes (x) x 8 row lanes (y)
constexpr size_t kCounterBytes = 1 << 20;   // up to 262144 norm instances (frames x samples) per call

// partial[(inst * nchunks + chunk) * 32 + g] = (sum, sumsq) ; the last block of an instance folds them (in
// chunk order, double precision) into stats[inst*32+g] = (mean, rstd) -> deterministic, no float atomics in HBM.
__global__ void __launch_bounds__(kStatsThreads) gn_stats_kernel(const __half* __restrict__ x, long long ld, int C,
                                                                 int rows_per_inst, int rows_per_chunk, int nchunks,
                                                                 float eps, float2* __restrict__ partial,
                                                                 unsigned int* __restrict__ counters,
                                                                 float2* __restrict__ stats) {
    griddep_wait();
    griddep_launch_small();
    extern __shared__ float sm[];          // [2*C] per-channel sum / sumsq
    float* s_sum = sm;
    float* s_sq = sm + C;
    __shared__ bool is_last;
    const int inst = blockIdx.y;
    const int chunk = blockIdx.x;
    const int tx = threadIdx.x & 31;
    const int ty = threadIdx.x >> 5;
    const int C8 = C >> 3;
    // cross-row-lane reduction through a fixed-order smem tree (no atomics: bit-reproducible run to run)
    __shared__ float red[8][2][256];       // [row lane][sum|sumsq][32 vectors x 8 channels of the current pass]
    const int r0 = chunk * rows_per_chunk;
    const int r1 = min(r0 + rows_per_chunk, rows_per_inst);
    const __half* base = x + static_cast<long long>(inst) * rows_per_inst * ld;
    for (int v0 = 0; v0 < C8; v0 += 32) {          // uniform trip count: barriers inside are safe
        const int vc = v0 + tx;
        float s[8], q[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
        if (vc < C8) {
#pragma unroll 4
            for (int r = r0 + ty; r < r1; r += 8) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + r * ld + vc * 8));
                const __half2* h2 = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(h2[e]);
                    s[2 * e] += f.x;
                    q[2 * e] += f.x * f.x;
                    s[2 * e + 1] += f.y;
                    q[2 * e + 1] += f.y * f.y;
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            red[ty][0][tx * 8 + e] = s[e];
            red[ty][1][tx * 8 + e] = q[e];
        }
        __syncthreads();
        {
            const int ch = threadIdx.x;                 // 256 threads <-> 256 channels of this pass
            const int gc = v0 * 8 + ch;
            if (gc < C) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int y = 0; y < 8; ++y) {
                    a += red[y][0][ch];
                    b += red[y][1][ch];
                }
                s_sum[gc] = a;
                s_sq[gc] = b;
            }
        }
        __syncthreads();
    }
    const int cpg = C / kGroups;
    if (threadIdx.x < kGroups) {
        float a = 0.f, b = 0.f;
        for (int c = 0; c < cpg; ++c) {
            a += s_sum[threadIdx.x * cpg + c];
            b += s_sq[threadIdx.x * cpg + c];
        }
        partial[(static_cast<long long>(inst) * nchunks + chunk) * kGroups + threadIdx.x] = make_float2(a, b);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(&counters[inst], 1u);
        is_last = (prev == static_cast<unsigned int>(nchunks - 1));
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fold the chunk partials: 8 thread-parts per group read strided chunks (independent L2 loads in flight), then a
        // fixed-order combine -> deterministic and no serial chain of nchunks dependent loads
        __shared__ double fold[8][kGroups][2];
        {
            const int gidx = threadIdx.x & 31, part = threadIdx.x >> 5;
            double a = 0.0, b = 0.0;
#pragma unroll 4
            for (int ch = part; ch < nchunks; ch += 8) {
                const float2 p = __ldcg(&partial[(static_cast<long long>(inst) * nchunks + ch) * kGroups + gidx]);
                a += p.x;
                b += p.y;
            }
            fold[part][gidx][0] = a;
            fold[part][gidx][1] = b;
        }
        __syncthreads();
        if (threadIdx.x < kGroups) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int part = 0; part < 8; ++part) {
                a += fold[part][threadIdx.x][0];
                b += fold[part][threadIdx.x][1];
            }
            const double n = static_cast<double>(rows_per_inst) * cpg;
            const double mean = a / n;
            double var = b / n - mean * mean;
            if (var < 0.0) var = 0.0;
            stats[inst * kGroups + threadIdx.x] =
                make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps))));
        }
        if (threadIdx.x == 0) counters[inst] = 0u;     // self-cleaning for the next launch
    }
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// grid = (row blocks, instances).  Each block first folds (mean, rstd, gamma, beta) of ITS instance into per-channel
// scale/shift in smem, then streams its rows: y = act(x * a[c] + b[c]) -- one FMA per element, 16-byte accesses.
__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x, long long ldx,
                                                       __half* __restrict__ y, long long ldy, int C,
                                                       int rows_per_inst, int rows_per_block,
                                                       const float2* __restrict__ stats,
                                                       const __half* __restrict__ gamma,
                                                       const __half* __restrict__ beta, int silu) {
    griddep_wait();
    griddep_launch_small();
    extern __shared__ float ab[];          // a[C] | b[C]
    const int inst = blockIdx.y;
    const int C8 = C >> 3;
    const int cpg = C / kGroups;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float2 ms = __ldg(stats + inst * kGroups + c / cpg);
        const float a = ms.y * __half2float(__ldg(gamma + c));
        ab[c] = a;
        ab[C + c] = __half2float(__ldg(beta + c)) - ms.x * a;
    }
    __syncthreads();
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(r0 + rows_per_block, rows_per_inst);
    const long long base_row = static_cast<long long>(inst) * rows_per_inst;
    const int total = (r1 - r0) * C8;
    constexpr int U = 4;                   // independent 16-byte loads in flight per thread
    for (int i0 = threadIdx.x; i0 < total; i0 += blockDim.x * U) {
        uint4 v[U];
        int rr[U], vv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * blockDim.x;
            rr[u] = i / C8;
            vv[u] = i - rr[u] * C8;
            if (i < total) v[u] = __ldg(reinterpret_cast<const uint4*>(x + (base_row + r0 + rr[u]) * ldx + vv[u] * 8));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i >= total) continue;
            const int vc = vv[u];
            const __half* xh = reinterpret_cast<const __half*>(&v[u]);
            const float4 a0 = *reinterpret_cast<const float4*>(ab + vc * 8);
            const float4 a1 = *reinterpret_cast<const float4*>(ab + vc * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(ab + C + vc * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(ab + C + vc * 8 + 4);
            float f[8] = {fmaf(__half2float(xh[0]), a0.x, b0.x), fmaf(__half2float(xh[1]), a0.y, b0.y),
                          fmaf(__half2float(xh[2]), a0.z, b0.z), fmaf(__half2float(xh[3]), a0.w, b0.w),
                          fmaf(__half2float(xh[4]), a1.x, b1.x), fmaf(__half2float(xh[5]), a1.y, b1.y),
                          fmaf(__half2float(xh[6]), a1.z, b1.z), fmaf(__half2float(xh[7]), a1.w, b1.w)};
            uint4 o;
            __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float u0 = f[2 * e], u1 = f[2 * e + 1];
                if (silu) {
                    u0 = silu_f(u0);
                    u1 = silu_f(u1);
                }
                oh[e] = __floats2half2_rn(u0, u1);
            }
            *reinterpret_cast<uint4*>(y + (base_row + r0 + rr[u]) * ldy + vc * 8) = o;
        }
    }
}

// one warp per row; C <= 2048, C % 8 == 0
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, long long ldx,
                                                        __half* __restrict__ y, long long ldy, long long rows, int C,
                                                        const __half* __restrict__ gamma,
                                                        const __half* __restrict__ beta, float eps) {
    griddep_wait();
    griddep_launch_small();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int C8 = C >> 3;
    constexpr int MAXV = 8;
    uint4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vc = lane + k * 32;
        if (vc < C8) {
            v[k] = __ldg(reinterpret_cast<const uint4*>(x + row * ldx + vc * 8));
            const __half2* h2 = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                s += f.x + f.y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vc = lane + k * 32;
        if (vc < C8) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                q += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vc = lane + k * 32;
        if (vc < C8) {
            const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gamma + vc * 8));
            const uint4 bv = __ldg(reinterpret_cast<const uint4*>(beta + vc * 8));
            const __half* xh = reinterpret_cast<const __half*>(&v[k]);
            const __half* gh = reinterpret_cast<const __half*>(&gv);
            const __half* bh = reinterpret_cast<const __half*>(&bv);
            uint4 o;
            __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
            for (int e = 0; e < 8; ++e)
                oh[e] = __float2half_rn((__half2float(xh[e]) - mean) * rstd * __half2float(gh[e]) + __half2float(bh[e]));
            *reinterpret_cast<uint4*>(y + row * ldy + vc * 8) = o;
        }
    }
}

__global__ void __launch_bounds__(256) ln_rowstats_kernel(const __half* __restrict__ x, long long ldx, long long rows, int C,
                                                          float eps, float2* __restrict__ out) {
    griddep_wait();
    griddep_launch_small();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int C8 = C >> 3;
    constexpr int MAXV = 8;
    uint4 v[MAXV];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vc = lane + k * 32;
        if (vc < C8) {
            v[k] = __ldg(reinterpret_cast<const uint4*>(x + row * ldx + vc * 8));
            const __half2* h2 = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                s += f.x + f.y;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < MAXV; ++k) {
        const int vc = lane + k * 32;
        if (vc < C8) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(h2[e]);
                q += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    if (lane == 0) out[row] = make_float2(mean, rsqrtf(q / C + eps));
}

}  // namespace

int gn_rows_per_chunk(int rows_per_inst, int n_inst, int num_sms) {
    // ~2 blocks per SM: fat blocks amortise the per-block latency chain (few dependent loads per thread otherwise)
    long long want = static_cast<long long>(num_sms) * 2;
    long long chunks_per_inst = (want + n_inst - 1) / n_inst;
    if (chunks_per_inst < 1) chunks_per_inst = 1;
    long long rpc = (rows_per_inst + chunks_per_inst - 1) / chunks_per_inst;
    if (rpc < 8) rpc = 8;
    rpc = (rpc + 7) / 8 * 8;
    return static_cast<int>(rpc);
}

size_t gn_workspace_bytes(int rows_per_inst, int n_inst, int num_sms) {
    const int rpc = gn_rows_per_chunk(rows_per_inst, n_inst, num_sms);
    const int nchunks = (rows_per_inst + rpc - 1) / rpc;
    // counters (fixed-size region at the START: their location must not depend on the call's shape, they have to
    // stay zero between launches) + stats + partials
    return kCounterBytes + static_cast<size_t>(n_inst) * kGroups * sizeof(float2) +
           static_cast<size_t>(n_inst) * nchunks * kGroups * sizeof(float2) + 256;
}

int groupnorm_silu(const __half* x, long long ldx, __half* y, long long ldy, long long rows, int C, int rows_per_inst,
                   const __half* gamma, const __half* beta, float eps, int silu, void* workspace, int num_sms,
                   cudaStream_t stream, int phase) {
    if (C % 32 != 0 || C % 8 != 0 || rows % rows_per_inst != 0) return -1;
    const int n_inst = static_cast<int>(rows / rows_per_inst);
    const int rpc = gn_rows_per_chunk(rows_per_inst, n_inst, num_sms);
    const int nchunks = (rows_per_inst + rpc - 1) / rpc;
    if (static_cast<size_t>(n_inst) * sizeof(unsigned int) > kCounterBytes) return -3;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    float2* stats = reinterpret_cast<float2*>(ws + kCounterBytes);
    float2* partial = stats + static_cast<size_t>(n_inst) * kGroups;
    if (phase != 2)
        launch_pdl(gn_stats_kernel, dim3(nchunks, n_inst), kStatsThreads, 2 * C * sizeof(float), stream, 
            x, ldx, C, rows_per_inst, rpc, nchunks, eps, partial, counters, stats);
    if (phase == 1) return cudaGetLastError() == cudaSuccess ? 0 : -2;
    // rows per apply block: ~4 blocks per SM overall, at least 4 rows
    long long want_blocks = static_cast<long long>(num_sms) * 4;
    long long per_inst = (want_blocks + n_inst - 1) / n_inst;
    if (per_inst < 1) per_inst = 1;
    long long rpb = (rows_per_inst + per_inst - 1) / per_inst;
    if (rpb < 4) rpb = 4;
    const int nblk = static_cast<int>((rows_per_inst + rpb - 1) / rpb);
    launch_pdl(gn_apply_kernel, dim3(nblk, n_inst), 256, 2 * C * sizeof(float), stream, x, ldx, y, ldy, C, rows_per_inst,
                                                                              static_cast<int>(rpb), stats, gamma, beta, silu);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int layernorm_rowstats(const __half* x, long long ldx, long long rows, int C, float eps, float2* out, cudaStream_t stream) {
    if (C % 8 != 0 || C > 2048) return -1;
    launch_pdl(ln_rowstats_kernel, static_cast<unsigned int>((rows + 7) / 8), 256, 0, stream, x, ldx, rows, C, eps, out);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int layernorm(const __half* x, long long ldx, __half* y, long long ldy, long long rows, int C, const __half* gamma,
              const __half* beta, float eps, cudaStream_t stream) {
    if (C % 8 != 0 || C > 2048) return -1;
    const long long blocks = (rows + 7) / 8;
    launch_pdl(layernorm_kernel, static_cast<unsigned int>(blocks), 256, 0, stream, x, ldx, y, ldy, rows, C, gamma, beta, eps);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // namespace t2v
