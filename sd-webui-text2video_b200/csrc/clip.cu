// OpenCLIP ViT-H-14 text tower as a pre-planned launch list -- the step right before the denoising loop.
// Replaces FrozenOpenCLIPEmbedder.encode_with_transformer (modelscope/clip_hardcode.py:112-119, :269-274) on top of
// open_clip's TextTransformer: token embedding + positional embedding, `layers_run` pre-LN residual attention blocks
// (nn.MultiheadAttention with the causal mask, GELU MLP; the reference stops one block early: layer = 'penultimate'),
// ln_final.  open_clip is NOT under /root/reference (pip dependency, unpinned by the reference's requirements.txt): the
// architecture restated here is open_clip's published ResidualAttentionBlock, pinned in tests against torch's own
// nn.MultiheadAttention / nn.LayerNorm modules.
//
// Same engine as the denoiser: LayerNorm folded into the consuming tcgen05 GEMM, residual adds in the GEMM epilogue; the
// 77-token causal attention and the GELU are small dedicated kernels (the tower runs once per prompt: 2 x 77 rows).
#include "../../include/t2v_b200.h"
#include "runtime.cuh"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>

using namespace t2v;

struct t2v_clip {
    t2v_clip_config cfg;
    ParamStore params;
    std::map<int, std::unique_ptr<Plan>> plans;       // key: batch
    std::map<Plan*, std::pair<int*, __half*>> io;      // plan -> (token staging, output tokens)
};

namespace t2v {
namespace {

// x[b, l, :] = token_embedding[tokens[b, l], :] + positional_embedding[l, :]
__global__ void clip_embed_kernel(const int* __restrict__ tokens, const __half* __restrict__ emb, const __half* __restrict__ pos,
                                  __half* __restrict__ x, int rows, int L, int W, int vocab) {
    griddep_wait();
    griddep_launch_small();
    const int W8 = W >> 3;
    const long long n = static_cast<long long>(rows) * W8;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / W8), vc = static_cast<int>(i - static_cast<long long>(r) * W8);
        int tok = tokens[r];
        tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
        const uint4 e = __ldg(reinterpret_cast<const uint4*>(emb + static_cast<long long>(tok) * W + vc * 8));
        const uint4 p = __ldg(reinterpret_cast<const uint4*>(pos + static_cast<long long>(r % L) * W + vc * 8));
        const __half2* eh = reinterpret_cast<const __half2*>(&e);
        const __half2* ph = reinterpret_cast<const __half2*>(&p);
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 a = __half22float2(eh[k]), b = __half22float2(ph[k]);
            oh[k] = __floats2half2_rn(a.x + b.x, a.y + b.y);
        }
        *reinterpret_cast<uint4*>(x + static_cast<long long>(r) * W + vc * 8) = o;
    }
}

// y = x * Phi(x) (nn.GELU, erf form), in place on [rows, C] fp16
__global__ void gelu_kernel(__half* __restrict__ x, long long n8) {
    griddep_wait();
    griddep_launch_small();
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n8; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        uint4 v = reinterpret_cast<uint4*>(x)[i];
        __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = __half22float2(h[k]);
            h[k] = __floats2half2_rn(0.5f * f.x * (1.0f + erff(f.x * 0.70710678118654752440f)),
                                     0.5f * f.y * (1.0f + erff(f.y * 0.70710678118654752440f)));
        }
        reinterpret_cast<uint4*>(x)[i] = v;
    }
}

// Causal self-attention over one short sequence per (sample, head): softmax(q k^T / sqrt(d) + causal mask) v, d = 64,
// L <= 128.  qkv [rows, 3W] as nn.MultiheadAttention's in_proj lays it out (q | k | v, head h at columns h*64 of each part).
// One block per (sample, head): K and V of the head in shared memory (rows padded to 66 halves: conflict-free column reads),
// one warp per query row; lanes own keys l, l+32, ... for the scores and output channels l, l+32 for P V.
constexpr int kClipD = 64;
__global__ void __launch_bounds__(128) clip_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ o, int L, int W, int heads) {
    griddep_wait();
    griddep_launch_small();
    extern __shared__ __half sm_kv[];
    const int b = blockIdx.x / heads, hd = blockIdx.x % heads;
    __half* sk = sm_kv;                         // [L][66]
    __half* sv = sm_kv + L * 66;                // [L][66]
    __shared__ float sq[4][kClipD];
    __shared__ float sp[4][128];
    const long long row0 = static_cast<long long>(b) * L;
    const int ld = 3 * W;
    for (int i = threadIdx.x; i < L * (kClipD / 2); i += blockDim.x) {
        const int r = i / (kClipD / 2), c2 = i % (kClipD / 2);
        const __half2 kk = *reinterpret_cast<const __half2*>(qkv + (row0 + r) * ld + W + hd * kClipD + c2 * 2);
        const __half2 vv = *reinterpret_cast<const __half2*>(qkv + (row0 + r) * ld + 2 * W + hd * kClipD + c2 * 2);
        *reinterpret_cast<__half2*>(sk + r * 66 + c2 * 2) = kk;
        *reinterpret_cast<__half2*>(sv + r * 66 + c2 * 2) = vv;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = 0.125f;                 // 64^-0.5
    for (int qi = warp; qi < L; qi += 4) {
        const __half* qp = qkv + (row0 + qi) * ld + hd * kClipD;
        sq[warp][lane] = __half2float(qp[lane]) * scale;             // nn.MultiheadAttention scales q before q k^T
        sq[warp][lane + 32] = __half2float(qp[lane + 32]) * scale;
        __syncwarp();
        float s[4];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int key = lane + 32 * j;
            s[j] = -INFINITY;
            if (key <= qi && key < L) {                               // causal: keys up to and including the query position
                float acc = 0.f;
                for (int d = 0; d < kClipD; ++d) acc = fmaf(sq[warp][d], __half2float(sk[key * 66 + d]), acc);
                s[j] = acc;
            }
            mx = fmaxf(mx, s[j]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float e = s[j] == -INFINITY ? 0.f : __expf(s[j] - mx);
            s[j] = e;
            sum += e;
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int j = 0; j < 4; ++j) sp[warp][lane + 32 * j] = s[j] * inv;
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int key = 0; key <= qi; ++key) {
            const float p = sp[warp][key];
            o0 = fmaf(p, __half2float(sv[key * 66 + lane]), o0);
            o1 = fmaf(p, __half2float(sv[key * 66 + lane + 32]), o1);
        }
        __half* op = o + (row0 + qi) * W + hd * kClipD;
        op[lane] = __float2half_rn(o0);
        op[lane + 32] = __float2half_rn(o1);
        __syncwarp();
    }
}

void expect_params(t2v_clip* m) {
    ParamStore& P = m->params;
    const t2v_clip_config& c = m->cfg;
    P.expect("token_embedding.weight", {c.vocab, c.width});
    P.expect("positional_embedding", {c.context, c.width});
    for (int i = 0; i < c.layers_run; ++i) {
        const std::string p = "transformer.resblocks." + std::to_string(i);
        P.expect(p + ".ln_1.weight", {c.width});
        P.expect(p + ".ln_1.bias", {c.width});
        P.expect(p + ".attn.in_proj_weight", {3 * c.width, c.width});
        P.expect(p + ".attn.in_proj_bias", {3 * c.width});
        P.expect(p + ".attn.out_proj.weight", {c.width, c.width});
        P.expect(p + ".attn.out_proj.bias", {c.width});
        P.expect(p + ".ln_2.weight", {c.width});
        P.expect(p + ".ln_2.bias", {c.width});
        P.expect(p + ".mlp.c_fc.weight", {4 * c.width, c.width});
        P.expect(p + ".mlp.c_fc.bias", {4 * c.width});
        P.expect(p + ".mlp.c_proj.weight", {c.width, 4 * c.width});
        P.expect(p + ".mlp.c_proj.bias", {c.width});
    }
    P.expect("ln_final.weight", {c.width});
    P.expect("ln_final.bias", {c.width});
}

int build(t2v_clip* m, Plan* plan, Arena* arena, bool dry, cudaStream_t stream, int B, int** tok_out, __half** out_tok) {
    Builder bld(plan, arena, dry, num_sms());
    NetCtx c{&m->params, &bld, stream, nullptr};
    const t2v_clip_config& cfg = m->cfg;
    const int L = cfg.context, W = cfg.width, heads = cfg.heads;
    const long long R = static_cast<long long>(B) * L;
    int* tokens = reinterpret_cast<int*>(bld.alloc_bytes(static_cast<size_t>(R) * sizeof(int)));
    *tok_out = tokens;
    Tok x = bld.alloc(R, W);
    {
        const __half* emb = prm(c, "token_embedding.weight");
        const __half* pos = prm(c, "positional_embedding");
        const Tok xx = x;
        const int vocab = cfg.vocab;
        bld.step([=](cudaStream_t s) {
            launch_pdl(clip_embed_kernel, dim3(static_cast<unsigned>((R * (W / 8) + 255) / 256)), dim3(256), 0, s, tokens, emb, pos, xx.p,
                       static_cast<int>(R), L, W, vocab);
            return launch_status("clip launch");
        }, 1, STEP_OTHER, 0.0, "clip embed");
    }
    for (int i = 0; i < cfg.layers_run; ++i) {
        const std::string p = "transformer.resblocks." + std::to_string(i);
        // x = x + out_proj(attention(in_proj(ln_1(x))))
        Tok qkv = ln_linear(c, x, p + ".ln_1", p + ".attn.in_proj", prm(c, p + ".attn.in_proj_weight"), prm(c, p + ".attn.in_proj_bias"),
                            3 * W, nullptr);
        Tok o = bld.alloc(R, W);
        {
            const Tok q = qkv, oo = o;
            const size_t smem = static_cast<size_t>(2) * L * 66 * sizeof(__half);
            bld.step([=](cudaStream_t s) {
                launch_pdl(clip_attention_kernel, dim3(static_cast<unsigned>(B * heads)), dim3(128), smem, s, q.p, oo.p, L, W, heads);
                return launch_status("clip launch");
            }, 1, STEP_ATTN, 4.0 * B * heads * static_cast<double>(L) * L * kClipD / 2, "clip causal attention");
        }
        bld.free(qkv);
        Tok y = linear(c, o, prm(c, p + ".attn.out_proj.weight"), W, prm(c, p + ".attn.out_proj.bias"), &x);
        bld.free(o);
        bld.free(x);
        x = y;
        // x = x + c_proj(gelu(c_fc(ln_2(x))))
        Tok h = ln_linear(c, x, p + ".ln_2", p + ".mlp.c_fc", prm(c, p + ".mlp.c_fc.weight"), prm(c, p + ".mlp.c_fc.bias"), 4 * W, nullptr);
        {
            const Tok hh = h;
            const long long n8 = R * (4 * W) / 8;
            bld.step([=](cudaStream_t s) {
                launch_pdl(gelu_kernel, dim3(static_cast<unsigned>((n8 + 255) / 256)), dim3(256), 0, s, hh.p, n8);
                return launch_status("clip launch");
            }, 1, STEP_OTHER, 0.0, "clip gelu");
        }
        Tok y2 = linear(c, h, prm(c, p + ".mlp.c_proj.weight"), W, prm(c, p + ".mlp.c_proj.bias"), &x);
        bld.free(h);
        bld.free(x);
        x = y2;
    }
    Tok z = layer_norm(c, x, "ln_final");
    bld.free(x);
    *out_tok = z.p;
    return bld.error;
}

Plan* get_plan(t2v_clip* m, int B, cudaStream_t stream) {
    auto it = m->plans.find(B);
    if (it != m->plans.end() && it->second->weights_version == m->params.version()) return it->second.get();
    if (it != m->plans.end()) {
        m->io.erase(it->second.get());
        m->plans.erase(it);
    }
    std::string miss;
    if (m->params.missing(&miss) > 0) {
        set_error("CLIP text tower parameters missing (e.g. '%s')", miss.c_str());
        return nullptr;
    }
    std::unique_ptr<Plan> plan(new Plan());
    Arena arena;
    int* tok = nullptr;
    __half* out = nullptr;
    {
        Plan scratch;
        arena.reset(nullptr, false);
        if (build(m, &scratch, &arena, true, stream, B, &tok, &out) != 0) return nullptr;
    }
    const size_t bytes = arena.peak() + (1 << 20);
    if (cudaMalloc(&plan->slab, bytes) != cudaSuccess) {
        set_error("CLIP activation slab cudaMalloc(%zu MB) failed", bytes >> 20);
        return nullptr;
    }
    plan->slab_bytes = bytes;
    arena.reset(plan->slab, false);
    if (build(m, plan.get(), &arena, false, stream, B, &tok, &out) != 0) return nullptr;
    plan->weights_version = m->params.version();
    Plan* raw = plan.get();
    m->io[raw] = {tok, out};
    m->plans[B] = std::move(plan);
    return raw;
}

__global__ void clip_out_kernel(const __half* __restrict__ z, void* __restrict__ out, int out_is_f32, long long n) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        if (out_is_f32) reinterpret_cast<float*>(out)[i] = __half2float(z[i]);
        else reinterpret_cast<__half*>(out)[i] = z[i];
    }
}

}  // namespace
}  // namespace t2v

extern "C" {

int t2v_clip_create(const t2v_clip_config* cfg, t2v_clip** out) {
    if (!cfg || !out) return -1;
    if (cfg->width % 64 != 0 || cfg->width / cfg->heads != 64 || cfg->context > 128 || cfg->layers_run < 1 || cfg->vocab < 1) {
        set_error("CLIP text tower: head width must be 64 (width / heads), context <= 128");
        return -2;
    }
    t2v_clip* m = new t2v_clip();
    m->cfg = *cfg;
    expect_params(m);
    *out = m;
    return 0;
}

void t2v_clip_destroy(t2v_clip* m) { delete m; }

int t2v_clip_set_param(t2v_clip* m, const char* name, const void* data, int dtype, int ndim, const int64_t* shape, void* stream) {
    return m->params.set(name, data, dtype, ndim, shape, reinterpret_cast<cudaStream_t>(stream));
}

int t2v_clip_param_info(t2v_clip* m, int index, char* name_out, size_t name_cap, int64_t* shape_out, int* ndim_out) {
    std::string name;
    std::vector<long long> shape;
    const int n = m->params.info(index, &name, &shape);
    if (n < 0) return -1;
    if (name_out && name_cap > 0) {
        strncpy(name_out, name.c_str(), name_cap - 1);
        name_out[name_cap - 1] = 0;
    }
    if (ndim_out) *ndim_out = static_cast<int>(shape.size());
    if (shape_out)
        for (size_t i = 0; i < shape.size() && i < 8; ++i) shape_out[i] = shape[i];
    return n;
}

int t2v_clip_encode(t2v_clip* m, const int* tokens, void* out, int out_is_f32, int B, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    Plan* plan = get_plan(m, B, stream);
    if (!plan) return -1;
    const auto& io = m->io[plan];
    const long long R = static_cast<long long>(B) * m->cfg.context;
    cudaMemcpyAsync(io.first, tokens, static_cast<size_t>(R) * sizeof(int), cudaMemcpyDeviceToDevice, stream);
    const int rc = run_plan(plan, stream, true);
    if (rc != 0) {
        set_error("CLIP launch failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
        return rc;
    }
    const long long n = R * m->cfg.width;
    clip_out_kernel<<<static_cast<unsigned>((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, stream>>>(io.second, out, out_is_f32, n);
    return launch_status("clip launch");
}

}  // extern "C"
