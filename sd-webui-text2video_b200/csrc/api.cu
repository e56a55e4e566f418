// C-ABI glue: library state + the kernel-level entry points of include/t2v_b200.h.
#include "../../include/t2v_b200.h"
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"

#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>

namespace t2v {

static char g_err[512] = "";
static int g_device = -1;
static int g_num_sms = 148;
static void* g_gn_ws = nullptr;
static size_t g_gn_ws_bytes = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int num_sms() { return g_num_sms; }
void clear_pending_error(const char* where) {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        fprintf(stderr, "[t2v_b200] %s: clearing a pending CUDA error left by an earlier call: %s (%s)\n", where, cudaGetErrorString(e),
                cudaGetErrorName(e));
}
int launch_status(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return -2;
}
bool pdl_enabled() {
    static const bool on = getenv("T2V_PDL") != nullptr;      // opt-in: measured 2.7 % SLOWER on the graphed forward
    return on;
}

static void* gn_scratch(size_t bytes) {
    if (bytes > g_gn_ws_bytes) {
        if (g_gn_ws) cudaFree(g_gn_ws);
        g_gn_ws = nullptr;
        if (cudaMalloc(&g_gn_ws, bytes) != cudaSuccess) return nullptr;
        cudaMemset(g_gn_ws, 0, bytes);
        g_gn_ws_bytes = bytes;
    }
    return g_gn_ws;
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int t2v_init(int device) {
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed");
        return -1;
    }
    if (prop.major != 10) {
        set_error("t2v_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return -2;
    }
    g_device = device;
    g_num_sms = prop.multiProcessorCount;
    if (gemm_init() != 0) {
        set_error("gemm_init failed (driver entry point / smem attribute)");
        return -3;
    }
    return 0;
}
const char* t2v_last_error(void) { return g_err; }
int t2v_num_sms(void) { return g_num_sms; }
const char* t2v_version(void) { return "t2v_b200 0.1 (sm_100a; tcgen05+TMA implicit GEMM)"; }

int t2v_op_gemm(const void* a, long long lda, int K, int nd, const int* dims, int ntaps, const int* tap_off,
                const void* w_packed, int n_alloc, int N, int b_batch_dim, int flags, void* out, long long ldo,
                const void* bias, int bias_rows, long long bias_stride, const void* residual, long long ldr,
                float alpha, int force_bn, int force_cg, void* stream) {
    GemmProblem p;
    memset(&p, 0, sizeof(p));
    p.a = reinterpret_cast<const __half*>(a);
    p.lda = lda;
    p.K = K;
    p.nd = nd;
    for (int d = 0; d < nd; ++d) p.dim[d] = dims[d];
    p.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t)
        for (int d = 0; d < nd; ++d) p.tap_off[t][d] = tap_off ? tap_off[t * nd + d] : 0;
    p.b = reinterpret_cast<const __half*>(w_packed);
    p.n_alloc = n_alloc;
    p.N = N;
    p.b_batch_dim = b_batch_dim;
    p.flags = flags & ~(GEMM_DBG_FORCE_BS | GEMM_DBG_NO_BS);
    p.force_bs = (flags & GEMM_DBG_FORCE_BS) ? 1 : ((flags & GEMM_DBG_NO_BS) ? -1 : 0);
    p.out = out;
    p.ldo = ldo;
    p.bias = reinterpret_cast<const __half*>(bias);
    p.bias_rows = bias_rows;
    p.bias_stride = bias_stride;
    p.residual = reinterpret_cast<const __half*>(residual);
    p.ldr = ldr;
    p.alpha = alpha;
    p.force_bn = force_bn;
    p.force_cg = force_cg;
    GemmPlan plan;
    int rc = gemm_plan(p, &plan, g_num_sms);
    if (rc != 0) {
        set_error("gemm_plan failed (%d)", rc);
        return rc;
    }
    rc = gemm_launch(plan, reinterpret_cast<cudaStream_t>(stream));
    if (rc != 0) set_error("gemm_launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int t2v_latent_blend(const float* image_latents, int image_frames, const double* noise, const double* weights, double* out,
                     double* mask_out, int BC, int F, long long hw, void* stream) {
    if ((image_frames != 1 && image_frames != F) || BC < 1 || F < 1 || hw < 1) {
        set_error("latent_blend: image_frames must be 1 or F");
        return -1;
    }
    return latent_blend(image_latents, image_frames, noise, weights, out, mask_out, BC, F, hw, reinterpret_cast<cudaStream_t>(stream));
}

int t2v_op_pack_conv_weight(const void* src, int src_is_f32, void* dst, int Cout, int Cin, int taps, int n_alloc,
                            int k_alloc, void* stream) {
    return pack_conv_weight(src, src_is_f32, reinterpret_cast<__half*>(dst), Cout, Cin, taps, n_alloc, k_alloc,
                            reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_pack_geglu_weight(const void* w, const void* b, int src_is_f32, void* wdst, void* bdst, int H, int K, int bn,
                             void* stream) {
    return pack_geglu_weight(w, b, src_is_f32, reinterpret_cast<__half*>(wdst), reinterpret_cast<__half*>(bdst), H, K, bn,
                             reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_groupnorm(const void* x, long long ldx, void* y, long long ldy, long long rows, int C, int rows_per_inst,
                     const void* gamma, const void* beta, float eps, int silu, void* stream) {
    const int n_inst = static_cast<int>(rows / rows_per_inst);
    void* ws = gn_scratch(gn_workspace_bytes(rows_per_inst, n_inst, g_num_sms));
    if (!ws) {
        set_error("groupnorm workspace allocation failed");
        return -1;
    }
    return groupnorm_silu(reinterpret_cast<const __half*>(x), ldx, reinterpret_cast<__half*>(y), ldy, rows, C,
                          rows_per_inst, reinterpret_cast<const __half*>(gamma), reinterpret_cast<const __half*>(beta),
                          eps, silu, ws, g_num_sms, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_layernorm(const void* x, long long ldx, void* y, long long ldy, long long rows, int C, const void* gamma,
                     const void* beta, float eps, void* stream) {
    return layernorm(reinterpret_cast<const __half*>(x), ldx, reinterpret_cast<__half*>(y), ldy, rows, C,
                     reinterpret_cast<const __half*>(gamma), reinterpret_cast<const __half*>(beta), eps,
                     reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_attention(const void* q, const void* k, const void* v, void* o, long long q_bs, long long q_ss,
                     long long k_bs, long long k_ss, long long v_bs, long long v_ss, long long o_bs, long long o_ss,
                     int batch, int heads, int sq, int skv, int kv_batch_div, float scale, void* stream) {
    AttnParams p;
    p.q = reinterpret_cast<const __half*>(q);
    p.k = reinterpret_cast<const __half*>(k);
    p.v = reinterpret_cast<const __half*>(v);
    p.o = reinterpret_cast<__half*>(o);
    p.q_bs = q_bs; p.q_ss = q_ss; p.k_bs = k_bs; p.k_ss = k_ss; p.v_bs = v_bs; p.v_ss = v_ss; p.o_bs = o_bs; p.o_ss = o_ss;
    p.batch = batch; p.heads = heads; p.sq = sq; p.skv = skv; p.head_dim = 64; p.kv_batch_div = kv_batch_div;
    p.scale = scale; p.b_inner = 1; p.q_bsi = p.k_bsi = p.v_bsi = p.o_bsi = 0;
    return attention(p, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_attention_hd(const void* q, const void* k, const void* v, void* o, long long q_bs, long long q_ss,
                        long long k_bs, long long k_ss, long long v_bs, long long v_ss, long long o_bs, long long o_ss,
                        int batch, int heads, int head_dim, int sq, int skv, int kv_batch_div, float scale, void* stream) {
    AttnParams p;
    memset(&p, 0, sizeof(p));
    p.q = reinterpret_cast<const __half*>(q);
    p.k = reinterpret_cast<const __half*>(k);
    p.v = reinterpret_cast<const __half*>(v);
    p.o = reinterpret_cast<__half*>(o);
    p.q_bs = q_bs; p.q_ss = q_ss; p.k_bs = k_bs; p.k_ss = k_ss; p.v_bs = v_bs; p.v_ss = v_ss; p.o_bs = o_bs; p.o_ss = o_ss;
    p.batch = batch; p.heads = heads; p.sq = sq; p.skv = skv; p.head_dim = head_dim; p.kv_batch_div = kv_batch_div;
    p.scale = scale; p.b_inner = 1;
    return head_dim == 64 ? attention(p, reinterpret_cast<cudaStream_t>(stream)) : attention_hd(p, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_attention_relpos(const void* q, const void* k, const void* v, void* o, const void* table_k, const void* table_v,
                            long long n_seq, long long seq_inner, long long bs_outer, long long bs_inner, long long ss,
                            long long o_bs_outer, long long o_bs_inner, long long o_ss, int heads, int head_dim, int T,
                            int max_rel, float scale, void* stream) {
    RelposParams p;
    memset(&p, 0, sizeof(p));
    p.q = reinterpret_cast<const __half*>(q);
    p.k = reinterpret_cast<const __half*>(k);
    p.v = reinterpret_cast<const __half*>(v);
    p.o = reinterpret_cast<__half*>(o);
    p.table_k = reinterpret_cast<const __half*>(table_k);
    p.table_v = reinterpret_cast<const __half*>(table_v);
    p.n_seq = n_seq; p.seq_inner = seq_inner; p.bs_outer = bs_outer; p.bs_inner = bs_inner; p.ss = ss;
    p.o_bs_outer = o_bs_outer; p.o_bs_inner = o_bs_inner; p.o_ss = o_ss;
    p.heads = heads; p.head_dim = head_dim; p.T = T; p.max_rel = max_rel; p.scale = scale;
    return attention_relpos(p, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_upsample2x(const void* x, void* y, int nframes, int h, int w, int C, void* stream) {
    return upsample2x(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(y), nframes, h, w, C,
                      reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_im2col_s2(const void* x, void* col, int nframes, int h, int w, int C, void* stream) {
    return im2col_s2(reinterpret_cast<const __half*>(x), reinterpret_cast<__half*>(col), nframes, h, w, C,
                     reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_time_sinusoid(const float* t, void* out, int B, int dim, void* stream) {
    return time_sinusoid(t, reinterpret_cast<__half*>(out), B, dim, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_op_small_linear(const void* x, long long ldx, const void* W, const void* bias, const void* addend, void* y,
                        long long ldy, int B, int N, int K, int silu_in, void* stream) {
    return small_linear(reinterpret_cast<const __half*>(x), ldx, reinterpret_cast<const __half*>(W),
                        reinterpret_cast<const __half*>(bias), reinterpret_cast<const __half*>(addend),
                        reinterpret_cast<__half*>(y), ldy, B, N, K, silu_in, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_ddim_step(const float* x, const void* eps_c, const void* eps_u, int eps_is_f32, float* x_out, long long n, long long chan_stride,
                  int C, int guided_channels, float g, int mode, float a0, float a1, float a2, float a3, float a4,
                  const float* noise, int cfg_fp16, void* stream) {
    DdimStepParams p;
    p.x = x; p.eps_c = eps_c; p.eps_u = eps_u; p.eps_is_f32 = eps_is_f32;
    p.x_out = x_out; p.n = n; p.chan_stride = chan_stride; p.C = C; p.guided_channels = guided_channels; p.g = g;
    p.mode = mode; p.a0 = a0; p.a1 = a1; p.a2 = a2; p.a3 = a3; p.a4 = a4; p.noise = noise; p.cfg_fp16 = cfg_fp16;
    return ddim_step(p, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_cfg_x0(const float* x, const void* eps_c, const void* eps_u, int eps_is_f32, float* x0, long long n, float g, float alpha,
               float sigma, int cfg_fp16, void* stream) {
    return cfg_x0(x, eps_c, eps_u, eps_is_f32, x0, n, g, alpha, sigma,
                  cfg_fp16, reinterpret_cast<cudaStream_t>(stream));
}
int t2v_lincomb(float* out, const float* const* src, const float* coef, int n_src, long long n, void* stream) {
    return lincomb(out, src, coef, n_src, n, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
