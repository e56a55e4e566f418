/* t2v_b200 -- C ABI of the B200-native text2video denoising path.
 *
 * The reference (kabachuha/sd-webui-text2video) has NO FFI boundary: its hot path is plain Python that calls
 * PyTorch library kernels.  This header is the boundary a maintainer would bind instead (ctypes stubs in
 * INTEGRATION.md); every entry point names the reference call it replaces (paths relative to
 * /root/reference/scripts).
 *
 * Conventions (mirroring the Python contract):
 *   - plain pointers and sizes only; device pointers unless stated otherwise; the CALLER owns every tensor passed
 *     in, the library owns only its packed-weight and workspace arenas (freed by *_destroy)
 *   - all work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*), no hidden synchronisation
 *   - return value 0 = success, negative = error (t2v_last_error() gives the text); the Python layer raises
 *     RuntimeError so failures surface as exceptions exactly like the reference (t2v_helpers/render.py:35-37)
 *   - a handle is bound to one device and is not thread-safe (the webui serialises callers, text2vid.py:82)
 */
#ifndef T2V_B200_H
#define T2V_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ library */
int t2v_init(int device);                 /* selects the device, resolves driver entry points, sizes grids */
const char* t2v_last_error(void);
int t2v_num_sms(void);
const char* t2v_version(void);

/* ------------------------------------------------------------------------------------------ denoiser
 * replaces modelscope/t2v_model.py::UNetSD (ctor :98-326, forward :386-501).                            */
typedef struct t2v_unet t2v_unet;

typedef struct {
    int in_dim, dim, context_dim, out_dim;
    int dim_mult[8];
    int n_mult;
    int num_heads;        /* heads of the stem TemporalTransformer (:171-179) */
    int head_dim;         /* must be 64 */
    int num_res_blocks;
    float attn_scales[8];
    int n_attn_scales;
    int arch;              /* 0: ModelScope UNetSD (modelscope/t2v_model.py:98-501)
                            * 1: VideoCrafter UNetModel (videocrafter/lvdm/models/modules/openaimodel3d.py:281-670; dim =
                            *    model_channels, dim_mult = channel_mult, attn_scales = 1/attention_resolutions, head width =
                            *    channels / num_heads (head_dim ignored), state_dict keys of `model.diffusion_model.*`) */
    int temporal_length;   /* arch 1: RelativePosition(max_relative_position = temporal_length), tables [2*L+1, d] */
} t2v_unet_config;

int t2v_unet_create(const t2v_unet_config* cfg, t2v_unet** out);
void t2v_unet_destroy(t2v_unet* u);
/* Hands one parameter of the reference state_dict (key names of SURVEY.md appendix D, e.g.
 * "input_blocks.1.0.in_layers.2.weight") to the library, which packs it into its own layout.
 * `data` is a DEVICE pointer to a contiguous tensor of `dtype` (0 = fp16, 1 = fp32) with `ndim` dims.
 * Replaces load_state_dict(strict=True) at modelscope/t2v_pipeline.py:95-101 (and is what the LoRA merger's
 * re-assigned .weight tensors are re-sent through, stable_lora/scripts/lora_processor.py:236-242).        */
int t2v_unet_set_param(t2v_unet* u, const char* name, const void* data, int dtype, int ndim, const int64_t* shape,
                       void* stream);
/* Number of parameters still missing (0 => ready); fills `name_out` with one missing key if non-null. */
int t2v_unet_missing_params(t2v_unet* u, char* name_out, size_t name_cap);
/* Enumerates the expected state_dict: index in [0, count) -> key name + shape (shape has room for 8 dims).
 * Returns the number of expected parameters, or -1 if `index` is out of range.  The Python mirror builds its
 * nn.Module tree (same names, nn.Linear / nn.Conv2d / nn.Conv3d leaves) from this list.                     */
int t2v_unet_param_info(t2v_unet* u, int index, char* name_out, size_t name_cap, int64_t* shape_out, int* ndim_out);
/* eps = UNetSD.forward(x, t, y) (t2v_model.py:386-459).
 *   x   [B, in_dim, F, h, w]  fp32 (x_is_f32 = 1) or fp16, NCFHW exactly as the samplers hold the latent
 *   t   [B] float32 (host converts int64 timesteps; UniPC already passes floats, uni_pc.py:248)
 *   ctx [B, L, context_dim] fp16
 *   out [B, out_dim, F, h, w] fp16 (out_is_f32 = 0) or fp32                                               */
int t2v_unet_forward(t2v_unet* u, const void* x, int x_is_f32, const float* t, const void* ctx, void* out,
                     int out_is_f32, int B, int F, int h, int w, int L, void* stream);
/* 2*MAC flop count of one forward at this shape (for roofline reporting). */
double t2v_unet_flops(t2v_unet* u, int B, int F, int h, int w, int L);
int t2v_unet_num_launches(t2v_unet* u);
/* Measurement aid: replays the plan of this shape once (inputs = whatever the last forward left in the staging
 * buffers) with a CUDA-event pair around every launch on `stream` and sums per kernel family:
 *   out[3k + 0] = milliseconds, out[3k + 1] = algorithmic flop, out[3k + 2] = launches, k = 0 implicit-GEMM (tcgen05),
 *   1 attention, 2 group/layer norm, 3 glue; out[12] = total ms.  Synchronises the stream (bench/tests only).   */
int t2v_unet_profile(t2v_unet* u, int B, int F, int h, int w, int L, void* stream, double* out13);
/* copies an internal activation (debug / parity taps): name = reference module path (e.g. "input_blocks.1.0"),
 * dst receives [(B F), C, h, w] fp16 as the reference module returns it. Returns element count or <0. */
long long t2v_unet_read_tap(t2v_unet* u, const char* name, void* dst, long long cap_elems, void* stream);
int t2v_unet_enable_taps(t2v_unet* u, int on);
/* shape of a tap of the most recent plan: rows x C token matrix viewed as [(rows / (h w)), C, h, w].  A frame-sharded clip
 * records spatial modules frame-sharded (this rank's frames, full h x w) and temporal modules pixel-sharded (all frames,
 * h = 1, w = this rank's pixel count).  Returns 0, or -1 if the tap does not exist. */
int t2v_unet_tap_info(t2v_unet* u, const char* name, long long* rows, int* C, int* h, int* w);

/* ------------------------------------------------------------------------------------------ LoRA hot-merge
 * replaces StableLoraProcessor.process_lora's weight surgery (stable_lora/stable_utils/lora_processor.py:50-96, :202-246):
 * instead of re-assigning `m.weight = nn.Parameter(W +- alpha * B @ A)` and re-shipping / re-packing the whole model, the
 * low-rank update is applied to the library's own copy of ONE weight,
 *     W <- fp16(W + fp16(fp16(B @ A) * alpha))        (the reference's roundings under autocast; several merges accumulate)
 * and only the packed variants derived from it (tap-major conv layout, fused q|k|v, GEGLU interleave, LayerNorm-folded
 * copies) are rebuilt in place -- buffer addresses, plans and captured CUDA graphs stay valid, the next forward just uses the
 * new weights.  lora_A [rank, cols] and lora_B [out, rank] are fp16 device pointers, cols = in * kernel taps of the weight;
 * temporal_mean = 1 for Conv3d (3,1,1) weights: lora_A has in * 9 columns, the product is viewed [out, in, 3, 3, 1] and
 * averaged over the second kernel axis (:86-94).  t2v_unet_lora_clear restores every merged weight from its base copy:
 * bit-identical to never having merged (the reference's `-=` undo leaves fp16 rounding residue; this does not).        */
int t2v_unet_lora_merge(t2v_unet* u, const char* weight_name, const void* lora_A, const void* lora_B, int rank, float alpha,
                        int temporal_mean, void* stream);
int t2v_unet_lora_clear(t2v_unet* u, void* stream);
int t2v_unet_lora_merged(t2v_unet* u);          /* number of weights currently carrying a merge */

/* ------------------------------------------------------------------------------------------ frame-sharded clip
 * ONE clip split over the GPUs of a node, one process per GPU (BASELINE config 4: 125 frames over 8 x B200).  Frames are
 * independent inside the spatial modules and coupled in TemporalConvBlock_v2 (t2v_model.py:1201-1212), TemporalTransformer
 * (:724, :734-738) and every 5-D GroupNorm; the library keeps activations frame-sharded in the spatial modules, transposes
 * them to a pixel-sharded layout around each temporal module with a kernel that writes straight into the peers' buffers
 * over NVLink (CUDA IPC mappings), and sums the 5-D GroupNorm statistics across ranks inside the statistics kernel.  No
 * NCCL call and no host synchronisation happens inside a forward; the caller's only collective is the exchange of the
 * fixed-size exports below (any byte all-gather: torch.distributed in the Python mirror) once per shape.
 *   1. t2v_unet_shard_setup(u, rank, nranks)                    every rank, once
 *   2. t2v_unet_shard_prepare(u, shape..., &mine)               builds this rank's plan, fills `mine`
 *   3. all-gather the exports in rank order -> all[nranks]
 *   4. t2v_unet_shard_connect(u, shape..., all)                 maps the peers' slabs
 *   5. t2v_unet_forward(u, x_local, ..., F = TOTAL frames ...)  x / out hold this rank's frames [B, C, F_local, h, w];
 *      every rank must issue the same sequence of forwards (the exchange kernels wait for their peers).              */
typedef struct {
    unsigned char comm_handle[64];        /* cudaIpcMemHandle_t of the rank's flag / GroupNorm exchange region */
    unsigned char slab_handle[64];        /* cudaIpcMemHandle_t of the plan's activation slab */
    int rank, nranks;
    int n_exchanges, n_groupnorms;
    long long dst_offset[192];            /* byte offset of every exchange's destination buffer inside the slab */
} t2v_shard_export;
int t2v_unet_shard_setup(t2v_unet* u, int rank, int nranks);
int t2v_unet_shard_prepare(t2v_unet* u, int B, int F, int h, int w, int L, void* stream, t2v_shard_export* out);
int t2v_unet_shard_connect(t2v_unet* u, int B, int F, int h, int w, int L, const t2v_shard_export* all, void* stream);
/* 1 if the plan of this shape exists for the current weights and is connected (a re-shipped parameter or a plan-cache
 * eviction drops the plan: prepare + connect again -- every rank takes the same decision, the inputs are identical) */
int t2v_unet_shard_connected(t2v_unet* u, int B, int F, int h, int w, int L);
/* device-side barrier over the ranks (after connect; all ranks must call it) */
int t2v_unet_shard_barrier(t2v_unet* u, void* stream);
/* this rank's frame range [begin, end) of an F-frame clip and the exchange count of the last forward */
int t2v_unet_shard_info(t2v_unet* u, int F, int* frame_begin, int* frame_end, int* n_exchanges);

/* ------------------------------------------------------------------------------------------ VAE decoder
 * replaces AutoencoderKL.decode (modelscope/t2v_model.py:1646-1649) + ldm Decoder (vendored twin
 * videocrafter/lvdm/models/modules/autoencoder_modules.py:484-596) and the per-frame loop of
 * t2v_pipeline.py:329-355 (all frames batched).                                                           */
typedef struct t2v_vae t2v_vae;
typedef struct {
    int ch;
    int ch_mult[8];
    int n_mult;
    int num_res_blocks;
    int z_channels;
    int out_ch;
    int embed_dim;
} t2v_vae_config;
int t2v_vae_create(const t2v_vae_config* cfg, t2v_vae** out);
void t2v_vae_destroy(t2v_vae* v);
int t2v_vae_set_param(t2v_vae* v, const char* name, const void* data, int dtype, int ndim, const int64_t* shape,
                      void* stream);
int t2v_vae_missing_params(t2v_vae* v, char* name_out, size_t name_cap);
int t2v_vae_param_info(t2v_vae* v, int index, char* name_out, size_t name_cap, int64_t* shape_out, int* ndim_out);
/* z [B, z_channels, F, h, w] fp32/fp16 latent as returned by the sampler; multiplied by `z_scale`
 * (1/0.18215, t2v_pipeline.py:348) on ingest.
 *   out_mode 0: float32 [B*F, 3, 8h, 8w] in [-1, 1]   (what AutoencoderKL.decode returns, per frame)
 *   out_mode 1: uint8   [B*F, 8h, 8w, 3] RGB, tensor2vid arithmetic (t2v_pipeline.py:447-460)          */
int t2v_vae_decode(t2v_vae* v, const void* z, int z_is_f32, float z_scale, void* out, int out_mode, int B, int F, int h,
                   int w, void* stream);
/* moments = quant_conv(Encoder(x)) of AutoencoderKL.encode (modelscope/t2v_model.py:1640-1644; ldm Encoder ≙
 * videocrafter/lvdm/models/modules/autoencoder_modules.py:382-482): x [N, 3, H, W] fp16/fp32 in [-1, 1] (device) ->
 * moments_out [N, 2*embed_dim, H/8, W/8] fp32 = (mean | logvar).  compute_latents (t2v_pipeline.py:148-194) keeps
 * mean * 0.18215.  Needs the `encoder.*` / `quant_conv.*` parameters (optional for decode-only use). */
int t2v_vae_encode(t2v_vae* v, const void* x, int x_is_f32, void* moments_out, int N, int H, int W, void* stream);
double t2v_vae_flops(t2v_vae* v, int nframes, int h, int w);

/* ------------------------------------------------------------------------------------------ text conditioning
 * replaces FrozenOpenCLIPEmbedder.encode_with_transformer (modelscope/clip_hardcode.py:112-119, :269-274): the OpenCLIP
 * ViT-H-14 text transformer (token + positional embedding, `layers_run` residual attention blocks with the causal mask --
 * 23 of the 24 for layer = 'penultimate' -- then ln_final; no text projection).  Parameter names are open_clip's
 * (`token_embedding.weight`, `positional_embedding`, `transformer.resblocks.N.{ln_1,attn.in_proj_weight,attn.in_proj_bias,
 * attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj}`, `ln_final`), i.e. the keys of open_clip_pytorch_model.bin without `visual.*`.
 * Prompt parsing / chunking / emphasis weights stay host Python (clip_hardcode.py:146-395).                      */
typedef struct t2v_clip t2v_clip;
typedef struct {
    int width;          /* 1024 */
    int heads;          /* 16 (head width must be 64) */
    int layers_run;     /* 23 = 24 resblocks, 'penultimate' */
    int context;        /* 77 */
    int vocab;          /* 49408 */
} t2v_clip_config;
int t2v_clip_create(const t2v_clip_config* cfg, t2v_clip** out);
void t2v_clip_destroy(t2v_clip* m);
int t2v_clip_set_param(t2v_clip* m, const char* name, const void* data, int dtype, int ndim, const int64_t* shape, void* stream);
int t2v_clip_param_info(t2v_clip* m, int index, char* name_out, size_t name_cap, int64_t* shape_out, int* ndim_out);
/* tokens [B, context] int32 (device) -> out [B, context, width] fp16 (out_is_f32 = 0) or fp32: ln_final(transformer(...)) */
int t2v_clip_encode(t2v_clip* m, const int* tokens, void* out, int out_is_f32, int B, void* stream);

/* ------------------------------------------------------------------------------------------ sampler steps
 * replace the per-step tensor arithmetic of scripts/samplers (ddim/gaussian_sampler.py:125-136,:269-283;
 * ddim/sampler.py:176-218; uni_pc/uni_pc.py:299-307,:378-391,:625-650).                                  */
int t2v_ddim_step(const float* x, const void* eps_c, const void* eps_u, int eps_is_f32, float* x_out, long long n,
                  long long chan_stride,
                  int C, int guided_channels, float g, int mode, float a0, float a1, float a2, float a3, float a4,
                  const float* noise, int cfg_fp16, void* stream);
int t2v_cfg_x0(const float* x, const void* eps_c, const void* eps_u, int eps_is_f32, float* x0, long long n, float g,
               float alpha, float sigma, int cfg_fp16, void* stream);
int t2v_lincomb(float* out, const float* const* src, const float* coef, int n_src, long long n, void* stream);

/* img2vid inpainting latent of process_modelscope.py:170-219: masked_latents = image_latents * (1 - mask) + latent_noise * mask
 * with mask[:, :, f] = weights[f] (the per-frame schedule of T2VAnimKeys), evaluated in fp64 like the reference's numpy code.
 *   image_latents [BC, image_frames, hw] fp32 (image_frames = 1: one encoded image shared by all frames, or F)
 *   noise, out, mask_out [BC, F, hw] fp64 (mask_out may be NULL); weights [F] fp64 -- all device pointers            */
int t2v_latent_blend(const float* image_latents, int image_frames, const double* noise, const double* weights, double* out,
                     double* mask_out, int BC, int F, long long hw, void* stream);

/* ------------------------------------------------------------------------------------------ kernel-level entry
 * points (used by the parity tests; the model-level calls above are built from exactly these launchers).   */
int t2v_op_gemm(const void* a, long long lda, int K, int nd, const int* dims, int ntaps, const int* tap_off,
                const void* w_packed, int n_alloc, int N, int b_batch_dim, int flags, void* out, long long ldo,
                const void* bias, int bias_rows, long long bias_stride, const void* residual, long long ldr,
                float alpha, int force_bn, int force_cg, void* stream);
int t2v_op_pack_conv_weight(const void* src, int src_is_f32, void* dst, int Cout, int Cin, int taps, int n_alloc,
                            int k_alloc, void* stream);
int t2v_op_pack_geglu_weight(const void* w, const void* b, int src_is_f32, void* wdst, void* bdst, int H, int K, int bn,
                             void* stream);
int t2v_op_groupnorm(const void* x, long long ldx, void* y, long long ldy, long long rows, int C, int rows_per_inst,
                     const void* gamma, const void* beta, float eps, int silu, void* stream);
int t2v_op_layernorm(const void* x, long long ldx, void* y, long long ldy, long long rows, int C, const void* gamma,
                     const void* beta, float eps, void* stream);
int t2v_op_attention(const void* q, const void* k, const void* v, void* o, long long q_bs, long long q_ss,
                     long long k_bs, long long k_ss, long long v_bs, long long v_ss, long long o_bs, long long o_ss,
                     int batch, int heads, int sq, int skv, int kv_batch_div, float scale, void* stream);
/* same for head_dim in {8,16,32,40,80,160} (64 dispatches to t2v_op_attention's kernels): CrossAttention.forward of the
 * VideoCrafter denoiser, videocrafter/lvdm/models/modules/attention_temporal.py:167-190 (8 heads of width C/8). */
int t2v_op_attention_hd(const void* q, const void* k, const void* v, void* o, long long q_bs, long long q_ss,
                        long long k_bs, long long k_ss, long long v_bs, long long v_ss, long long o_bs, long long o_ss,
                        int batch, int heads, int head_dim, int sq, int skv, int kv_batch_div, float scale, void* stream);
/* TemporalCrossAttention.forward with RelativePosition tables (attention_temporal.py:46-65, :107-144), context = x:
 * sequences of T <= 32 frames; sequence s of n_seq lives at (s / seq_inner) * bs_outer + (s % seq_inner) * bs_inner, its
 * frames `ss` elements apart; head h at column h * head_dim; tables [2*max_rel+1, head_dim] fp16 (2*max_rel+1 <= 48),
 * frame distances beyond +-max_rel use the end rows of the tables, as the reference's clamp. */
int t2v_op_attention_relpos(const void* q, const void* k, const void* v, void* o, const void* table_k, const void* table_v,
                            long long n_seq, long long seq_inner, long long bs_outer, long long bs_inner, long long ss,
                            long long o_bs_outer, long long o_bs_inner, long long o_ss, int heads, int head_dim, int T,
                            int max_rel, float scale, void* stream);
int t2v_op_upsample2x(const void* x, void* y, int nframes, int h, int w, int C, void* stream);
int t2v_op_im2col_s2(const void* x, void* col, int nframes, int h, int w, int C, void* stream);
int t2v_op_time_sinusoid(const float* t, void* out, int B, int dim, void* stream);
int t2v_op_small_linear(const void* x, long long ldx, const void* W, const void* bias, const void* addend, void* y,
                        long long ldy, int B, int N, int K, int silu_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* T2V_B200_H */
