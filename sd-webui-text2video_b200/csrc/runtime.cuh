// Host runtime shared by the denoiser and the VAE decoder: parameter store, packed-weight cache, a lifetime-aware
// activation arena, and the "plan" -- a flat list of pre-encoded kernel launches (tensor maps encoded once per
// shape) that one forward replays on a stream without any host-side shape logic.
#pragma once
#include "common.cuh"
#include "gemm_tc.cuh"
#include "kernels.cuh"
#include "shard.cuh"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

namespace t2v {

// ----------------------------------------------------------------------------------------- parameters
struct Param {
    std::vector<long long> shape;
    __half* data = nullptr;      // fp16 copy in the ORIGINAL (PyTorch) layout, library-owned (the EFFECTIVE weight: base + merged LoRAs)
    __half* base = nullptr;      // copy of the shipped weight, kept from the first LoRA merge on (lora_clear restores it bit for bit)
    long long elems = 0;
    bool expected = false;
    bool set = false;
};

// How one packed variant is (re)built from its sources: replayed in creation order when a source changes WITHOUT a new
// weights version (LoRA hot-merge), so packed buffers keep their addresses and every plan / captured graph stays valid.
struct PackRecipe {
    std::string key;
    std::vector<std::string> sources;      // parameter names and / or keys of other packed variants
    std::function<int(cudaStream_t)> run;
};

class ParamStore {
public:
    ~ParamStore();
    void expect(const std::string& name, std::vector<long long> shape);
    int set(const std::string& name, const void* src, int dtype, int ndim, const int64_t* shape, cudaStream_t s);
    int missing(std::string* one) const;
    int info(int index, std::string* name, std::vector<long long>* shape) const;   // returns count, -1 if out of range
    const Param& get(const std::string& name) const;     // aborts via set_error + null data if absent
    bool has(const std::string& name) const { return params_.count(name) != 0; }
    // packed variants, created lazily and cached until any parameter changes
    __half* packed(const std::string& key) const;
    __half* new_packed(const std::string& key, long long elems);
    void invalidate_packed();
    unsigned long long version() const { return version_; }
    // packed-variant recipes (see PackRecipe) and the name a device pointer is known under ("" if unknown)
    void add_recipe(const std::string& key, std::vector<std::string> sources, std::function<int(cudaStream_t)> run);
    std::string key_of(const void* p) const;
    // LoRA hot-merge (stable_lora/stable_utils/lora_processor.py:50-96): data = fp16(data + fp16(fp16(B @ A) * alpha)), then only
    // the packed variants that depend on `name` are rebuilt in place.  temporal_mean: Conv3d (3,1,1) weights, the product is
    // viewed [out, in, 3, 3, 1] and averaged over the second kernel axis (:86-94).
    int lora_merge(const std::string& name, const __half* lora_A, const __half* lora_B, int rank, float alpha, int temporal_mean,
                   cudaStream_t s);
    int lora_clear(cudaStream_t s);          // every merged weight back to its base copy (bit-identical to never merging)
    int merged_count() const;

private:
    int repack(const std::vector<std::string>& dirty, cudaStream_t s);
    std::map<std::string, Param> params_;
    std::map<std::string, __half*> packed_;
    std::vector<PackRecipe> recipes_;
    unsigned long long version_ = 0;
};

// ----------------------------------------------------------------------------------------- arena
// Offsets are handed out by a first-fit free list while the plan is being built (the build order IS the execution
// order, so build-time lifetimes are run-time lifetimes).  A dry pass measures the peak, then the real pass runs
// against one cudaMalloc'ed slab.
class Arena {
public:
    void reset(char* base, bool no_reuse);
    char* alloc(size_t bytes);
    void free(char* p);
    size_t peak() const { return peak_; }
    long count() const { return count_; }

private:
    struct Blk { size_t off, size; };
    char* base_ = nullptr;
    bool no_reuse_ = false;
    long count_ = 0, limit_ = -1;
    size_t top_ = 0, peak_ = 0;
    std::vector<Blk> free_;
    std::map<size_t, size_t> live_;
};

struct Tok {                      // channels-last token matrix view
    __half* p = nullptr;
    long long rows = 0;
    int C = 0;
    long long ld = 0;
};

using Step = std::function<int(cudaStream_t)>;
enum StepKind : int { STEP_GEMM = 0, STEP_ATTN = 1, STEP_NORM = 2, STEP_OTHER = 3, STEP_NKINDS = 4 };
struct StepRec {
    Step fn;
    int kind;
    double flops;
    std::string label;
};

// Frame-sharded plans (shard.cuh): host-side state the exchange steps read at LAUNCH time -- the peers' mapped slabs and
// the byte offset of every exchange's destination buffer inside each peer's slab are only known after the ranks have
// swapped their exports (t2v_unet_shard_connect), which happens after the plan is built and before its first replay.
struct PlanShard {
    int fb[SHARD_MAX_RANKS + 1] = {0};                 // frame partition of the clip
    int n_xchg = 0, n_gn = 0;
    long long dst_off[SHARD_MAX_XCHG] = {0};           // this rank's destination offsets (bytes into its slab)
    long long peer_dst_off[SHARD_MAX_RANKS][SHARD_MAX_XCHG] = {{0}};
    char* peer_slab[SHARD_MAX_RANKS] = {nullptr};      // peer_slab[rank] = own slab
    bool connected = false;
    int own_rank = 0;
    ~PlanShard();
};

struct Plan {
    std::vector<StepRec> steps;
    std::shared_ptr<PlanShard> shard;
    char* slab = nullptr;
    size_t slab_bytes = 0;
    double flops = 0.0;
    int launches = 0;
    unsigned long long weights_version = 0;
    std::map<std::string, std::pair<Tok, std::pair<int, int>>> taps;    // name -> (tokens, (h, w))
    cudaGraphExec_t graph = nullptr;
    int eager_runs = 0;
    ~Plan();
};

// Helper that records launches into a plan (or only simulates allocations when `dry`).
class Builder {
public:
    Builder(Plan* plan, Arena* arena, bool dry, int sms) : plan_(plan), arena_(arena), dry_(dry), sms_(sms) {}
    bool dry() const { return dry_; }
    Tok alloc(long long rows, int C, int ld = 0);
    void* alloc_bytes(size_t bytes);
    void free(const Tok& t) { arena_->free(reinterpret_cast<char*>(t.p)); }
    void free_bytes(void* p) { arena_->free(reinterpret_cast<char*>(p)); }
    int gemm(GemmProblem& p);                                     // plans + records; returns 0 or <0
    void step(Step s, int launches = 1, int kind = STEP_OTHER, double flops = 0.0, const char* label = "");
    void add_flops(double f) { plan_->flops += f; }
    int sms() const { return sms_; }
    int error = 0;

private:
    Plan* plan_;
    Arena* arena_;
    bool dry_;
    int sms_;
};

// ----------------------------------------------------------------------------------------- layer helpers
struct NetCtx {
    ParamStore* params;
    Builder* b;
    cudaStream_t stream;      // weight-packing kernels are enqueued here while the plan is built
    void* gn_ws;              // zero-initialised groupnorm workspace (partials, stats, counters)
    const ShardPeers* shard_peers = nullptr;   // frame-sharded clip: live peer table (filled by connect), else null
    PlanShard* plan_shard = nullptr;
};
int round_up(int v, int m);
// packed-weight accessors (created on first use, cached in the ParamStore until a parameter changes)
const __half* w_conv(NetCtx& c, const std::string& name, int taps, int n_alloc = 0, int k_alloc = 0);
const __half* w_conv_kmajor(NetCtx& c, const std::string& name);
const __half* w_cat(NetCtx& c, const std::vector<std::string>& names);
struct Geglu { const __half* w; const __half* b; int bn; };
Geglu w_geglu(NetCtx& c, const std::string& prefix, int H, int K, int bn);
const __half* prm(NetCtx& c, const std::string& name);
// recorded ops
GemmProblem base_problem(const Tok& a, int K, const __half* w, int n_alloc, int N, const Tok& out);
Tok linear(NetCtx& c, const Tok& x, const __half* w, int N, const __half* bias, const Tok* residual, int K = 0);
// shard_total_rows > 0: 5-D norm of a pixel-sharded clip -- the statistics span `shard_total_rows` rows per sample over all ranks
Tok group_norm(NetCtx& c, const Tok& x, const std::string& prefix, long long rows_per_inst, float eps, bool silu,
               long long shard_total_rows = 0);
Tok layer_norm(NetCtx& c, const Tok& x, const std::string& prefix);
// y = Linear(LayerNorm(x)) (+residual) with the normalisation folded into the GEMM: a row-statistics kernel (reads x once)
// + one GEMM on the RAW rows whose epilogue applies rstd * (acc - mean * colsum) + bias'.  `w_src` [N, K] / `bias_src` are
// the (already concatenated / GEGLU-interleaved) weights; the gamma/beta-folded copy is cached under `key`.
Tok ln_linear(NetCtx& c, const Tok& x, const std::string& ln_prefix, const std::string& key, const __half* w_src,
              const __half* bias_src, int N, const Tok* residual, int flags = 0, int force_bn = 0);
Tok conv3x3(NetCtx& c, const Tok& x, const std::string& wname, const __half* bias, int bias_rows, long long bias_stride,
            int N, int hcur, int wcur, const Tok* residual, int n_alloc = 0);

// Runs the plan once with a CUDA-event pair around every step; out[kind*3 + {0,1,2}] = {ms, flops, launches}, out[12] = total ms.
// If T2V_PROFILE_DUMP names a file, one line per launch (index, kind, ms, flop, label) is written there.
int profile_plan(Plan* plan, cudaStream_t stream, double* out13);
// Replays the plan on `stream`.  The first call runs launch by launch; the second call captures the launch list into a
// CUDA graph (through a private capture stream: the caller's may be the legacy default stream) and from then on a
// forward is ONE cudaGraphLaunch -- the ~1k launches stop costing host time.  T2V_NO_GRAPH=1 disables it.
int run_plan(Plan* plan, cudaStream_t stream, bool allow_graph);

// conv taps helpers over row dims (w, h, frames) and (pixels, frames, samples)
void taps_3x3(GemmProblem& p);
void taps_temporal(GemmProblem& p);

}  // namespace t2v
