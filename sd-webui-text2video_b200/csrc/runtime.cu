#include "runtime.cuh"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace t2v {

// ----------------------------------------------------------------------------------------- ParamStore
ParamStore::~ParamStore() {
    for (auto& kv : params_)
        if (kv.second.data) cudaFree(kv.second.data);
    invalidate_packed();
}

void ParamStore::expect(const std::string& name, std::vector<long long> shape) {
    Param& p = params_[name];
    p.shape = std::move(shape);
    p.elems = 1;
    for (long long s : p.shape) p.elems *= s;
    p.expected = true;
}

int ParamStore::set(const std::string& name, const void* src, int dtype, int ndim, const int64_t* shape,
                    cudaStream_t s) {
    auto it = params_.find(name);
    if (it == params_.end() || !it->second.expected) {
        set_error("unexpected parameter '%s' (strict load, as load_state_dict(strict=True))", name.c_str());
        return -1;
    }
    Param& p = it->second;
    if (static_cast<int>(p.shape.size()) != ndim) {
        set_error("parameter '%s': rank %d, expected %d", name.c_str(), ndim, static_cast<int>(p.shape.size()));
        return -2;
    }
    for (int i = 0; i < ndim; ++i)
        if (p.shape[i] != shape[i]) {
            set_error("parameter '%s': dim %d is %lld, expected %lld", name.c_str(), i, static_cast<long long>(shape[i]),
                      p.shape[i]);
            return -2;
        }
    if (dtype != 0 && dtype != 1) {
        set_error("parameter '%s': dtype must be 0 (fp16) or 1 (fp32)", name.c_str());
        return -3;
    }
    if (!p.data) {
        // pad allocations to 16 B multiples so vector loads of tails stay in bounds
        if (cudaMalloc(&p.data, ((p.elems + 7) / 8 * 8) * sizeof(__half)) != cudaSuccess) {
            set_error("cudaMalloc failed for parameter '%s'", name.c_str());
            return -4;
        }
    }
    int rc = convert_to_f16(src, dtype, p.data, p.elems, s);
    if (rc != 0) return rc;
    p.set = true;
    ++version_;
    invalidate_packed();
    return 0;
}

int ParamStore::missing(std::string* one) const {
    int n = 0;
    for (auto& kv : params_)
        if (kv.second.expected && !kv.second.set) {
            if (n == 0 && one) *one = kv.first;
            ++n;
        }
    return n;
}

const Param& ParamStore::get(const std::string& name) const {
    static Param empty;
    auto it = params_.find(name);
    if (it == params_.end() || !it->second.set) {
        set_error("parameter '%s' is not loaded", name.c_str());
        return empty;
    }
    return it->second;
}

__half* ParamStore::packed(const std::string& key) const {
    auto it = packed_.find(key);
    return it == packed_.end() ? nullptr : it->second;
}

__half* ParamStore::new_packed(const std::string& key, long long elems) {
    __half* p = nullptr;
    if (cudaMalloc(&p, ((elems + 7) / 8 * 8) * sizeof(__half)) != cudaSuccess) {
        set_error("cudaMalloc failed for packed weight '%s'", key.c_str());
        return nullptr;
    }
    packed_[key] = p;
    return p;
}

void ParamStore::invalidate_packed() {
    for (auto& kv : packed_) cudaFree(kv.second);
    packed_.clear();
}

// ----------------------------------------------------------------------------------------- Arena
void Arena::reset(char* base, bool no_reuse) {
    base_ = base;
    no_reuse_ = no_reuse;
    top_ = 0;
    peak_ = 0;
    free_.clear();
    live_.clear();
}

char* Arena::alloc(size_t bytes) {
    bytes = (bytes + 1023) & ~static_cast<size_t>(1023);
    if (bytes == 0) bytes = 1024;
    if (!no_reuse_) {
        int best = -1;
        for (int i = 0; i < static_cast<int>(free_.size()); ++i)
            if (free_[i].size >= bytes && (best < 0 || free_[i].size < free_[best].size)) best = i;
        if (best >= 0) {
            const size_t off = free_[best].off;
            if (free_[best].size == bytes) free_.erase(free_.begin() + best);
            else {
                free_[best].off += bytes;
                free_[best].size -= bytes;
            }
            live_[off] = bytes;
            return base_ + off;
        }
    }
    const size_t off = top_;
    top_ += bytes;
    peak_ = std::max(peak_, top_);
    live_[off] = bytes;
    return base_ + off;
}

void Arena::free(char* p) {
    if (no_reuse_ || p == nullptr && base_ != nullptr) return;
    const size_t off = static_cast<size_t>(p - base_);
    auto it = live_.find(off);
    if (it == live_.end()) return;
    Blk b{off, it->second};
    live_.erase(it);
    // insert sorted + coalesce
    auto pos = std::lower_bound(free_.begin(), free_.end(), b, [](const Blk& a, const Blk& c) { return a.off < c.off; });
    pos = free_.insert(pos, b);
    const int i = static_cast<int>(pos - free_.begin());
    if (i + 1 < static_cast<int>(free_.size()) && free_[i].off + free_[i].size == free_[i + 1].off) {
        free_[i].size += free_[i + 1].size;
        free_.erase(free_.begin() + i + 1);
    }
    if (i > 0 && free_[i - 1].off + free_[i - 1].size == free_[i].off) {
        free_[i - 1].size += free_[i].size;
        free_.erase(free_.begin() + i);
    }
    // shrink the bump pointer when the tail is free
    if (!free_.empty() && free_.back().off + free_.back().size == top_) {
        top_ = free_.back().off;
        free_.pop_back();
    }
}

Plan::~Plan() {
    if (graph) cudaGraphExecDestroy(graph);
    if (slab) cudaFree(slab);
}

// ----------------------------------------------------------------------------------------- Builder
Tok Builder::alloc(long long rows, int C, int ld) {
    Tok t;
    t.rows = rows;
    t.C = C;
    t.ld = ld > 0 ? ld : C;
    t.p = reinterpret_cast<__half*>(arena_->alloc(static_cast<size_t>(rows) * t.ld * sizeof(__half)));
    return t;
}

void* Builder::alloc_bytes(size_t bytes) { return arena_->alloc(bytes); }

int Builder::gemm(GemmProblem& p) {
    double rows = 1;
    for (int d = 0; d < p.nd; ++d) rows *= p.dim[d];
    plan_->flops += 2.0 * rows * p.N * p.K * p.ntaps;
    if (dry_) {
        plan_->launches += 1;
        return 0;
    }
    GemmPlan gp;
    int rc = gemm_plan(p, &gp, sms_);
    if (rc != 0) {
        set_error("gemm_plan failed (%d): rows %.0f N %d K %d taps %d", rc, rows, p.N, p.K, p.ntaps);
        error = rc;
        return rc;
    }
    plan_->steps.push_back([gp](cudaStream_t s) { return gemm_launch(gp, s); });
    plan_->launches += 1;
    return 0;
}

void Builder::step(Step s, int launches) {
    plan_->launches += launches;
    if (!dry_) plan_->steps.push_back(std::move(s));
}

void taps_3x3(GemmProblem& p) {
    p.ntaps = 9;
    for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
            const int t = ky * 3 + kx;
            p.tap_off[t][0] = kx - 1;
            p.tap_off[t][1] = ky - 1;
            p.tap_off[t][2] = 0;
            p.tap_off[t][3] = 0;
        }
}

void taps_temporal(GemmProblem& p) {
    p.ntaps = 3;
    for (int kt = 0; kt < 3; ++kt) {
        p.tap_off[kt][0] = 0;
        p.tap_off[kt][1] = kt - 1;
        p.tap_off[kt][2] = 0;
        p.tap_off[kt][3] = 0;
    }
}

}  // namespace t2v
