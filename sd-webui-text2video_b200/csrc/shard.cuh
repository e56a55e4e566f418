// Frame-sharded denoiser: device-side exchange between the GPUs of one clip over NVLink peer memory (CUDA IPC mappings of
// the peers' activation slabs).  One process per GPU; no NCCL call and no host synchronisation inside a forward.
//
// A clip of F frames is split over P ranks.  Frames are independent inside the spatial modules
// (ResBlock conv2d, SpatialTransformer, Down/Upsample) and coupled in TemporalConvBlock_v2 (3-tap conv along f,
// modelscope/t2v_model.py:1201-1212), TemporalTransformer (attention along f, :734-738) and every 5-D GroupNorm (:724,
// :1202-1211).  Two layouts of the SAME token matrix are used:
//   FS (frame-sharded) : rows (b, f in own frames, all h*w pixels)   -- spatial modules
//   PS (pixel-sharded) : rows (b, all F frames, p in own pixel range) -- temporal modules: the conv taps, the attention
//                        sequences and the conv's zero padding are then entirely local
// and an all-to-all "transpose" kernel moves between them: every rank PUSHES the (own frames x peer's pixels) blocks
// straight into the peer's destination buffer with 16-byte stores over NVLink, bracketed by flag handshakes
// (ready: "my destination buffer is dead, you may write" / done: "my block has landed").  The 2 x 32 per-sample
// GroupNorm sums of a 5-D norm in PS layout are exchanged the same way from inside the statistics kernel.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace t2v {

constexpr int SHARD_MAX_RANKS = 8;
constexpr int SHARD_MAX_XCHG = 192;       // layout exchanges per forward (ModelScope UNetSD: 78)
constexpr int SHARD_MAX_GN = 160;         // cross-rank GroupNorm reductions per forward (105)
constexpr int SHARD_MAX_INST = 4;         // samples per forward (B)

// Lives in every rank's IPC-shared communication region.  Flags carry the forward's epoch (monotonic), so nothing is ever
// reset: slot k of forward e is valid once flag >= e.
struct ShardComm {
    unsigned int epoch;                                              // written by the local rank only
    unsigned int pad[31];
    unsigned int ready[SHARD_MAX_XCHG][SHARD_MAX_RANKS];             // [k][src]: src's destination buffer of exchange k is free
    unsigned int done[SHARD_MAX_XCHG][SHARD_MAX_RANKS];              // [k][src]: src's block of exchange k has landed here
    unsigned int gn_flag[SHARD_MAX_GN][SHARD_MAX_INST][SHARD_MAX_RANKS];
    double2 gn_part[SHARD_MAX_GN][SHARD_MAX_INST][SHARD_MAX_RANKS][32];   // (sum, sumsq) per group from every rank
    unsigned int block_counter[SHARD_MAX_XCHG][SHARD_MAX_RANKS];     // local: blocks finished per (exchange, destination)
};

struct ShardPeers {
    ShardComm* comm[SHARD_MAX_RANKS];     // comm[r] = this rank's own region
    int rank, nranks;
};

// cross-rank part of a 5-D GroupNorm's statistics (norm.cu): nranks <= 1 -> plain local norm
struct GnShard {
    ShardPeers peers;
    int slot;
    long long total_rows_per_inst;        // rows of one sample over ALL ranks (F * h * w)
};

// FS <-> PS transpose.  Row r of a [rows, C] fp16 matrix with pitch ld.  frame partition fb[0..P], pixel partition pb[0..P].
struct XchgParams {
    ShardPeers peers;
    int slot;
    int to_ps;                            // 1: FS -> PS, 0: PS -> FS
    int B, F, P, C;                       // samples, total frames, pixels per frame at this level, channels
    long long ld_src, ld_dst;
    const __half* src;                    // local source matrix
    __half* dst[SHARD_MAX_RANKS];         // destination matrix on every rank (peer-mapped pointers; dst[rank] is local)
    int fb[SHARD_MAX_RANKS + 1];
    int pb[SHARD_MAX_RANKS + 1];
};

int shard_exchange(const XchgParams& p, int num_sms, cudaStream_t stream);
int shard_bump_epoch(ShardComm* local, cudaStream_t stream);
// device barrier over all ranks (used once after connect and by tests): returns after every rank has arrived
int shard_barrier(const ShardPeers& peers, int slot, cudaStream_t stream);

inline void shard_partition(int n, int parts, int* bounds) {      // balanced contiguous ranges, larger ones first
    int off = 0;
    for (int r = 0; r < parts; ++r) {
        bounds[r] = off;
        off += n / parts + (r < n % parts ? 1 : 0);
    }
    bounds[parts] = off;
}

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Bounded: a peer that never arrives (a rank died, or the ranks issued different launch sequences) must surface as a trapped
// kernel -> a CUDA error in this process, never as a GPU spinning forever.  2^29 polls of >= 64 ns each are more than half a
// minute, far beyond any legitimate skew between the ranks' graphs (plan construction happens before the first launch).
__device__ __forceinline__ void spin_until_ge(const unsigned int* p, unsigned int e) {
    for (unsigned int spins = 0; static_cast<int>(ld_acquire_sys(p) - e) < 0; ++spins) {
        __nanosleep(64);
        if (spins > (1u << 29)) __trap();
    }
}
#endif

}  // namespace t2v
