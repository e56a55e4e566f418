// FS <-> PS layout exchange over NVLink peer memory + epoch / barrier helpers (see shard.cuh).
// Roofline: NVLink-bound.  Algorithmic bytes per launch and rank = rows * C * 2 * (P-1)/P sent and the same received
// (e.g. 125 f x 32 x 32 x 320 ch, B = 2, 8 ranks: 17.9 MB each way per level-0 exchange).
#include "shard.cuh"

#include "common.cuh"

namespace t2v {

namespace {

constexpr int kXThreads = 256;

__global__ void __launch_bounds__(kXThreads) shard_exchange_kernel(const XchgParams p, int nblk) {
    const int me = p.peers.rank, nr = p.peers.nranks;
    ShardComm* mine = p.peers.comm[me];
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&mine->epoch);
    const int k = p.slot;
    if (blockIdx.x == 0) {
        // ---- waiter block (scheduled first): announce that OUR destination buffer is dead (stream order: every kernel that used the memory it
        // occupies has completed), then hold the kernel open until every peer's block has landed in it
        if (threadIdx.x < nr && static_cast<int>(threadIdx.x) != me) st_release_sys(&p.peers.comm[threadIdx.x]->ready[k][me], e);
        if (threadIdx.x < nr && static_cast<int>(threadIdx.x) != me) spin_until_ge(&mine->done[k][threadIdx.x], e);
        return;
    }
    const int s = (blockIdx.x - 1) / nblk;           // destination rank of this block
    const int j = (blockIdx.x - 1) - s * nblk;
    const int C8 = p.C >> 3;
    const int nf_me = p.fb[me + 1] - p.fb[me], np_me = p.pb[me + 1] - p.pb[me];
    const int nf_s = p.fb[s + 1] - p.fb[s], np_s = p.pb[s + 1] - p.pb[s];
    const int nf = p.to_ps ? nf_me : nf_s;           // frames in the (me -> s) block
    const int np = p.to_ps ? np_s : np_me;           // pixels in it
    const long long nvec = static_cast<long long>(p.B) * nf * np * C8;
    if (s != me) {
        if (threadIdx.x == 0) spin_until_ge(&mine->ready[k][s], e);      // s has released its destination buffer
        __syncthreads();
    }
    const __half* src = p.src;
    __half* dst = p.dst[s];
    // vector index -> (sample, frame, pixel, 16-byte chunk) without integer divisions in the copy loop: the three divisors are
    // block constants, so thread 0 derives round-up magic numbers once (Granlund-Montgomery: q = (t + ((n - t) >> s1)) >> s2 with
    // t = umulhi(m, n), exact for every 32-bit n); the 64-bit div/mod chain cost ~300 instructions per 16 bytes before.
    __shared__ uint32_t fd[3][3];
    if (threadIdx.x < 3) {
        uint32_t d = threadIdx.x == 0 ? static_cast<uint32_t>(C8) : (threadIdx.x == 1 ? static_cast<uint32_t>(np) : static_cast<uint32_t>(nf));
        if (d == 0u) d = 1u;          // empty block (nvec = 0): the constants are never used
        uint32_t l = 0;
        while ((1u << l) < d) ++l;
        fd[threadIdx.x][0] = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - d)) / d) + 1u;
        fd[threadIdx.x][1] = l < 1u ? l : 1u;
        fd[threadIdx.x][2] = l - (l < 1u ? l : 1u);
    }
    __syncthreads();
    auto fdiv = [&](uint32_t n, int i) {
        const uint32_t t = __umulhi(fd[i][0], n);
        return (t + ((n - t) >> fd[i][1])) >> fd[i][2];
    };
    const uint32_t nv32 = static_cast<uint32_t>(nvec);
    const uint32_t stride = static_cast<uint32_t>(nblk) * kXThreads;
    const long long rbase_s = p.to_ps ? static_cast<long long>(p.pb[s]) : 0, rbase_d = p.to_ps ? 0 : static_cast<long long>(p.pb[me]);
    constexpr int U = 4;
    for (uint32_t v0 = static_cast<uint32_t>(j) * kXThreads + threadIdx.x; v0 < nv32; v0 += U * stride) {
        uint4 val[U];
        long long doff[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t v = v0 + u * stride;
            doff[u] = -1;
            if (v < nv32) {
                const uint32_t row = fdiv(v, 0);
                const uint32_t vc = v - row * static_cast<uint32_t>(C8);
                const uint32_t bf = fdiv(row, 1);
                const uint32_t pl = row - bf * static_cast<uint32_t>(np);
                const uint32_t b = fdiv(bf, 2);
                const uint32_t fl = bf - b * static_cast<uint32_t>(nf);
                long long srow, drow;
                if (p.to_ps) {
                    srow = (static_cast<long long>(b) * nf_me + fl) * p.P + rbase_s + pl;
                    drow = (static_cast<long long>(b) * p.F + p.fb[me] + fl) * np_s + pl;
                } else {
                    srow = (static_cast<long long>(b) * p.F + p.fb[s] + fl) * np_me + pl;
                    drow = (static_cast<long long>(b) * nf_s + fl) * p.P + rbase_d + pl;
                }
                val[u] = __ldg(reinterpret_cast<const uint4*>(src + srow * p.ld_src + vc * 8));
                doff[u] = drow * p.ld_dst + vc * 8;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (doff[u] >= 0) *reinterpret_cast<uint4*>(dst + doff[u]) = val[u];
    }
    if (s != me) {
        // publish: all of this block's stores, then count the block; the last block of destination s raises done[k][me] on s
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int prev = atomicAdd(&mine->block_counter[k][s], 1u);
            if (prev == static_cast<unsigned int>(nblk - 1)) {
                mine->block_counter[k][s] = 0u;                              // self-cleaning for the next forward
                __threadfence_system();
                st_release_sys(&p.peers.comm[s]->done[k][me], e);
            }
        }
    }
}

__global__ void shard_epoch_kernel(ShardComm* c) { c->epoch = c->epoch + 1u; }

__global__ void shard_barrier_kernel(const ShardPeers peers, int slot) {
    const int me = peers.rank;
    const unsigned int e = *reinterpret_cast<volatile unsigned int*>(&peers.comm[me]->epoch);
    const int t = threadIdx.x;
    if (t < peers.nranks && t != me) {
        st_release_sys(&peers.comm[t]->done[slot][me], e);
        spin_until_ge(&peers.comm[me]->done[slot][t], e);
    }
}

}  // namespace

int shard_exchange(const XchgParams& p, int num_sms, cudaStream_t stream) {
    if (p.peers.nranks < 2 || p.peers.nranks > SHARD_MAX_RANKS || p.slot < 0 || p.slot >= SHARD_MAX_XCHG || (p.C & 7) != 0 ||
        (p.ld_src & 7) != 0 || (p.ld_dst & 7) != 0)
        return -1;
    const int nr = p.peers.nranks;
    int max_nf = 0, max_np = 0;
    for (int r = 0; r < nr; ++r) {
        max_nf = max(max_nf, p.fb[r + 1] - p.fb[r]);
        max_np = max(max_np, p.pb[r + 1] - p.pb[r]);
    }
    const long long vecs = static_cast<long long>(p.B) * max_nf * max_np * (p.C >> 3);
    if (vecs >= (1LL << 31)) return -1;      // the copy loop indexes 16-byte vectors with 32 bits
    long long nblk = (vecs + 4 * kXThreads - 1) / (4 * kXThreads);
    const long long cap = (2LL * num_sms + nr - 1) / nr;
    if (nblk > cap) nblk = cap;
    if (nblk < 1) nblk = 1;
    shard_exchange_kernel<<<static_cast<unsigned>(nblk * nr + 1), kXThreads, 0, stream>>>(p, static_cast<int>(nblk));
    return launch_status("shard launch");
}

int shard_bump_epoch(ShardComm* local, cudaStream_t stream) {
    shard_epoch_kernel<<<1, 1, 0, stream>>>(local);
    return launch_status("shard launch");
}

int shard_barrier(const ShardPeers& peers, int slot, cudaStream_t stream) {
    shard_barrier_kernel<<<1, 32, 0, stream>>>(peers, slot);
    return launch_status("shard launch");
}

}  // namespace t2v
