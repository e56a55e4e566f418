// Spatial self-attention on the 5th-gen tensor cores: softmax(Q K^T * scale) V, head_dim 64, long sequences
// (S = h*w tokens of one frame, t2v_model.py:540-584 with the (b, hw, c) token layout of :639-658).
//
// One CTA owns 256 queries (two 128-row tiles) of one (frame, head) and streams the keys/values in 128-row tiles:
//
//   warp 8      TMA producer   Q (2 x 16 KB, once), then K_j / V_j tiles into a 4-stage ring -- straight out of the fused
//                              [tokens, 3C] QKV matrix: the head is a column offset of the tensor map, rows past the end
//                              of the frame are TMA zero fill.
//   warp 9      MMA issuer     S_t  = Q_t K_j^T  4 x tcgen05.mma M128 N128 K16, both operands from shared memory,
//                                                fp32 scores in TMEM columns t*128 + [0,128)
//                              O_t += P_t V_j    8 x tcgen05.mma M128 N64 K16: A = P_t read from TMEM (fp16 pairs, columns
//                                                256 + t*64 + [0,64)), B = V_j as it lies in the token matrix ([key][d] rows)
//                                                through an MN-major shared-memory descriptor -- no transposed copy of V, no
//                                                shared-memory round trip for P; fp32 O_t in TMEM columns 384 + t*64 + [0,64)
//   warps 0-3   softmax, query tile 0 } thread = query row (TMEM lane): the 128 scores of the row are pulled into registers
//   warps 4-7   softmax, query tile 1 } in one go and the TMEM copy released at once (S_t(j+1) is computed while the
//                              exponentials of S_t(j) are evaluated); row max, exp2 on the MUFU, fp16 pairs stored back
//                              to TMEM with tcgen05.st.
//   setmaxnreg moves registers from the TMA/MMA warpgroup (72) to the softmax warpgroups (216).
//
// O_t accumulates inside the tensor core across key tiles.  The exponent offset m of a row is therefore only moved
// (and O_t, l rescaled by exp2((m_old - m_new) c) through a TMEM read-modify-write) when the running max outgrew it by
// more than 2^8: until then P = exp2(S c - m c) <= 256 is exact in fp16's range and the final O / l is unchanged
// (every term carries the same factor).  In steady state the softmax warps never touch O_t.
//
// The two query tiles run out of phase: while one group evaluates exponentials the tensor core works on the other
// tile's S / P.V; both share every K/V tile (one L2 read per 256 queries).  Shared-memory traffic per key tile is
// Q,K operand reads 64 KB + V reads 32 KB + TMA fill 32 KB (128 B/clk/SM), exponentials 32768 / (16/clk/SM).
//
// Numerics (= torch SDPA fused kernels the reference dispatches to, t2v_model.py:561-569): fp16 operands, fp32 scores,
// fp32 online softmax with the scale folded into exp2, P rounded to fp16 for P.V, fp32 output accumulation,
// normalised by the fp32 row sum at the end.
#include <cuda.h>

#include <cstdio>

#include "gemm_tc.cuh"
#include "common.cuh"
#include "kernels.cuh"
#include "ptx.cuh"

namespace t2v {

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;                 // query rows per tile (= TMEM lanes)
constexpr int NT = 2;                   // query tiles per CTA
constexpr int BKV = 128;                // keys per iteration
constexpr int ST = 4;                   // K/V ring stages
constexpr int TILE_BYTES = 128 * 128;   // 128 rows x 64 fp16
constexpr int SMEM_Q = 0;
constexpr int SMEM_K = SMEM_Q + NT * TILE_BYTES;
constexpr int SMEM_V = SMEM_K + ST * TILE_BYTES;
constexpr int SMEM_BAR = SMEM_V + ST * TILE_BYTES;
constexpr int SMEM_TOTAL = SMEM_BAR + 256 + 1024;          // + alignment slack
constexpr int NTHREADS = 384;           // warpgroups: softmax tile 0 | softmax tile 1 | TMA, MMA (+2 idle warps)
constexpr int REGS_SOFTMAX = 216;       // 2 x 128 x 216 + 128 x 72 = 384 x 168
constexpr int REGS_OTHER = 72;

constexpr int TMEM_S = 0;               // S_t : fp32 scores, columns t*128 + [0, 128)
constexpr int TMEM_P = 256;             // P_t : fp16 probabilities, two keys per column, columns 256 + t*64 + [0, 64)
constexpr int TMEM_O = 384;             // O_t : fp32 running output, columns 384 + t*64 + [0, 64)
constexpr float RESCALE_LOG2 = 8.f;     // the exponent offset lags the true running max by at most 2^8 (P <= 256 in fp16)

struct Args {
    __half* o;
    long long o_bs, o_ss;
    int sq, skv, kv_batch_div, n_kv;
    float sl2;                          // scale * log2(e)
};

__global__ void __launch_bounds__(NTHREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, const Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_BAR);
    uint64_t* bar_q = bars;                    // Q tiles landed
    uint64_t* kv_full = bars + 1;              // [ST] K_j, V_j landed
    uint64_t* kv_empty = kv_full + ST;         // [ST] all MMAs reading the stage retired
    uint64_t* s_full = kv_empty + ST;          // [NT] S_t(j) in TMEM
    uint64_t* s_free = s_full + NT;            // [NT] S_t(j) copied to registers: S_t(j+1) may be issued
    uint64_t* p_full = s_free + NT;            // [NT] P_t(j) in TMEM (and O_t rescaled if the row max moved)
    uint64_t* pv_full = p_full + NT;           // [NT] O_t += P_t(j) V_j retired: P_t may be overwritten, O_t read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_full + NT);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * (NT * BQ);
    const int head = blockIdx.y;
    const int b = blockIdx.z;
    const int n_kv = a.n_kv;

    if (threadIdx.x == 0) {
        mbar_init(bar_q, 1);
        for (int s = 0; s < ST; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
        }
        for (int t = 0; t < NT; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&s_free[t], 4);          // one arrival per softmax warp
            mbar_init(&p_full[t], 4);
            mbar_init(&pv_full[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == 9) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    griddep_wait();        // the prologue above overlaps the previous kernel's tail (PDL, common.cuh)

    if (warp >= 8) {
        setmaxnreg_dec<REGS_OTHER>();
        if (warp == 8) {
            // -------------------------------------------------------------- TMA producer
            if (elect_one()) {
                tma_prefetch_desc(&map_q);
                tma_prefetch_desc(&map_k);
                tma_prefetch_desc(&map_v);
                mbar_expect_tx(bar_q, NT * TILE_BYTES);
                for (int t = 0; t < NT; ++t)
                    tma_load_3d(smem + SMEM_Q + t * TILE_BYTES, &map_q, bar_q, head * HD, q0 + t * BQ, b);
                const int bkv = b / a.kv_batch_div;
                for (int j = 0; j < n_kv; ++j) {
                    const int s = j % ST;
                    if (j >= ST) mbar_wait(&kv_empty[s], ((j / ST) - 1) & 1);
                    mbar_expect_tx(&kv_full[s], 2 * TILE_BYTES);
                    tma_load_3d(smem + SMEM_K + s * TILE_BYTES, &map_k, &kv_full[s], head * HD, j * BKV, bkv);
                    tma_load_3d(smem + SMEM_V + s * TILE_BYTES, &map_v, &kv_full[s], head * HD, j * BKV, bkv);
                }
                griddep_launch();          // all loads issued: dependents may be scheduled as SMs drain
            }
        } else if (warp == 9) {
            // -------------------------------------------------------------- MMA issuer
            if (elect_one()) {
                constexpr uint32_t idesc_s = umma_idesc_f16(BQ, BKV);
                constexpr uint32_t idesc_pv = umma_idesc_f16(BQ, HD) | UMMA_IDESC_B_MN_MAJOR;
                const uint32_t sq_addr = smem_u32(smem + SMEM_Q);
                const uint32_t sk_addr = smem_u32(smem + SMEM_K);
                const uint32_t sv_addr = smem_u32(smem + SMEM_V);
                auto issue_s = [&](int t, int s) {
                    const uint64_t dq = umma_desc_k_sw128(sq_addr + t * TILE_BYTES);
                    const uint64_t dk = umma_desc_k_sw128(sk_addr + s * TILE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)
                        umma_f16(tmem_base + TMEM_S + t * BKV, dq + 2 * ks, dk + 2 * ks, idesc_s, ks > 0);
                    umma_commit(&s_full[t]);
                };
                auto issue_pv = [&](int t, int j) {
                    const int s = j % ST;
#pragma unroll
                    for (int ks = 0; ks < BKV / 16; ++ks) {
                        // A = P_t from TMEM: 16 keys = 8 packed columns per k-step; B = V: 16 key rows = 2048 B per k-step
                        const uint64_t dv = umma_desc_mn_sw128(sv_addr + s * TILE_BYTES + ks * 2048);
                        umma_f16_ts(tmem_base + TMEM_O + t * HD, tmem_base + TMEM_P + t * (BKV / 2) + ks * 8, dv, idesc_pv,
                                    (j > 0 || ks > 0) ? 1u : 0u);
                    }
                    umma_commit(&pv_full[t]);
                };
                mbar_wait(bar_q, 0);
                for (int j = 0; j < n_kv; ++j) {
                    const int s = j % ST;
                    mbar_wait(&kv_full[s], (j / ST) & 1);
                    tc_fence_after();
                    for (int t = 0; t < NT; ++t) {
                        if (j > 0) {
                            mbar_wait(&s_free[t], (j - 1) & 1);
                            tc_fence_after();
                        }
                        issue_s(t, s);
                    }
                    if (j > 0) {
                        for (int t = 0; t < NT; ++t) {
                            mbar_wait(&p_full[t], (j - 1) & 1);
                            tc_fence_after();
                            issue_pv(t, j - 1);
                        }
                        umma_commit(&kv_empty[(j - 1) % ST]);
                    }
                }
                for (int t = 0; t < NT; ++t) {
                    mbar_wait(&p_full[t], (n_kv - 1) & 1);
                    tc_fence_after();
                    issue_pv(t, n_kv - 1);
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax / output (thread = query row)
        setmaxnreg_inc<REGS_SOFTMAX>();
        const int t = warp >> 2;
        const int row = (warp & 3) * 32 + lane;                      // row inside the tile = TMEM lane
        const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tS = tmem_base + lane_base + TMEM_S + t * BKV;
        const uint32_t tP = tmem_base + lane_base + TMEM_P + t * (BKV / 2);
        const uint32_t tO = tmem_base + lane_base + TMEM_O + t * HD;
        const float sl2 = a.sl2;
        if (t == 1) named_bar_arrive(2, 2 * BQ);             // MUFU turn-taking: group 0 goes first
        float m_used = -INFINITY;        // exponent offset in use (<= true running max, lags it by at most RESCALE_LOG2)
        float l_run = 0.f;

        for (int j = 0; j < n_kv; ++j) {
            const int valid = a.skv - j * BKV;                       // >= BKV: the whole tile is real keys
            mbar_wait(&s_full[t], j & 1);
            tc_fence_after();
            uint32_t s[BKV];
#pragma unroll
            for (int c = 0; c < BKV / 32; ++c) tmem_ld_32x32_p(tS + c * 32, s + c * 32);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[t]);                  // S_t(j+1) may overwrite the TMEM copy now

            // ---- row max (padding keys of a ragged last tile -> -inf); four independent chains
            if (valid < BKV) {
#pragma unroll
                for (int i = 0; i < BKV; ++i)
                    if (i >= valid) s[i] = 0xff800000u;
            }
            float mx4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                mx4[c] = fmaxf(__uint_as_float(s[c * 32]), __uint_as_float(s[c * 32 + 1]));
#pragma unroll
                for (int i = 2; i < 32; i += 2)
                    mx4[c] = fmax3(mx4[c], __uint_as_float(s[c * 32 + i]), __uint_as_float(s[c * 32 + i + 1]));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));

            // ---- move the exponent offset only when the row max outgrew it by 2^RESCALE_LOG2 (warp-uniform decision:
            //      the TMEM accesses are warp-collective); rescaling a row that did not need it is exact (corr <= 1)
            if (j == 0) {
                m_used = mx;
            } else if (__any_sync(0xffffffffu, (mx - m_used) * sl2 > RESCALE_LOG2)) {
                mbar_wait(&pv_full[t], (j - 1) & 1);                 // O_t quiescent once PV_t(j-1) retired
                tc_fence_after();
                const float m_new = fmaxf(m_used, mx);
                const float corr = ex2_approx((m_used - m_new) * sl2);
                m_used = m_new;
                l_run *= corr;
                uint32_t r[HD];
                tmem_ld_32x32_p(tO, r);
                tmem_ld_32x32_p(tO + 32, r + 32);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < HD; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
                tmem_st_32x32_p(tO, r);
                tmem_st_32x32_p(tO + 32, r + 32);
            }
            // ---- P = exp2(S * c - m * c) -> fp16 pairs -> TMEM (A operand of the P.V MMA).  The MUFU is the scarcest unit
            //      of the whole kernel (16 exp2/clk/SM): the two softmax groups take turns on it, so that one group's
            //      loads / max / stores always run under the other group's exponentials instead of both stalling on it.
            const float msc = m_used * sl2;
            float sum0 = 0.f, sum1 = 0.f;
            named_bar_sync(2 + t, 2 * BQ);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t pk[BKV / 4];
#pragma unroll
                for (int i = 0; i < BKV / 2; i += 2) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(s[h * 64 + i]), sl2, -msc));    // exp2(-inf) = 0
                    const float p1 = ex2_approx(fmaf(__uint_as_float(s[h * 64 + i + 1]), sl2, -msc));
                    sum0 += p0;
                    sum1 += p1;
                    const __half2 hh = __floats2half2_rn(p0, p1);
                    pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                if (h == 0 && j > 0) {                               // P_t(j-1) consumed once PV_t(j-1) retired
                    mbar_wait(&pv_full[t], (j - 1) & 1);
                    tc_fence_after();
                }
                if (h == 1 && !(t == 1 && j == n_kv - 1)) named_bar_arrive(2 + (t ^ 1), 2 * BQ);   // hand the MUFU over
                tmem_st_32x32_p(tP + h * 32, pk);
            }
            l_run += sum0 + sum1;
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
        }
        // ---- normalise, store
        mbar_wait(&pv_full[t], (n_kv - 1) & 1);
        tc_fence_after();
        const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
        const int qrow = q0 + t * BQ + row;
        __half* orow = a.o + static_cast<long long>(b) * a.o_bs + static_cast<long long>(qrow) * a.o_ss + head * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(tO + c * 32, r);
            tmem_ld_wait();
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const __half2 h = __floats2half2_rn(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
                w[i >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            }
            if (qrow < a.sq) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    U32x8 v;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v.v[i] = w[q * 8 + i];
                    stg_256(orow + c * 32 + q * 16, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

bool g_attr_set = false;

}  // namespace

bool attention_tc_eligible(const AttnParams& p) {
    if (p.head_dim != HD || p.b_inner != 1 || p.kv_batch_div < 1) return false;
    if (p.sq < 2 * BQ || p.skv < BKV) return false;                 // short sequences stay on the warp-MMA kernel
    const long long strides[] = {p.q_bs, p.q_ss, p.k_bs, p.k_ss, p.v_bs, p.v_ss};
    for (long long s : strides)
        if (s <= 0 || (s & 7) != 0) return false;                   // TMA: 16 B multiples
    if ((p.o_bs & 15) != 0 || (p.o_ss & 15) != 0) return false;     // 32 B output stores
    const uintptr_t ptrs[] = {reinterpret_cast<uintptr_t>(p.q), reinterpret_cast<uintptr_t>(p.k),
                              reinterpret_cast<uintptr_t>(p.v)};
    for (uintptr_t x : ptrs)
        if (x & 15) return false;
    if (reinterpret_cast<uintptr_t>(p.o) & 31) return false;
    if (p.heads > 65535 || p.batch > 65535) return false;
    return true;
}

int attention_tc_plan(const AttnParams& p, AttnTcPlan* plan) {
    if (!attention_tc_eligible(p)) return -1;
    if (!g_attr_set) {
        if (cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL) != cudaSuccess) {
            fprintf(stderr, "[t2v_b200] attention_tc: cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            return -2;
        }
        g_attr_set = true;
    }
    const unsigned box[3] = {HD, BQ, 1};
    const int kvb = (p.batch + p.kv_batch_div - 1) / p.kv_batch_div;
    struct {
        CUtensorMap* m;
        const __half* base;
        long long bs, ss;
        int S, nb;
    } maps[3] = {{&plan->map_q, p.q, p.q_bs, p.q_ss, p.sq, p.batch},
                 {&plan->map_k, p.k, p.k_bs, p.k_ss, p.skv, kvb},
                 {&plan->map_v, p.v, p.v_bs, p.v_ss, p.skv, kvb}};
    for (auto& m : maps) {
        const unsigned long long dims[3] = {static_cast<unsigned long long>(p.heads) * HD, static_cast<unsigned long long>(m.S),
                                            static_cast<unsigned long long>(m.nb)};
        const unsigned long long str[2] = {static_cast<unsigned long long>(m.ss) * 2, static_cast<unsigned long long>(m.bs) * 2};
        if (tma_encode_f16(m.m, m.base, 3, dims, str, box) != 0) return -3;
    }
    plan->o = p.o;
    plan->o_bs = p.o_bs;
    plan->o_ss = p.o_ss;
    plan->sq = p.sq;
    plan->skv = p.skv;
    plan->kv_batch_div = p.kv_batch_div;
    plan->batch = p.batch;
    plan->heads = p.heads;
    plan->sl2 = p.scale * 1.4426950408889634f;
    return 0;
}

int attention_tc_launch(const AttnTcPlan& pl, cudaStream_t stream) {
    Args a;
    a.o = pl.o;
    a.o_bs = pl.o_bs;
    a.o_ss = pl.o_ss;
    a.sq = pl.sq;
    a.skv = pl.skv;
    a.kv_batch_div = pl.kv_batch_div;
    a.n_kv = (pl.skv + BKV - 1) / BKV;
    a.sl2 = pl.sl2;
    dim3 grid((pl.sq + NT * BQ - 1) / (NT * BQ), pl.heads, pl.batch);
    launch_pdl(attention_tc_kernel, grid, NTHREADS, SMEM_TOTAL, stream, pl.map_q, pl.map_k, pl.map_v, a);
    return launch_status("attention_tc launch");
}

}  // namespace t2v
