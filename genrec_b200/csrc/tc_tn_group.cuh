// genrec_b200 - grouped weight-gradient GEMM:  for each problem p:  Out_p[M_p, N_p] += A_p^T B_p   (A_p stored [K, M_p],
// B_p stored [K, N_p], both row-major bf16; K = number of tokens).  One persistent launch covers all the weight
// gradients of an HSTU layer (dWp, dW1, dW2): every problem is split along K into enough work items to fill the chip,
// a CTA walks its items back to back so the atomic epilogue of one item overlaps the TMA/MMA main loop of the next
// (TMEM accumulators are double-buffered), and partial tiles are merged with 16-byte vector reductions
// (red.global.add.v4.f32) straight into the flat fp32 gradient buffer.
#pragma once
#include "tc_gemm.cuh"

namespace grb {

constexpr int TN_MAX_PROBLEMS = 4;

struct TnProblem {
    int M, N, K;
    int num_m, num_n, splits, kb_total, kb_per_split;
    int work_begin;  // first global work-item index of this problem
    float* out;
    int ldo;
};
struct TnGroupParams {
    CUtensorMap tmA[TN_MAX_PROBLEMS];
    CUtensorMap tmB[TN_MAX_PROBLEMS];
    TnProblem p[TN_MAX_PROBLEMS];
    int nprob;
    int total_work;
};


struct TnItem {
    int prob, m0, n0, kb0, kb1;
};
GRB_DEVINL TnItem tn_decode(const TnGroupParams& P, int w) {
    int pi = 0;
#pragma unroll
    for (int i = 1; i < TN_MAX_PROBLEMS; ++i)
        if (i < P.nprob && w >= P.p[i].work_begin) pi = i;
    const TnProblem& q = P.p[pi];
    const int local = w - q.work_begin;
    const int split = local % q.splits, tile = local / q.splits;
    TnItem it;
    it.prob = pi;
    it.m0 = (tile / q.num_n) * TC_BM;
    it.n0 = (tile % q.num_n) * TC_BN;
    it.kb0 = split * q.kb_per_split;
    it.kb1 = min(q.kb_total, it.kb0 + q.kb_per_split);
    return it;
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_tn_group_kernel(const __grid_constant__ TnGroupParams P) {
    extern __shared__ unsigned char tn_smem_raw[];
    unsigned char* base = tn_smem_raw + ((1024u - (smem_u32(tn_smem_raw) & 1023u)) & 1023u)   /* offset from the __shared__ array: keeps the shared address space (LDS / STS) */;
    constexpr int STAGES = 6;
    unsigned char* sA = base;
    unsigned char* sB = base + STAGES * TC_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * STAGES * TC_TILE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < P.nprob; ++i) { tma_prefetch_desc(&P.tmA[i]); tma_prefetch_desc(&P.tmB[i]); }
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], TC_EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // prologue above overlaps the previous kernel's tail

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
                const TnItem it = tn_decode(P, w);
                const CUtensorMap* ta = &P.tmA[it.prob];
                const CUtensorMap* tb = &P.tmB[it.prob];
                for (int kb = it.kb0; kb < it.kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], 2 * TC_TILE_BYTES);
                    unsigned char* a_dst = sA + stage * TC_TILE_BYTES;
                    unsigned char* b_dst = sB + stage * TC_TILE_BYTES;
                    const int k0 = kb * TC_BK;
                    tma_load_2d(a_dst, ta, it.m0, k0, &full_bar[stage]);
                    tma_load_2d(a_dst + TC_TILE_BYTES / 2, ta, it.m0 + 64, k0, &full_bar[stage]);
                    tma_load_2d(b_dst, tb, it.n0, k0, &full_bar[stage]);
                    tma_load_2d(b_dst + TC_TILE_BYTES / 2, tb, it.n0 + 64, k0, &full_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(TC_BM, TC_BN, 1, 1);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
                const TnItem it = tn_decode(P, w);
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * TC_BN;
                for (int kb = it.kb0; kb < it.kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sA + stage * TC_TILE_BYTES);
                    const uint32_t b_addr = smem_u32(sB + stage * TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)
                        umma_bf16(d_tmem, umma_desc(a_addr + k * 2048, TC_TILE_BYTES / 2, 1024), umma_desc(b_addr + k * 2048, TC_TILE_BYTES / 2, 1024),
                                  idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int sub = warp & 3, cq = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        for (int w = blockIdx.x; w < P.total_work; w += gridDim.x) {
            const TnItem it = tn_decode(P, w);
            const TnProblem& q = P.p[it.prob];
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const int row = it.m0 + sub * 32 + lane;
#pragma unroll 1
            for (int c = cq * TC_EPI_CPW; c < (cq + 1) * TC_EPI_CPW; ++c) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * TC_BN + c * 32), v);
                const int col0 = it.n0 + c * 32;
                if (row < q.M && col0 < q.N) {
                    float* dst = q.out + (size_t)row * q.ldo + col0;
                    if (col0 + 32 <= q.N && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    } else {
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i < q.N) atomicAdd(dst + i, v[i]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

constexpr int TN_SMEM_BYTES = 2 * 6 * TC_TILE_BYTES + 1024 + 256;

struct TnSpec {
    const bf16* A; const bf16* B; float* out;
    int M, N, K, lda, ldb, ldo;
};

inline cudaError_t launch_tc_tn_group(const TnSpec* specs, int n, int num_sms, cudaStream_t st) {
    if (n < 1 || n > TN_MAX_PROBLEMS) return cudaErrorInvalidValue;
    TnGroupParams P;
    memset(&P, 0, sizeof(P));
    P.nprob = n;
    // give every problem a share of (2 work items per SM) proportional to its FLOPs
    double total = 0;
    for (int i = 0; i < n; ++i) total += (double)specs[i].M * specs[i].N * specs[i].K;
    int work = 0;
    for (int i = 0; i < n; ++i) {
        const TnSpec& s = specs[i];
        if (!make_tmap_bf16(&P.tmA[i], s.A, s.K, s.M, s.lda, 64, TC_BK) || !make_tmap_bf16(&P.tmB[i], s.B, s.K, s.N, s.ldb, 64, TC_BK))
            return cudaErrorInvalidValue;
        TnProblem& q = P.p[i];
        q.M = s.M; q.N = s.N; q.K = s.K; q.out = s.out; q.ldo = s.ldo;
        q.num_m = (s.M + TC_BM - 1) / TC_BM;
        q.num_n = (s.N + TC_BN - 1) / TC_BN;
        q.kb_total = (s.K + TC_BK - 1) / TC_BK;
        const int tiles = q.num_m * q.num_n;
        int want_items = (int)(2.0 * num_sms * ((double)s.M * s.N * s.K / total) + 0.5);
        int splits = (want_items + tiles - 1) / tiles;
        if (splits < 1) splits = 1;
        int max_splits = q.kb_total / 4 > 0 ? q.kb_total / 4 : 1;  // at least 4 k-blocks per item
        if (splits > max_splits) splits = max_splits;
        q.kb_per_split = (q.kb_total + splits - 1) / splits;
        q.splits = (q.kb_total + q.kb_per_split - 1) / q.kb_per_split;
        q.work_begin = work;
        work += tiles * q.splits;
    }
    P.total_work = work;
    static bool attr_set_dev[64] = {false};
    int attr_dev = 0;
    cudaGetDevice(&attr_dev);
    bool& attr_set = attr_set_dev[attr_dev & 63];   // the attribute is per device
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_tn_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TN_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int grid = work < num_sms ? work : num_sms;
    launch_k(tc_tn_group_kernel, grid, TC_THREADS, TN_SMEM_BYTES, st, P);
    return cudaGetLastError();
}

}  // namespace grb
