// genrec_b200 - fused tied-embedding logits + cross-entropy (forward loss AND d(loss)/d(logits)) on tcgen05.
//
// Replaces `logits = x @ E^T ; loss = cross_entropy(logits, targets, ignore_index=0)` (genrec/models/hstu.py:137-146) and
// the autograd softmax-backward.  The [T, C] logits tensor is never written: one CTA owns a block of 128 token rows and
// sweeps the C classes twice with the same TMA -> tcgen05.mma pipeline (the token tile stays resident in shared memory,
// the embedding table streams from L2):
//   sweep 0 : S = X E_n^T in TMEM  ->  per-row running (max, sum exp) + the target logit, one thread per row
//   sweep 1 : S recomputed         ->  g = (exp(S - lse) - onehot(target)) * inv_count  ->  bf16 -> smem -> TMA store
// so HBM sees one write of dlogits (bf16) and nothing else; the 2x K=D recompute is cheap (D <= 256).
// dlogits then feeds the two gradient GEMMs (dX = g E, dE = g^T X) of tc_gemm.cuh.
#pragma once
#include "tc_gemm.cuh"

namespace grb {

constexpr int CE_BSTAGES = 5;     // ring of 16 KB (128 classes x 64 k) slices of the table
constexpr int CE_THREADS = TC_THREADS;   // TMA, MMA, 16 epilogue warps

template <int KB>
constexpr int ce_smem_bytes() {
    return KB * TC_TILE_BYTES + CE_BSTAGES * TC_TILE_BYTES + 2 * 32768 + 8 * 128 * 4 + 1024 + 256;
}

struct CeShape {
    int T, C, ldl;       // tokens, classes, leading dimension of dlogits (multiple of 8, >= C)
    int num_m, num_n;
};

template <int KB>  // k-blocks of 64: D = 64 * KB
__global__ void __launch_bounds__(CE_THREADS, 1)
    tc_ce_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmG,
                 CeShape sh, const long long* __restrict__ targets, const float* __restrict__ inv_count, float* __restrict__ loss) {
    extern __shared__ unsigned char ce_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ce_smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sX = base;                                   // KB x 16 KB, resident per row block
    unsigned char* sE = sX + KB * TC_TILE_BYTES;                // ring
    unsigned char* sOut0 = sE + CE_BSTAGES * TC_TILE_BYTES;     // 2 x 32 KB staging
    float* s_part = reinterpret_cast<float*>(sOut0 + 2 * 32768);  // [4 column quarters][{max, sum}][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_part) + 8 * 128 * 4);
    uint64_t* efull = bars;                      // [CE_BSTAGES]
    uint64_t* eempty = bars + CE_BSTAGES;        // [CE_BSTAGES]
    uint64_t* tfull = bars + 2 * CE_BSTAGES;     // [2]
    uint64_t* tempty = bars + 2 * CE_BSTAGES + 2;  // [2]
    uint64_t* xfull = bars + 2 * CE_BSTAGES + 4;   // [1]
    uint64_t* xempty = bars + 2 * CE_BSTAGES + 5;  // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * CE_BSTAGES + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmE);
        tma_prefetch_desc(&tmG);
        for (int s = 0; s < CE_BSTAGES; ++s) { mbar_init(&efull[s], 1); mbar_init(&eempty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], TC_EPI_WARPS); }
        mbar_init(xfull, 1);
        mbar_init(xempty, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_block = 2 * sh.num_n;  // two sweeps

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, xphase = 0;
            for (int blk = blockIdx.x; blk < sh.num_m; blk += gridDim.x) {
                mbar_wait(xempty, xphase ^ 1);
                mbar_expect_tx(xfull, KB * TC_TILE_BYTES);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(sX + kb * TC_TILE_BYTES, &tmX, kb * 64, blk * 128, xfull);
                xphase ^= 1;
                for (int tile = 0; tile < tiles_per_block; ++tile) {
                    const int n0 = (tile % sh.num_n) * 128;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&eempty[stage], phase ^ 1);
                        mbar_expect_tx(&efull[stage], TC_TILE_BYTES);
                        tma_load_2d(sE + stage * TC_TILE_BYTES, &tmE, kb * 64, n0, &efull[stage]);
                        if (++stage == CE_BSTAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc(128, 128, 0, 0);
            int stage = 0; uint32_t phase = 0, xphase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int blk = blockIdx.x; blk < sh.num_m; blk += gridDim.x) {
                mbar_wait(xfull, xphase);
                xphase ^= 1;
                tc_fence_after();
                for (int tile = 0; tile < tiles_per_block; ++tile) {
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 128;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&efull[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sX + kb * TC_TILE_BYTES);
                        const uint32_t b_addr = smem_u32(sE + stage * TC_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, umma_desc(a_addr + k * 32, 16, 1024), umma_desc(b_addr + k * 32, 16, 1024), idesc,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&eempty[stage]);
                        if (++stage == CE_BSTAGES) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                umma_commit(xempty);  // every MMA that reads this row block has retired -> the producer may overwrite sX
            }
        }
    } else {
        const int sub = warp & 3, chalf = (warp - 2) >> 2;   // chalf: 32-column quarter 0..3 of the tile
        const int r = sub * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        const float ic = *inv_count;
        for (int blk = blockIdx.x; blk < sh.num_m; blk += gridDim.x) {
            const int row = blk * 128 + r;
            const int t = row < sh.T ? (int)targets[row] : 0;
            const float icr = t != 0 ? ic : 0.f;   // ignore_index = 0 (and rows past the end)
            float m_run = -INFINITY, s_run = 0.f, tl = 0.f;
            // ------------------------------------------------------------------ sweep 0: statistics
            for (int n = 0; n < sh.num_n; ++n) {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = chalf; c < chalf + 1; ++c) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * 128 + c * 32), v);
                    const int col0 = n * 128 + c * 32;
                    if (col0 + 32 > sh.C) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i >= sh.C) v[i] = -INFINITY;
                    }
                    if ((unsigned)(t - col0) < 32u) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i == t - col0) tl = v[i];
                    }
                    float cm = v[0];
#pragma unroll
                    for (int i = 1; i < 32; ++i) cm = fmaxf(cm, v[i]);
                    const float m_new = fmaxf(m_run, cm);
                    if (m_new > -INFINITY) {
                        float cs = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; ++i) cs += __expf(v[i] - m_new);
                        s_run = s_run * __expf(m_run - m_new) + cs;
                        m_run = m_new;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            // combine the two column halves of every row
            s_part[(chalf * 2 + 0) * 128 + r] = m_run;
            s_part[(chalf * 2 + 1) * 128 + r] = s_run;
            float* s_tl = s_part;  // re-used after the barrier below
            epi_bar_sync();
            float mm = -INFINITY;
#pragma unroll
            for (int q = 0; q < 4; ++q) mm = fmaxf(mm, s_part[(q * 2 + 0) * 128 + r]);
            float ssum = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) ssum += s_part[(q * 2 + 1) * 128 + r] * __expf(s_part[(q * 2 + 0) * 128 + r] - mm);
            const float lse = mm + __logf(ssum);
            // loss: the half that saw the target column contributes -tl, half 0 contributes +lse
            float contrib = (chalf == 0 ? lse : 0.f) - tl;
            contrib = warp_sum(contrib * icr);
            if (lane == 0 && contrib != 0.f) atomicAdd(loss, contrib);
            epi_bar_sync();  // everybody has read s_part before the next block overwrites it
            (void)s_tl;
            // ------------------------------------------------------------------ sweep 1: gradient tiles
            if (warp == 2 && lane == 0) tma_store_wait_read();  // both staging buffers are free (previous block's stores drained)
            epi_bar_sync();
            for (int n = 0; n < sh.num_n; ++n) {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                unsigned char* sOut = sOut0 + acc * 32768;
#pragma unroll 1
                for (int c = chalf; c < chalf + 1; ++c) {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * 128 + c * 32), v);
                    const int col0 = n * 128 + c * 32;
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __expf(v[i] - lse) * icr;
                    if (col0 + 32 > sh.C) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i >= sh.C) v[i] = 0.f;
                    }
                    if ((unsigned)(t - col0) < 32u) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i == t - col0) v[i] -= icr;
                    }
                    unsigned char* dst = sOut + (c >> 1) * 16384 + r * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 u;
                        u.x = pack_bf16(v[8 * j], v[8 * j + 1]); u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                        u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<uint4*>(dst + ((((c & 1) * 4 + j) ^ (r & 7)) << 4)) = u;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                fence_proxy_async();
                epi_bar_sync();
                if (warp == 2 && lane == 0) {
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (n * 128 + h * 64 < sh.ldl) tma_store_2d(&tmG, sOut + h * 16384, n * 128 + h * 64, blk * 128);
                    tma_store_commit();
                    tma_store_wait_read1();
                }
                epi_bar_sync();
            }
        }
        if (warp == 2 && lane == 0) tma_store_wait_read();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

// X [T, D] bf16, E [C, D] bf16 -> G [T, ldl] bf16 (columns >= C zeroed), loss += sum_rows(lse - logit[target]) * inv_count
template <int KB>
inline cudaError_t launch_tc_ce(const bf16* X, const bf16* E, bf16* G, int T, int C, int ldl, const long long* targets, const float* inv_count,
                                float* loss, int num_sms, cudaStream_t st) {
    CUtensorMap tmX, tmE, tmG;
    const int D = 64 * KB;
    bool ok = make_tmap_bf16(&tmX, X, T, D, D, 64, 128) && make_tmap_bf16(&tmE, E, C, D, D, 64, 128) && make_tmap(&tmG, G, false, T, ldl, ldl, 64, 128);
    if (!ok) return cudaErrorInvalidValue;
    CeShape sh;
    sh.T = T; sh.C = C; sh.ldl = ldl;
    sh.num_m = (T + 127) / 128;
    sh.num_n = (C + 127) / 128;
    auto kern = tc_ce_kernel<KB>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ce_smem_bytes<KB>());
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    int grid = sh.num_m < num_sms ? sh.num_m : num_sms;
    kern<<<grid, CE_THREADS, ce_smem_bytes<KB>(), st>>>(tmX, tmE, tmG, sh, targets, inv_count, loss);
    return cudaGetLastError();
}

}  // namespace grb
