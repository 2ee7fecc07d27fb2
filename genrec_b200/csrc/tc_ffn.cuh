// genrec_b200 - fused position-wise feed-forward, forward:   y = x1 + drop( drop(silu(xn W1^T + b1)) W2^T + b2 )
// (genrec/models/hstu.py:209-218, :278), one tcgen05 kernel instead of two GEMM launches.
//
// One CTA owns a block of 128 token rows.  Its LN2 output tile xn[128, D] stays resident in shared memory; the hidden
// dimension 4D is walked in chunks of 128:
//     S_c  = xn  W1[c]^T                (TMEM, double-buffered)                                      first MMA
//     z_c  = bf16(S_c + b1[c])   -> global (saved for the backward, straight from registers)
//     h_c  = drop(silu(z_c))     -> bf16, 128B-swizzled staging tile -> TMA store (saved for dW2 = dy^T h)
//     Y   += h_c W2[:, c]^T             the SAME staging tile is the K-major A operand of the second MMA; Y lives in TMEM
// and after the last chunk   y = x1 + drop(Y + b2)   leaves from registers as fp32.
// The hidden activation is therefore written once (for the backward) and never re-read in the forward: against the two
// separate GEMMs this saves the 2 x T x 4D bytes re-read of h, one launch, and - because a CTA now carries 4x the work per
// row block - most of the fixed cost and tile-quantisation loss of the N = D GEMM (200 tiles on 148 SMs).
//
// Warp roles as in tc_ce.cuh: warp 0 TMA producer (weights stream from L2 through one in-order ring of 16 KB slots, in
// exactly the order the MMA warp consumes them), warp 1 MMA issuer, 16 epilogue warps (4 per TMEM sub-partition).
#pragma once
#include "tc_gemm.cuh"

namespace grb {

constexpr int FFN_EPI_WARPS = 16;
constexpr int FFN_THREADS = 64 + 32 * FFN_EPI_WARPS;
constexpr int FFN_RING = 6;   // 16 KB slots, consumed strictly in order: W1 chunk = KB slots, W2 chunk = 2 slots
GRB_DEVINL void ffn_group_sync(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(32 * FFN_EPI_WARPS / 2) : "memory"); }   // the 8 warps of one epilogue group

template <int KB>
struct FfnSmem {
    static constexpr int kBytes = KB * TC_TILE_BYTES + FFN_RING * TC_TILE_BYTES + 2 * 32768 + 5 * 64 * KB * 4 + 1024 + 256;
};

struct FfnShape {
    int T, num_m;
};
struct FfnEpiArgs {
    const float* b1;       // [4D]
    const float* b2;       // [D]
    const float* x1;       // [T, D] fp32 residual
    bf16* z1;              // [T, 4D] pre-activation (saved)
    float* y;              // [T, D]
    Dropout drop_hid, drop_out;
};

template <int KB>  // D = 64 * KB  (KB = 1, 2)
__global__ void __launch_bounds__(FFN_THREADS, 1)
    tc_ffn_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                      const __grid_constant__ CUtensorMap tmH, FfnShape sh, FfnEpiArgs ea) {
    constexpr int D = 64 * KB, NC = 4 * D / 128, NS = FFN_RING;
    extern __shared__ unsigned char ffn_smem_raw[];
    unsigned char* base = ffn_smem_raw + ((1024u - (smem_u32(ffn_smem_raw) & 1023u)) & 1023u)   /* offset from the __shared__ array: keeps the shared address space (LDS / STS) */;
    unsigned char* sX = base;                                   // KB x 16 KB: xn tile, K-major, resident per row block
    unsigned char* sW = sX + KB * TC_TILE_BYTES;                // ring of weight slices
    unsigned char* sH0 = sW + NS * TC_TILE_BYTES;               // 2 x 32 KB staging: h chunk = TMA-store source + A operand of MMA 2
    float* s_b1 = reinterpret_cast<float*>(sH0 + 2 * 32768);         // [4D] first-layer bias, then [D] second-layer bias
    float* s_b2 = s_b1 + 4 * D;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_b2 + D);
    uint64_t* wfull = bars;                 // [NS]  TMA -> MMA
    uint64_t* wempty = bars + NS;           // [NS]  MMA -> TMA
    uint64_t* tfull = bars + 2 * NS;        // [2]   S accumulator ready
    uint64_t* tempty = bars + 2 * NS + 2;   // [2]   S accumulator drained
    uint64_t* xfull = bars + 2 * NS + 4;
    uint64_t* xempty = bars + 2 * NS + 5;
    uint64_t* gfull = bars + 2 * NS + 6;    // [2]   h staging tile written (epilogue -> MMA)
    uint64_t* gempty = bars + 2 * NS + 8;   // [2]   second MMA has read the staging tile
    uint64_t* yfull = bars + 2 * NS + 10;   //       Y accumulator complete
    uint64_t* yempty = bars + 2 * NS + 11;  //       Y accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmH);
        for (int s = 0; s < NS; ++s) { mbar_init(&wfull[s], 1); mbar_init(&wempty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1); mbar_init(&tempty[a], FFN_EPI_WARPS / 2);   // accumulator a belongs to epilogue group a
            mbar_init(&gfull[a], 1); mbar_init(&gempty[a], 1);
        }
        mbar_init(xfull, 1);
        mbar_init(xempty, 1);
        mbar_init(yfull, 1);
        mbar_init(yempty, FFN_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    for (int i = threadIdx.x; i < 5 * D; i += FFN_THREADS) s_b1[i] = i < 4 * D ? ea.b1[i] : ea.b2[i - 4 * D];   // parameters: not written by the previous kernel
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();
    ea.drop_hid.resolve();
    ea.drop_out.resolve();
    const uint32_t tmem_y = tmem_base + 256;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, xphase = 0;
            auto load_w1 = [&](int c) {      // W1 rows [128c, 128c + 128), all of K = D : KB slices [128 n][64 k]
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&wempty[stage], phase ^ 1);
                    mbar_expect_tx(&wfull[stage], TC_TILE_BYTES);
                    tma_load_2d(sW + stage * TC_TILE_BYTES, &tmW1, kb * 64, c * 128, &wfull[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            };
            auto load_w2 = [&](int c) {      // W2 rows [0, D) (n), hidden columns [128c, 128c + 128) (k): 2 slices [D n][64 k]
                for (int kb = 0; kb < 2; ++kb) {
                    mbar_wait(&wempty[stage], phase ^ 1);
                    mbar_expect_tx(&wfull[stage], D * 128);
                    tma_load_2d(sW + stage * TC_TILE_BYTES, &tmW2, c * 128 + kb * 64, 0, &wfull[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
            };
            for (int blk = blockIdx.x; blk < sh.num_m; blk += gridDim.x) {
                mbar_wait(xempty, xphase ^ 1);
                mbar_expect_tx(xfull, KB * TC_TILE_BYTES);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(sX + kb * TC_TILE_BYTES, &tmX, kb * 64, blk * 128, xfull);
                xphase ^= 1;
                // same order as the MMA warp consumes: S(0), then for c >= 1: S(c), Y-part(c-1), and Y-part(NC-1) last
                load_w1(0);
                for (int c = 1; c < NC; ++c) { load_w1(c); load_w2(c - 1); }
                load_w2(NC - 1);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc1 = umma_idesc(128, 128, 0, 0);   // S = xn W1[c]^T   (both K-major)
            constexpr uint32_t idesc2 = umma_idesc(128, D, 0, 0);     // Y += h_c W2[:, c]^T
            int stage = 0; uint32_t phase = 0, xphase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int gb = 0; uint32_t gphase = 0, yphase = 0;
            auto gemm2 = [&](int buf, bool first) {
                const uint32_t a_addr = smem_u32(sH0 + buf * 32768);
                for (int kb = 0; kb < 2; ++kb) {     // 128 hidden columns of the chunk = two staging boxes of 64
                    mbar_wait(&wfull[stage], phase);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(sW + stage * TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_y, umma_desc(a_addr + kb * 16384 + k * 32, 16, 1024), umma_desc(b_addr + k * 32, 16, 1024), idesc2,
                                  (first && kb == 0 && k == 0) ? 0u : 1u);
                    umma_commit(&wempty[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                umma_commit(&gempty[buf]);
            };
            for (int blk = blockIdx.x; blk < sh.num_m; blk += gridDim.x) {
                mbar_wait(xfull, xphase);
                xphase ^= 1;
                tc_fence_after();
                for (int c = 0; c < NC; ++c) {
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 128;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&wfull[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sX + kb * TC_TILE_BYTES);
                        const uint32_t b_addr = smem_u32(sW + stage * TC_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, umma_desc(a_addr + k * 32, 16, 1024), umma_desc(b_addr + k * 32, 16, 1024), idesc1,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&wempty[stage]);
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    if (c == NC - 1) umma_commit(xempty);   // the last S MMA of this row block: sX may be refilled once it retires
                    if (c >= 1) {
                        if (c == 1) { mbar_wait(yempty, yphase ^ 1); tc_fence_after(); }
                        mbar_wait(&gfull[gb], gphase);
                        tc_fence_after();
                        gemm2(gb, c == 1);
                        if (gb == 1) gphase ^= 1;
                        gb ^= 1;
                    }
                }
                if (NC == 1) { mbar_wait(yempty, yphase ^ 1); tc_fence_after(); }
                mbar_wait(&gfull[gb], gphase);
                tc_fence_after();
                gemm2(gb, NC == 1);
                if (gb == 1) gphase ^= 1;
                gb ^= 1;
                umma_commit(yfull);
                yphase ^= 1;
            }
        }
    } else {
        // ===================================================================== epilogue: two groups of 8 warps
        // Group g owns the chunks c = g, g + 2, ... : accumulator g, staging buffer g.  While one group converts a chunk
        // the other one is reading its accumulator or writing its staging tile, so the TMEM-read, MUFU and shared-memory
        // phases of neighbouring chunks overlap instead of running back to back.
        const int ew = warp - 2, grp = ew >> 3;
        const int sub = warp & 3;             // TMEM sub-partition of this warp
        const int ch = (ew & 7) >> 2;         // which 64-column half of a chunk this warp converts (two 32-column passes)
        const int cq = ew >> 2;               // column quarter for the final Y pass (all 16 warps)
        const int r = sub * 32 + lane;
        const bool leader = (ew & 7) == 0 && lane == 0;
        uint32_t tphase = 0, gphase = 0, yphase = 0;
        for (int blk = blockIdx.x; blk < sh.num_m; blk += gridDim.x) {
            const int row = blk * 128 + r;
            const bool live = row < sh.T;
            for (int c = grp; c < NC; c += 2) {
                uint32_t hp[2][16];
                mbar_wait(&tfull[grp], tphase);
                tc_fence_after();
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int col0 = c * 128 + ch * 64 + q * 32;      // hidden column of this pass's first element
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(grp * 128 + ch * 64 + q * 32), v);
                    if (q == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[grp]);   // accumulator drained (it lives in registers now)
                    }
                    // z = bf16(acc + b1) -> global ; h = drop(silu(z)) -> packed bf16
                    uint32_t zp[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 b = *reinterpret_cast<const float2*>(s_b1 + col0 + i);
                        zp[i >> 1] = pack_bf16(v[i] + b.x, v[i + 1] + b.y);
                        const float2 zr = unpack_bf16(zp[i >> 1]);
                        float h0 = siluf(zr.x), h1 = siluf(zr.y);
                        ea.drop_hid.apply2p(h0, h1, (uint32_t)row, (uint32_t)((col0 + i) >> 1));
                        hp[q][i >> 1] = pack_bf16(h0, h1);
                    }
                    if (live) {
                        uint4* zdst = reinterpret_cast<uint4*>(ea.z1 + (size_t)row * (4 * D) + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) zdst[j] = make_uint4(zp[4 * j], zp[4 * j + 1], zp[4 * j + 2], zp[4 * j + 3]);
                    }
                }
                tphase ^= 1;
                // this group's previous bulk store (chunk c - 2) must have finished READING the staging tile, and the second
                // MMA of that chunk must have retired, before the tile is rewritten
                if (leader) tma_store_wait_read();
                mbar_wait(&gempty[grp], gphase ^ 1);
                ffn_group_sync(grp);
                unsigned char* dst = sH0 + grp * 32768 + ch * 16384 + r * 128;
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4*>(dst + (((q * 4 + j) ^ (r & 7)) << 4)) =
                            make_uint4(hp[q][4 * j], hp[q][4 * j + 1], hp[q][4 * j + 2], hp[q][4 * j + 3]);
                fence_proxy_async();   // generic-proxy smem writes -> visible to TMA and to tcgen05.mma (async proxy)
                ffn_group_sync(grp);
                if (leader) {
                    mbar_arrive(&gfull[grp]);
                    tma_store_2d(&tmH, sH0 + grp * 32768, c * 128, blk * 128);
                    tma_store_2d(&tmH, sH0 + grp * 32768 + 16384, c * 128 + 64, blk * 128);
                    tma_store_commit();
                }
                gphase ^= 1;
            }
            // ---------------------------------------------------------------------- y = x1 + drop(Y + b2)
            if (cq * 32 < D) {
                const int ycol = cq * 32;
                float res[32];
                if (live) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 f = *(reinterpret_cast<const float4*>(ea.x1 + (size_t)row * D + ycol) + j);
                        res[4 * j] = f.x; res[4 * j + 1] = f.y; res[4 * j + 2] = f.z; res[4 * j + 3] = f.w;
                    }
                }
                mbar_wait(yfull, yphase);
                tc_fence_after();
                float v[32];
                tmem_ld32(tmem_y + ((uint32_t)(sub * 32) << 16) + (uint32_t)ycol, v);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(yempty);
                if (live) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 b = *reinterpret_cast<const float2*>(s_b2 + ycol + i);
                        float y0 = v[i] + b.x, y1 = v[i + 1] + b.y;
                        ea.drop_out.apply2p(y0, y1, (uint32_t)row, (uint32_t)((ycol + i) >> 1));
                        v[i] = res[i] + y0;
                        v[i + 1] = res[i + 1] + y1;
                    }
                    float4* ydst = reinterpret_cast<float4*>(ea.y + (size_t)row * D + ycol);
#pragma unroll
                    for (int j = 0; j < 8; ++j) ydst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
            } else {
                mbar_wait(yfull, yphase);
                if (lane == 0) mbar_arrive(yempty);
            }
            yphase ^= 1;
        }
        if (leader) tma_store_wait_read();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// xn [T, D] bf16, W1 [4D, D] bf16, W2 [D, 4D] bf16 -> z1, h [T, 4D] bf16 (saved), y [T, D] fp32
template <int KB>
inline cudaError_t launch_tc_ffn_fwd(const bf16* xn, const bf16* w1, const bf16* w2, bf16* h, int T, const FfnEpiArgs& ea, int num_sms, cudaStream_t st) {
    constexpr int D = 64 * KB;
    CUtensorMap tmX, tmW1, tmW2, tmH;
    bool ok = make_tmap_bf16(&tmX, xn, T, D, D, 64, 128) && make_tmap_bf16(&tmW1, w1, 4 * D, D, D, 64, 128) &&
              make_tmap_bf16(&tmW2, w2, D, 4 * D, 4 * D, 64, D) && make_tmap(&tmH, h, false, T, 4 * D, 4 * D, 64, 128);
    if (!ok) return cudaErrorInvalidValue;
    FfnShape sh;
    sh.T = T;
    sh.num_m = (T + 127) / 128;
    auto kern = tc_ffn_fwd_kernel<KB>;
    static bool attr_set_dev[64] = {false};
    int attr_dev = 0;
    cudaGetDevice(&attr_dev);
    bool& attr_set = attr_set_dev[attr_dev & 63];   // the attribute is per device
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem<KB>::kBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const int grid = sh.num_m < num_sms ? sh.num_m : num_sms;
    launch_k(kern, grid, FFN_THREADS, FfnSmem<KB>::kBytes, st, tmX, tmW1, tmW2, tmH, sh, ea);
    return cudaGetLastError();
}

}  // namespace grb
