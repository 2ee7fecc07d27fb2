// genrec_b200 - fused tied-embedding logits + cross-entropy: loss, d(loss)/d(logits) AND d(loss)/d(x) on tcgen05.
//
// Replaces `logits = x @ E^T ; loss = cross_entropy(logits, targets, ignore_index=0)` (genrec/models/hstu.py:137-146), the
// autograd softmax-backward and the `dlogits @ E` GEMM.  The [T, C] logits tensor is never written: one CTA owns a block
// of 128 token rows (token tile resident in shared memory) and sweeps the C classes twice with the same TMA -> tcgen05.mma
// pipeline, the embedding table streaming from L2:
//   sweep 0 : S = X E_n^T in TMEM  ->  per-row running MAX + the target logit, one thread per (row, 32 columns) - no exponentials
//   sweep 1 : S recomputed         ->  G' = exp(S - max) * inv_count (UNNORMALISED: the row sum is only known once this sweep
//                                      has seen every class) + the row sum  ->  bf16, 128B-swizzled staging tile
//   Every logit therefore costs ONE exponential (the kernel is MUFU-bound: the first version spent one per sweep).  The
//   normalisation 1 / sum_row commutes with both products that consume G': ce_finish_kernel scales dX' = G' E by it, subtracts the
//   one-hot term, and hands the dE GEMM x / sum_row instead of x.
//                                      -> TMA store to dlogits (consumed by the dE = G^T X GEMM)
//                                      -> (D <= 128) the SAME staging tile is the K-major A operand of a second MMA
//                                         dX[128, D] += G_tile E_n, whose B operand is the E_n tile already in shared memory
//                                         read through an MN-major descriptor; dX accumulates in TMEM over all classes.
// HBM sees one bf16 write of dlogits (620 MB at cfg-2) and one fp32 write of dX; the K = D recompute is cheap (D <= 256).
//
// Load balance: T/128 row blocks rarely divide by the SM count (200 blocks on 148 SMs = 2 rounds at 68 % occupancy), so a
// work item is HALF the class range of a row block: items (2b, 2b+1) always run in the same round on neighbouring CTAs
// (the grid is persistent and even-sized), exchange their per-row (max, sum, target-logit) partials through global memory
// with a release/acquire flag after sweep 0, and each finishes sweep 1 on its own half; dX partials meet in a pre-zeroed
// fp32 buffer through 16-byte vector reductions.  400 half-items on 148 CTAs = 2.7 rounds of half the length.
#pragma once
#include "tc_gemm.cuh"

namespace grb {

constexpr int CE_EPI_WARPS = 16;                 // 4 per TMEM sub-partition, one 32-column quarter of the class tile each
constexpr int CE_THREADS = 64 + 32 * CE_EPI_WARPS;   // TMA, MMA, epilogue warps
GRB_DEVINL void ce_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * CE_EPI_WARPS) : "memory"); }

template <int KB>
constexpr int ce_smem_bytes() {   // fused dX (KB <= 2) keeps the table slices of three class tiles alive: 3 * KB ring slots
    return KB * TC_TILE_BYTES + (KB <= 2 ? 3 * KB : 5) * TC_TILE_BYTES + 2 * 32768 + 12 * 128 * 4 + 1024 + 256;
}

// Kernel modes.  CE_STORE_G and CE_KEEP_G are the row-stationary CE pass (token rows resident, classes streaming) with / without the
// bf16 G' tile going to HBM; CE_ACCUM_T is the same pipeline with the operands exchanged (a block of 128 CLASS rows resident, the token
// tiles streaming, the exponent shift a per-COLUMN array): its "dX" accumulator is dE[class, :] = sum_tokens G[token, class] x[token, :],
// so the tied-table gradient needs no [T, C] tensor in HBM either - it recomputes S once more instead (K = D is cheap).
// CE_ONE_SWEEP is CE_KEEP_G without the statistics sweep: the exponent shift of a row only has to be NEAR its maximum, not equal to it
// (the normalisation by the row sum happens later, in ce_finish_kernel), so the kernel takes
//     shift = max( max over the first class tile ,  logit of the row's target class )
// - the first term from ONE probe tile that both halves of a row block compute (bit-identical, no exchange between the halves), the
// second a 128-term dot product per row.  shift <= the row maximum, so nothing underflows that the exact shift would keep; the result
// is exact as long as no class beats BOTH the target and 128 probe classes by more than ~98 nats (fp32 / bf16 range of
// exp(s - shift) / count) - a token whose own loss exceeds 98 nats.  Beyond that the row's G' overflows to inf and loss and gradients
// come out inf / NaN: loud, never silently wrong; GRB_CE=exact selects the two-sweep mode, which has no such limit.
// One S recompute and one epilogue pass less: a third of the kernel's tensor and element-wise work.
enum { CE_STORE_G = 0, CE_KEEP_G = 1, CE_ACCUM_T = 2, CE_ONE_SWEEP = 3 };

struct CeShape {
    int T, C, ldl;       // rows of the resident operand, rows of the streaming operand, leading dimension of dlogits (multiple of 8, >= C)
    int num_m, num_n;
    int n_half;          // class tiles [0, n_half) belong to half 0, [n_half, num_n) to half 1
    int nsplit;          // CE_ACCUM_T: work items per resident block (each sweeps 1/nsplit of the streaming tiles); 2 otherwise
    float4* stats;       // [2 * num_m][128] {max, sum, target logit, -}  partials published after sweep 0
    unsigned* flags;     // [2 * num_m] zero before launch
};
GRB_DEVINL void red_add_v4_ce(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// tcgen05.mma with the A operand in tensor memory (lane = row, one 32-bit column = two consecutive bf16 along K)
GRB_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
GRB_DEVINL void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
GRB_DEVINL void st_release_u32(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
GRB_DEVINL unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int MODE>
GRB_DEVINL void ce_item(int w, const CeShape& sh, int& blk, int& nb, int& ne) {
    if (MODE == CE_ACCUM_T) {
        blk = w / sh.nsplit;
        const int p = w - blk * sh.nsplit;
        nb = (int)((long long)p * sh.num_n / sh.nsplit);
        ne = (int)((long long)(p + 1) * sh.num_n / sh.nsplit);
    } else {
        blk = w >> 1;
        nb = (w & 1) ? sh.n_half : 0;
        ne = (w & 1) ? sh.num_n : sh.n_half;
    }
}

template <int KB, int MODE>  // k-blocks of 64: D = 64 * KB
__global__ void __launch_bounds__(CE_THREADS, 1)
    tc_ce_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmG,
                 CeShape sh, const long long* __restrict__ targets, const float* __restrict__ inv_count,
                 float* __restrict__ row_sums /* [2][T]: sum_c G'[row, c] of each class half */,
                 float2* __restrict__ row_stats /* [T]: {row max, target logit} */,
                 const float* __restrict__ col_shift /* CE_ACCUM_T: [num_n * 128] exponent shift of each streaming row (+inf = skip) */,
                 const bf16* __restrict__ table /* CE_ONE_SWEEP: the class table [C, D] (rows gathered for the target logits) */,
                 float* __restrict__ dx_out /* [T, 64*KB] fp32 +=, written only when the second MMA is compiled in (KB <= 2) */) {
    constexpr bool FUSE_DX = KB <= 2;
    static_assert(MODE == CE_STORE_G || FUSE_DX, "without the G' store the second MMA is the only consumer of the tile");
    constexpr int SWEEPS = (MODE == CE_ACCUM_T || MODE == CE_ONE_SWEEP) ? 1 : 2;
    constexpr bool ONE = MODE == CE_ONE_SWEEP;
    constexpr int NS = KB <= 2 ? 3 * KB : 5;
    constexpr int D = 64 * KB;
    extern __shared__ unsigned char ce_smem_raw[];
    unsigned char* base = ce_smem_raw + ((1024u - (smem_u32(ce_smem_raw) & 1023u)) & 1023u)   /* offset from the __shared__ array: keeps the shared address space (LDS / STS) */;
    unsigned char* sX = base;                                   // KB x 16 KB, resident per row block
    unsigned char* sE = sX + KB * TC_TILE_BYTES;                // ring of table slices [128 classes][64 d]
    unsigned char* sOut0 = sE + NS * TC_TILE_BYTES;             // 2 x 32 KB staging (G tiles)
    float* s_part = reinterpret_cast<float*>(sOut0 + 2 * 32768);  // [4 column quarters][{max, sum, target logit}][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(s_part) + 12 * 128 * 4);
    uint64_t* efull = bars;                 // [NS]   TMA -> MMA
    uint64_t* eempty = bars + NS;           // [NS]   MMA -> TMA
    uint64_t* tfull = bars + 2 * NS;        // [2]    S accumulator ready
    uint64_t* tempty = bars + 2 * NS + 2;   // [2]    S accumulator drained
    uint64_t* xfull = bars + 2 * NS + 4;    // [1]
    uint64_t* xempty = bars + 2 * NS + 5;   // [1]
    uint64_t* gfull = bars + 2 * NS + 6;    // [2]    G staging tile written (epilogue -> MMA)
    uint64_t* gempty = bars + 2 * NS + 8;   // [2]    second MMA has read the staging tile
    uint64_t* dxfull = bars + 2 * NS + 10;  // [1]    dX accumulator complete
    uint64_t* dxempty = bars + 2 * NS + 11; // [1]    dX accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NS + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmE);
        if (MODE == CE_STORE_G) tma_prefetch_desc(&tmG);
        for (int s = 0; s < NS; ++s) { mbar_init(&efull[s], 1); mbar_init(&eempty[s], 1); }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1); mbar_init(&tempty[a], CE_EPI_WARPS);
            mbar_init(&gfull[a], MODE == CE_STORE_G ? 1 : CE_EPI_WARPS); mbar_init(&gempty[a], 1);
        }
        mbar_init(xfull, 1);
        mbar_init(xempty, 1);
        mbar_init(dxfull, 1);
        mbar_init(dxempty, CE_EPI_WARPS);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();  // prologue above overlaps the previous kernel's tail
    const uint32_t tmem_dx = tmem_base + 256;
    // CE_KEEP_G / CE_ACCUM_T: the G tile never touches shared memory - the epilogue packs it to bf16 pairs in tensor memory (2 x 64
    // columns) and the second MMA takes its A operand from there.  With both operands in shared memory that MMA alone would read
    // 96 KB per class tile (two N = 64 instructions re-reading A); the kernel is bound by shared-memory bandwidth, not by the tensor pipe.
    const uint32_t tmem_g = tmem_base + 384;
    const int num_items = (MODE == CE_ACCUM_T ? sh.nsplit : 2) * sh.num_m;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0, xphase = 0;
            for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
                int blk, nb, ne;
                ce_item<MODE>(w, sh, blk, nb, ne);
                const int ntile = ne - nb;
                mbar_wait(xempty, xphase ^ 1);
                mbar_expect_tx(xfull, KB * TC_TILE_BYTES);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(sX + kb * TC_TILE_BYTES, &tmX, kb * 64, blk * 128, xfull);
                xphase ^= 1;
                for (int tile = ONE ? -1 : 0; tile < SWEEPS * ntile; ++tile) {
                    const int n0 = tile < 0 ? 0 : (nb + tile % ntile) * 128;      // tile -1: the probe (class tile 0) of CE_ONE_SWEEP
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&eempty[stage], phase ^ 1);
                        mbar_expect_tx(&efull[stage], TC_TILE_BYTES);
                        tma_load_2d(sE + stage * TC_TILE_BYTES, &tmE, kb * 64, n0, &efull[stage]);
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc1 = umma_idesc(128, 128, 0, 0);   // S  = X   E_n^T  (both K-major)
            constexpr uint32_t idesc2 = umma_idesc(128, 64, 0, 1);    // dX += G   E_n    (A K-major staging, B MN-major table slice)
            int stage = 0; uint32_t phase = 0, xphase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            int gbuf = 0; uint32_t gphase = 0, dxphase = 0;
            // dX += G(n) E_n : reads staging buffer `buf` (two 64-class halves) and the KB ring slices that held tile n
            auto gemm2 = [&](int first_stage, int buf, bool first) {
                if (MODE != CE_STORE_G) {
                    // A = G tile in tensor memory; B = the KB adjacent ring slices of this class tile read MN-major as ONE [128 k][64 KB n]
                    // operand (the next 64-wide n block is one slice = TC_TILE_BYTES away; a tile never wraps the ring: NS = 3 KB)
                    constexpr uint32_t idesc2t = umma_idesc(128, 64 * KB, 0, 1);
                    const uint32_t b_addr = smem_u32(sE + first_stage * TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 8; ++k)   // 128 classes = 8 k-steps of 16 = 8 packed columns each
                        umma_bf16_ts(tmem_dx, tmem_g + buf * 64 + k * 8, umma_desc(b_addr + k * 2048, TC_TILE_BYTES, 1024), idesc2t,
                                     (first && k == 0) ? 0u : 1u);
#pragma unroll
                    for (int kb = 0; kb < KB; ++kb) umma_commit(&eempty[first_stage + kb]);
                    umma_commit(&gempty[buf]);
                    return;
                }
                const uint32_t a_addr = smem_u32(sOut0 + buf * 32768);
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) {
                    int st = first_stage + kb;
                    if (st >= NS) st -= NS;
                    const uint32_t b_addr = smem_u32(sE + st * TC_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {   // 128 classes = 8 k-steps of 16: staging half k/4, 32 B per step inside the row
                        const uint64_t ad = umma_desc(a_addr + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
                        const uint64_t bd = umma_desc(b_addr + k * 2048, 16, 1024);   // 16 class-rows of 128 B per step
                        umma_bf16(tmem_dx + kb * 64, ad, bd, idesc2, (first && k == 0) ? 0u : 1u);
                    }
                    umma_commit(&eempty[st]);      // this slice of the table is no longer needed
                }
                umma_commit(&gempty[buf]);
            };
            for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
                int blk_, nb_, ne_;
                ce_item<MODE>(w, sh, blk_, nb_, ne_);
                const int ntile = ne_ - nb_;
                mbar_wait(xfull, xphase);
                xphase ^= 1;
                tc_fence_after();
                int prev_stage = -1, prev_buf = 0;
                bool first_g = true;
                for (int tile = ONE ? -1 : 0; tile < SWEEPS * ntile; ++tile) {
                    const bool sweep1 = MODE == CE_ACCUM_T || (ONE ? tile >= 0 : tile >= ntile);
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * 128;
                    const int tile_stage = stage;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(&efull[stage], phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_u32(sX + kb * TC_TILE_BYTES);
                        const uint32_t b_addr = smem_u32(sE + stage * TC_TILE_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(d_tmem, umma_desc(a_addr + k * 32, 16, 1024), umma_desc(b_addr + k * 32, 16, 1024), idesc1,
                                      (kb > 0 || k > 0) ? 1u : 0u);
                        if (!(FUSE_DX && sweep1)) umma_commit(&eempty[stage]);   // sweep 1 keeps the slice for the dX MMA
                        if (++stage == NS) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    if (FUSE_DX && sweep1) {
                        if (prev_stage >= 0) {   // dX MMA of the PREVIOUS class tile, once its G tile has been staged
                            if (first_g) { mbar_wait(dxempty, dxphase ^ 1); tc_fence_after(); }
                            mbar_wait(&gfull[prev_buf], gphase);
                            tc_fence_after();
                            gemm2(prev_stage, prev_buf, first_g);
                            first_g = false;
                            if (prev_buf == 1) gphase ^= 1;
                        }
                        prev_stage = tile_stage;
                        prev_buf = gbuf;
                        gbuf ^= 1;
                    }
                }
                if (FUSE_DX && ntile > 0) {
                    if (first_g) { mbar_wait(dxempty, dxphase ^ 1); tc_fence_after(); }
                    mbar_wait(&gfull[prev_buf], gphase);
                    tc_fence_after();
                    gemm2(prev_stage, prev_buf, first_g);
                    if (prev_buf == 1) gphase ^= 1;
                    umma_commit(dxfull);
                    dxphase ^= 1;
                }
                umma_commit(xempty);  // every MMA that reads this row block has retired -> the producer may overwrite sX
            }
        }
    } else {
        // ===================================================================== epilogue (16 warps)
        const int sub = warp & 3, cq = (warp - 2) >> 2;   // TMEM sub-partition ; 32-column quarter of the class tile
        const int r = sub * 32 + lane;
        int acc = 0; uint32_t acc_phase = 0;
        int gbuf = 0; uint32_t gphase = 0, dxphase = 0, xphase_e = 0;
        const float ic = MODE == CE_ACCUM_T ? 0.f : *inv_count;
        for (int w = blockIdx.x; w < num_items; w += gridDim.x) {
            int blk, nb, ne;
            ce_item<MODE>(w, sh, blk, nb, ne);
            const int half = w & 1;
            const int row = blk * 128 + r;
            float gshift = 0.f, g_sum = 0.f;   // exponent shift of this row ; sum of this thread's G' entries (this half, this quarter)
            if (ONE) { mbar_wait(xfull, xphase_e); xphase_e ^= 1; }   // the epilogue reads the token tile itself (row norms)
            int t = 0;
            float tl = 0.f;
            if (MODE != CE_ACCUM_T) {
            t = row < sh.T ? (int)targets[row] : 0;
            const float icr = t != 0 ? ic : 0.f;   // ignore_index = 0 (and rows past the end)
            float m_run = -INFINITY;
            // ------------------------------------------------------------------ sweep 0: statistics (CE_ONE_SWEEP: the probe tile only)
            for (int n = ONE ? 0 : nb; n < (ONE ? 1 : ne); ++n) {
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * 128 + cq * 32), v);
                    // the accumulator is in registers: hand it back to the MMA warp before the element-wise work
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                    const int col0 = n * 128 + cq * 32;
                    if (col0 + 32 > sh.C) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i >= sh.C) v[i] = -INFINITY;
                    }
                    if (!ONE && (unsigned)(t - col0) < 32u) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i == t - col0) tl = v[i];
                    }
                    float cm = v[0];
#pragma unroll
                    for (int i = 1; i < 32; ++i) cm = fmaxf(cm, v[i]);
                    m_run = fmaxf(m_run, cm);
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            // combine the four column quarters of every row (this half of the classes) ...
            s_part[(cq * 3 + 0) * 128 + r] = m_run;
            s_part[(cq * 3 + 2) * 128 + r] = tl;
            if (ONE) {
                // this quarter's 32 terms of x_row . E[target]: x from the resident token tile (16-byte chunk j of a row sits at
                // j ^ (row & 7) inside its 128-byte swizzle row), E[target] from global memory (L2-resident table)
                float dot = 0.f;
                if (cq * 32 < D) {
                    const bf16* erow = table + (size_t)t * D + cq * 32;
                    const unsigned char* xrow = sX + (cq >> 1) * TC_TILE_BYTES + r * 128;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint4 xu = *reinterpret_cast<const uint4*>(xrow + ((((cq & 1) * 4 + c) ^ (r & 7)) << 4));
                        const uint4 eu = *reinterpret_cast<const uint4*>(erow + c * 8);
                        const float2 x0 = unpack_bf16(xu.x), x1 = unpack_bf16(xu.y), x2 = unpack_bf16(xu.z), x3 = unpack_bf16(xu.w);
                        const float2 e0 = unpack_bf16(eu.x), e1 = unpack_bf16(eu.y), e2 = unpack_bf16(eu.z), e3 = unpack_bf16(eu.w);
                        dot = fmaf(x0.x, e0.x, dot); dot = fmaf(x0.y, e0.y, dot); dot = fmaf(x1.x, e1.x, dot); dot = fmaf(x1.y, e1.y, dot);
                        dot = fmaf(x2.x, e2.x, dot); dot = fmaf(x2.y, e2.y, dot); dot = fmaf(x3.x, e3.x, dot); dot = fmaf(x3.y, e3.y, dot);
                    }
                }
                s_part[(cq * 3 + 1) * 128 + r] = dot;
            }
            ce_bar_sync();
            float mh = -INFINITY, tlh = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                mh = fmaxf(mh, s_part[(q * 3 + 0) * 128 + r]);
                tlh += s_part[(q * 3 + 2) * 128 + r];
            }
            if (ONE) {
                const float tdot = (s_part[1 * 128 + r] + s_part[4 * 128 + r]) + (s_part[7 * 128 + r] + s_part[10 * 128 + r]);
                const float mm = fmaxf(mh, tdot);
                gshift = mm * kLog2e - __log2f(icr);
                if (half == 0 && cq == 0 && row < sh.T) row_stats[row] = make_float2(mm, 0.f);
                tl = 0.f;
            } else {
            // ... publish them, and pick up the partner item's maximum: both halves must scale their exponentials alike, their dX'
            // partials and G' tiles are summed
            if (cq == 0) sh.stats[(size_t)w * 128 + r] = make_float4(mh, 0.f, tlh, 0.f);
            __threadfence();
            ce_bar_sync();
            if (warp == 2 && lane == 0) {
                st_release_u32(sh.flags + w, 1u);
                while (ld_acquire_u32(sh.flags + (w ^ 1)) == 0u) { }
            }
            ce_bar_sync();
            const float4 pp = __ldcg(sh.stats + (size_t)(w ^ 1) * 128 + r);
            const float mm = fmaxf(mh, pp.x);
            gshift = mm * kLog2e - __log2f(icr);               // icr == 0 (ignored row) -> +inf -> every entry of G' is 2^-inf = 0
            if (half == 0 && cq == 0 && row < sh.T) row_stats[row] = make_float2(mm, tlh + pp.z);
            }
            }
            // ------------------------------------------------------------------ sweep 1: gradient tiles
            if (MODE == CE_STORE_G && warp == 2 && lane == 0) tma_store_wait_read();  // both staging buffers are free of pending bulk stores
            if (MODE != CE_ACCUM_T) ce_bar_sync();             // (also: everybody has read s_part)
            for (int n = nb; n < ne; ++n) {
                if (MODE == CE_ACCUM_T && n + 1 < ne)   // next tile's shifts -> L1 while this tile is processed
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(col_shift + (n + 1) * 128 + cq * 32));
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                if (FUSE_DX) { mbar_wait(&gempty[gbuf], gphase ^ 1); tc_fence_after(); }   // the dX MMA that read this G buffer two tiles ago is done
                unsigned char* sOut = sOut0 + gbuf * 32768;
                {
                    float v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * 128 + cq * 32), v);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);   // accumulator drained (it lives in registers now)
                    const int col0 = n * 128 + cq * 32;
                    if (ONE && (unsigned)(t - col0) < 32u) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (i == t - col0) tl = v[i];
                    }
                    if (MODE == CE_ACCUM_T) {   // the shift belongs to the streaming (token) row = this tile's column
                        const float4* cs = reinterpret_cast<const float4*>(col_shift + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 s4 = __ldg(cs + j);
                            v[4 * j] = ex2_fast(fmaf(v[4 * j], kLog2e, -s4.x));
                            v[4 * j + 1] = ex2_fast(fmaf(v[4 * j + 1], kLog2e, -s4.y));
                            v[4 * j + 2] = ex2_fast(fmaf(v[4 * j + 2], kLog2e, -s4.z));
                            v[4 * j + 3] = ex2_fast(fmaf(v[4 * j + 3], kLog2e, -s4.w));
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = ex2_fast(fmaf(v[i], kLog2e, -gshift));   // exp(s - max) * (1/count), folded into the exponent
                    }
                    if (col0 + 32 > sh.C) {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (col0 + i >= sh.C) v[i] = 0.f;
                    }
                    if (MODE != CE_ACCUM_T) {   // row sum of G' (four partial chains keep the adds off the critical path)
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) { a0 += v[i]; a1 += v[i + 1]; a2 += v[i + 2]; a3 += v[i + 3]; }
                        g_sum += (a0 + a1) + (a2 + a3);
                    }
                    if (MODE != CE_STORE_G) {
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
                        tmem_st16(tmem_g + ((uint32_t)(sub * 32) << 16) + (uint32_t)(gbuf * 64 + cq * 16), pk);
                    } else {
                    unsigned char* dst = sOut + (cq >> 1) * 16384 + r * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 u;
                        u.x = pack_bf16(v[8 * j], v[8 * j + 1]); u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                        u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<uint4*>(dst + ((((cq & 1) * 4 + j) ^ (r & 7)) << 4)) = u;
                    }
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                if (MODE != CE_STORE_G) {
                    // each warp hands its 32 x 32 piece of the G tile to the MMA warp on its own: no CTA-wide barrier per class tile
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&gfull[gbuf]);
                } else {
                fence_proxy_async();   // generic-proxy smem writes -> visible to TMA and to tcgen05.mma (async proxy)
                ce_bar_sync();
                if (warp == 2 && lane == 0) {
                    if (FUSE_DX) mbar_arrive(&gfull[gbuf]);
#pragma unroll
                    for (int h = 0; h < 2; ++h)
                        if (n * 128 + h * 64 < sh.ldl) tma_store_2d(&tmG, sOut + h * 16384, n * 128 + h * 64, blk * 128);
                    tma_store_commit();
                    tma_store_wait_read1();
                }
                ce_bar_sync();   // the staging buffer written two tiles ago has been read by its bulk store
                }
                if (gbuf == 1) gphase ^= 1;
                gbuf ^= 1;
            }
            // row sums of this half -> global (four quarters combined through shared memory; slot 1 of s_part is free now)
            if (MODE != CE_ACCUM_T) {
                s_part[(cq * 3 + 1) * 128 + r] = g_sum;
                if (ONE) s_part[(cq * 3 + 2) * 128 + r] = tl;
                ce_bar_sync();
                if (cq == 0 && row < sh.T) {
                    row_sums[(size_t)half * sh.T + row] = (s_part[1 * 128 + r] + s_part[4 * 128 + r]) + (s_part[7 * 128 + r] + s_part[10 * 128 + r]);
                    if (ONE)   // target logit of this half's classes (0 when the target lives in the other half)
                        row_sums[(size_t)(2 + half) * sh.T + row] = (s_part[2 * 128 + r] + s_part[5 * 128 + r]) + (s_part[8 * 128 + r] + s_part[11 * 128 + r]);
                }
                if (ONE) ce_bar_sync();   // the next item's probe reuses slots 1 and 2 of s_part
            }
            if (FUSE_DX && ne > nb) {
                // ------------------------------------------------------------------ dX' partial of this half: TMEM -> += fp32 global
                mbar_wait(dxfull, dxphase);
                dxphase ^= 1;
                tc_fence_after();
                if (cq * 32 < D) {
                    float v[32];
                    tmem_ld32(tmem_dx + ((uint32_t)(sub * 32) << 16) + (uint32_t)(cq * 32), v);
                    if (row < sh.T) {
                        float* dst = dx_out + (size_t)row * D + cq * 32;
#pragma unroll
                        for (int j = 0; j < 8; ++j) red_add_v4_ce(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(dxempty);
            }
        }
        if (MODE == CE_STORE_G && warp == 2 && lane == 0) tma_store_wait_read();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// X [T, D] bf16, E [C, D] bf16 -> row_sums [2][T], row_stats [T] {max, target logit}, (D <= 128) dx' [T, D] fp32 = G' E with
// G' = exp(s - rowmax) * inv_count, and - store_g - G' [T, ldl] bf16 (columns >= C zeroed) for a dE GEMM.  Returns through *fused_dx
// whether dx' was produced.  ce_finish_kernel (rowwise.cuh) turns these into the loss, d loss / d x and the operands of the dE pass.
inline size_t ce_scratch_bytes(int T) { return (size_t)2 * ((T + 127) / 128) * (128 * sizeof(float4) + sizeof(unsigned)) + 256; }
template <int KB, int MODE>
inline cudaError_t ce_set_attr() {
    static bool attr_set_dev[64] = {false};
    int attr_dev = 0;
    cudaGetDevice(&attr_dev);
    bool& attr_set = attr_set_dev[attr_dev & 63];   // the attribute is per device
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_ce_kernel<KB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, ce_smem_bytes<KB>());
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    return cudaSuccess;
}
// scratch: ce_scratch_bytes(T) bytes of device memory (partials + flags); dx must hold [T, D] fp32 and is zero-filled here
// mode: CE_STORE_G, CE_KEEP_G or CE_ONE_SWEEP (the latter two need D <= 128; row_sums must be [4][T] for CE_ONE_SWEEP: sums of the two
// halves, then their target-logit partials)
template <int KB>
inline cudaError_t launch_tc_ce(const bf16* X, const bf16* E, bf16* G, int mode, int T, int C, int ldl, const long long* targets,
                                const float* inv_count, float* row_sums, float2* row_stats, float* dx, bool* fused_dx, void* scratch,
                                int num_sms, cudaStream_t st) {
    CUtensorMap tmX, tmE, tmG;
    const int D = 64 * KB;
    if (KB > 2) mode = CE_STORE_G;
    const bool store_g = mode == CE_STORE_G;
    bool ok = make_tmap_bf16(&tmX, X, T, D, D, 64, 128) && make_tmap_bf16(&tmE, E, C, D, D, 64, 128) && make_tmap(&tmG, G, false, T, ldl, ldl, 64, 128);
    if (!ok) return cudaErrorInvalidValue;
    CeShape sh;
    sh.T = T; sh.C = C; sh.ldl = ldl;
    sh.num_m = (T + 127) / 128;
    sh.num_n = (C + 127) / 128;
    sh.n_half = (sh.num_n + 1) / 2;
    sh.nsplit = 2;
    sh.stats = reinterpret_cast<float4*>(scratch);
    sh.flags = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(scratch) + (size_t)2 * sh.num_m * 128 * sizeof(float4));
    cudaError_t me = cudaSuccess;
    if (mode != CE_ONE_SWEEP) me = cudaMemsetAsync(sh.flags, 0, (size_t)2 * sh.num_m * sizeof(unsigned), st);
    if (me != cudaSuccess) return me;
    *fused_dx = KB <= 2;
    if (KB <= 2) {
        me = cudaMemsetAsync(dx, 0, (size_t)T * D * sizeof(float), st);
        if (me != cudaSuccess) return me;
    }
    // the two halves of a row block spin on each other: both must be resident at the same time -> even, persistent grid
    int grid = 2 * sh.num_m < num_sms ? 2 * sh.num_m : (num_sms & ~1);
    if (store_g) {
        me = ce_set_attr<KB, CE_STORE_G>();
        if (me != cudaSuccess) return me;
        launch_k(tc_ce_kernel<KB, CE_STORE_G>, grid, CE_THREADS, ce_smem_bytes<KB>(), st, tmX, tmE, tmG, sh, targets, inv_count, row_sums,
                 row_stats, (const float*)nullptr, E, dx);
    } else if (mode == CE_ONE_SWEEP) {
        constexpr int M = KB <= 2 ? CE_ONE_SWEEP : CE_STORE_G;
        me = ce_set_attr<KB, M>();
        if (me != cudaSuccess) return me;
        grid = 2 * sh.num_m < num_sms ? 2 * sh.num_m : num_sms;     // no exchange between the halves: any grid will do
        launch_k(tc_ce_kernel<KB, M>, grid, CE_THREADS, ce_smem_bytes<KB>(), st, tmX, tmE, tmG, sh, targets, inv_count, row_sums, row_stats,
                 (const float*)nullptr, E, dx);
    } else {
        constexpr int M = KB <= 2 ? CE_KEEP_G : CE_STORE_G;
        me = ce_set_attr<KB, M>();
        if (me != cudaSuccess) return me;
        launch_k(tc_ce_kernel<KB, M>, grid, CE_THREADS, ce_smem_bytes<KB>(), st, tmX, tmE, tmG, sh, targets, inv_count, row_sums, row_stats,
                 (const float*)nullptr, E, dx);
    }
    return cudaGetLastError();
}

// dE [C, D] fp32 += G^T X without G in HBM (CE_ACCUM_T): G[t, c] = 2^(s_tc log2e - col_shift[t]) with s = X E^T recomputed, col_shift
// [ceil(T/128)*128] = log2e * logsumexp_t - log2(inv_count_t) from ce_finish_kernel (+inf for ignored tokens and the padding).
// One work item = a block of 128 classes x 1/nsplit of the token tiles; its accumulator leaves through 16-byte vector reductions.
inline int ce_accum_nsplit(int num_m, int num_n, int num_sms) {
    int best = 1; long long best_cost = -1;
    for (int ns = 1; ns <= 16 && ns <= num_n; ++ns) {
        const long long rounds = ((long long)num_m * ns + num_sms - 1) / num_sms;
        const long long cost = rounds * ((num_n + ns - 1) / ns + 2);   // +2: resident load and accumulator flush of every item
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = ns; }
    }
    return best;
}
template <int KB>
inline cudaError_t launch_tc_ce_accum_t(const bf16* X, const bf16* E, const float* col_shift, float* dE, int T, int C, int num_sms,
                                        cudaStream_t st) {
    static_assert(KB <= 2, "second MMA needs D <= 128");
    CUtensorMap tmRes, tmStream;
    const int D = 64 * KB;
    bool ok = make_tmap_bf16(&tmRes, E, C, D, D, 64, 128) && make_tmap_bf16(&tmStream, X, T, D, D, 64, 128);
    if (!ok) return cudaErrorInvalidValue;
    CeShape sh;
    sh.T = C; sh.C = T; sh.ldl = 0;            // resident rows = classes, streaming rows = tokens
    sh.num_m = (C + 127) / 128;
    sh.num_n = (T + 127) / 128;
    sh.n_half = 0;
    sh.nsplit = ce_accum_nsplit(sh.num_m, sh.num_n, num_sms);
    sh.stats = nullptr; sh.flags = nullptr;
    cudaError_t me = ce_set_attr<KB, CE_ACCUM_T>();
    if (me != cudaSuccess) return me;
    const int items = sh.num_m * sh.nsplit;
    const int grid = items < num_sms ? items : num_sms;
    launch_k(tc_ce_kernel<KB, CE_ACCUM_T>, grid, CE_THREADS, ce_smem_bytes<KB>(), st, tmRes, tmStream, tmStream, sh, (const long long*)nullptr,
             (const float*)nullptr, (float*)nullptr, (float2*)nullptr, col_shift, (const bf16*)nullptr, dE);
    return cudaGetLastError();
}

}  // namespace grb
