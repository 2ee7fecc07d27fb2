// genrec_b200 - HSTU pointwise (SiLU) attention on the Blackwell tensor path: TMA tiles -> tcgen05.mma -> TMEM -> SiLU warps.
//
//   S[b,h,i,j] = Q_i . K_j + Wpos[pb0, h] + Wtime[tb(|ts_i - ts_j|), h]      valid = (j <= i) and not pad[b,j]
//   A = valid ? silu(S) : 0        O = A V                                   (reference: genrec/models/hstu.py:244-267)
//   backward: dA = dO V^T, dS = dA * silu'(S), dQ = dS K, dK = dS^T Q, dV = A^T dO, bias-table gradients = histogram of dS
//
// Nothing of size [L, L] exists outside the SM: the time bucket of every (i, j) cell is derived INSIDE the kernels from the
// two timestamps (integer thresholds that reproduce the reference's fp32 log/0.693 expression bit-exactly), once per
// (query tile, key tile) and shared by the heads the CTA works on.  Per sequence a tiny pre-pass (hstu_seq_prep_kernel)
// rebases the int64 timestamps to int32 offsets from the first valid event when the sequence spans < 2^31 ticks (always
// true for second-resolution data); sequences that do not fit take a 64-bit path inside the same kernels.
//
// Geometry.  A "box" is 64 channels of the [T, 4D] projection output = 64 / DH heads; every shared-memory tile is a
// 128-row x 64-channel bf16 box in the 128B-swizzled layout TMA writes and UMMA descriptors read (16 KB).  One CTA owns
//   forward : 128 query rows of one sequence x one box, looping over key tiles kt <= qt
//   backward: 128 key rows x one box, looping over query tiles qt >= kt (dK, dV accumulate in TMEM; dQ tiles are reduced
//             into an fp32 [T, D] buffer with 16-byte vector reductions)
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warps 2-9 = element-wise warps.
// An element-wise thread owns (row r = TMEM lane, 32-column chunk c) of a 128 x 64 half tile: S (and dA) arrive by
// tcgen05.ld.32x32b, the accumulator is released to the MMA warp as soon as the values are in registers, and P (dS) goes
// back to shared memory as the K-major / MN-major A operand of the second-stage MMAs.  The bucket bytes of a thread's own
// cells live in registers for the whole (qt, kt) iteration, so no [128,128] index tile is ever staged.
#pragma once
#include "attn_hstu.cuh"
#include "tc_gemm.cuh"

namespace grb {

constexpr int ATC_THREADS = 320;
constexpr int ATC_EW_WARPS = 8;
constexpr int ATC_BOX_BYTES = 16384;      // [128 rows][64 bf16], SWIZZLE_128B
constexpr int ATC_TILE2_BYTES = 32768;    // [128 rows][128 bf16] as two boxes (keys 0..63 | 64..127)
constexpr int ATC_TBL_LD = 68;

struct HstuTcArgs {
    const long long* ts;        // [B, L] int64 or null (no temporal term)
    const int* rel32;           // [B, L] timestamp offsets (valid when wide[b] == 0)
    const uint8_t* wide;        // [B] 1: the sequence spans >= 2^31 ticks -> 64-bit bucket path
    const uint8_t* pad;         // [B, L] 1 = padded key
    const long long* thr64;     // [65] time-bucket thresholds (thr64[64] = INT64_MAX)
    const float* wpos;          // [H]: the one live row of the position table (uniform position buckets)
    const float* wtime;         // [ntime, H] or null
    int ntime;
    int B, L, H, D;
    // forward
    bf16* o; int ldo;
    // backward
    const bf16* zk; const bf16* zv; int ldz;     // pre-activations (silu' factor of the K / V gradients)
    bf16* dk; bf16* dv; int lddz;                 // gradients w.r.t. the K / V pre-activations
    float* dq_acc;                                // [T, D] fp32, zero on entry: dQ (w.r.t. the activation) accumulates here
    float* dwpos;                                 // [H]   (+=)
    float* dwtime;                                // [ntime, H] (+=) or null
};

// ------------------------------------------------------------------------------------------------ small PTX helpers
GRB_DEVINL void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
GRB_DEVINL void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
GRB_DEVINL void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
template <int ID, int N>
GRB_DEVINL void nbar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory"); }
GRB_DEVINL void nbar_sync_dyn(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
GRB_DEVINL void nbar_arrive_dyn(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// mbarrier wait with a suspend-time hint: a waiting warp sleeps in hardware instead of re-polling (the polls of 8 element-wise
// warps, of the TMA lane and of the MMA lane would otherwise take issue slots from the working warps and from the CTA that shares the
// SM).  The hint only bounds the sleep: an arrive that completes the phase wakes the waiter.
GRB_DEVINL void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP_S:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra WAIT_DONE_S;\n"
        "bra WAIT_LOOP_S;\n"
        "WAIT_DONE_S:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(20000u)
        : "memory");
}
// The TMA lane and the MMA lane are single threads that mostly wait; polling with an explicit nanosleep keeps them from taking the
// issue slots of the element-wise warps that share their schedulers (profile: the two lanes were ~20 % of all issued
// instructions).  The accumulators are released early (see below), so the MMA lane has ~1 us of slack before its next issue.
template <int NS>
GRB_DEVINL void mbar_wait_poll(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(NS);
    }
}
GRB_DEVINL uint64_t atc_kmaj(uint32_t addr) { return umma_desc(addr, 16, 1024); }          // K-major box (rows of 128 B)
GRB_DEVINL uint64_t atc_mnmaj(uint32_t addr) { return umma_desc(addr, ATC_BOX_BYTES, 1024); }  // MN-major, next 64-wide block one box away

// ------------------------------------------------------------------------------------------------ per-sequence pre-pass
// rel[b, i] = int32(ts[b, i] - min_i ts[b, :]) ; wide[b] = 1 when max - min >= 2^31 (then rel is unused).  Padded positions take
// part: a padded QUERY row still attends to earlier real keys with its own timestamp (hstu.py:400 does not look at the pad mask), so
// with the usual "pads carry 0" convention the span is the largest real timestamp - fine for second-resolution data until 2038.
__global__ void __launch_bounds__(256) hstu_seq_prep_kernel(const long long* __restrict__ ts, const uint8_t* __restrict__ pad, int L,
                                                           int* __restrict__ rel, uint8_t* __restrict__ wide) {
    pdl_wait();
    (void)pad;
    __shared__ long long s_min[8], s_max[8];
    const int b = blockIdx.x;
    const long long* t = ts + (size_t)b * L;
    long long mn = 0x7fffffffffffffffLL, mx = -0x7fffffffffffffffLL - 1;
    for (int i = threadIdx.x; i < L; i += 256) { const long long v = t[i]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, mn, o), c = __shfl_xor_sync(0xffffffffu, mx, o);
        mn = a < mn ? a : mn; mx = c > mx ? c : mx;
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    mn = s_min[0]; mx = s_max[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) { mn = s_min[w] < mn ? s_min[w] : mn; mx = s_max[w] > mx ? s_max[w] : mx; }
    // span computed in unsigned arithmetic: mx - mn of two int64 may not fit int64
    const unsigned long long span = (unsigned long long)mx - (unsigned long long)mn;
    const bool w64 = span >= (1ull << 31);
    if (threadIdx.x == 0) wide[b] = w64 ? 1 : 0;
    for (int i = threadIdx.x; i < L; i += 256) rel[(size_t)b * L + i] = w64 ? 0 : (int)(t[i] - mn);
}

// ------------------------------------------------------------------------------------------------ bucket bytes of 32 cells
// cells (row i, keys j0 .. j0+31) ; bk[k >> 2] byte (k & 3) = time bucket of cell k, or 64 when the cell is masked.
// s_rel / s_ts / s_pad point at the first of the 32 keys in shared memory.
struct AtcBk8 {
    uint32_t w[8];
};
// 32-bit path (the sequence spans < 2^31 ticks): inlined, this is the hot one
template <bool MASKED>
GRB_DEVINL void atc_buckets_narrow(uint32_t (&bk)[8], int i, int j0, int ri, const int* s_rel, const uint8_t* s_pad, const uint32_t* s_thr32,
                                   int ntime, int L, bool row_ok) {
    const int ntm1 = ntime - 1;
    const bool clamp = ntm1 < 31;     // 32-bit differences never reach bucket 32 (warp-uniform)
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        uint32_t word = 0;
        const int4 r4 = *reinterpret_cast<const int4*>(s_rel + 4 * w);      // four keys per broadcast LDS.128
        const int rj[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int k = w * 4 + q;
            int b = 0;
            if (ntime > 0) {
                // e = floor(log2 |d|) (-1 for d == 0: thr32[0] == 0 then gives bucket 0) ; bucket = e + (|d| >= thr[e + 1])
                const int d = ri - rj[q];
                const unsigned a = (unsigned)(d < 0 ? -d : d);
                const int e = 31 - __clz(a);
                b = e + (a >= s_thr32[e + 1] ? 1 : 0);
                if (clamp) b = b < ntm1 ? b : ntm1;
            }
            if (MASKED) {
                const int j = j0 + k;
                const bool valid = row_ok && j <= i && j < L && s_pad[k] == 0;
                b = valid ? b : 64;
            }
            word |= (uint32_t)b << (8 * q);
        }
        bk[w] = word;
    }
}
// 64-bit path: a sequence that spans >= 2^31 ticks (millisecond clocks, synthetic extremes).  Out of line on purpose: it is
// rare and its 64-bit arithmetic would otherwise be inlined at every call site (instruction-cache footprint of the hot loop).
// g_ts points at the timestamps of the 32 keys in GLOBAL memory.
__device__ __noinline__ AtcBk8 atc_buckets_wide(int i, int j0, long long ti, const long long* g_ts, const uint8_t* s_pad,
                                                const long long* thr64, int ntime, int L, int row_ok) {
    AtcBk8 out;
    for (int w = 0; w < 8; ++w) {
        uint32_t word = 0;
        for (int q = 0; q < 4; ++q) {
            const int k = w * 4 + q;
            const int j = j0 + k;
            int b = 0;
            if (ntime > 0 && j < L) b = time_bucket_dev(ti - g_ts[k], thr64, ntime);
            const bool valid = row_ok && j <= i && j < L && s_pad[k] == 0;
            b = valid ? b : 64;
            word |= (uint32_t)b << (8 * q);
        }
        out.w[w] = word;
    }
    return out;
}
// bucket bytes of one (row, 32-key chunk): classification -> all masked / narrow fast / narrow masked / wide
GRB_DEVINL bool atc_chunk_buckets(uint32_t (&bk)[8], int q_first, int lane, int i, int j0, int ri, long long ti, bool wide, const int* s_rel,
                                  const long long* g_ts, const uint8_t* s_pad, const uint32_t* s_thr32, const long long* thr64, int ntime, int L,
                                  bool row_ok);
GRB_DEVINL void atc_buckets_all_masked(uint32_t (&bk)[8]) {
#pragma unroll
    for (int w = 0; w < 8; ++w) bk[w] = 0x40404040u;
}

// per-head bias table: tbl[v] = Wpos[h] + Wtime[v, h] (v < ntime) ; tbl[64] = mask
GRB_DEVINL void atc_build_table(float* tbl, const HstuTcArgs& a, int h, int t, int nthreads) {
    for (int v = t; v < 64; v += nthreads) {
        float x = a.wpos[h];
        if (a.wtime != nullptr && v < a.ntime) x += a.wtime[(size_t)v * a.H + h];
        tbl[v] = x;
    }
    if (t == 0) tbl[64] = ATT_MASK_BIAS;
}

// store 32 bf16 of row r, key columns half*64 + c*32 .. +31 into a [128 x 128] two-box tile
GRB_DEVINL void atc_store_chunk(unsigned char* tile, int r, int half, int c, const float (&v)[32]) {
    unsigned char* dst = tile + half * ATC_BOX_BYTES + r * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        u.x = pack_bf16(v[8 * j], v[8 * j + 1]); u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
        u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
        *reinterpret_cast<uint4*>(dst + (((c * 4 + j) ^ (r & 7)) << 4)) = u;
    }
}
GRB_DEVINL void atc_store_chunk_zero(unsigned char* tile, int r, int half, int c) {
    unsigned char* dst = tile + half * ATC_BOX_BYTES + r * 128;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + (((c * 4 + j) ^ (r & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
}

// classification of a (32-row block, 32-key chunk) pair, warp-uniform
struct AtcChunkClass {
    bool all_masked, needs_mask;
};
GRB_DEVINL AtcChunkClass atc_classify(int q_first, int k_first, int L, const uint8_t* s_pad_chunk, int lane) {
    AtcChunkClass cc;
    cc.all_masked = q_first >= L || k_first >= L || k_first > q_first + 31;
    const bool any_pad = __any_sync(0xffffffffu, s_pad_chunk[lane] != 0);
    cc.needs_mask = any_pad || (k_first + 31 > q_first) || (q_first + 31 >= L) || (k_first + 31 >= L);
    return cc;
}

GRB_DEVINL bool atc_chunk_buckets(uint32_t (&bk)[8], int q_first, int lane, int i, int j0, int ri, long long ti, bool wide, const int* s_rel,
                                  const long long* g_ts, const uint8_t* s_pad, const uint32_t* s_thr32, const long long* thr64, int ntime, int L,
                                  bool row_ok) {
    const AtcChunkClass cl = atc_classify(q_first, j0, L, s_pad, lane);
    if (cl.all_masked) {
        atc_buckets_all_masked(bk);
    } else if (wide) {
        const AtcBk8 r = atc_buckets_wide(i, j0, ti, g_ts, s_pad, thr64, ntime, L, row_ok ? 1 : 0);
#pragma unroll
        for (int w = 0; w < 8; ++w) bk[w] = r.w[w];
    } else if (cl.needs_mask) {
        atc_buckets_narrow<true>(bk, i, j0, ri, s_rel, s_pad, s_thr32, ntime, L, row_ok);
    } else {
        atc_buckets_narrow<false>(bk, i, j0, ri, s_rel, s_pad, s_thr32, ntime, L, row_ok);
    }
    return cl.all_masked;
}

// ================================================================================================ forward
template <int DH>
struct AtcFwdSmem {
    static constexpr int HB = 64 / DH;
    static constexpr int kQ = 0;
    static constexpr int kK = ATC_BOX_BYTES;                       // one stage: two CTAs share an SM and cover each other's loads
    static constexpr int kV = kK + ATC_BOX_BYTES;
    static constexpr int kP = kV + ATC_BOX_BYTES;
    static constexpr int kSmall = kP + ATC_TILE2_BYTES;
    // small area: rel[2][128] int | pad[2][128] | thr32[36] | tbl[HB][ATC_TBL_LD] | barriers
    static constexpr int kRel = kSmall;
    static constexpr int kPad = kRel + 2 * 128 * 4;
    static constexpr int kThr32 = kPad + 2 * 128;
    static constexpr int kTbl = kThr32 + 36 * 4;
    static constexpr int kBars = kTbl + HB * ATC_TBL_LD * 4;
    static constexpr int kBytes = kBars + 16 * 8 + 16;
};

template <int DH>
__global__ void __launch_bounds__(ATC_THREADS, 2) hstu_attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmP, HstuTcArgs a, int nqt) {
    using SM = AtcFwdSmem<DH>;
    constexpr int HB = SM::HB, KS = DH / 16;
    extern __shared__ unsigned char atc_smem_raw[];
    unsigned char* base = atc_smem_raw + ((1024u - (smem_u32(atc_smem_raw) & 1023u)) & 1023u)   /* offset from the __shared__ array: keeps the shared address space (LDS / STS) */;
    unsigned char* sQ = base + SM::kQ;
    unsigned char* sK = base + SM::kK;
    unsigned char* sV = base + SM::kV;
    unsigned char* sP = base + SM::kP;
    int* s_rel = reinterpret_cast<int*>(base + SM::kRel);
    uint8_t* s_pad = base + SM::kPad;
    uint32_t* s_thr32 = reinterpret_cast<uint32_t*>(base + SM::kThr32);
    float* s_tbl = reinterpret_cast<float*>(base + SM::kTbl);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + SM::kBars);
    uint64_t* q_full = bars;            // TMA -> MMA
    uint64_t* kv_full = bars + 1;       // TMA -> MMA (K and V boxes of one key tile)
    uint64_t* kv_empty = bars + 3;      // MMA -> TMA
    uint64_t* s_full = bars + 5;        // MMA -> EW : a 128 x 64 half tile of S is in TMEM
    uint64_t* s_free = bars + 6;        // EW -> MMA : ... and has been read into registers
    uint64_t* p_full = bars + 7;        // EW -> MMA : the P tile of one head is in shared memory
    uint64_t* p_empty = bars + 8;       // MMA -> EW : ... and has been consumed by P V
    uint64_t* o_full = bars + 9;        // MMA -> EW : the O accumulators are final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item: heaviest (largest qt) first
    const int nbox = a.D / 64;
    int item = blockIdx.x;
    const int box = item % nbox; item /= nbox;
    const int b = item % a.B; item /= a.B;
    const int qt = nqt - 1 - item;
    const int L = a.L;
    const int q0 = qt * 128;
    const long long tok0 = (long long)b * L;
    const int nkt = qt + 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmP);
        mbar_init(q_full, 1);
        mbar_init(kv_full, 1);
        mbar_init(kv_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(s_free, ATC_EW_WARPS);
        mbar_init(p_full, ATC_EW_WARPS);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_S = tmem, tm_O = tmem + 64;
    pdl_wait();

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            const int row_q = (int)tok0 + q0;
            mbar_expect_tx(q_full, ATC_BOX_BYTES);
            tma_load_2d(sQ, &tmP, 2 * a.D + box * 64, row_q, q_full);
            for (int kt = 0; kt < nkt; ++kt) {
                mbar_wait_poll<400>(kv_empty, (kt & 1) ^ 1);
                mbar_expect_tx(kv_full, 2 * ATC_BOX_BYTES);
                tma_load_2d(sK, &tmP, 3 * a.D + box * 64, (int)tok0 + kt * 128, kv_full);
                tma_load_2d(sV, &tmP, a.D + box * 64, (int)tok0 + kt * 128, kv_full);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc(128, 64, 0, 0);    // S half = Q_h K_h[half]^T
            constexpr uint32_t idesc_pv = umma_idesc(128, DH, 0, 1);   // O_h += P V_h   (V: MN-major, DH columns of the box)
            const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
            mbar_wait_poll<100>(q_full, 0);
            int u = 0, n = 0;                 // units issued, heads whose P V has been issued
            int pend_hb = -1, pend_kt = 0;    // head whose P V is still to be issued
            auto issue_pv = [&](int hb, int kt) {
                mbar_wait_poll<100>(p_full, n & 1);
                tc_fence_after();
                const uint32_t v_addr = smem_u32(sV);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tm_O + hb * DH, atc_kmaj(p_addr + (ks >> 2) * ATC_BOX_BYTES + (ks & 3) * 32),
                              atc_mnmaj(v_addr + hb * DH * 2 + ks * 2048), idesc_pv, (kt > 0 || ks > 0) ? 1u : 0u);
                umma_commit(p_empty);
                if (hb == HB - 1) umma_commit(kv_empty);   // last reader of this key tile's K / V boxes
                ++n;
            };
            const uint32_t k_addr = smem_u32(sK);
            for (int kt = 0; kt < nkt; ++kt) {
                // the single K/V stage is refilled only after the last P V of the previous key tile: flush it before waiting
                if (pend_hb >= 0) { issue_pv(pend_hb, pend_kt); pend_hb = -1; }
                mbar_wait_poll<100>(kv_full, kt & 1);
                tc_fence_after();
                for (int hb = 0; hb < HB; ++hb) {
                    for (int half = 0; half < 2; ++half) {
                        if (u > 0) mbar_wait_poll<100>(s_free, (u - 1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int s = 0; s < KS; ++s)
                            umma_bf16(tm_S, atc_kmaj(q_addr + (hb * KS + s) * 32), atc_kmaj(k_addr + half * 8192 + (hb * KS + s) * 32), idesc_s,
                                      s > 0 ? 1u : 0u);
                        umma_commit(s_full);
                        ++u;
                        if (half == 0 && pend_hb >= 0) { issue_pv(pend_hb, pend_kt); pend_hb = -1; }
                    }
                    pend_hb = hb; pend_kt = kt;
                }
            }
            issue_pv(pend_hb, pend_kt);
            umma_commit(o_full);
        }
    } else {
        // ===================================================================== element-wise warps
        const int sub = warp & 3, c = (warp - 2) >> 2;
        const int r = sub * 32 + lane;
        const int ew_t = threadIdx.x - 64;
        const int i = q0 + r;
        const bool row_ok = i < L;
        const bool has_time = a.ts != nullptr && a.ntime > 0;
        const int ntime = has_time ? a.ntime : 0;
        const bool wide = has_time && a.wide[b] != 0;
        // thresholds + bias tables
        for (int k = ew_t; k < 36; k += 256) s_thr32[k] = k <= 31 ? (uint32_t)a.thr64[k] : 0xffffffffu;
        for (int hb = 0; hb < HB; ++hb) atc_build_table(s_tbl + hb * ATC_TBL_LD, a, box * HB + hb, ew_t, 256);
        const long long* g_ts = a.ts != nullptr ? a.ts + tok0 : nullptr;      // wide path reads key timestamps from global memory
        const int ri = (has_time && !wide && row_ok) ? a.rel32[tok0 + i] : 0;
        const long long ti = (wide && row_ok) ? a.ts[tok0 + i] : 0;
        // key metadata of tile kt -> buffer kt & 1 (thread k < 128 stages key k)
        int k_rel = 0; uint8_t k_pad = 1;
        auto fetch_keys = [&](int kt) {
            const int j = kt * 128 + ew_t;
            k_rel = 0; k_pad = 1;
            if (ew_t < 128 && j < L) {
                k_pad = a.pad[tok0 + j];
                if (has_time && !wide) k_rel = a.rel32[tok0 + j];
            }
        };
        auto stage_keys = [&](int kt) {
            if (ew_t < 128) {
                s_rel[(kt & 1) * 128 + ew_t] = k_rel;
                s_pad[(kt & 1) * 128 + ew_t] = k_pad;
            }
        };
        fetch_keys(0);
        stage_keys(0);
        nbar_sync<1, 256>();
        int u = 0, n = 0;
        for (int kt = 0; kt < nkt; ++kt) {
            if (kt + 1 < nkt) fetch_keys(kt + 1);   // in flight during this tile's work
            const int kb = (kt & 1) * 128;
            const int k0 = kt * 128;
            // bucket bytes of this thread's two chunks (half 0: keys c*32.., half 1: keys 64 + c*32..)
            uint32_t bk[2][8];
            bool masked_all[2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int cc0 = half * 64 + c * 32;
                masked_all[half] = atc_chunk_buckets(bk[half], q0 + sub * 32, lane, i, k0 + cc0, ri, ti, wide, s_rel + kb + cc0,
                                                     g_ts + k0 + cc0, s_pad + kb + cc0, s_thr32, a.thr64, ntime, L, row_ok);
            }
            for (int hb = 0; hb < HB; ++hb) {
                const float* tbl = s_tbl + hb * ATC_TBL_LD;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    mbar_wait_sleep(s_full, u & 1);
                    tc_fence_after();
                    float s[32];
                    if (!masked_all[half]) {
                        tmem_ld32_nowait(tm_S + ((uint32_t)(sub * 32) << 16) + c * 32, s);
                        tmem_wait_ld();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free);
                    ++u;
                    if (half == 0 && n > 0) mbar_wait_sleep(p_empty, (n - 1) & 1);   // P V of the previous head has read the P tile
                    if (!masked_all[half]) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const uint32_t bb = (bk[half][k >> 2] >> (8 * (k & 3))) & 0xffu;
                            s[k] = siluf(s[k] + tbl[bb]);
                        }
                        atc_store_chunk(sP, r, half, c, s);
                    } else {
                        atc_store_chunk_zero(sP, r, half, c);
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
                ++n;
            }
            if (kt + 1 < nkt) {
                stage_keys(kt + 1);
                nbar_sync<1, 256>();
            }
        }
        // epilogue: O box [128 x 64] fp32 in TMEM -> bf16 rows
        mbar_wait_sleep(o_full, 0);
        tc_fence_after();
        float o[32];
        tmem_ld32_nowait(tm_O + ((uint32_t)(sub * 32) << 16) + c * 32, o);
        tmem_wait_ld();
        if (row_ok) store_bf16x32(a.o + (size_t)(tok0 + i) * a.ldo + box * 64 + c * 32, o, 32);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 128);
    }
}

// ================================================================================================ backward
template <int DH>
struct AtcBwdSmem {
    static constexpr int HB = 64 / DH;
    static constexpr int kK = 0;
    static constexpr int kV = ATC_BOX_BYTES;
    static constexpr int kQ = 2 * ATC_BOX_BYTES;                   // [2 stages]
    static constexpr int kDO = kQ + 2 * ATC_BOX_BYTES;             // [2 stages]
    static constexpr int kP = kDO + 2 * ATC_BOX_BYTES;
    static constexpr int kDS = kP + ATC_TILE2_BYTES;
    static constexpr int kHist = kDS + ATC_TILE2_BYTES;            // [HB][32 bins][256 element-wise threads] fp32: private bins
    static constexpr int kRel = kHist + HB * 32 * 256 * 4;
    static constexpr int kPad = kRel + 128 * 4;
    static constexpr int kThr32 = kPad + 128;
    static constexpr int kTbl = kThr32 + 36 * 4;
    static constexpr int kBars = kTbl + HB * ATC_TBL_LD * 4;
    static constexpr int kBytes = kBars + 20 * 8 + 16;
};

template <int DH>
__global__ void __launch_bounds__(ATC_THREADS, 1)
    hstu_attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmDO, HstuTcArgs a, int nqt) {
    using SM = AtcBwdSmem<DH>;
    constexpr int HB = SM::HB, KS = DH / 16;
    extern __shared__ unsigned char atc_smem_raw[];
    unsigned char* base = atc_smem_raw + ((1024u - (smem_u32(atc_smem_raw) & 1023u)) & 1023u)   /* offset from the __shared__ array: keeps the shared address space (LDS / STS) */;
    unsigned char* sK = base + SM::kK;
    unsigned char* sV = base + SM::kV;
    unsigned char* sQ = base + SM::kQ;
    unsigned char* sDO = base + SM::kDO;
    unsigned char* sP = base + SM::kP;
    unsigned char* sDS = base + SM::kDS;
    float* s_hist = reinterpret_cast<float*>(base + SM::kHist);
    int* s_rel = reinterpret_cast<int*>(base + SM::kRel);
    uint8_t* s_pad = base + SM::kPad;
    uint32_t* s_thr32 = reinterpret_cast<uint32_t*>(base + SM::kThr32);
    float* s_tbl = reinterpret_cast<float*>(base + SM::kTbl);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + SM::kBars);
    // the 64-bit (wide) path reads the key timestamps and the thresholds straight from global memory: rare, and shared memory
    // is needed for the private histogram bins
    uint64_t* kv_full = bars;            // TMA -> MMA
    uint64_t* qdo_full = bars + 1;       // [2]
    uint64_t* qdo_empty = bars + 3;      // [2]
    uint64_t* sda_full = bars + 5;       // MMA -> EW : S and dA half tiles in TMEM
    uint64_t* sda_free = bars + 6;       // EW -> MMA : ... read into registers
    uint64_t* pds_full = bars + 7;       // EW -> MMA : P and dS tiles of a head in shared memory
    uint64_t* pds_empty = bars + 8;      // MMA -> EW : second-stage MMAs have consumed them
    uint64_t* dq_full = bars + 9;        // [2] MMA -> EW : a dQ tile is in TMEM
    uint64_t* dq_free = bars + 11;       // [2] EW -> MMA : ... and has been reduced to global memory
    uint64_t* dkdv_full = bars + 13;     // MMA -> EW : dK / dV accumulators final
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nbox = a.D / 64;
    int item = blockIdx.x;                 // heaviest (kt = 0) first
    const int box = item % nbox; item /= nbox;
    const int b = item % a.B; item /= a.B;
    const int kt = item;
    const int L = a.L;
    const int k0 = kt * 128;
    const long long tok0 = (long long)b * L;
    const long long* s_ts = a.ts != nullptr ? a.ts + tok0 + k0 : nullptr;     // key timestamps of this tile (wide path only, guarded by j < L)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmP);
        tma_prefetch_desc(&tmDO);
        mbar_init(kv_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&qdo_full[s], 1); mbar_init(&qdo_empty[s], 1);
            mbar_init(&dq_full[s], 1); mbar_init(&dq_free[s], ATC_EW_WARPS);
        }
        mbar_init(sda_full, 1);
        mbar_init(sda_free, ATC_EW_WARPS);
        mbar_init(pds_full, ATC_EW_WARPS);
        mbar_init(pds_empty, 1);
        mbar_init(dkdv_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tm_S = tmem, tm_dA = tmem + 64, tm_dQ = tmem + 128 /* 2 x DH */, tm_dK = tmem + 256, tm_dV = tmem + 320;
    pdl_wait();

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            mbar_expect_tx(kv_full, 2 * ATC_BOX_BYTES);
            tma_load_2d(sK, &tmP, 3 * a.D + box * 64, (int)tok0 + k0, kv_full);
            tma_load_2d(sV, &tmP, a.D + box * 64, (int)tok0 + k0, kv_full);
            for (int qt = kt, it = 0; qt < nqt; ++qt, ++it) {
                const int st = it & 1;
                mbar_wait_poll<400>(&qdo_empty[st], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&qdo_full[st], 2 * ATC_BOX_BYTES);
                tma_load_2d(sQ + st * ATC_BOX_BYTES, &tmP, 2 * a.D + box * 64, (int)tok0 + qt * 128, &qdo_full[st]);
                tma_load_2d(sDO + st * ATC_BOX_BYTES, &tmDO, box * 64, (int)tok0 + qt * 128, &qdo_full[st]);
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc_s = umma_idesc(128, 64, 0, 0);     // S / dA half tiles
            constexpr uint32_t idesc_t = umma_idesc(128, DH, 1, 1);     // dV += P^T dO , dK += dS^T Q   (A, B MN-major)
            constexpr uint32_t idesc_q = umma_idesc(128, DH, 0, 1);     // dQ = dS K                     (A K-major, B MN-major)
            const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), p_addr = smem_u32(sP), ds_addr = smem_u32(sDS);
            mbar_wait_poll<100>(kv_full, 0);
            int u = 0, n = 0;
            int pend_hb = -1, pend_it = 0;
            auto second_stage = [&](int hb, int it) {
                const int st = it & 1;
                const uint32_t q_addr = smem_u32(sQ + st * ATC_BOX_BYTES), do_addr = smem_u32(sDO + st * ATC_BOX_BYTES);
                const int buf = n & 1;
                mbar_wait_poll<100>(pds_full, n & 1);
                mbar_wait_poll<100>(&dq_free[buf], ((n >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t acc = it > 0 ? 1u : 0u;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tm_dV + hb * DH, atc_mnmaj(p_addr + ks * 2048), atc_mnmaj(do_addr + hb * DH * 2 + ks * 2048), idesc_t,
                              (acc || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tm_dK + hb * DH, atc_mnmaj(ds_addr + ks * 2048), atc_mnmaj(q_addr + hb * DH * 2 + ks * 2048), idesc_t,
                              (acc || ks > 0) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma_bf16(tm_dQ + buf * DH, atc_kmaj(ds_addr + (ks >> 2) * ATC_BOX_BYTES + (ks & 3) * 32),
                              atc_mnmaj(k_addr + hb * DH * 2 + ks * 2048), idesc_q, ks > 0 ? 1u : 0u);
                umma_commit(pds_empty);
                umma_commit(&dq_full[buf]);
                if (hb == HB - 1) umma_commit(&qdo_empty[st]);
                ++n;
            };
            for (int qt = kt, it = 0; qt < nqt; ++qt, ++it) {
                const int st = it & 1;
                mbar_wait_poll<100>(&qdo_full[st], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t q_addr = smem_u32(sQ + st * ATC_BOX_BYTES), do_addr = smem_u32(sDO + st * ATC_BOX_BYTES);
                for (int hb = 0; hb < HB; ++hb) {
                    for (int half = 0; half < 2; ++half) {
                        if (u > 0) mbar_wait_poll<100>(sda_free, (u - 1) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int s = 0; s < KS; ++s)
                            umma_bf16(tm_S, atc_kmaj(q_addr + (hb * KS + s) * 32), atc_kmaj(k_addr + half * 8192 + (hb * KS + s) * 32), idesc_s,
                                      s > 0 ? 1u : 0u);
#pragma unroll
                        for (int s = 0; s < KS; ++s)
                            umma_bf16(tm_dA, atc_kmaj(do_addr + (hb * KS + s) * 32), atc_kmaj(v_addr + half * 8192 + (hb * KS + s) * 32), idesc_s,
                                      s > 0 ? 1u : 0u);
                        umma_commit(sda_full);
                        ++u;
                        if (half == 0 && pend_hb >= 0) { second_stage(pend_hb, pend_it); pend_hb = -1; }
                    }
                    pend_hb = hb; pend_it = it;
                }
            }
            second_stage(pend_hb, pend_it);
            umma_commit(dkdv_full);
        }
    } else {
        // ===================================================================== element-wise warps
        const int sub = warp & 3, c = (warp - 2) >> 2;
        const int r = sub * 32 + lane;
        const int ew_t = threadIdx.x - 64;
        const bool has_time = a.ts != nullptr && a.ntime > 0;
        const int ntime = has_time ? a.ntime : 0;
        const bool wide = has_time && a.wide[b] != 0;
        for (int k = ew_t; k < 36; k += 256) s_thr32[k] = k <= 31 ? (uint32_t)a.thr64[k] : 0xffffffffu;
        for (int hb = 0; hb < HB; ++hb) atc_build_table(s_tbl + hb * ATC_TBL_LD, a, box * HB + hb, ew_t, 256);
        for (int k = ew_t; k < HB * 32 * 256; k += 256) s_hist[k] = 0.f;
        if (ew_t < 128) {
            const int j = k0 + ew_t;
            int kr = 0; uint8_t kp = 1;
            if (j < L) {
                kp = a.pad[tok0 + j];
                if (has_time && !wide) kr = a.rel32[tok0 + j];
            }
            s_rel[ew_t] = kr; s_pad[ew_t] = kp;
        }
        nbar_sync<1, 256>();
        int u = 0, n = 0;
        float pos_acc[HB];          // wide path only: sum of dS per head
#pragma unroll
        for (int hb = 0; hb < HB; ++hb) pos_acc[hb] = 0.f;
        // dQ tile of head number m (0-based over the CTA's lifetime) -> fp32 reduction into dq_acc
        auto drain_dq = [&](int m) {
            const int buf = m & 1;
            const int m_it = m / HB, m_hb = m % HB;
            const int qi = (kt + m_it) * 128 + r;
            mbar_wait_sleep(&dq_full[buf], (m >> 1) & 1);
            tc_fence_after();
            constexpr int W = DH / 2;     // columns per thread
            float v[W];
            if constexpr (W == 16) tmem_ld16_nowait(tm_dQ + buf * DH + ((uint32_t)(sub * 32) << 16) + c * W, reinterpret_cast<float(&)[16]>(v));
            else tmem_ld32_nowait(tm_dQ + buf * DH + ((uint32_t)(sub * 32) << 16) + c * W, reinterpret_cast<float(&)[32]>(v));
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dq_free[buf]);
            if (qi < L) {
                float* dst = a.dq_acc + (size_t)(tok0 + qi) * a.D + box * 64 + m_hb * DH + c * W;
#pragma unroll
                for (int x = 0; x < W; x += 4) red_add_v4(dst + x, v[x], v[x + 1], v[x + 2], v[x + 3]);
            }
        };
        int ri_next = 0; long long ti_next = 0;
        auto fetch_row = [&](int qt) {
            const int i = qt * 128 + r;
            ri_next = 0; ti_next = 0;
            if (i < L) {
                if (has_time && !wide) ri_next = a.rel32[tok0 + i];
                if (wide) ti_next = a.ts[tok0 + i];
            }
        };
        fetch_row(kt);
        for (int qt = kt, it = 0; qt < nqt; ++qt, ++it) {
            const int q0 = qt * 128;
            const int i = q0 + r;
            const bool row_ok = i < L;
            const int ri = ri_next; const long long ti = ti_next;
            if (qt + 1 < nqt) fetch_row(qt + 1);
            uint32_t bk[2][8];
            bool masked_all[2];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int cc0 = half * 64 + c * 32;
                masked_all[half] = atc_chunk_buckets(bk[half], q0 + sub * 32, lane, i, k0 + cc0, ri, ti, wide, s_rel + cc0, s_ts + cc0, s_pad + cc0,
                                                     s_thr32, a.thr64, ntime, L, row_ok);
            }
#pragma unroll
            for (int hb = 0; hb < HB; ++hb) {
                const float* tbl = s_tbl + hb * ATC_TBL_LD;
                float* hist = s_hist + hb * 32 * 256 + ew_t;     // this thread's private column of bins
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    mbar_wait_sleep(sda_full, u & 1);
                    tc_fence_after();
                    float s[32], da[32];
                    if (!masked_all[half]) {
                        tmem_ld32_nowait(tm_S + ((uint32_t)(sub * 32) << 16) + c * 32, s);
                        tmem_ld32_nowait(tm_dA + ((uint32_t)(sub * 32) << 16) + c * 32, da);
                        tmem_wait_ld();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(sda_free);
                    ++u;
                    if (half == 0 && n > 0) mbar_wait_sleep(pds_empty, (n - 1) & 1);
                    if (!masked_all[half]) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const uint32_t bb = (bk[half][k >> 2] >> (8 * (k & 3))) & 0xffu;
                            const float x = s[k] + tbl[bb];
                            const float sg = sigmoidf_fast(x);
                            s[k] = x * sg;                                      // A
                            da[k] = da[k] * (sg * (1.f + x * (1.f - sg)));      // dS (exactly 0 on masked cells)
                        }
                        atc_store_chunk(sP, r, half, c, s);
                        atc_store_chunk(sDS, r, half, c, da);
                    } else {
                        atc_store_chunk_zero(sP, r, half, c);
                        atc_store_chunk_zero(sDS, r, half, c);
                    }
                    // bias-table gradients.  Every element-wise thread owns a private column of 32 bins per head in shared memory
                    // (plain read-modify-write, bank-conflict free, no atomics).  Consecutive keys of a row mostly fall into the
                    // same log bucket, so runs are summed in a register and only run ends touch shared memory - a handful of
                    // updates per 32 cells instead of a chain of 32 dependent read-modify-writes.
                    if (!masked_all[half]) {
                        if (!wide) {
                            unsigned prev = bk[half][0] & 31u;
                            float acc = 0.f;
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                const unsigned bb = (bk[half][k >> 2] >> (8 * (k & 3))) & 31u;   // masked cells (64) add an exact 0 to bin 0
                                if (bb != prev) {
                                    hist[prev * 256] += acc;
                                    prev = bb;
                                    acc = 0.f;
                                }
                                acc += da[k];
                            }
                            hist[prev * 256] += acc;
                        } else {
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                const uint32_t bb = (bk[half][k >> 2] >> (8 * (k & 3))) & 0xffu;
                                if (da[k] != 0.f) {
                                    pos_acc[hb] += da[k];
                                    if (a.dwtime != nullptr && ntime > 0 && bb < 64u) atomicAdd(a.dwtime + (size_t)bb * a.H + box * HB + hb, da[k]);
                                }
                            }
                        }
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(pds_full);
                if (n > 0) drain_dq(n - 1);
                ++n;
            }
        }
        drain_dq(n - 1);
        // epilogue: dK / dV boxes [128 keys x 64] -> x silu'(z) -> bf16
        mbar_wait_sleep(dkdv_full, 0);
        tc_fence_after();
        {
            const int j = k0 + r;
            float gk[32], gv[32];
            tmem_ld32_nowait(tm_dK + ((uint32_t)(sub * 32) << 16) + c * 32, gk);
            tmem_ld32_nowait(tm_dV + ((uint32_t)(sub * 32) << 16) + c * 32, gv);
            tmem_wait_ld();
            if (j < L) {
                const size_t zo = (size_t)(tok0 + j) * a.ldz + box * 64 + c * 32;
                const size_t go = (size_t)(tok0 + j) * a.lddz + box * 64 + c * 32;
                float z[32];
                if (a.zk != nullptr) {
                    load_bf16x32(a.zk + zo, z, 32);
#pragma unroll
                    for (int x = 0; x < 32; ++x) gk[x] *= dsiluf(z[x]);
                }
                if (a.zv != nullptr) {
                    load_bf16x32(a.zv + zo, z, 32);
#pragma unroll
                    for (int x = 0; x < 32; ++x) gv[x] *= dsiluf(z[x]);
                }
                store_bf16x32(a.dk + go, gk, 32);
                store_bf16x32(a.dv + go, gv, 32);
            }
        }
        // bias-table gradients -> global
        nbar_sync<1, 256>();
        if (!wide) {
            const int ew_warp = warp - 2;
            for (int e = ew_warp; e < HB * 32; e += ATC_EW_WARPS) {
                const int hb = e >> 5, v = e & 31;
                const float* row = s_hist + (size_t)e * 256;
                float sum = row[lane] + row[lane + 32] + row[lane + 64] + row[lane + 96] + row[lane + 128] + row[lane + 160] +
                            row[lane + 192] + row[lane + 224];
                sum = warp_sum(sum);
                if (lane == 0 && sum != 0.f) {
                    const int h = box * HB + hb;
                    atomicAdd(a.dwpos + h, sum);
                    if (a.dwtime != nullptr && ntime > 0 && v < ntime) atomicAdd(a.dwtime + (size_t)v * a.H + h, sum);
                }
            }
        } else {
#pragma unroll
            for (int hb = 0; hb < HB; ++hb) {
                const float sum = warp_sum(pos_acc[hb]);
                if (lane == 0 && sum != 0.f) atomicAdd(a.dwpos + box * HB + hb, sum);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// dzp[Q columns] = bf16(dq_acc * silu'(zq))   (dq_acc holds dQ w.r.t. the Q activation)
__global__ void __launch_bounds__(256) hstu_dq_finish_kernel(const float* __restrict__ dq_acc, const bf16* __restrict__ zq, int ldz,
                                                            bf16* __restrict__ dq, int lddq, size_t T, int D) {
    pdl_wait();
    const size_t n8 = T * (size_t)(D / 8);
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n8; e += (size_t)gridDim.x * blockDim.x) {
        const size_t t = e / (D / 8);
        const int col = (int)(e % (D / 8)) * 8;
        const float4 a0 = *reinterpret_cast<const float4*>(dq_acc + t * D + col);
        const float4 a1 = *reinterpret_cast<const float4*>(dq_acc + t * D + col + 4);
        float g[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (zq != nullptr) {
            const uint4 zz = *reinterpret_cast<const uint4*>(zq + t * ldz + col);
            const float2 z0 = unpack_bf16(zz.x), z1 = unpack_bf16(zz.y), z2 = unpack_bf16(zz.z), z3 = unpack_bf16(zz.w);
            g[0] *= dsiluf(z0.x); g[1] *= dsiluf(z0.y); g[2] *= dsiluf(z1.x); g[3] *= dsiluf(z1.y);
            g[4] *= dsiluf(z2.x); g[5] *= dsiluf(z2.y); g[6] *= dsiluf(z3.x); g[7] *= dsiluf(z3.y);
        }
        uint4 o;
        o.x = pack_bf16(g[0], g[1]); o.y = pack_bf16(g[2], g[3]); o.z = pack_bf16(g[4], g[5]); o.w = pack_bf16(g[6], g[7]);
        *reinterpret_cast<uint4*>(dq + t * lddq + col) = o;
    }
}

// Test hook: the bucket / mask byte of every (i, j) cell, produced by the routine the attention kernels use (same
// classification, same fast / masked / wide variants).  grid (ceil(L / 32), ceil(L / 128), B), block 128.
__global__ void __launch_bounds__(128) hstu_bucket_bytes_debug_kernel(HstuTcArgs a, uint8_t* __restrict__ out) {
    pdl_wait();
    __shared__ __align__(16) int s_rel[32];
    __shared__ uint8_t s_pad[32];
    __shared__ uint32_t s_thr32[36];
    const int b = blockIdx.z, q0 = blockIdx.y * 128, j0 = blockIdx.x * 32;
    const int L = a.L, t = threadIdx.x, sub = t >> 5, lane = t & 31;
    const long long tok0 = (long long)b * L;
    const bool has_time = a.ts != nullptr && a.ntime > 0;
    const int ntime = has_time ? a.ntime : 0;
    const bool wide = has_time && a.wide[b] != 0;
    for (int k = t; k < 36; k += 128) s_thr32[k] = k <= 31 ? (uint32_t)a.thr64[k] : 0xffffffffu;
    if (t < 32) {
        const int j = j0 + t;
        int kr = 0; uint8_t kp = 1;
        if (j < L) {
            kp = a.pad[tok0 + j];
            if (has_time && !wide) kr = a.rel32[tok0 + j];
        }
        s_rel[t] = kr; s_pad[t] = kp;
    }
    __syncthreads();
    const int i = q0 + t;
    const bool row_ok = i < L;
    const int ri = (has_time && !wide && row_ok) ? a.rel32[tok0 + i] : 0;
    const long long ti = (wide && row_ok) ? a.ts[tok0 + i] : 0;
    uint32_t bk[8];
    atc_chunk_buckets(bk, q0 + sub * 32, lane, i, j0, ri, ti, wide, s_rel, a.ts != nullptr ? a.ts + tok0 + j0 : nullptr, s_pad, s_thr32, a.thr64,
                      ntime, L, row_ok);
    if (row_ok)
        for (int k = 0; k < 32; ++k)
            if (j0 + k < L) out[(size_t)(tok0 + i) * L + j0 + k] = (uint8_t)((bk[k >> 2] >> (8 * (k & 3))) & 0xffu);
}

}  // namespace grb
