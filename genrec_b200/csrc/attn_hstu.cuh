// genrec_b200 - HSTU pointwise (SiLU) attention, forward and backward, first-generation mma.sync path.
//
//   S[b,h,i,j] = Q_i . K_j + Wpos[pb(i-j), h] + Wtime[tb(|ts_i - ts_j|), h]
//   valid      = (j <= i) and not pad[b,j]
//   A          = valid ? silu(S) : 0          O = A V
// (reference: genrec/models/hstu.py:244-267).  No [L,L] tensor ever reaches HBM.
//
// Layout: Q/K/V/dO/O are row-major [T = B*L, ld] bf16 with head h at columns h*DH .. h*DH+DH-1.
// One CTA = 4 warps = 64 query rows (fwd, dQ) or 64 key rows (dK/dV); KV (resp. Q) tiles of 64 stream through smem.
#pragma once
#include "common.cuh"

namespace grb {

constexpr int ATT_BLK = 64;       // rows per CTA and per streamed tile
constexpr int ATT_THREADS = 128;  // 4 warps x 16 rows
constexpr int ATT_MAX_BUCKETS = 64;

struct HstuBiasArgs {
    const float* wpos;           // [npos, H]
    const uint8_t* pos_bucket;   // [L]   bucket of delta = i - j >= 0  (host: reference bucketing of clamp(j - i, 0) -> all 0)
    const float* wtime;          // [ntime, H] or null
    const long long* time_thr;   // [65] thr[k] = min |dt| whose reference bucket >= k ; thr[64] = INT64_MAX
    const long long* ts;         // [B, L] or null
    int npos, ntime;
};

struct HstuAttnArgs {
    const bf16* q; const bf16* k; const bf16* v;   // forward operands (activations after SiLU)
    int ldq, ldk, ldv;
    const uint8_t* pad;  // [B, L] 1 = padded key
    int B, L, H;
    HstuBiasArgs bias;
    // forward
    bf16* o; int ldo;
    // backward
    const bf16* d_o; int lddo;
    const bf16* zq; const bf16* zk; const bf16* zv; int ldz;    // pre-activations (nullable -> no silu' factor)
    bf16* dq; bf16* dk; bf16* dv; int lddq;                       // gradients w.r.t. pre-activations (or activations if z null)
    float* dwpos;   // [npos, H]  accumulated (atomicAdd)
    float* dwtime;  // [ntime, H] accumulated
};

GRB_DEVINL int time_bucket_dev(long long dt, const long long* s_thr, int ntime) {
    long long d = dt < 0 ? -dt : dt;
    d = d < 1 ? 1 : d;
    int e = 63 - __clzll(d);
    int b = e + (d >= s_thr[e + 1] ? 1 : 0);
    return min(b, ntime - 1);
}

template <int DH>
struct AttSmem {
    static constexpr int LD = DH + 8;
    bf16 tile[4][ATT_BLK * LD];  // roles differ per kernel
    long long ts_tile[ATT_BLK];
    long long thr[ATT_MAX_BUCKETS + 1];
    float wpos[ATT_MAX_BUCKETS];
    float wtime[ATT_MAX_BUCKETS];
    uint8_t pad_tile[ATT_BLK];
};

// cooperative 64 x DH tile load (rows row0.. of one batch element, zero-filled beyond L)
template <int DH>
GRB_DEVINL void att_load_tile(bf16* s, const bf16* g, int ld, long long tok0, int row0, int L, int col0, int tid) {
    constexpr int LD = DH + 8;
    constexpr int CH = DH / 8;  // 16-byte chunks per row
    for (int c = tid; c < ATT_BLK * CH; c += ATT_THREADS) {
        int r = c / CH, kc = (c % CH) * 8;
        bool ok = (row0 + r) < L;
        const bf16* src = ok ? g + (size_t)(tok0 + row0 + r) * ld + col0 + kc : g;
        cp_async16(s + r * LD + kc, src, ok ? 16 : 0);
    }
}

template <int DH>
GRB_DEVINL void att_load_bias_tables(AttSmem<DH>& sm, const HstuBiasArgs& b, int h, int H, int tid) {
    for (int i = tid; i < ATT_MAX_BUCKETS; i += ATT_THREADS) {
        sm.wpos[i] = i < b.npos ? b.wpos[i * H + h] : 0.f;
        sm.wtime[i] = (b.wtime && i < b.ntime) ? b.wtime[i * H + h] : 0.f;
    }
    for (int i = tid; i <= ATT_MAX_BUCKETS; i += ATT_THREADS) sm.thr[i] = b.time_thr ? b.time_thr[i] : 0x7fffffffffffffffLL;
}

// A-operand fragments of a 16 x DH slab (rows wrow..wrow+15 of an smem tile)
template <int DH>
GRB_DEVINL void att_load_afrag(uint32_t (&f)[DH / 16][4], const bf16* tile, int wrow, int lane) {
    constexpr int LD = DH + 8;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) ldsm_x4(f[ks], tile + (wrow + lane_a_row(lane)) * LD + ks * 16 + lane_a_col(lane));
}

// acc[8][4] (16 rows x 64 cols) = Afrag(16 x DH) * Tile^T   where Tile is [64][DH] (k = DH contiguous)
template <int DH>
GRB_DEVINL void att_mma_nt(float (&acc)[8][4], const uint32_t (&af)[DH / 16][4], const bf16* tile, int lane) {
    constexpr int LD = DH + 8;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) {
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
            uint32_t r[4];
            ldsm_x4(r, tile + (j2 * 16 + lane_b_row(lane)) * LD + ks * 16 + lane_b_col(lane));
            mma_bf16(acc[2 * j2], af[ks], r[0], r[1]);
            mma_bf16(acc[2 * j2 + 1], af[ks], r[2], r[3]);
        }
    }
}

// out[DH/8][4] (16 rows x DH cols) += P(16 x 64, as 4 k16 A-fragments) * Tile   where Tile is [64][DH] (n = DH contiguous)
template <int DH>
GRB_DEVINL void att_mma_nn(float (&out)[DH / 8][4], const uint32_t (&pf)[4][4], const bf16* tile, int lane) {
    constexpr int LD = DH + 8;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int n2 = 0; n2 < DH / 16; ++n2) {
            uint32_t r[4];
            ldsm_x4_t(r, tile + (kk * 16 + lane_a_row(lane)) * LD + n2 * 16 + lane_a_col(lane));
            mma_bf16(out[2 * n2], pf[kk], r[0], r[1]);
            mma_bf16(out[2 * n2 + 1], pf[kk], r[2], r[3]);
        }
    }
}

GRB_DEVINL void att_pack_p(uint32_t (&pf)[4][4], const float (&s)[8][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        pf[kk][0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pf[kk][1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pf[kk][2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pf[kk][3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    }
}

// ============================================================================================ forward
// tile roles: 0 = Q, 1 = K, 2 = V
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) hstu_attn_fwd_kernel(HstuAttnArgs a, const uint8_t* __restrict__ posb_g) {
    extern __shared__ __align__(16) unsigned char att_smem_raw[];
    AttSmem<DH>& sm = *reinterpret_cast<AttSmem<DH>*>(att_smem_raw);
    uint8_t* s_posb = att_smem_raw + sizeof(AttSmem<DH>);  // [L]
    constexpr int LD = DH + 8;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int L = a.L, q0 = qt * ATT_BLK;
    const long long tok0 = (long long)b * L;
    const bool has_time = a.bias.wtime != nullptr && a.bias.ts != nullptr;

    att_load_bias_tables<DH>(sm, a.bias, h, a.H, tid);
    for (int i = tid; i < L; i += ATT_THREADS) s_posb[i] = posb_g[i];
    att_load_tile<DH>(sm.tile[0], a.q, a.ldq, tok0, q0, L, h * DH, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    uint32_t qf[DH / 16][4];
    att_load_afrag<DH>(qf, sm.tile[0], warp * 16, lane);
    const int i0 = q0 + warp * 16 + g, i1 = i0 + 8;
    long long ts_i0 = 0, ts_i1 = 0;
    if (has_time) {
        if (i0 < L) ts_i0 = a.bias.ts[tok0 + i0];
        if (i1 < L) ts_i1 = a.bias.ts[tok0 + i1];
    }

    float o[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
#pragma unroll
        for (int r = 0; r < 4; ++r) o[n][r] = 0.f;

    for (int kt = 0; kt <= qt; ++kt) {
        const int k0 = kt * ATT_BLK;
        __syncthreads();  // previous tile fully consumed
        att_load_tile<DH>(sm.tile[1], a.k, a.ldk, tok0, k0, L, h * DH, tid);
        att_load_tile<DH>(sm.tile[2], a.v, a.ldv, tok0, k0, L, h * DH, tid);
        cp_async_commit();
        if (tid < ATT_BLK) {
            int j = k0 + tid;
            sm.pad_tile[tid] = (j < L) ? a.pad[tok0 + j] : 1;
            sm.ts_tile[tid] = (has_time && j < L) ? a.bias.ts[tok0 + j] : 0;
        }
        cp_async_wait<0>();
        __syncthreads();

        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) s[n][r] = 0.f;
        att_mma_nt<DH>(s, qf, sm.tile[1], lane);

#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int jl = n * 8 + 2 * t + (r & 1);
                const int j = k0 + jl;
                const int i = (r < 2) ? i0 : i1;
                const bool valid = (j <= i) && (i < L) && (sm.pad_tile[jl] == 0);
                float val = 0.f;
                if (valid) {
                    float bias = sm.wpos[s_posb[i - j]];
                    if (has_time) bias += sm.wtime[time_bucket_dev(((r < 2) ? ts_i0 : ts_i1) - sm.ts_tile[jl], sm.thr, a.bias.ntime)];
                    val = siluf(s[n][r] + bias);
                }
                s[n][r] = val;
            }
        }
        uint32_t pf[4][4];
        att_pack_p(pf, s);
        att_mma_nn<DH>(o, pf, sm.tile[2], lane);
    }

#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
        int col = h * DH + n * 8 + 2 * t;
        if (i0 < L) *reinterpret_cast<uint32_t*>(a.o + (size_t)(tok0 + i0) * a.ldo + col) = pack_bf16(o[n][0], o[n][1]);
        if (i1 < L) *reinterpret_cast<uint32_t*>(a.o + (size_t)(tok0 + i1) * a.ldo + col) = pack_bf16(o[n][2], o[n][3]);
    }
}

// ============================================================================================ backward: dQ
// tile roles: 0 = Q, 1 = K, 2 = V, 3 = dO
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) hstu_attn_bwd_dq_kernel(HstuAttnArgs a, const uint8_t* __restrict__ posb_g) {
    extern __shared__ __align__(16) unsigned char att_smem_raw[];
    AttSmem<DH>& sm = *reinterpret_cast<AttSmem<DH>*>(att_smem_raw);
    uint8_t* s_posb = att_smem_raw + sizeof(AttSmem<DH>);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int L = a.L, q0 = qt * ATT_BLK;
    const long long tok0 = (long long)b * L;
    const bool has_time = a.bias.wtime != nullptr && a.bias.ts != nullptr;

    att_load_bias_tables<DH>(sm, a.bias, h, a.H, tid);
    for (int i = tid; i < L; i += ATT_THREADS) s_posb[i] = posb_g[i];
    att_load_tile<DH>(sm.tile[0], a.q, a.ldq, tok0, q0, L, h * DH, tid);
    att_load_tile<DH>(sm.tile[3], a.d_o, a.lddo, tok0, q0, L, h * DH, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    uint32_t qf[DH / 16][4], dof[DH / 16][4];
    att_load_afrag<DH>(qf, sm.tile[0], warp * 16, lane);
    att_load_afrag<DH>(dof, sm.tile[3], warp * 16, lane);
    const int i0 = q0 + warp * 16 + g, i1 = i0 + 8;
    long long ts_i0 = 0, ts_i1 = 0;
    if (has_time) {
        if (i0 < L) ts_i0 = a.bias.ts[tok0 + i0];
        if (i1 < L) ts_i1 = a.bias.ts[tok0 + i1];
    }
    float dq[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
#pragma unroll
        for (int r = 0; r < 4; ++r) dq[n][r] = 0.f;

    for (int kt = 0; kt <= qt; ++kt) {
        const int k0 = kt * ATT_BLK;
        __syncthreads();
        att_load_tile<DH>(sm.tile[1], a.k, a.ldk, tok0, k0, L, h * DH, tid);
        att_load_tile<DH>(sm.tile[2], a.v, a.ldv, tok0, k0, L, h * DH, tid);
        cp_async_commit();
        if (tid < ATT_BLK) {
            int j = k0 + tid;
            sm.pad_tile[tid] = (j < L) ? a.pad[tok0 + j] : 1;
            sm.ts_tile[tid] = (has_time && j < L) ? a.bias.ts[tok0 + j] : 0;
        }
        cp_async_wait<0>();
        __syncthreads();

        float s[8][4], da[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) s[n][r] = 0.f, da[n][r] = 0.f;
        att_mma_nt<DH>(s, qf, sm.tile[1], lane);    // S  = Q K^T
        att_mma_nt<DH>(da, dof, sm.tile[2], lane);  // dA = dO V^T
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int jl = n * 8 + 2 * t + (r & 1);
                const int j = k0 + jl;
                const int i = (r < 2) ? i0 : i1;
                const bool valid = (j <= i) && (i < L) && (sm.pad_tile[jl] == 0);
                float val = 0.f;
                if (valid) {
                    float bias = sm.wpos[s_posb[i - j]];
                    if (has_time) bias += sm.wtime[time_bucket_dev(((r < 2) ? ts_i0 : ts_i1) - sm.ts_tile[jl], sm.thr, a.bias.ntime)];
                    val = da[n][r] * dsiluf(s[n][r] + bias);
                }
                s[n][r] = val;  // dS
            }
        }
        uint32_t pf[4][4];
        att_pack_p(pf, s);
        att_mma_nn<DH>(dq, pf, sm.tile[1], lane);  // dQ += dS K
    }

#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
        int col = h * DH + n * 8 + 2 * t;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            int i = half ? i1 : i0;
            if (i >= L) continue;
            float v0 = dq[n][2 * half], v1 = dq[n][2 * half + 1];
            if (a.zq) {
                float2 z = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.zq + (size_t)(tok0 + i) * a.ldz + col));
                v0 *= dsiluf(z.x);
                v1 *= dsiluf(z.y);
            }
            *reinterpret_cast<uint32_t*>(a.dq + (size_t)(tok0 + i) * a.lddq + col) = pack_bf16(v0, v1);
        }
    }
}

// ============================================================================================ backward: dK, dV, bias tables
// CTA owns 64 keys; tile roles: 0 = K (own), 1 = V (own), 2 = Q (streamed), 3 = dO (streamed)
// dynamic smem tail: s_posb[L] (padded to 16) then lane-private histograms  hist_t[4][ntime][32], hist_p[4][npos][32]
template <int DH>
__global__ void __launch_bounds__(ATT_THREADS) hstu_attn_bwd_dkdv_kernel(HstuAttnArgs a, const uint8_t* __restrict__ posb_g,
                                                                         int posb_bytes) {
    extern __shared__ __align__(16) unsigned char att_smem_raw[];
    AttSmem<DH>& sm = *reinterpret_cast<AttSmem<DH>*>(att_smem_raw);
    uint8_t* s_posb = att_smem_raw + sizeof(AttSmem<DH>);
    float* hist_t = reinterpret_cast<float*>(s_posb + posb_bytes);
    const int ntime = a.bias.ntime, npos = a.bias.npos;
    float* hist_p = hist_t + 4 * ntime * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int L = a.L, k0 = kt * ATT_BLK;
    const long long tok0 = (long long)b * L;
    const bool has_time = a.bias.wtime != nullptr && a.bias.ts != nullptr;
    const int nqt = (L + ATT_BLK - 1) / ATT_BLK;

    att_load_bias_tables<DH>(sm, a.bias, h, a.H, tid);
    for (int i = tid; i < L; i += ATT_THREADS) s_posb[i] = posb_g[i];
    for (int i = tid; i < 4 * (ntime + npos) * 32; i += ATT_THREADS) hist_t[i] = 0.f;
    att_load_tile<DH>(sm.tile[0], a.k, a.ldk, tok0, k0, L, h * DH, tid);
    att_load_tile<DH>(sm.tile[1], a.v, a.ldv, tok0, k0, L, h * DH, tid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    uint32_t kf[DH / 16][4], vf[DH / 16][4];
    att_load_afrag<DH>(kf, sm.tile[0], warp * 16, lane);
    att_load_afrag<DH>(vf, sm.tile[1], warp * 16, lane);
    const int j0 = k0 + warp * 16 + g, j1 = j0 + 8;
    long long ts_j0 = 0, ts_j1 = 0;
    bool ok_j0 = false, ok_j1 = false;  // key exists and is not padding
    if (j0 < L) { ok_j0 = a.pad[tok0 + j0] == 0; if (has_time) ts_j0 = a.bias.ts[tok0 + j0]; }
    if (j1 < L) { ok_j1 = a.pad[tok0 + j1] == 0; if (has_time) ts_j1 = a.bias.ts[tok0 + j1]; }

    float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
    for (int n = 0; n < DH / 8; ++n)
#pragma unroll
        for (int r = 0; r < 4; ++r) dk[n][r] = 0.f, dv[n][r] = 0.f;
    float* my_ht = hist_t + (warp * ntime) * 32 + lane;
    float* my_hp = hist_p + (warp * npos) * 32 + lane;

    for (int qt = kt; qt < nqt; ++qt) {
        const int q0 = qt * ATT_BLK;
        __syncthreads();
        att_load_tile<DH>(sm.tile[2], a.q, a.ldq, tok0, q0, L, h * DH, tid);
        att_load_tile<DH>(sm.tile[3], a.d_o, a.lddo, tok0, q0, L, h * DH, tid);
        cp_async_commit();
        if (tid < ATT_BLK) {
            int i = q0 + tid;
            sm.ts_tile[tid] = (has_time && i < L) ? a.bias.ts[tok0 + i] : 0;
        }
        cp_async_wait<0>();
        __syncthreads();

        float st[8][4], dat[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n)
#pragma unroll
            for (int r = 0; r < 4; ++r) st[n][r] = 0.f, dat[n][r] = 0.f;
        att_mma_nt<DH>(st, kf, sm.tile[2], lane);   // S^T  = K Q^T   (rows = keys, cols = queries)
        att_mma_nt<DH>(dat, vf, sm.tile[3], lane);  // dA^T = V dO^T
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int il = n * 8 + 2 * t + (r & 1);
                const int i = q0 + il;
                const int j = (r < 2) ? j0 : j1;
                const bool valid = (j <= i) && (i < L) && ((r < 2) ? ok_j0 : ok_j1);
                float av = 0.f, dsv = 0.f;
                if (valid) {
                    int pbk = s_posb[i - j];
                    float x = st[n][r] + sm.wpos[pbk];
                    int tbk = 0;
                    if (has_time) {
                        tbk = time_bucket_dev(sm.ts_tile[il] - ((r < 2) ? ts_j0 : ts_j1), sm.thr, ntime);
                        x += sm.wtime[tbk];
                    }
                    float sg = sigmoidf_fast(x);
                    av = x * sg;
                    dsv = dat[n][r] * (sg * (1.f + x * (1.f - sg)));
                    my_hp[pbk * 32] += dsv;
                    if (has_time) my_ht[tbk * 32] += dsv;
                }
                st[n][r] = av;
                dat[n][r] = dsv;
            }
        }
        uint32_t pf[4][4];
        att_pack_p(pf, st);
        att_mma_nn<DH>(dv, pf, sm.tile[3], lane);  // dV += A^T dO
        att_pack_p(pf, dat);
        att_mma_nn<DH>(dk, pf, sm.tile[2], lane);  // dK += dS^T Q
    }

#pragma unroll
    for (int n = 0; n < DH / 8; ++n) {
        int col = h * DH + n * 8 + 2 * t;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            int j = half ? j1 : j0;
            if (j >= L) continue;
            float k0v = dk[n][2 * half], k1v = dk[n][2 * half + 1];
            float v0v = dv[n][2 * half], v1v = dv[n][2 * half + 1];
            size_t zo = (size_t)(tok0 + j) * a.ldz + col;
            if (a.zk) {
                float2 z = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.zk + zo));
                k0v *= dsiluf(z.x);
                k1v *= dsiluf(z.y);
            }
            if (a.zv) {
                float2 z = unpack_bf16(*reinterpret_cast<const uint32_t*>(a.zv + zo));
                v0v *= dsiluf(z.x);
                v1v *= dsiluf(z.y);
            }
            size_t go = (size_t)(tok0 + j) * a.lddq + col;
            *reinterpret_cast<uint32_t*>(a.dk + go) = pack_bf16(k0v, k1v);
            *reinterpret_cast<uint32_t*>(a.dv + go) = pack_bf16(v0v, v1v);
        }
    }

    // reduce the lane-private histograms: warp w sums buckets w, w+4, ... over (4 warps x 32 lanes)
    __syncthreads();
    for (int bk = warp; bk < npos; bk += 4) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) v += hist_p[(w * npos + bk) * 32 + lane];
        v = warp_sum(v);
        if (lane == 0 && v != 0.f) atomicAdd(a.dwpos + bk * a.H + h, v);
    }
    if (has_time && a.dwtime) {
        for (int bk = warp; bk < ntime; bk += 4) {
            float v = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) v += hist_t[(w * ntime + bk) * 32 + lane];
            v = warp_sum(v);
            if (lane == 0 && v != 0.f) atomicAdd(a.dwtime + bk * a.H + h, v);
        }
    }
}

}  // namespace grb
