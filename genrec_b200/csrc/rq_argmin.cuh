// genrec_b200 - RQ-VAE residual nearest-codebook search (reference: genrec/models/rqvae.py:185-199, :397-412).
//
//   for level l:  id_l = argmin_k ( |r|^2 + |c_k|^2 - 2 r.c_k )  (first index on ties) ;  r <- r - c_{id_l}
//
// No tensor cores by specification: FP32 FMA on the CUDA cores.  One thread owns one item row (its residual lives in
// registers for all levels); the level's codebook sits in shared memory and every lane reads the SAME code word at the
// same time (a broadcast, conflict-free), so the inner loop is 4 FFMA per 16-byte LDS.  Two code words are processed
// per iteration for ILP.  HBM traffic per item: D*4 B in, levels*8 B out (+ optional embeddings / residuals / loss).
#pragma once
#include "common.cuh"

namespace grb {

constexpr int RQ_THREADS = 256;

struct RqArgs {
    const float* x;          // [N, D] latent
    const float* codebooks;  // [levels, K, D]
    long long* ids;          // [N, levels]
    float* emb;              // nullable [N, D, levels]   (reference layout, rqvae.py:408)
    float* res;              // nullable [N, D, levels]   residual BEFORE each level (rqvae.py:400, :409)
    float* loss;             // nullable [N]  sum_l (1 + w) * |r_l - e_l|^2   (loss.py:75-77)
    float* res_out;          // nullable [N, D] final residual
    long long N;
    int K, levels;
    float commitment;
};

// ROWS item rows per thread (register tile): one broadcast LDS.128 of a code word feeds 4 * ROWS FFMA, which moves the
// inner loop from shared-memory-issue bound (ROWS = 1: 1 LDS per 4 FFMA) to FMA bound (ROWS = 2).
template <int D, int ROWS>
__global__ void __launch_bounds__(RQ_THREADS) rq_residual_argmin_kernel(RqArgs a) {
    pdl_wait();
    extern __shared__ __align__(16) float rq_smem[];
    float* cb = rq_smem;             // [K][D]
    float* cn = rq_smem + a.K * D;   // [K] squared norms
    const int tid = threadIdx.x;
    long long row[ROWS];
    bool live[ROWS];
    float r[ROWS][D];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        row[q] = ((long long)blockIdx.x * ROWS + q) * RQ_THREADS + tid;
        live[q] = row[q] < a.N;
        if (live[q]) {
#pragma unroll
            for (int j = 0; j < D; j += 4) {
                float4 v = *reinterpret_cast<const float4*>(a.x + row[q] * D + j);
                r[q][j] = v.x; r[q][j + 1] = v.y; r[q][j + 2] = v.z; r[q][j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < D; ++j) r[q][j] = 0.f;
        }
    }
    float loss[ROWS];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) loss[q] = 0.f;

    for (int l = 0; l < a.levels; ++l) {
        __syncthreads();
        const float* g = a.codebooks + (size_t)l * a.K * D;
        for (int i = tid * 4; i < a.K * D; i += RQ_THREADS * 4)
            *reinterpret_cast<float4*>(cb + i) = *reinterpret_cast<const float4*>(g + i);
        __syncthreads();
        for (int k = tid; k < a.K; k += RQ_THREADS) {
            float s = 0.f;
            for (int j = 0; j < D; ++j) s = fmaf(cb[k * D + j], cb[k * D + j], s);
            cn[k] = s;
        }
        __syncthreads();

        float xn[ROWS], best[ROWS];
        int best_k[ROWS];
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            xn[q] = 0.f;
#pragma unroll
            for (int j = 0; j < D; ++j) xn[q] = fmaf(r[q][j], r[q][j], xn[q]);
            if (a.res && live[q]) {
#pragma unroll
                for (int j = 0; j < D; ++j) a.res[(row[q] * D + j) * a.levels + l] = r[q][j];
            }
            best[q] = INFINITY;
            best_k[q] = 0;
        }
        for (int k = 0; k < a.K; k += 2) {
            float d0[ROWS], d1[ROWS];
#pragma unroll
            for (int q = 0; q < ROWS; ++q) d0[q] = d1[q] = 0.f;
            const float4* c0 = reinterpret_cast<const float4*>(cb + k * D);
            const float4* c1 = reinterpret_cast<const float4*>(cb + (k + 1) * D);
#pragma unroll
            for (int j = 0; j < D / 4; ++j) {
                const float4 u = c0[j], w = c1[j];
#pragma unroll
                for (int q = 0; q < ROWS; ++q) {
                    d0[q] = fmaf(r[q][4 * j], u.x, d0[q]); d0[q] = fmaf(r[q][4 * j + 1], u.y, d0[q]);
                    d0[q] = fmaf(r[q][4 * j + 2], u.z, d0[q]); d0[q] = fmaf(r[q][4 * j + 3], u.w, d0[q]);
                    d1[q] = fmaf(r[q][4 * j], w.x, d1[q]); d1[q] = fmaf(r[q][4 * j + 1], w.y, d1[q]);
                    d1[q] = fmaf(r[q][4 * j + 2], w.z, d1[q]); d1[q] = fmaf(r[q][4 * j + 3], w.w, d1[q]);
                }
            }
            const float n0 = cn[k], n1 = cn[k + 1];
#pragma unroll
            for (int q = 0; q < ROWS; ++q) {
                const float dist0 = (xn[q] + n0) - 2.f * d0[q];
                const float dist1 = (xn[q] + n1) - 2.f * d1[q];
                if (dist0 < best[q]) { best[q] = dist0; best_k[q] = k; }
                if (dist1 < best[q]) { best[q] = dist1; best_k[q] = k + 1; }
            }
        }
#pragma unroll
        for (int q = 0; q < ROWS; ++q) {
            float sq = 0.f;
            const float* cw = cb + best_k[q] * D;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const float e = cw[j];
                if (a.emb && live[q]) a.emb[(row[q] * D + j) * a.levels + l] = e;
                r[q][j] -= e;
                sq = fmaf(r[q][j], r[q][j], sq);
            }
            loss[q] += sq + a.commitment * sq;
            if (live[q]) a.ids[row[q] * a.levels + l] = best_k[q];
        }
    }
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
        if (!live[q]) continue;
        if (a.loss) a.loss[row[q]] = loss[q];
        if (a.res_out) {
#pragma unroll
            for (int j = 0; j < D; j += 4)
                *reinterpret_cast<float4*>(a.res_out + row[q] * D + j) = make_float4(r[q][j], r[q][j + 1], r[q][j + 2], r[q][j + 3]);
        }
    }
}

}  // namespace grb
