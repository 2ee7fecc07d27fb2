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

template <int D>
__global__ void __launch_bounds__(RQ_THREADS) rq_residual_argmin_kernel(RqArgs a) {
    extern __shared__ __align__(16) float rq_smem[];
    float* cb = rq_smem;             // [K][D]
    float* cn = rq_smem + a.K * D;   // [K] squared norms
    const int tid = threadIdx.x;
    const long long row = (long long)blockIdx.x * RQ_THREADS + tid;
    const bool live = row < a.N;

    float r[D];
    if (live) {
#pragma unroll
        for (int j = 0; j < D; j += 4) {
            float4 v = *reinterpret_cast<const float4*>(a.x + row * D + j);
            r[j] = v.x; r[j + 1] = v.y; r[j + 2] = v.z; r[j + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < D; ++j) r[j] = 0.f;
    }
    float loss = 0.f;

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

        float xn = 0.f;
#pragma unroll
        for (int j = 0; j < D; ++j) xn = fmaf(r[j], r[j], xn);
        if (a.res && live) {
#pragma unroll
            for (int j = 0; j < D; ++j) a.res[(row * D + j) * a.levels + l] = r[j];
        }

        float best = INFINITY;
        int best_k = 0;
        for (int k = 0; k < a.K; k += 2) {
            float d0 = 0.f, d1 = 0.f;
            const float4* c0 = reinterpret_cast<const float4*>(cb + k * D);
            const float4* c1 = reinterpret_cast<const float4*>(cb + (k + 1) * D);
#pragma unroll
            for (int j = 0; j < D / 4; ++j) {
                float4 u = c0[j], w = c1[j];
                d0 = fmaf(r[4 * j], u.x, d0); d0 = fmaf(r[4 * j + 1], u.y, d0);
                d0 = fmaf(r[4 * j + 2], u.z, d0); d0 = fmaf(r[4 * j + 3], u.w, d0);
                d1 = fmaf(r[4 * j], w.x, d1); d1 = fmaf(r[4 * j + 1], w.y, d1);
                d1 = fmaf(r[4 * j + 2], w.z, d1); d1 = fmaf(r[4 * j + 3], w.w, d1);
            }
            float dist0 = (xn + cn[k]) - 2.f * d0;
            float dist1 = (xn + cn[k + 1]) - 2.f * d1;
            if (dist0 < best) { best = dist0; best_k = k; }
            if (dist1 < best) { best = dist1; best_k = k + 1; }
        }
        float sq = 0.f;
        const float* cw = cb + best_k * D;
#pragma unroll
        for (int j = 0; j < D; ++j) {
            float e = cw[j];
            if (a.emb && live) a.emb[(row * D + j) * a.levels + l] = e;
            r[j] -= e;
            sq = fmaf(r[j], r[j], sq);
        }
        loss += sq + a.commitment * sq;
        if (live) a.ids[row * a.levels + l] = best_k;
    }
    if (live) {
        if (a.loss) a.loss[row] = loss;
        if (a.res_out) {
#pragma unroll
            for (int j = 0; j < D; j += 4)
                *reinterpret_cast<float4*>(a.res_out + row * D + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
    }
}

}  // namespace grb
