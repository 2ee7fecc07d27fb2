// genrec_b200 - data-parallel optimizer step as ONE pass over NVLink peer memory: all-reduce + Adam + parameter broadcast.
//
// Replaces `all_reduce(flat_grad)` (NCCL) followed by the fused Adam kernel.  The flat fp32 gradient, the fp32 master parameters
// and their bf16 operand mirror of every rank live in symmetric (peer-mapped, and where the fabric allows multicast-mapped)
// memory.  Rank r owns the slice [r n/W, (r+1) n/W) of the flat buffers:
//   1. dp_barrier_kernel   every rank's backward has finished writing its gradient (release/acquire flags over NVLink)
//   2. dp_adam_kernel      g = sum over ranks of the slice - one `multimem.ld_reduce` per 16 bytes (the NVSwitch adds the
//                          eight copies in the fabric), or W peer loads without multicast ; Adam on the slice (the moments of a
//                          slice exist on its owner only) ; the new parameters go to EVERY rank: `multimem.st` of the fp32
//                          master and of the bf16 mirror (or W peer stores)
//   3. dp_barrier_kernel   all parameter stores have landed and every rank is done reading my gradient -> it may be zeroed
// Bytes on the wire per rank and step: n/W * 4 read (reduced in the switch) + n/W * 6 written (multicast) - against
// 2 (W-1)/W n * 4 each way for a ring all-reduce; no second kernel touches the parameters.
#pragma once
#include "common.cuh"

namespace grb {

struct DpAdamArgs {
    float* p; float* g; float* m; float* v; bf16* mirror;          // this rank's flat buffers (m, v: only the owned slice is live)
    const float* mc_g; float* mc_p; bf16* mc_mirror;              // multicast addresses of g / p / mirror, or null
    const float* const* peer_g; float* const* peer_p; bf16* const* peer_mirror;   // device arrays [world] of peer pointers
    size_t n;                 // elements, a multiple of 8 * world
    int rank, world;
    const float* state;       // {step, 1 - b1^step, 1 - b2^step}
    float lr, beta1, beta2, eps, weight_decay, grad_scale;
};

GRB_DEVINL float4 multimem_ld_reduce_add(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
GRB_DEVINL void multimem_st_v4(void* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <bool MC>
__global__ void __launch_bounds__(256) dp_adam_kernel(DpAdamArgs a) {
    pdl_wait();
    const float bc1 = a.state[1], bc2 = a.state[2];
    const float step_size = a.lr / bc1;
    const float inv_sqrt_bc2 = rsqrtf(bc2);
    const size_t per = a.n / a.world;                       // multiple of 8
    const size_t lo = (size_t)a.rank * per;
    for (size_t e = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8; e < per; e += (size_t)gridDim.x * blockDim.x * 8) {
        const size_t i = lo + e;
        float g[8];
        if (MC) {
            const float4 g0 = multimem_ld_reduce_add(a.mc_g + i), g1 = multimem_ld_reduce_add(a.mc_g + i + 4);
            g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = 0.f;
            for (int r = 0; r < a.world; ++r) {
                const float4 g0 = *reinterpret_cast<const float4*>(a.peer_g[r] + i), g1 = *reinterpret_cast<const float4*>(a.peer_g[r] + i + 4);
                g[0] += g0.x; g[1] += g0.y; g[2] += g0.z; g[3] += g0.w; g[4] += g1.x; g[5] += g1.y; g[6] += g1.z; g[7] += g1.w;
            }
        }
        float p[8], m[8], v[8];
        *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(a.p + i);
        *reinterpret_cast<float4*>(p + 4) = *reinterpret_cast<const float4*>(a.p + i + 4);
        *reinterpret_cast<float4*>(m) = *reinterpret_cast<const float4*>(a.m + i);
        *reinterpret_cast<float4*>(m + 4) = *reinterpret_cast<const float4*>(a.m + i + 4);
        *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(a.v + i);
        *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(a.v + i + 4);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float gk = g[k] * a.grad_scale;
            if (a.weight_decay != 0.f) gk += a.weight_decay * p[k];
            m[k] = a.beta1 * m[k] + (1.f - a.beta1) * gk;
            v[k] = a.beta2 * v[k] + (1.f - a.beta2) * gk * gk;
            p[k] -= step_size * (m[k] / (sqrtf(v[k]) * inv_sqrt_bc2 + a.eps));
        }
        *reinterpret_cast<float4*>(a.m + i) = *reinterpret_cast<float4*>(m);
        *reinterpret_cast<float4*>(a.m + i + 4) = *reinterpret_cast<float4*>(m + 4);
        *reinterpret_cast<float4*>(a.v + i) = *reinterpret_cast<float4*>(v);
        *reinterpret_cast<float4*>(a.v + i + 4) = *reinterpret_cast<float4*>(v + 4);
        float4 mir;
        {
            uint32_t* u = reinterpret_cast<uint32_t*>(&mir);
            u[0] = pack_bf16(p[0], p[1]); u[1] = pack_bf16(p[2], p[3]); u[2] = pack_bf16(p[4], p[5]); u[3] = pack_bf16(p[6], p[7]);
        }
        if (MC) {
            multimem_st_v4(a.mc_p + i, *reinterpret_cast<float4*>(p));
            multimem_st_v4(a.mc_p + i + 4, *reinterpret_cast<float4*>(p + 4));
            multimem_st_v4(a.mc_mirror + i, mir);
        } else {
            for (int r = 0; r < a.world; ++r) {
                *reinterpret_cast<float4*>(a.peer_p[r] + i) = *reinterpret_cast<float4*>(p);
                *reinterpret_cast<float4*>(a.peer_p[r] + i + 4) = *reinterpret_cast<float4*>(p + 4);
                *reinterpret_cast<float4*>(a.peer_mirror[r] + i) = mir;
            }
        }
    }
}

// One CTA; thread t < world talks to peer t.  sig: this rank's flag array [2 channels][world] in symmetric memory (zero at
// start); peer_sig[t]: rank t's array; epoch: local counters [2], bumped here so that a captured CUDA graph replays correctly.
__global__ void __launch_bounds__(32) dp_barrier_kernel(unsigned* const* peer_sig, unsigned* sig, unsigned* epoch, int rank, int world, int channel) {
    pdl_wait();
    __shared__ unsigned e_s;
    if (threadIdx.x == 0) e_s = ++epoch[channel];
    __syncthreads();
    const unsigned e = e_s;
    const int t = threadIdx.x;
    if (t < world) {
        __threadfence_system();      // everything this GPU wrote before (earlier kernels on the stream) is ordered before the flag
        unsigned* dst = peer_sig[t] + channel * world + rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(e) : "memory");
        const unsigned* src = sig + channel * world + t;
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        } while ((int)(v - e) < 0);
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace grb
