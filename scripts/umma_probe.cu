// Hardware probe for the UMMA operand forms the tcgen05 attention kernels rely on (not part of the library).
//   mode 0: B MN-major, SWIZZLE_128B, N = 64 (whole 64-wide atom)                  - the form tc_ce.cuh already uses
//   mode 1: B MN-major, SWIZZLE_128B, N = 32, columns  0..31 of the 64-wide atom   (descriptor start + 0 B)
//   mode 2: B MN-major, SWIZZLE_128B, N = 32, columns 32..63 of the 64-wide atom   (descriptor start + 64 B)
//   mode 3: A from TMEM (tcgen05.st of packed bf16 pairs), B as mode 0
//   mode 4: B MN-major, SWIZZLE_64B tile of 32 columns (TMA SWIZZLE_64B box), N = 32
//   mode 5: A from TMEM, B as mode 2
//   mode 6: A K-major k-slices 2,3 only (K = 32 out of a 64-wide box), B K-major rows = N = 128 (the S = Q_h K_h^T form), head 1
//   mode 7: A MN-major (M = 128 = two 64-atoms, LBO) x B MN-major N = 32 sub-atom offset 64 B  (the dV / dK form)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o scripts/umma_probe scripts/umma_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../genrec_b200/csrc/tc_gemm.cuh"

using namespace grb;

GRB_DEVINL uint64_t umma_desc_sw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
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
GRB_DEVINL void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// A_g [128][64] bf16 (k contiguous) ; At_g [64 k][128 m] bf16 (m contiguous) ; B_g [64 k][64 n] bf16 (n contiguous) ; Bk_g [128 n][64 k]
__global__ void __launch_bounds__(128, 1)
    probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAt, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmB64, const __grid_constant__ CUtensorMap tmBk, const bf16* __restrict__ A_g, int mode,
                 float* __restrict__ out /* [128][128] */) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;               // 16 KB  [128 m][64 k]  K-major SW128
    unsigned char* sAt = base + 16384;      // 16 KB  2 x [64 k][64 m] MN-major SW128 (m blocks 8 KB apart)
    unsigned char* sB = base + 32768;       // 8 KB   [64 k][64 n]  MN-major SW128
    unsigned char* sB64 = base + 40960;     // 2 x 4 KB [64 k][32 n]  SW64 (n blocks 0..31, 32..63)
    unsigned char* sBk = base + 49152;      // 16 KB  [128 n][64 k]  K-major SW128
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 65536);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], 16384 + 16384 + 8192 + 8192 + 16384);
        tma_load_2d(sA, &tmA, 0, 0, &bars[0]);
        tma_load_2d(sAt, &tmAt, 0, 0, &bars[0]);
        tma_load_2d(sAt + 8192, &tmAt, 64, 0, &bars[0]);
        tma_load_2d(sB, &tmB, 0, 0, &bars[0]);
        tma_load_2d(sB64, &tmB64, 0, 0, &bars[0]);
        tma_load_2d(sB64 + 4096, &tmB64, 32, 0, &bars[0]);
        tma_load_2d(sBk, &tmBk, 0, 0, &bars[0]);
    }
    // A into TMEM columns 128.. as packed bf16 pairs: lane = row, column c holds k = 2c, 2c+1
    if (mode == 3 || mode == 5) {
        const int r = warp * 32 + lane;
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = *reinterpret_cast<const uint32_t*>(A_g + r * 64 + 2 * c);
        tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 128, v);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t a = smem_u32(sA), at = smem_u32(sAt), b = smem_u32(sB), b64 = smem_u32(sB64), bk = smem_u32(sBk);
        if (mode == 0 || mode == 1 || mode == 2) {
            const int N = mode == 0 ? 64 : 32;
            const uint32_t idesc = umma_idesc(128, N, 0, 1);
            const uint32_t off = mode == 2 ? 64 : 0;
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, umma_desc(a + k * 32, 16, 1024), umma_desc(b + off + k * 2048, 8192, 1024), idesc, k > 0);
        } else if (mode == 3 || mode == 5) {
            const int N = mode == 3 ? 64 : 32;
            const uint32_t idesc = umma_idesc(128, N, 0, 1);
            const uint32_t off = mode == 5 ? 64 : 0;
            for (int k = 0; k < 4; ++k) umma_bf16_ts(tmem, tmem + 128 + k * 8, umma_desc(b + off + k * 2048, 8192, 1024), idesc, k > 0);
        } else if (mode == 4) {
            const uint32_t idesc = umma_idesc(128, 32, 0, 1);
            // SW64 MN-major: rows of 64 B, 8-row groups 512 B apart; second 32-column block used (columns 32..63)
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, umma_desc(a + k * 32, 16, 1024), umma_desc_sw(b64 + 4096 + k * 1024, 4096, 512, 4), idesc, k > 0);
        } else if (mode == 6) {
            const uint32_t idesc = umma_idesc(128, 128, 0, 0);
            for (int k = 2; k < 4; ++k) umma_bf16(tmem, umma_desc(a + k * 32, 16, 1024), umma_desc(bk + k * 32, 16, 1024), idesc, k > 2);
        } else if (mode == 7) {
            const uint32_t idesc = umma_idesc(128, 32, 1, 1);
            for (int k = 0; k < 4; ++k)
                umma_bf16(tmem, umma_desc(at + k * 2048, 8192, 1024), umma_desc(b + 64 + k * 2048, 8192, 1024), idesc, k > 0);
        }
        umma_commit(&bars[1]);
    }
    __syncthreads();
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    {
        const int r = warp * 32 + lane;
        for (int c = 0; c < 4; ++c) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c * 32, v);
            for (int i = 0; i < 32; ++i) out[r * 128 + c * 32 + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 256);
    }
}

static bool make_tmap_sw(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols, uint32_t box_rows,
                         CUtensorMapSwizzle sw) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int main() {
    std::vector<float> A(128 * 64), B(64 * 64), Bk(128 * 64);
    std::vector<bf16> Ab(128 * 64), Atb(64 * 128), Bb(64 * 64), Bkb(128 * 64);
    srand(1);
    auto rnd = []() { return (float)((rand() % 17) - 8) / 8.f; };   // exactly representable in bf16, sums exact in fp32
    for (int i = 0; i < 128 * 64; ++i) { A[i] = rnd(); Ab[i] = __float2bfloat16(A[i]); }
    for (int m = 0; m < 128; ++m) for (int k = 0; k < 64; ++k) Atb[k * 128 + m] = Ab[m * 64 + k];
    for (int i = 0; i < 64 * 64; ++i) { B[i] = rnd(); Bb[i] = __float2bfloat16(B[i]); }
    for (int i = 0; i < 128 * 64; ++i) { Bk[i] = rnd(); Bkb[i] = __float2bfloat16(Bk[i]); }
    bf16 *dA, *dAt, *dB, *dBk;
    float* dOut;
    cudaMalloc(&dA, Ab.size() * 2); cudaMalloc(&dAt, Atb.size() * 2); cudaMalloc(&dB, Bb.size() * 2); cudaMalloc(&dBk, Bkb.size() * 2);
    cudaMalloc(&dOut, 128 * 128 * 4);
    cudaMemcpy(dA, Ab.data(), Ab.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dAt, Atb.data(), Atb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bb.data(), Bb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dBk, Bkb.data(), Bkb.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap tmA, tmAt, tmB, tmB64, tmBk;
    bool ok = make_tmap_sw(&tmA, dA, 128, 64, 64, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_tmap_sw(&tmAt, dAt, 64, 128, 128, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_tmap_sw(&tmB, dB, 64, 64, 64, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B) &&
              make_tmap_sw(&tmB64, dB, 64, 64, 64, 32, 64, CU_TENSOR_MAP_SWIZZLE_64B) &&
              make_tmap_sw(&tmBk, dBk, 128, 64, 64, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (!ok) { printf("tensor map creation failed\n"); return 1; }
    const int smem = 65536 + 1024 + 256;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    std::vector<float> out(128 * 128);
    int fails = 0;
    for (int mode = 0; mode < 8; ++mode) {
        cudaMemset(dOut, 0xff, 128 * 128 * 4);
        probe_kernel<<<1, 128, smem>>>(tmA, tmAt, tmB, tmB64, tmBk, dA, mode, dOut);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost);
        int N = (mode == 0 || mode == 3) ? 64 : (mode == 6 ? 128 : 32);
        int n0 = (mode == 2 || mode == 4 || mode == 5 || mode == 7) ? 32 : 0;
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                if (mode == 6) { for (int k = 32; k < 64; ++k) ref += (double)A[m * 64 + k] * Bk[n * 64 + k]; }
                else { for (int k = 0; k < 64; ++k) ref += (double)A[m * 64 + k] * B[k * 64 + n0 + n]; }
                maxerr = fmax(maxerr, fabs(ref - out[m * 128 + n]));
            }
        printf("mode %d: max |err| = %g  %s\n", mode, maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
        fails += maxerr < 1e-3 ? 0 : 1;
    }
    printf("probe done, %d failing modes\n", fails);
    return 0;
}
