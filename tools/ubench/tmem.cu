// Micro-benchmark + correctness spike: tcgen05.mma kind::i8 (u8 x s8 -> s32, M128 N256 K32, operands in 128-byte-swizzled shared
// memory written by ordinary stores) feeding an epilogue of 16 warps that reads EVERY int32 of the accumulator back (tcgen05.ld
// 32x32b.x64) and runs the exact GEMM's per-lane-sum work on it (IADD + FFMA + FADD per value, 2 FMUL per four values).
// Question: how many cycles per quant block (128 channels x 64 rows x 4 lane sums) can one SM sustain -- the SIMT dp4a kernel
// needs ~2300 (0.28 cycles per (row, channel, block) item).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench/tmem tools/ubench/tmem.cu && tools/ubench/tmem
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int EPI_WARPS = 16, NT = (EPI_WARPS + 1) * 32;
constexpr int BM = 128, BN = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]),
          "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]),
          "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Smem {
    alignas(1024) uint8_t A[BM * 128];      // 128 channels x 128 B (4 quant blocks), SWIZZLE_128B
    alignas(1024) uint8_t B[BN * 128];      // 256 N-rows (= 64 rows x 4 masked lane copies) x 128 B
    float ad[4][64];                        // activation scale per (block, row)
    int4 nb[4][64];                         // per (block, row): bias - 7 * lane code sums (Q4)
    uint64_t bars[4];
    uint32_t tmem_slot;
};
__host__ __device__ inline int sw128(int r, int kbyte) { return (r >> 3) * 1024 + (r & 7) * 128 + (((kbyte >> 4) ^ (r & 7)) << 4) + (kbyte & 15); }

// MODE 0: MMA + tcgen05.ld only; MODE 1: + the per-value arithmetic of int32 lane sums; MODE 2: the arithmetic if the lane sums
// arrived as exact fp32 (kind::f16 MMA on fp16 copies of the codes): FMUL + FADD per value, one FMUL per four values
template <int MODE>
__global__ void __launch_bounds__(NT, 1) k(const uint8_t* gA, const int8_t* gB, int* dump, float* outf, long long* cyc) {
    extern __shared__ unsigned char raw[];
    Smem& sm = *reinterpret_cast<Smem*>(raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < BM * 128; i += NT) sm.A[sw128(i >> 7, i & 127)] = gA[i];
    for (int i = tid; i < BN * 128; i += NT) sm.B[sw128(i >> 7, i & 127)] = (uint8_t)gB[i];
    for (int i = tid; i < 4 * 64; i += NT) { sm.ad[i >> 6][i & 63] = 1e-3f * (float)(1 + (i & 7)); sm.nb[i >> 6][i & 63] = make_int4(0x4b400000 - i, 0x4b400000 + i, 0x4b400000, 0x4b400000 - 3); }
    const uint32_t bar0 = smem_u32(&sm.bars[0]);
    auto tfull = [&](int a) { return bar0 + 8u * a; };
    auto tempty = [&](int a) { return bar0 + 8u * (2 + a); };
    if (tid == 0) {
        for (int a = 0; a < 2; a++) { mbar_init(tfull(a), 1); mbar_init(tempty(a), EPI_WARPS * 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores to A / B visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_slot;
    // u8 x s8 -> s32: c_format 2 at bits 4-5, a_format 0 (u8) bits 7-9, b_format 1 (s8) bits 10-12, N >> 3 at 17, M >> 4 at 24
    constexpr uint32_t IDESC = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const long long t0 = clock64();
    if (warp == EPI_WARPS) {
        if (lane == 0) {
            const uint32_t sa = smem_u32(sm.A), sb = smem_u32(sm.B);
            for (int it = 0; it < ITERS; it++) {
                const int as = it & 1, kb = it & 3;
                mbar_wait(tempty(as), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                umma_i8(tmem + (uint32_t)(as * BN), umma_desc_sw128(sa + kb * 32), umma_desc_sw128(sb + kb * 32), IDESC);
                umma_commit(tfull(as));
            }
        }
        __syncwarp();
    } else {
        const int lg = warp & 3, cg = warp >> 2;          // TMEM lane quarter; 64-column group = 16 rows x 4 lane sums
        const float dw = 0.01f * (float)(1 + (tid & 15));
        float acc[16][4];
#pragma unroll
        for (int r = 0; r < 16; r++)
#pragma unroll
            for (int l = 0; l < 4; l++) acc[r][l] = 0.0f;
        int isum = 0;
        for (int it = 0; it < ITERS; it++) {
            const int as = it & 1, kb = it & 3;
            mbar_wait(tfull(as), (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BN + cg * 64);
#pragma unroll
            for (int h = 0; h < 4; h++) {
                uint32_t d[16];
                tmem_ld16(taddr + h * 16, d);
                if (h == 3) { tc_fence_before(); mbar_arrive(tempty(as)); }
                if (MODE == 0 && it < 4 && blockIdx.x == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j++) dump[((size_t)it * BM + lg * 32 + lane) * BN + cg * 64 + h * 16 + j] = (int)d[j];
                }
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 16; j++) isum ^= (int)d[j];
                } else if (MODE == 2) {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const int rr = h * 4 + r;
                        const float s = __fmul_rn(sm.ad[kb][cg * 16 + rr], dw);
                        acc[rr][0] = __fadd_rn(acc[rr][0], __fmul_rn(__uint_as_float(d[4 * r + 0]), s));
                        acc[rr][1] = __fadd_rn(acc[rr][1], __fmul_rn(__uint_as_float(d[4 * r + 1]), s));
                        acc[rr][2] = __fadd_rn(acc[rr][2], __fmul_rn(__uint_as_float(d[4 * r + 2]), s));
                        acc[rr][3] = __fadd_rn(acc[rr][3], __fmul_rn(__uint_as_float(d[4 * r + 3]), s));
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const int rr = h * 4 + r;
                        const float s = __fmul_rn(sm.ad[kb][cg * 16 + rr], dw);
                        const float ms = __fmul_rn(-12582912.0f, s);
                        const int4 nbv = sm.nb[kb][cg * 16 + rr];
                        acc[rr][0] = __fadd_rn(acc[rr][0], fmaf(__int_as_float((int)d[4 * r + 0] + nbv.x), s, ms));
                        acc[rr][1] = __fadd_rn(acc[rr][1], fmaf(__int_as_float((int)d[4 * r + 1] + nbv.y), s, ms));
                        acc[rr][2] = __fadd_rn(acc[rr][2], fmaf(__int_as_float((int)d[4 * r + 2] + nbv.z), s, ms));
                        acc[rr][3] = __fadd_rn(acc[rr][3], fmaf(__int_as_float((int)d[4 * r + 3] + nbv.w), s, ms));
                    }
                }
            }
        }
        float f = (float)isum;
#pragma unroll
        for (int r = 0; r < 16; r++) f += (acc[r][0] + acc[r][1]) + (acc[r][2] + acc[r][3]);
        outf[blockIdx.x * 512 + tid] = f;
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) *cyc = clock64() - t0;
    if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

template <int MODE>
static int run(const char* name, const uint8_t* dA, const int8_t* dB, const std::vector<uint8_t>& hA, const std::vector<int8_t>& hB) {
    int* dump; float* of; long long* c;
    cudaMalloc(&dump, 4 * BM * BN * 4); cudaMalloc(&of, 148 * 512 * 4); cudaMalloc(&c, 8);
    cudaMemset(dump, 0xff, 4 * BM * BN * 4);
    const size_t smem = sizeof(Smem) + 1024;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<MODE><<<148, NT, smem>>>(dA, dB, dump, of, c);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return 1; }
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    std::vector<int> d(4 * BM * BN);
    cudaMemcpy(d.data(), dump, d.size() * 4, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int kb = 0; kb < (MODE == 0 ? 4 : 0); kb++)
        for (int m = 0; m < BM; m++)
            for (int n = 0; n < BN; n++) {
                int ref = 0;
                for (int kk = 0; kk < 32; kk++) ref += (int)hA[m * 128 + kb * 32 + kk] * (int)hB[n * 128 + kb * 32 + kk];
                if (ref != d[((size_t)kb * BM + m) * BN + n]) { if (bad < 5) printf("  mismatch kb %d m %d n %d: got %d want %d\n", kb, m, n, d[((size_t)kb * BM + m) * BN + n], ref); bad++; }
            }
    const double per = (double)h / ITERS;
    printf("%-34s %.0f cycles per quant block per SM (128 ch x 64 rows x 4 lanes) = %.3f cycles per item; int32 results %s (%lld wrong)\n", name, per,
           per / (128.0 * 64.0), bad ? "WRONG" : "exact", bad);
    return bad ? 1 : 0;
}

int main() {
    std::vector<uint8_t> hA(BM * 128);
    std::vector<int8_t> hB(BN * 128);
    srand(5);
    for (auto& v : hA) v = (uint8_t)(rand() % 16);                   // Q4 codes
    for (size_t i = 0; i < hB.size(); i++) {
        const int n = (int)(i >> 7), kbyte = (int)(i & 127), l = n & 3, w = (kbyte & 31) >> 2;    // word w of the block belongs to lane w & 3
        hB[i] = ((w & 3) == l) ? (int8_t)(rand() % 255 - 127) : 0;     // masked copy l keeps lane l's eight codes
    }
    uint8_t* dA; int8_t* dB;
    cudaMalloc(&dA, hA.size()); cudaMalloc(&dB, hB.size());
    cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
    int rc = run<0>("i8 MMA + tcgen05.ld only", dA, dB, hA, hB);
    rc |= run<1>("i8 MMA + ld + int32 lane-sum arithmetic", dA, dB, hA, hB);
    rc |= run<2>("MMA + ld + fp32 lane-sum arithmetic", dA, dB, hA, hB);
    return rc;
}
