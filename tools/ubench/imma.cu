// Micro-benchmark: legacy warp-level integer MMA (mma.sync m16n8k32 u8 x s8 -> s32) on B200, alone and mixed with the FP32
// work the exact GEMM does per lane sum (FFMA + FADD), to see whether it can replace the half-rate IDP.4A.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096;
__device__ __forceinline__ void imma(int (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2], const int (&c)[4]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]));
}
template <int MODE>
__global__ void k(int* out, float* outf, long long* cyc, unsigned seed, float fs) {
    unsigned a[4], b[4][2];
    int c[4] = {0x4b400000, 0x4b400000, 0x4b400000, 0x4b400000};
    for (int i = 0; i < 4; i++) { a[i] = seed * (i + 1) + threadIdx.x; for (int j = 0; j < 2; j++) b[i][j] = seed * (i + 7 + j) ^ threadIdx.x; }
    float acc[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
    int sum = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            int d[4];
            imma(d, a, b[m], c);
            if (MODE == 0) { sum += d[0] ^ d[1] ^ d[2] ^ d[3]; }
            else {
                // per lane sum: FFMA (bias trick) + FADD, plus 2 scale FMULs per pair
                const float s0 = fs * (float)(m + 1), ms0 = -12582912.0f * s0;
#pragma unroll
                for (int j = 0; j < 4; j++) acc[m][j] = __fadd_rn(acc[m][j], fmaf(__int_as_float(d[j]), s0, ms0));
            }
            a[0] += (unsigned)it;        // keep the MMAs from being hoisted
        }
    }
    const long long t1 = clock64();
    float fsum = 0;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) fsum += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum; outf[blockIdx.x * blockDim.x + threadIdx.x] = fsum;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(const char* name) {
    int* o; float* of; long long* c;
    cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&of, 148 * 1024 * 4); cudaMalloc(&c, 8);
    for (int nt : {128, 256, 512, 1024}) {
        k<MODE><<<148, nt>>>(o, of, c, 12345u, 1e-3f);
        cudaError_t e = cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / (ITERS * 4) / (nt / 128.0);
        printf("%-28s warps/SMSP %d: %.2f cycles per MMA per SMSP (%s)  => %.0f int8 MACs/cycle/SM\n", name, nt / 128, per, cudaGetErrorString(e), 4096.0 * 4 / per);
    }
}
int main() { run<0>("IMMA m16n8k32 alone"); run<1>("IMMA + 4x(FFMA+FADD)+2 FMUL"); return 0; }
