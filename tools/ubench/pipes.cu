// Micro-benchmark: issue rate of the instructions the exact GEMM is made of (B200, sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 4096, CH = 16;
template <int OP>
__global__ void k(float* out, int* outi, long long* cyc, float fa, float fb, int ia, int ib) {
    float f[CH]; int x[CH];
    for (int i = 0; i < CH; i++) { f[i] = fa + i + threadIdx.x; x[i] = ia + i + threadIdx.x; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) f[i] = fmaf(f[i], fa, fb);
            if (OP == 1) f[i] = __fadd_rn(f[i], fb);
            if (OP == 2) f[i] = __fmul_rn(f[i], fa);
            if (OP == 3) x[i] = __dp4a(x[i], ia, ib);
            if (OP == 4) { if (i & 1) f[i] = fmaf(f[i], fa, fb); else x[i] = __dp4a(x[i], ia, ib); }
            if (OP == 5) { if (i & 1) f[i] = __fadd_rn(f[i], fb); else x[i] = __dp4a(x[i], ia, ib); }
            if (OP == 6) x[i] = (x[i] >> 4) & ib;
            if (OP == 7) { if (i & 1) f[i] = __fadd_rn(f[i], fb); else f[i] = fmaf(f[i], fa, fb); }
            if (OP == 8) { if ((i & 3) == 3) x[i] = (x[i] ^ ia) & ib; else if (i & 1) f[i] = __fadd_rn(f[i], fb); else x[i] = __dp4a(x[i], ia, ib); }
            if (OP == 9) x[i] = x[i] * ia + ib;
        }
    }
    const long long t1 = clock64();
    float s = 0; int si = 0;
    for (int i = 0; i < CH; i++) { s += f[i]; si += x[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; outi[blockIdx.x * blockDim.x + threadIdx.x] = si;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name) {
    float* o; int* oi; long long* c;
    cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&oi, 148 * 1024 * 4); cudaMalloc(&c, 8);
    for (int nt : {128, 256, 512, 1024}) {
        k<OP><<<148, nt>>>(o, oi, c, 1.0001f, 0.5f, 0x01020304, 0x0f0f0f0f);
        cudaDeviceSynchronize();
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h / (ITERS * CH) / (nt / 128.0);
        printf("%-22s warps/SMSP %d: %.3f cycles per warp-instruction per SMSP\n", name, nt / 128, per);
    }
}
int main() {
    run<0>("FFMA"); run<1>("FADD"); run<2>("FMUL"); run<3>("IDP4A"); run<4>("IDP4A+FFMA 1:1"); run<5>("IDP4A+FADD 1:1"); run<6>("SHF+LOP3");
    run<7>("FADD+FFMA 1:1"); run<8>("2 IDP:1 FADD:1 LOP"); run<9>("IMAD");
    return 0;
}
