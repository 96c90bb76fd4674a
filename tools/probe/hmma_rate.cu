// developer probe (GPU box): issue rate of legacy mma.sync m16n8k8 TF32 on sm_100a, per SM, for 4/8/16 warps per CTA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ACC>
__global__ void k(float* out, long long* cyc, int iters) {
    unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
    float c[ACC][4] = {};
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < ACC; ++j) mma(c[j], a, b);
    }
    __syncthreads();
    long long t1 = clock64();
    float s = 0; for (int j = 0; j < ACC; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int acc : {1, 2, 3}) {      // dependent-chain latency: one warp, `acc` independent accumulators
        const int iters = 4000;
        if (acc == 1) k<1><<<1, 32>>>(out, cyc, iters); else if (acc == 2) k<2><<<1, 32>>>(out, cyc, iters); else k<3><<<1, 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h0; cudaMemcpy(&h0, cyc, 8, cudaMemcpyDeviceToHost);
        printf("1 warp, %d accumulator chain(s): %.1f cycles per HMMA issue slot\n", acc, (double)h0 / (iters * acc));
    }
    for (int warps : {1, 4, 8, 16, 32}) {
        const int iters = 2000;
        k<8><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        k<8><<<148, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double per = (double)h[0] / (iters * 8.0 * warps);
        printf("warps/CTA %2d: %lld cycles for %d HMMA per warp -> %.2f cycles per HMMA per SM (%.0f MAC/cycle/SM) err=%s\n", warps, h[0], iters * 8, per, 1024.0 / per, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
