// probe: can a 576-thread (18-warp) CTA launch with 112 registers per thread, or are registers allocated per 4 warps (20 x 32 x 112 > 64K)?
#include <cstdio>
#include <cuda_runtime.h>
template <int N> __global__ void __maxnreg__(N) k(float* out, int n) {
    float a[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) a[i] = out[(threadIdx.x + i * 7) % n];
    for (int r = 0; r < n; ++r) {
#pragma unroll
        for (int i = 0; i < 64; ++i) a[i] = fmaf(a[i], a[(i + 1) & 63], (float)r);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += a[i];
    out[threadIdx.x] = s;
}
template <int N> void probe(int threads) {
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<N>);
    int nb = -1; cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<N>, threads, 0);
    float* d; cudaMalloc(&d, 4096 * 4); cudaMemset(d, 0, 4096 * 4);
    k<N><<<1, threads>>>(d, 4); cudaError_t le = cudaGetLastError(); cudaDeviceSynchronize();
    printf("maxnreg %d: numRegs %d, threads %d -> occupancy %d (%s), launch: %s\n", N, fa.numRegs, threads, nb, cudaGetErrorString(e), cudaGetErrorString(le));
    cudaFree(d);
}
int main() { probe<96>(576); probe<104>(576); probe<112>(576); probe<112>(512); probe<120>(512); probe<128>(512); probe<96>(640); probe<104>(640); return 0; }
