// dff_legacy_inst.cu -- the mma.sync fused kernel (dff_kernel.cuh): fallback for shapes outside the tcgen05 kernel's shared-memory
// budget (57..64 beads at hidden 96 / 128) and the A/B reference of the tests (DFF_CONFIG=legacy|wide|tall|duo).
// One configuration per translation unit (-DDFF_LHP=64|128 -DDFF_LCFG=0 wide | 1 tall | 2 duo) for a parallel build.
#include "dff_kernel.cuh"

#define DFF_CAT_(a, b) dff_legacy_launch_##a##_##b
#define DFF_CAT(a, b) DFF_CAT_(a, b)

#if DFF_LCFG == 0
using LC = dff::Cfg<DFF_LHP, 64, 1>;
constexpr int kMinB = 1;
#elif DFF_LCFG == 1
using LC = dff::Cfg<DFF_LHP, 32, 2>;
constexpr int kMinB = 1;
#else
using LC = dff::Cfg<DFF_LHP, 32, 1, 2, 32>;
constexpr int kMinB = (dff::kThreads == 256 ? 2 : 1);
#endif

extern "C" cudaError_t DFF_CAT(DFF_LHP, DFF_LCFG)(const dff::ModelDev* M, const dff::StepArgs* A, int grid, cudaStream_t stream) {
    auto kern = dff::dff_fused_kernel<LC, kMinB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LC::kSmemBytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, dff::kThreads, LC::kSmemBytes, stream>>>(*M, *A);
    return cudaGetLastError();
}
