// dff_tc_inst.cu -- ONE instantiation of the fused tcgen05 kernel per translation unit, so that the build can compile the
// configurations in parallel (nvcc -DDFF_PN=.. -DDFF_HP=.. -DDFF_R=.. -DDFF_ATT=..).  Exports a C-linkage launcher
// dff_tc_launch_<PN>_<HP>_<R>_<ATT> that dff_b200.cu picks from its configuration table.
#include "dff_kernel_tc.cuh"

#define DFF_CAT_(a, b, c, d) dff_tc_launch_##a##_##b##_##c##_##d
#define DFF_CAT(a, b, c, d) DFF_CAT_(a, b, c, d)

extern "C" cudaError_t DFF_CAT(DFF_PN, DFF_HP, DFF_R, DFF_ATT)(const dff::ModelDev* M, const dff::StepArgs* A, const dff::v2::TcArgs* T,
                                                                int grid, cudaStream_t stream) {
    using C = dff::v2::TcCfg<DFF_PN, DFF_HP, DFF_R, DFF_ATT>;
    auto kern = dff::v2::dff_fused_tc_kernel<C>;
    // set on every call: a few hundred ns, and correct for any device ordinal / any number of host threads
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmemBytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, dff::v2::kTcThreads, C::kSmemBytes, stream>>>(*M, *A, *T);
    return cudaGetLastError();
}
