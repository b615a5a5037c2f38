// Kernel launches of the ECO launchers (eco_cg.cu, eco_loc.cu) behind two names, so that the launchers' host code -- argument validation,
// launch plan, workspace carving, parameter binding -- compiles verbatim as host C++ in the CPU test tier (tests/cpu_emul/eco_abi_emul.cpp:
// the `-m gpu` test file of these entry points then runs against the kernel sources on the CPU).  Device code is not affected.
#pragma once

#ifdef B200_CPU_EMUL

#define B200_LAUNCH_KERNEL(kern, gx, gy, block, smem, st, ...) \
    ::cpu_emul::launch_blocks(kern, (unsigned)(gx), (unsigned)(gy), 1u, (unsigned)(block), (size_t)(smem), __VA_ARGS__)

template <class K, class P>
static int b200_launch_cooperative(K kern, int grid, int block, size_t smem, cudaStream_t, P& params) {
    ::cpu_emul::launch_coop(kern, (unsigned)grid, (unsigned)block, smem, params);
    return 0;
}

#else

#define B200_LAUNCH_KERNEL(kern, gx, gy, block, smem, st, ...) kern<<<dim3((unsigned)(gx), (unsigned)(gy)), (unsigned)(block), (smem), (st)>>>(__VA_ARGS__)

// one cooperative launch of a kernel that takes its parameter block by value; the dynamic shared-memory limit raised to what the plan asks for
template <class K, class P>
static int b200_launch_cooperative(K kern, int grid, int block, size_t smem, cudaStream_t st, P& params) {
    B200_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&params};
    B200_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3((unsigned)grid), dim3((unsigned)block), args, smem, st));
    return 0;
}

#endif
