// Translation unit of the tensor-core multi-query kernel (gsb_tensor.cuh), apart from gsb_api.cu so
// that the two compile in parallel.
#define GSB_TENSOR_IMPL
#include "gsb_tensor.cuh"

namespace gsb
{

cudaError_t tensor_kernel_launch(const TensorParams& tp, int grid, cudaStream_t st)
{
    void* kernel = reinterpret_cast<void*>(scan_tensor_kernel);
    const uint32_t smem = tc_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess)
        return e;
    TensorParams copy = tp;
    void* args[] = {&copy};
    return cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kTcThreads), args, smem, st);
}

} // namespace gsb
