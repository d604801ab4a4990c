// Does a 640-thread CTA get through setmaxnreg 48 / 56 / 160 (the tensor-core kernel's split)?  Prints "ok".
#include <cstdio>
#include <cuda_runtime.h>
template <int A, int B, int C>
__global__ void __launch_bounds__(640, 1) k(float* out, const float* in, int n)
{
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(A));
        out[threadIdx.x] = in[threadIdx.x];
    } else if (warp < 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(B));
        out[threadIdx.x] = in[threadIdx.x] * 2.0f;
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(C));
        float a[100];
#pragma unroll
        for (int i = 0; i < 100; i++)
            a[i] = in[i * 640 + threadIdx.x];
        for (int it = 0; it < n; it++) {
#pragma unroll
            for (int i = 0; i < 100; i++)
                a[i] = a[i] * a[(i + 7) % 100] + 1.0f;
        }
        float s = 0;
#pragma unroll
        for (int i = 0; i < 100; i++)
            s += a[i];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}
template <int A, int B, int C> int run(float* out, float* in)
{
    k<A, B, C><<<148, 640>>>(out, in, 10);
    cudaError_t e = cudaDeviceSynchronize();
    std::printf("setmaxnreg %d/%d/%d: %s\n", A, B, C, cudaGetErrorString(e));
    std::fflush(stdout);
    return e != cudaSuccess;
}
int main()
{
    float *in, *out;
    cudaMalloc(&in, 640 * 128 * 4);
    cudaMalloc(&out, 640 * 4);
    cudaMemset(in, 0, 640 * 128 * 4);
    int rc = 0;
    rc |= run<48, 56, 152>(out, in);
    rc |= run<48, 56, 160>(out, in);
    return rc;
}
