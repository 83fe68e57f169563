// Where do the ~6 us between dependent graph kernel nodes go?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a launch_gap.cu -o launch_gap
#include <cstdio>
#include <cuda_runtime.h>
struct Big { char b[2048]; };
__global__ void k_empty(int* p) { if (p && threadIdx.x == 1000) p[0] = 1; }
__global__ void k_big(const __grid_constant__ Big b, int* p) { if (p && threadIdx.x == 1000) p[0] = b.b[5]; }
__global__ void __launch_bounds__(192, 2) k_tmem(int* p, int cols) {
    __shared__ unsigned slot;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)), "r"((unsigned)cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"((unsigned)cols) : "memory");
    if (p && threadIdx.x == 1000) p[0] = 1;
}
__global__ void k_spin(int* p, long long cycles) {   // every CTA runs `cycles` clocks, then writes 64 KB to global memory
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) {}
    if (p) for (int i = threadIdx.x; i < 16384; i += blockDim.x) p[(size_t)blockIdx.x * 16384 + i] = i;
}
template <class F> float time_graph(F launch, int reps = 20) {
    cudaStream_t s; cudaStreamCreate(&s);
    launch(s); cudaStreamSynchronize(s);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < reps; ++i) launch(s);
    cudaStreamEndCapture(s, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s); cudaGraphLaunch(ge, s); cudaEventRecord(e1, s); cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / reps;
}
int main() {
    int* d; cudaMalloc(&d, 1024 * 65536);
    cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    printf("empty <<<1,32>>>                     %6.2f us\n", time_graph([&](cudaStream_t s) { k_empty<<<1, 32, 0, s>>>(d); }));
    printf("empty <<<320,192,0>>>                %6.2f us\n", time_graph([&](cudaStream_t s) { k_empty<<<320, 192, 0, s>>>(d); }));
    printf("empty <<<320,192,100KB>>>            %6.2f us\n", time_graph([&](cudaStream_t s) { k_empty<<<320, 192, 100 * 1024, s>>>(d); }));
    printf("empty <<<1035,192,100KB>>>           %6.2f us\n", time_graph([&](cudaStream_t s) { k_empty<<<1035, 192, 100 * 1024, s>>>(d); }));
    printf("2KB params <<<320,192,100KB>>>       %6.2f us\n", time_graph([&](cudaStream_t s) { Big b{}; k_big<<<320, 192, 100 * 1024, s>>>(b, d); }));
    printf("cluster(2) empty <<<320,192,100KB>>> %6.2f us\n", time_graph([&](cudaStream_t s) {
        cudaLaunchConfig_t c{}; c.gridDim = dim3(320); c.blockDim = dim3(192); c.dynamicSmemBytes = 100 * 1024; c.stream = s;
        cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = 2; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        c.attrs = a; c.numAttrs = 1; cudaLaunchKernelEx(&c, k_empty, d); }));
    printf("tmem alloc 128 <<<320,192,100KB>>>   %6.2f us\n", time_graph([&](cudaStream_t s) { k_tmem<<<320, 192, 100 * 1024, s>>>(d, 128); }));
    printf("tmem alloc 256 <<<320,192,100KB>>>   %6.2f us\n", time_graph([&](cudaStream_t s) { k_tmem<<<320, 192, 100 * 1024, s>>>(d, 256); }));
    for (long long cyc : {0LL, 10000LL, 20000LL, 40000LL}) {
        printf("spin %6lld clk <<<296,192,100KB>>> + 64KB stores  %6.2f us (spin alone = %.2f us @1.965GHz)\n", cyc,
               time_graph([&](cudaStream_t s) { k_spin<<<296, 192, 100 * 1024, s>>>(d, cyc); }), cyc / 1965.0);
    }
    printf("spin 20000 clk <<<592,192,100KB>>> (2 waves)          %6.2f us\n", time_graph([&](cudaStream_t s) { k_spin<<<592, 192, 100 * 1024, s>>>(d, 20000); }));
    return 0;
}
