// Probe: under which launch shapes do two kernels on different streams co-reside on B200 SMs?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/overlap_probe tools/overlap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ unsigned char dyn[];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int ID>
__global__ void spin(unsigned long long ns, int *sink) {
    const unsigned long long t0 = gtime();
    dyn[threadIdx.x] = (unsigned char)threadIdx.x;
    while (gtime() - t0 < ns) { }
    if (dyn[threadIdx.x] == 77 && ns == 1) sink[0] = 1;
}
static float run(int ctasA, int thrA, size_t smA, int ctasB, int thrB, size_t smB, bool bfirst, int carve, cudaStream_t s1, cudaStream_t s2, int *sink) {
    cudaFuncSetAttribute(spin<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
    cudaFuncSetAttribute(spin<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 222 * 1024);
    cudaFuncSetAttribute(spin<0>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaFuncSetAttribute(spin<1>, cudaFuncAttributePreferredSharedMemoryCarveout, carve);
    cudaEvent_t e0, e1, ef, ej;
    cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&ef); cudaEventCreate(&ej);
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1);
    cudaEventRecord(ef, s1);
    cudaStreamWaitEvent(s2, ef, 0);
    if (bfirst) { if (ctasB) spin<1><<<ctasB, thrB, smB, s2>>>(1000000ull, sink); if (ctasA) spin<0><<<ctasA, thrA, smA, s1>>>(1000000ull, sink); }
    else { if (ctasA) spin<0><<<ctasA, thrA, smA, s1>>>(1000000ull, sink); if (ctasB) spin<1><<<ctasB, thrB, smB, s2>>>(1000000ull, sink); }
    cudaEventRecord(ej, s2);
    cudaStreamWaitEvent(s1, ej, 0);
    cudaEventRecord(e1, s1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
    return ms;
}
int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int *sink; cudaMalloc(&sink, 4);
    cudaStream_t s1, s2;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    int sm = 0; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d; each kernel spins 1.0 ms; ~1 ms = overlapped, ~2 ms = serialised\n", sm);
    const int carves[3] = {-1, 100, 50};
    for (int ci = 0; ci < 3; ++ci) {
        const int cv = carves[ci];
        printf("carveout %d\n", cv);
        printf("  A alone 3/SM x 64K            : %.3f\n", run(sm * 3, 256, 65536, 0, 0, 0, false, cv, s1, s2, sink));
        printf("  A 3/SM x 64K + B 1/SM x 216K  : %.3f\n", run(sm * 3, 256, 65536, sm, 256, 216 * 1024, false, cv, s1, s2, sink));
        printf("  A 2/SM x 64K + B 1/SM x 81K   : %.3f\n", run(sm * 2, 256, 65536, sm, 96, 81 * 1024, false, cv, s1, s2, sink));
        printf("  same, B first                 : %.3f\n", run(sm * 2, 256, 65536, sm, 96, 81 * 1024, true, cv, s1, s2, sink));
        printf("  A 2/SM x 64K + B 1/SM x 27K   : %.3f\n", run(sm * 2, 256, 65536, sm, 32, 27 * 1024, false, cv, s1, s2, sink));
        printf("  A 3/SM x 64K + B 1/SM x 27K   : %.3f\n", run(sm * 3, 256, 65536, sm, 32, 27 * 1024, false, cv, s1, s2, sink));
        printf("  A 3/SM x 40K + B 1/SM x 81K   : %.3f\n", run(sm * 3, 256, 40960, sm, 96, 81 * 1024, false, cv, s1, s2, sink));
        printf("  A 1/SM x 64K + B 20 x 216K    : %.3f\n", run(sm, 256, 65536, 20, 256, 216 * 1024, false, cv, s1, s2, sink));
        printf("  A 3/SM x 64K (128 SMs) + B 20 x 216K : %.3f\n", run(128 * 3, 256, 65536, 20, 256, 216 * 1024, true, cv, s1, s2, sink));
        printf("  A 3/SM x 48K + B 1/SM x 81K   : %.3f\n", run(sm * 3, 256, 48 * 1024, sm, 96, 81 * 1024, false, cv, s1, s2, sink));
        printf("  A 2/SM x 64K + B 2/SM x 40K   : %.3f\n", run(sm * 2, 256, 65536, sm * 2, 96, 40 * 1024, false, cv, s1, s2, sink));
    }
    return 0;
}
