// pcie_bw.cu -- measured PCIe rooflines on this box: pinned H2D, D2H, and both at once.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)
int main()
{
    const size_t N = 1ull << 30;
    void *h0, *h1, *d0, *d1;
    CK(cudaHostAlloc(&h0, N, cudaHostAllocPortable)); CK(cudaHostAlloc(&h1, N, cudaHostAllocPortable));
    CK(cudaMalloc(&d0, N)); CK(cudaMalloc(&d1, N));
    cudaStream_t s0, s1; cudaStreamCreate(&s0); cudaStreamCreate(&s1);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto run = [&](const char *name, bool up, bool down, size_t chunk) {
        for (int rep = 0; rep < 2; rep++) {
            cudaDeviceSynchronize();
            cudaEventRecord(a, 0);
            cudaStreamWaitEvent(s0, a, 0); cudaStreamWaitEvent(s1, a, 0);
            for (size_t o = 0; o < N; o += chunk) {
                if (up) cudaMemcpyAsync((char *)d0 + o, (char *)h0 + o, chunk, cudaMemcpyHostToDevice, s0);
                if (down) cudaMemcpyAsync((char *)h1 + o, (char *)d1 + o, chunk, cudaMemcpyDeviceToHost, s1);
            }
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, s0); cudaEventRecord(e1, s1);
            cudaStreamWaitEvent(0, e0, 0); cudaStreamWaitEvent(0, e1, 0);
            cudaEventRecord(b, 0); CK(cudaEventSynchronize(b));
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (rep) printf("%-26s chunk %4zu MB: %6.1f GB/s per direction\n", name, chunk >> 20, N / ms / 1e6);
        }
    };
    run("H2D only", true, false, 48u << 20);
    run("D2H only", false, true, 48u << 20);
    run("H2D + D2H concurrently", true, true, 48u << 20);
    run("H2D + D2H concurrently", true, true, 256u << 20);
    return 0;
}
