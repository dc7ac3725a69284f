// ring_bw.cu -- microbenchmark: how fast can one B200 stream a batch of 4000x3000 RGBA
// images with the access pattern of k_stream (512-column slabs x row bands)?
//   mode 0: plain grid-stride LDG.128 read (sum) -- upper bound for reads
//   mode 1: slab/band pattern, direct LDG.128 per row with U rows in flight per thread
//   mode 2: slab/band pattern, TMA 1-D bulk ring (R rows per transaction, S stages)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ring_bw ring_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(b)), "r"(ph), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void tma1d(void *dst, const void *src, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(n), "r"(smem_u32(b)) : "memory");
}

__global__ void k_linear(const uint4 *__restrict__ p, size_t n, uint32_t *out)
{
    uint32_t acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldcs(p + i);
        acc += v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}

struct Geo { int W, H, pitch, tiles, bands, tile_w, rows_per_band; };

template <int U>
__global__ void __launch_bounds__(128) k_slab_ldg(const uint8_t *__restrict__ base, size_t img_bytes, Geo g, uint32_t *out)
{
    const int item = blockIdx.x;
    const int img = item / (g.tiles * g.bands), r = item % (g.tiles * g.bands);
    const int band = r / g.tiles, tile = r % g.tiles;
    const int c = tile * g.tile_w + threadIdx.x * 4;
    const int y0 = band * g.rows_per_band, y1 = min(g.H, y0 + g.rows_per_band);
    const uint8_t *p = base + img * img_bytes + (size_t)y0 * g.pitch + (size_t)c * 4;
    uint32_t acc = 0;
    if (c + 4 <= g.W) {
        for (int y = y0; y < y1; y += U) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = (y + u < y1) ? __ldcs((const uint4 *)(p + (size_t)u * g.pitch)) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < U; u++) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            p += (size_t)U * g.pitch;
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// R rows per transaction group, S stages.  128 consumer threads + 1 producer warp.
template <int R, int S, bool WRITE>
__global__ void __launch_bounds__(160) k_slab_tma(const uint8_t *__restrict__ base, uint8_t *__restrict__ dstbase, size_t img_bytes, Geo g, uint32_t *out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint4 (*ring)[R][128] = reinterpret_cast<uint4 (*)[R][128]>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)S * R * 2048);
    uint64_t *empty = full + S;
    const int item = blockIdx.x;
    const int img = item / (g.tiles * g.bands), r = item % (g.tiles * g.bands);
    const int band = r / g.tiles, tile = r % g.tiles;
    const int cx0 = tile * g.tile_w;
    const int y0 = band * g.rows_per_band, y1 = min(g.H, y0 + g.rows_per_band);
    const int ngroups = (y1 - y0 + R - 1) / R;
    const uint32_t row_bytes = (uint32_t)(min(512, g.W - cx0) * 4);
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 128) {
        if (threadIdx.x == 128) {
            const uint8_t *p = base + img * img_bytes + (size_t)y0 * g.pitch + (size_t)cx0 * 4;
            int s = 0; uint32_t ph = 1;
            for (int i = 0; i < ngroups; i++) {
                if (i >= S) mbar_wait(&empty[s], ph);
                const int nr = min(R, y1 - y0 - i * R);
                mbar_expect(&full[s], row_bytes * nr);
                for (int k = 0; k < nr; k++) tma1d(&ring[s][k][0], p + (size_t)(i * R + k) * g.pitch, row_bytes, &full[s]);
                if (++s == S) { s = 0; ph ^= 1; }
            }
        }
        return;
    }
    uint32_t acc = 0;
    int s = 0; uint32_t ph = 0;
    const int c = cx0 + threadIdx.x * 4;
    uint8_t *d = dstbase + img * img_bytes + (size_t)y0 * g.pitch + (size_t)c * 4;
    for (int i = 0; i < ngroups; i++) {
        mbar_wait(&full[s], ph);
        const int nr = min(R, y1 - y0 - i * R);
        uint4 v[R];
#pragma unroll
        for (int k = 0; k < R; k++) v[k] = ring[s][k][threadIdx.x];
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
#pragma unroll
        for (int k = 0; k < R; k++) {
            if (k < nr) {
                acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
                if (WRITE && c + 4 <= g.W && threadIdx.x * 4 < g.tile_w) __stcs((uint4 *)(d + (size_t)(i * R + k) * g.pitch), v[k]);
            }
        }
        if (++s == S) { s = 0; ph ^= 1; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F> float timeit(F f, int iters = 5)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < iters; i++) f();
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

template <int R, int S, bool WRITE>
void run_tma(const char *name, const uint8_t *src, uint8_t *dst, size_t img_bytes, Geo g, int nimg, uint32_t *out)
{
    size_t sm = (size_t)S * R * 2048 + 2 * S * 8 + 64;
    CK(cudaFuncSetAttribute(k_slab_tma<R, S, WRITE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_slab_tma<R, S, WRITE>, 160, sm);
    float ms = timeit([&] { k_slab_tma<R, S, WRITE><<<nimg * g.tiles * g.bands, 160, sm>>>(src, dst, img_bytes, g, out); });
    double bytes = (double)nimg * g.W * g.H * 4 * (WRITE ? 2 : 1);
    printf("%-28s R=%d S=%2d ctas/sm=%d  %8.3f ms  %7.1f GB/s  (%.1f us/img)\n", name, R, S, nb, ms, bytes / ms / 1e6, 1e3 * ms / nimg);
}

int main(int argc, char **argv)
{
    const int nimg = argc > 1 ? atoi(argv[1]) : 32;
    Geo g; g.W = 4000; g.H = 3000; g.pitch = 16000; g.tile_w = 448; g.tiles = 9;
    g.bands = argc > 2 ? atoi(argv[2]) : 6; g.rows_per_band = (g.H + g.bands - 1) / g.bands;
    size_t img_bytes = (size_t)g.pitch * g.H;
    uint8_t *src, *dst; uint32_t *out;
    CK(cudaMalloc(&src, img_bytes * nimg)); CK(cudaMalloc(&dst, img_bytes * nimg)); CK(cudaMalloc(&out, 64));
    CK(cudaMemset(src, 1, img_bytes * nimg));
    double bytes = (double)nimg * g.W * g.H * 4;
    {
        float ms = timeit([&] { k_linear<<<148 * 8, 256>>>((const uint4 *)src, img_bytes * nimg / 16, out); });
        printf("%-28s %8.3f ms  %7.1f GB/s\n", "linear LDG.128 read", ms, bytes / ms / 1e6);
        ms = timeit([&] { cudaMemcpyAsync(dst, src, img_bytes * nimg, cudaMemcpyDeviceToDevice); });
        printf("%-28s %8.3f ms  %7.1f GB/s (R+W)\n", "cudaMemcpy D2D", ms, 2 * bytes / ms / 1e6);
    }
    {
        int items = nimg * g.tiles * g.bands;
        float ms = timeit([&] { k_slab_ldg<1><<<items, 128>>>(src, img_bytes, g, out); });
        printf("%-28s U=1  %8.3f ms  %7.1f GB/s\n", "slab LDG", ms, bytes / ms / 1e6);
        ms = timeit([&] { k_slab_ldg<4><<<items, 128>>>(src, img_bytes, g, out); });
        printf("%-28s U=4  %8.3f ms  %7.1f GB/s\n", "slab LDG", ms, bytes / ms / 1e6);
        ms = timeit([&] { k_slab_ldg<8><<<items, 128>>>(src, img_bytes, g, out); });
        printf("%-28s U=8  %8.3f ms  %7.1f GB/s\n", "slab LDG", ms, bytes / ms / 1e6);
    }
    run_tma<1, 8, false>("slab TMA read", src, dst, img_bytes, g, nimg, out);
    run_tma<1, 16, false>("slab TMA read", src, dst, img_bytes, g, nimg, out);
    run_tma<2, 8, false>("slab TMA read", src, dst, img_bytes, g, nimg, out);
    run_tma<4, 4, false>("slab TMA read", src, dst, img_bytes, g, nimg, out);
    run_tma<4, 8, false>("slab TMA read", src, dst, img_bytes, g, nimg, out);
    run_tma<8, 4, false>("slab TMA read", src, dst, img_bytes, g, nimg, out);
    run_tma<1, 8, true>("slab TMA copy", src, dst, img_bytes, g, nimg, out);
    run_tma<4, 4, true>("slab TMA copy", src, dst, img_bytes, g, nimg, out);
    run_tma<4, 8, true>("slab TMA copy", src, dst, img_bytes, g, nimg, out);
    return 0;
}
