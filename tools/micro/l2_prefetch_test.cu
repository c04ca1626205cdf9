// Developer micro-benchmark: does cp.async.bulk.prefetch.L2 (UBLKPF.L2) / prefetch.global.L2 actually land data in L2?
// Kernel P prefetches a buffer, kernel R reads it (sum) and is timed; compared with a cold read and a warm (just read) read.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void reader(const uint4* p, size_t n16, unsigned* out)
{
    unsigned acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678) out[0] = acc;
}
__global__ void prefetch_bulk(const uint8_t* p, size_t bytes, int piece)
{
    size_t per = (bytes / gridDim.x + 15) & ~(size_t)15;
    size_t lo = per * blockIdx.x, hi = lo + per < bytes ? lo + per : bytes;
    for (size_t o = lo + (size_t)threadIdx.x * piece; o < hi; o += (size_t)blockDim.x * piece) {
        unsigned n = (unsigned)(hi - o < (size_t)piece ? hi - o : piece) & ~15u;
        if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + o), "r"(n) : "memory");
    }
}
__global__ void prefetch_line(const uint8_t* p, size_t bytes)
{
    for (size_t o = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 128; o < bytes; o += (size_t)gridDim.x * blockDim.x * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
}
__global__ void flusher(uint4* p, size_t n16) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(1, 2, 3, 4); }

int main()
{
    const size_t bytes = 30u << 20, fl = 512u << 20;
    uint8_t *buf, *junk; unsigned* out;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&junk, fl)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(buf, 1, bytes));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto flush = [&]() { flusher<<<1184, 256>>>((uint4*)junk, fl / 16); cudaDeviceSynchronize(); };
    auto read_time = [&]() { float ms; cudaEventRecord(e0); reader<<<148 * 4, 512>>>((const uint4*)buf, bytes / 16, out); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); return ms * 1e3f; };
    for (int rep = 0; rep < 2; rep++) {
        flush(); float cold = read_time(); float warm = read_time();
        printf("read 30 MB: cold %.2f us (%.0f GB/s)  warm(L2) %.2f us (%.0f GB/s)\n", cold, bytes / cold / 1e3, warm, bytes / warm / 1e3);
        for (int piece : {128, 1024, 8192, 65536}) {
            flush();
            float ms; cudaEventRecord(e0); prefetch_bulk<<<148, 32>>>(buf, bytes, piece); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            cudaDeviceSynchronize();
            // give the prefetches time to land
            for (volatile int spin = 0; spin < 200000; spin++) {}
            float t = read_time();
            printf("  after bulk prefetch piece %6d (issue kernel %.2f us): read %.2f us (%.0f GB/s)\n", piece, ms * 1e3, t, bytes / t / 1e3);
        }
        flush();
        prefetch_line<<<148 * 2, 256>>>(buf, bytes); cudaDeviceSynchronize();
        float t = read_time();
        printf("  after prefetch.global.L2 per 128 B line: read %.2f us (%.0f GB/s)\n", t, bytes / t / 1e3);
    }
    return 0;
}
