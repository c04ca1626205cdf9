// Developer micro-benchmark: what does a kernel launch that only READS a 2-60 MB matrix from HBM cost on this GPU, back to back in a
// CUDA graph (the GEMV's speed of light including launch boundaries)?  nvcc -O3 -gencode arch=compute_100a,code=sm_100a read_bw.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct u32x8 { uint32_t v[8]; };
__device__ __forceinline__ u32x8 ldg256(const void* p)
{
    u32x8 r;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}

// MODE 0: grid-stride 256-bit loads, U in flight per thread.  MODE 1: CTA-contiguous slices (each CTA owns bytes/grid), U in flight.
// PF: the CTA first queues its whole slice with TMA bulk L2 prefetches (MODE 1 only).
template <int MODE, int U, bool PF, bool PDL>
__global__ void __launch_bounds__(512) reader(const uint8_t* p, size_t bytes, unsigned* out)
{
    if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    unsigned acc = 0;
    const size_t n32 = bytes / 32;
    if (MODE == 0) {
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
        for (; i + (U - 1) * stride < n32; i += U * stride) {
            u32x8 r[U];
#pragma unroll
            for (int u = 0; u < U; u++) r[u] = ldg256(p + (i + u * stride) * 32);
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc ^= r[u].v[j];
        }
        for (; i < n32; i += stride) {
            u32x8 r = ldg256(p + i * 32);
#pragma unroll
            for (int j = 0; j < 8; j++) acc ^= r.v[j];
        }
    } else {
        const size_t per = (n32 + gridDim.x - 1) / gridDim.x;
        const size_t lo = per * blockIdx.x, hi = lo + per < n32 ? lo + per : n32;
        if (PF && threadIdx.x < 32) {
            for (size_t o = lo * 32 + (size_t)threadIdx.x * 8192; o < hi * 32; o += 32 * 8192) {
                unsigned n = (unsigned)(hi * 32 - o < 8192 ? hi * 32 - o : 8192) & ~15u;
                if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + o), "r"(n) : "memory");
            }
        }
        size_t i = lo + threadIdx.x;
        for (; i + (U - 1) * blockDim.x < hi; i += U * blockDim.x) {
            u32x8 r[U];
#pragma unroll
            for (int u = 0; u < U; u++) r[u] = ldg256(p + (i + u * blockDim.x) * 32);
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int j = 0; j < 8; j++) acc ^= r[u].v[j];
        }
        for (; i < hi; i += blockDim.x) {
            u32x8 r = ldg256(p + i * 32);
#pragma unroll
            for (int j = 0; j < 8; j++) acc ^= r.v[j];
        }
    }
    if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (acc == 0x12345678) out[0] = acc;
}

template <typename K>
float run_graph(K kern, int grid, int threads, bool pdl, const std::vector<uint8_t*>& bufs, size_t bytes, unsigned* out, int replays)
{
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal));
    for (size_t i = 0; i < bufs.size(); i++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(threads);
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = pdl ? 1 : 0;
        const uint8_t* p = bufs[i];
        CK(cudaLaunchKernelEx(&cfg, kern, p, bytes, out));
    }
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    for (int i = 0; i < 2; i++) CK(cudaGraphLaunch(ge, s));
    CK(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    CK(cudaEventRecord(e0, s));
    for (int i = 0; i < replays; i++) CK(cudaGraphLaunch(ge, s));
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaGraphExecDestroy(ge);
    cudaGraphDestroy(g);
    cudaStreamDestroy(s);
    return ms * 1e3f / (replays * bufs.size());
}

int main()
{
    unsigned* out;
    CK(cudaMalloc(&out, 4));
    const size_t sizes[] = {2097152 + 32768, 8388608 + 131072, 12582912 + 196608, 29360128 + 458752, 58720256 + 917504};
    for (size_t bytes : sizes) {
        const int nbuf = (int)((1ull << 30) / bytes) > 64 ? 64 : (int)((1ull << 30) / bytes);
        std::vector<uint8_t*> bufs(nbuf);
        for (auto& b : bufs) {
            CK(cudaMalloc(&b, bytes));
            CK(cudaMemset(b, 1, bytes));
        }
        printf("%.1f MB x %d buffers: ideal at 6531.6 GB/s = %.2f us\n", bytes / 1e6, nbuf, bytes / 6531.6e3);
#define RUN(name, kern, grid, pdl)                                                                      \
    {                                                                                                   \
        float t = run_graph(kern, grid, 512, pdl, bufs, bytes, out, 20);                                \
        printf("   %-58s %7.2f us/launch  %7.1f GB/s\n", name, t, bytes / t / 1e3);                      \
    }
        RUN("grid-stride U4, 148 CTAs", (reader<0, 4, false, false>), 148, false);
        RUN("grid-stride U4, 296 CTAs", (reader<0, 4, false, false>), 296, false);
        RUN("grid-stride U4, 592 CTAs", (reader<0, 4, false, false>), 592, false);
        RUN("grid-stride U8, 148 CTAs", (reader<0, 8, false, false>), 148, false);
        RUN("grid-stride U8, 296 CTAs", (reader<0, 8, false, false>), 296, false);
        RUN("grid-stride U4, 296 CTAs, PDL", (reader<0, 4, false, true>), 296, true);
        RUN("grid-stride U8, 148 CTAs, PDL", (reader<0, 8, false, true>), 148, true);
        RUN("CTA slices U4, 148 CTAs", (reader<1, 4, false, false>), 148, false);
        RUN("CTA slices U4, 148 CTAs + bulk L2 prefetch", (reader<1, 4, true, false>), 148, false);
        RUN("CTA slices U4, 148 CTAs + bulk L2 prefetch, PDL", (reader<1, 4, true, true>), 148, true);
        RUN("CTA slices U2, 148 CTAs + bulk L2 prefetch, PDL", (reader<1, 2, true, true>), 148, true);
        RUN("CTA slices U8, 148 CTAs, PDL", (reader<1, 8, false, true>), 148, true);
        for (auto b : bufs) cudaFree(b);
    }
    return 0;
}
