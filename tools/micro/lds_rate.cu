// Developer micro-benchmark: shared-memory lookup INSTRUCTION rate per SM (is a conflict-free LDS.32 one instruction per clock, or is
// there an LSU issue floor?), the same for LDS.64 / LDS.128, for PRMT alone and for the PRMT + LDS pair of the decode GEMV's table lookup.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a lds_rate.cu -o lds_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// MODE 0: LDS.32  1: LDS.64  2: LDS.128  3: PRMT only  4: PRMT + LDS.32 (address depends on the previous result: latency chain per slot,
// ILP slots per thread)  5: LDS.32 with a fixed address per thread (no address arithmetic at all)
template <int MODE, int ILP>
__global__ void __launch_bounds__(1024) k(unsigned* out, long long* cyc, int iters, unsigned seed)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = (i * 2654435761u) ^ seed;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t v[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) v[j] = (tid * 7 + j * 13) * 2654435761u ^ seed;
    const uint32_t lane_base = base + lane * 4;           // MODE 0/4/5: row = byte, word = lane (conflict-free)
    const uint32_t lane_base8 = base + (lane & 15) * 8;    // LDS.64: 16 lanes x 8 B per phase
    const uint32_t lane_base16 = base + (lane & 7) * 16;   // LDS.128: 8 lanes x 16 B per phase
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            if (MODE == 0 || MODE == 4) {
                const uint32_t addr = __byte_perm(v[j], lane_base, 0x7604);  // byte 0 of v -> bits 8..15
                uint32_t r;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
                v[j] = MODE == 4 ? r : (v[j] + r);
            } else if (MODE == 5) {
                uint32_t r;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(lane_base + (uint32_t)j * 256));
                v[j] ^= r;
            } else if (MODE == 1) {
                const uint32_t addr = lane_base8 + ((v[j] & 0xffu) << 7);  // 128-byte rows, 512 rows
                uint32_t r0, r1;
                asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
                v[j] += r0 ^ r1;
            } else if (MODE == 2) {
                const uint32_t addr = lane_base16 + ((v[j] & 0xffu) << 7);
                uint32_t r0, r1, r2, r3;
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
                v[j] += r0 ^ r1 ^ r2 ^ r3;
            } else {
                v[j] = __byte_perm(v[j], lane_base + it, 0x7604 ^ (j & 1));
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) acc ^= v[j];
    out[blockIdx.x * blockDim.x + tid] = acc;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int ILP>
static int run(const char* name, int threads)
{
    unsigned* out;
    long long* cyc;
    CK(cudaMalloc(&out, 148 * 1024 * 4));
    CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaFuncSetAttribute(k<MODE, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int iters = 2000;
    k<MODE, ILP><<<148, threads, 65536>>>(out, cyc, iters, 1);
    k<MODE, ILP><<<148, threads, 65536>>>(out, cyc, iters, 2);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double warp_insts = (double)iters * ILP * (threads / 32);
    printf("%-34s threads %4d ILP %2d: %8.0f clk, %.3f clk per warp-instruction per SM\n", name, threads, ILP, avg, avg / warp_insts);
    cudaFree(out);
    cudaFree(cyc);
    return 0;
}

int main()
{
    for (int threads : {256, 512, 1024}) {
        run<0, 8>("PRMT + LDS.32 (independent)", threads);
        run<4, 8>("PRMT + LDS.32 (chained)", threads);
        run<5, 8>("LDS.32 fixed address", threads);
        run<1, 8>("LDS.64", threads);
        run<2, 8>("LDS.128", threads);
        run<3, 8>("PRMT only", threads);
    }
    run<0, 16>("PRMT + LDS.32 (independent)", 512);
    run<5, 16>("LDS.32 fixed address", 512);
    return 0;
}
