// Developer micro-benchmark: what does ONE SM sustain on the decode GEMV's core instruction mix -- per 2-KB sub-tile and warp
// 64 x (PRMT + LDS.32 table lookup) + 16 x mma.sync.m16n8k16 (+ 4 x LDS.128 for the packed bytes) -- with 16 warps and nothing else
// in the loop?  The distance between this number and the kernel's 132 clk / sub-tile / SM is what trimming its bookkeeping can win.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a mix_rate.cu -o mix_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void hmma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int SEL> __device__ __forceinline__ uint32_t lookup(uint32_t w, uint32_t lane_base)
{
    uint32_t v;
    const uint32_t addr = __byte_perm(w, lane_base, 0x7604 | (SEL << 4));
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// MODE bit 0: 4 x LDS.128 per sub-tile (packed bytes from a 2-KB shared-memory tile); bit 1: no HMMA (xor instead); bit 2: no lookups
template <int MODE, int AHEAD, int U = 1>
__global__ void __launch_bounds__(512) k(unsigned* out, long long* cyc, int subtiles, unsigned seed)
{
    extern __shared__ __align__(1024) uint8_t sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (65536 + 16 * 2048) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = (i * 2654435761u) ^ seed;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t lane_base = base + lane * 4;
    const uint32_t tile = base + 65536 + warp * 2048 + lane * 16;
    uint32_t wa[8], wb[8], xr[32];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        wa[i] = (tid * 31 + i) * 2654435761u ^ seed;
        wb[i] = (tid * 17 + i) * 2246822519u ^ seed;
    }
#pragma unroll
    for (int i = 0; i < 32; i++) xr[i] = 0x3f803f80u + i;
    uint32_t f[AHEAD + 1][4];
    auto fetch = [&](uint32_t (&d)[4], int j) {
        const uint32_t va = wa[j >> 1], vb = wb[j >> 1];
        if (MODE & 4) {
            d[0] = va; d[1] = vb; d[2] = va >> 1; d[3] = vb >> 1;
        } else if (j & 1) {
            d[0] = lookup<2>(va, lane_base); d[1] = lookup<2>(vb, lane_base); d[2] = lookup<3>(va, lane_base); d[3] = lookup<3>(vb, lane_base);
        } else {
            d[0] = lookup<0>(va, lane_base); d[1] = lookup<0>(vb, lane_base); d[2] = lookup<1>(va, lane_base); d[3] = lookup<1>(vb, lane_base);
        }
    };
#pragma unroll
    for (int j = 0; j < AHEAD; j++) fetch(f[j], j);
    float acc = 0.0f;
    const long long t0 = clock64();
    for (int s0 = 0; s0 < subtiles; s0 += U) {
#pragma unroll
      for (int us = 0; us < U; us++) {
        const int s = s0 + us;
        float ce[4] = {0, 0, 0, 0}, co[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 16; j++) {
            fetch(f[(j + AHEAD) % (AHEAD + 1)], (j + AHEAD) % 16);
            uint32_t(&a4)[4] = f[j % (AHEAD + 1)];
            if (MODE & 2) {
                ce[j & 3] = __uint_as_float((__float_as_uint(ce[j & 3]) ^ a4[0] ^ a4[1] ^ a4[2] ^ a4[3] ^ xr[2 * j]) & 0x3fffffffu);
            } else if (j & 1) hmma(co, a4[0], a4[1], a4[2], a4[3], xr[2 * j], xr[2 * j + 1]);
            else hmma(ce, a4[0], a4[1], a4[2], a4[3], xr[2 * j], xr[2 * j + 1]);
            if (MODE & 1) {
                if (j == 8) {
                    uint4 a0, b0;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w) : "r"(tile));
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b0.x), "=r"(b0.y), "=r"(b0.z), "=r"(b0.w) : "r"(tile + 512));
                    wa[0] = a0.x ^ s; wa[1] = a0.y; wa[2] = a0.z; wa[3] = a0.w; wb[0] = b0.x ^ s; wb[1] = b0.y; wb[2] = b0.z; wb[3] = b0.w;
                }
                if (j == 12) {
                    uint4 a0, b0;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w) : "r"(tile + 1024));
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b0.x), "=r"(b0.y), "=r"(b0.z), "=r"(b0.w) : "r"(tile + 1536));
                    wa[4] = a0.x ^ s; wa[5] = a0.y; wa[6] = a0.z; wa[7] = a0.w; wb[4] = b0.x ^ s; wb[5] = b0.y; wb[6] = b0.z; wb[7] = b0.w;
                }
            }
        }
        acc += ce[0] + co[0] + ce[1] + co[1] + ce[2] + co[2] + ce[3] + co[3];
      }
    }
    const long long t1 = clock64();
    uint32_t r = __float_as_uint(acc);
#pragma unroll
    for (int j = 0; j < AHEAD; j++) r ^= f[j][0] ^ f[j][1] ^ f[j][2] ^ f[j][3];
    out[blockIdx.x * blockDim.x + tid] = r;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int AHEAD, int U = 1>
static int run(const char* name, int threads)
{
    unsigned* out;
    long long* cyc;
    CK(cudaMalloc(&out, 148 * 512 * 4));
    CK(cudaMalloc(&cyc, 148 * 8));
    const int smem = 65536 + 16 * 2048;
    CK(cudaFuncSetAttribute(k<MODE, AHEAD, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int subtiles = 400;
    k<MODE, AHEAD, U><<<148, threads, smem>>>(out, cyc, subtiles, 1);
    k<MODE, AHEAD, U><<<148, threads, smem>>>(out, cyc, subtiles, 2);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double per = avg / ((double)subtiles * (threads / 32));
    printf("%-44s warps %2d ahead %d: %6.1f clk per sub-tile per SM  -> %5.1f B/clk/SM, %.2f TB/s at 1.965 GHz x 148\n", name, threads / 32, AHEAD, per,
           2048.0 / per, 2048.0 / per * 1.965e9 * 148 / 1e12);
    cudaFree(out);
    cudaFree(cyc);
    return 0;
}

int main()
{
    run<1, 3, 1>("lookups + HMMA + packed, body x1 (~2.5 KB)", 512);
    run<1, 3, 2>("lookups + HMMA + packed, body x2 (~5 KB)", 512);
    run<1, 3, 4>("lookups + HMMA + packed, body x4 (~10 KB)", 512);
    run<1, 3, 8>("lookups + HMMA + packed, body x8 (~20 KB)", 512);
    run<1, 3, 16>("lookups + HMMA + packed, body x16 (~40 KB)", 512);
    for (int threads : {256, 384, 512}) {  // 8 / 12 / 16 warps: can fewer, fatter warps keep the shared-memory pipe as busy?
        run<0, 3>("lookups + HMMA", threads);
        run<1, 3>("lookups + HMMA + packed LDS.128", threads);
        run<3, 3>("lookups + packed LDS.128, no HMMA", threads);
        run<5, 3>("HMMA + packed LDS.128, no lookups", threads);
        run<1, 2>("lookups + HMMA + packed LDS.128", threads);
        run<1, 5>("lookups + HMMA + packed LDS.128", threads);
    }
    return 0;
}
