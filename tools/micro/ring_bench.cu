// Developer harness for the persistent ring GEMV (quantizations_b200/csrc/q4_gemv_ring.cuh), no Python / torch: builds synthetic
// NF4 + double-quant statistics for the Llama-3-8B layer shapes, checks the kernel against a naive CUDA GEMV of the same packed
// format, and times a decode token's worth of stages (L layers x qkv / o / gate-up / down, rotating over distinct weights > L2).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I quantizations_b200/csrc tools/micro/ring_bench.cu -o ring_bench
//   ring_bench [--nc 16] [--chain 4] [--layers 32] [--pool 6] [--iters 20] [--no-split] [--no-pdl] [--slots N] [--trace] [--check-only]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <cmath>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "q4_gemv_ring.cuh"
#include "q4_tma.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

using bf16 = __nv_bfloat16;
using namespace q4;

__global__ void fill_bytes(uint8_t* p, size_t n, uint32_t seed)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = (uint8_t)h;
    }
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float lo, float hi)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = lo + (hi - lo) * (float)(h >> 8) / 16777216.0f;
    }
}
__global__ void fill_bf16(bf16* p, size_t n, uint32_t seed)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed;
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        p[i] = __float2bfloat16(((float)(h >> 8) / 16777216.0f - 0.5f) * 4.0f);
    }
}

// naive reference: one warp per row, fp32
__global__ void ref_gemv(const bf16* x, const uint8_t* B, const uint8_t* qabs, const float* code2, const float* absmax2, float offset,
                         const float* code, const bf16* bias, float* out, int rows, int K)
{
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    float acc = 0.0f;
    const int bpr = K / 64;
    for (int b = lane; b < bpr; b += 32) {
        const size_t blk = (size_t)r * bpr + b;
        const float am = __fadd_rn(__fmul_rn(code2[qabs[blk]], absmax2[blk >> 8]), offset);
        float s = 0.0f;
        for (int i = 0; i < 32; i++) {
            const uint8_t byte = B[blk * 32 + i];
            const float c0 = __bfloat162float(__float2bfloat16(code[byte >> 4])), c1 = __bfloat162float(__float2bfloat16(code[byte & 15]));
            s += __bfloat162float(x[b * 64 + 2 * i]) * c0 + __bfloat162float(x[b * 64 + 2 * i + 1]) * c1;
        }
        acc += s * am;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = acc + (bias ? __bfloat162float(bias[r]) : 0.0f);
}

// h = silu(gate) * up with the staging glue's arithmetic (F.silu rounded to bf16, then the product rounded to bf16)
__global__ void swiglu_ref(const bf16* gate, const bf16* up, bf16* h, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float g = __bfloat162float(gate[i]), u = __bfloat162float(up[i]);
    const float sg = __bfloat162float(__float2bfloat16(__fdividef(g, 1.0f + __expf(-g))));
    h[i] = __float2bfloat16(sg * u);
}

struct Mat {
    int rows, K;
    uint8_t* B;
    uint8_t* qabs;
    float* absmax2;
    float* offset;  // device scalar
    bf16* out;
};

static const float kNf4[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f, -0.28444138169288635f,
                               -0.18477343022823334f, -0.09105003625154495f, 0.0f, 0.07958029955625534f, 0.16093020141124725f,
                               0.24611230194568634f, 0.33791524171829224f, 0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

template <int NC, int WPS>
static cudaError_t launch_nc(const ring::Args& a, int grid, size_t smem, bool pdl, cudaStream_t s)
{
    auto kern = ring::gemv_ring_kernel<bf16, true, NC, WPS>;
    static bool set = false;
    if (!set) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(ring::ring_threads(NC));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a);
}

static int g_wps = 2;
static cudaError_t launch(int nc, const ring::Args& a, int grid, size_t smem, bool pdl, cudaStream_t s)
{
    switch (nc * 10 + g_wps) {
        case 81: return launch_nc<8, 1>(a, grid, smem, pdl, s);
        case 162: return launch_nc<16, 2>(a, grid, smem, pdl, s);
        default: printf("unsupported --nc / --wps\n"); exit(1);
    }
}

int main(int argc, char** argv)
{
    int nc = 16, chain = 4, layers = 32, pool = 6, iters = 20;
    bool split = true, pdl = true, trace = false, check_only = false;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "--nc")) nc = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--chain")) chain = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--wps")) g_wps = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--layers")) layers = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--pool")) pool = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--iters")) iters = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--no-split")) split = false;
        else if (!strcmp(argv[i], "--no-pdl")) pdl = false;
        else if (!strcmp(argv[i], "--trace")) trace = true;
        else if (!strcmp(argv[i], "--check-only")) check_only = true;
    }
    if (chain < 1 || chain > ring::kMaxStages) { printf("--chain 1..%d\n", ring::kMaxStages); return 1; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int G = prop.multiProcessorCount;
    printf("device %s, %d SMs; nc %d chain %d layers %d pool %d split %d pdl %d\n", prop.name, G, nc, chain, layers, pool, (int)split, (int)pdl);

    // tables
    float h_code2[256];
    for (int i = 0; i < 256; i++) h_code2[i] = (i - 127) / 128.0f * 0.9f;
    float *d_code, *d_code2;
    CK(cudaMalloc(&d_code, 64));
    CK(cudaMalloc(&d_code2, 1024));
    CK(cudaMemcpy(d_code, kNf4, 64, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_code2, h_code2, 1024, cudaMemcpyHostToDevice));
    uint32_t* d_lut;
    CK(cudaMalloc(&d_lut, kLutBytes));
    gemv_lut_build_kernel<bf16><<<kLutBytes / 4 / 256, 256>>>(d_code, d_code2, d_lut);
    CK(cudaGetLastError());

    // one decoder layer = 4 stage kinds (grouped qkv, o, grouped gate/up, down)
    const int shapes[4][2] = {{6144, 4096}, {4096, 4096}, {28672, 4096}, {4096, 14336}};
    std::vector<Mat> mats;
    size_t packed_total = 0;
    for (int l = 0; l < pool; l++)
        for (int k = 0; k < 4; k++) {
            Mat m;
            m.rows = shapes[k][0];
            m.K = shapes[k][1];
            const size_t n = (size_t)m.rows * m.K;
            CK(cudaMalloc(&m.B, n / 2));
            CK(cudaMalloc(&m.qabs, n / 64));
            CK(cudaMalloc(&m.absmax2, (n / 64 + 255) / 256 * 4));
            CK(cudaMalloc(&m.offset, 4));
            CK(cudaMalloc(&m.out, m.rows * 2));
            fill_bytes<<<1024, 256>>>(m.B, n / 2, 17 * l + k);
            fill_bytes<<<256, 256>>>(m.qabs, n / 64, 1000 + 17 * l + k);
            fill_f32<<<64, 256>>>(m.absmax2, (n / 64 + 255) / 256, 77 + l, 0.01f, 0.05f);
            const float off = 0.03f + 0.001f * k;
            CK(cudaMemcpy(m.offset, &off, 4, cudaMemcpyHostToDevice));
            mats.push_back(m);
            packed_total += n / 2;
        }
    bf16 *x4096, *x14336, *bias;
    CK(cudaMalloc(&x4096, 4096 * 2));
    CK(cudaMalloc(&x14336, 14336 * 2));
    CK(cudaMalloc(&bias, 28672 * 2));
    fill_bf16<<<16, 256>>>(x4096, 4096, 5);
    fill_bf16<<<16, 256>>>(x14336, 14336, 6);
    fill_bf16<<<16, 256>>>(bias, 28672, 7);
    unsigned* ws;
    CK(cudaMalloc(&ws, ring::kWsBytes));
    CK(cudaMemset(ws, 0, ring::kWsBytes));
    unsigned long long* d_trace = nullptr;
    const size_t trace_n = (size_t)ring::kMaxStages * G * ring::kTraceSlots;
    if (trace) {
        CK(cudaMalloc(&d_trace, trace_n * 8 * 64));
        CK(cudaMemset(d_trace, 0, trace_n * 8 * 64));
    }
    CK(cudaDeviceSynchronize());
    printf("pool: %zu MB packed\n", packed_total >> 20);

    // stage `pos` of a launch over mats[idx[..]]: stages after the first consume the previous stage's (tagged) output; a gate/up
    // stage followed by a K = rows / 2 stage runs in pair mode and publishes silu(gate) * up
    auto fill_stage = [&](ring::Stage& st, const std::vector<int>& idx, int pos, bool with_bias) {
        const Mat& m = mats[idx[pos]];
        memset(&st, 0, sizeof(st));
        const bool has_next = pos + 1 < (int)idx.size();
        const Mat* nx = has_next ? &mats[idx[pos + 1]] : nullptr;
        st.pair = (nx && nx->K * 2 == m.rows) ? 1 : 0;
        st.half = m.rows / 2;
        st.publish = !nx ? 0 : (st.pair ? 2 : 1);
        if (nx && !st.pair && nx->K > m.rows) { printf("chain: stage %d needs %d activations, the previous one has %d rows\n", pos + 1, nx->K, m.rows); exit(1); }
        st.x_tagged = pos > 0 ? 1 : 0;
        if (!ring::make_weight_map(&st.map, m.B, m.rows, m.K, st.pair != 0)) {
            printf("tensor map failed\n");
            exit(1);
        }
        st.x = m.K == 4096 ? x4096 : x14336;
        st.s.qabsmax = m.qabs;
        st.s.code2 = d_code2;
        st.s.absmax2 = m.absmax2;
        st.s.offset = m.offset;
        st.s.shift2 = 8;
        for (int i = 0; i < kMaxMats; i++) {
            st.offsets[i] = nullptr;
            st.row_end[i] = 0x7fffffff;
        }
        st.offsets[0] = m.offset;
        st.out = m.out;
        st.bias = with_bias ? bias : nullptr;
        st.bias_stage = -1;
        ring::plan_stage(st, m.rows, m.K, G, split, nc / g_wps);
    };
    auto make_args = [&](ring::Args& a, const std::vector<int>& idx, bool with_bias, size_t& smem) {
        memset(&a, 0, sizeof(a));
        a.n = (int)idx.size();
        for (int i = 0; i < a.n; i++) fill_stage(a.st[i], idx, i, with_bias);
        a.lut = d_lut;
        a.code = d_code;
        a.ws = ws;
        smem = ring::plan_launch(a, 227 * 1024);
        if (!smem) { printf("smem plan failed\n"); exit(1); }
    };

    // ---- correctness: every stage kind, alone and chained, against the naive kernel
    {
        float* d_ref;
        bf16* d_h;
        CK(cudaMalloc(&d_ref, 28672 * 4));
        CK(cudaMalloc(&d_h, 28672 * 2));
        std::vector<float> ref(28672);
        std::vector<bf16> got(28672);
        for (int pass = 0; pass < 3; pass++) {
            ring::Args a;
            size_t smem;
            std::vector<int> idx;
            if (pass == 0) idx = {0, 1, 2, 3};
            else if (pass == 1) idx = {3, 1};
            else idx = {1, 2, 3};  // o -> gate/up -> down, down's residual = o's output through its tagged copy
            if ((int)idx.size() > chain) idx.resize(chain);
            make_args(a, idx, pass == 1, smem);
            if (pass == 2 && a.n == 3) {
                a.st[2].bias = mats[idx[0]].out;
                a.st[2].bias_stage = 0;
            }
            for (int i = 0; i < a.n; i++) CK(cudaMemset(mats[idx[i]].out, 0xff, mats[idx[i]].rows * 2));
            CK(launch(nc, a, G, smem, false, 0));
            CK(cudaDeviceSynchronize());
            printf("check pass %d: smem %zu x_bytes %d part_bytes %d\n", pass, smem, a.x_bytes, a.part_bytes);
            for (int i = 0; i < a.n; i++) {
                const Mat& m = mats[idx[i]];
                float off;
                CK(cudaMemcpy(&off, m.offset, 4, cudaMemcpyDeviceToHost));
                const bf16* xin = m.K == 4096 ? x4096 : x14336;
                if (i > 0) {  // the stage consumed the previous stage's output as the kernel stored it
                    const Mat& pm = mats[idx[i - 1]];
                    if (a.st[i - 1].pair) {
                        swiglu_ref<<<(m.K + 255) / 256, 256>>>(pm.out, pm.out + pm.rows / 2, d_h, m.K);
                        xin = d_h;
                    } else xin = pm.out;
                }
                ref_gemv<<<(m.rows + 7) / 8, 256>>>(xin, m.B, m.qabs, d_code2, m.absmax2, off, d_code,
                                                    pass == 1 ? bias : (pass == 2 && i == 2 ? mats[idx[0]].out : nullptr), d_ref, m.rows, m.K);
                CK(cudaMemcpy(ref.data(), d_ref, m.rows * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(got.data(), m.out, m.rows * 2, cudaMemcpyDeviceToHost));
                double worst = 0, scale = 0;
                int bad = -1;
                for (int r = 0; r < m.rows; r++) {
                    const double d = fabs((double)__bfloat162float(got[r]) - ref[r]);
                    if (!(d <= worst)) { worst = d; bad = r; }
                    if (fabs(ref[r]) > scale) scale = fabs(ref[r]);
                }
                printf("  stage %d (%dx%d, pair %d publish %d, active %d per %d rem %d gran %d): max err %.3g / scale %.3g = %.3g at row %d %s\n", i, m.rows, m.K,
                       a.st[i].pair, a.st[i].publish, a.st[i].active, a.st[i].per, a.st[i].rem, a.st[i].gran, worst, scale, worst / scale, bad, worst / scale < 8e-3 ? "ok" : "FAIL");
            }
        }
        CK(cudaFree(d_ref));
    }
    if (check_only) return 0;

    // ---- timing: `layers` layers x 4 stages, `chain` stages per launch, rotating over the pool
    std::vector<ring::Args> launches;
    std::vector<size_t> smems;
    {
        std::vector<int> seq;
        for (int l = 0; l < layers; l++)
            for (int k = 0; k < 4; k++) seq.push_back((l % pool) * 4 + k);
        for (size_t i = 0; i < seq.size(); i += chain) {
            std::vector<int> idx(seq.begin() + i, seq.begin() + (i + chain < seq.size() ? i + chain : seq.size()));
            ring::Args a;
            size_t smem;
            make_args(a, idx, false, smem);
            launches.push_back(a);
            smems.push_back(smem);
        }
    }
    size_t algo = 0;
    for (int k = 0; k < 4; k++) {
        // grouped stages are several Linears: bytes as bench.py counts them (per Linear: packed + stats + tables + x + y)
        const size_t n = (size_t)shapes[k][0] * shapes[k][1];
        const int nlin = k == 0 ? 3 : (k == 2 ? 2 : 1);
        algo += n / 2 + n / 64 + 4 * ((n + 16383) / 16384) + (size_t)nlin * (1024 + 64 + 4 + 2 * shapes[k][1]) + 2 * (size_t)shapes[k][0];
    }
    algo *= layers;
    cudaStream_t s;
    CK(cudaStreamCreate(&s));
    auto run_step = [&](unsigned long long* tr) {
        for (size_t i = 0; i < launches.size(); i++) {
            launches[i].trace = tr ? tr + i * trace_n : nullptr;
            CK(launch(nc, launches[i], G, smems[i], pdl, s));
        }
    };
    for (int i = 0; i < 3; i++) run_step(nullptr);
    CK(cudaStreamSynchronize(s));
    cudaGraph_t graph;
    cudaGraphExec_t gexec;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    run_step(nullptr);
    CK(cudaStreamEndCapture(s, &graph));
    CK(cudaGraphInstantiate(&gexec, graph, 0));
    for (int i = 0; i < 3; i++) CK(cudaGraphLaunch(gexec, s));
    CK(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e9f, sum = 0;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0, s));
        for (int i = 0; i < iters; i++) CK(cudaGraphLaunch(gexec, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        ms /= iters;
        sum += ms;
        if (ms < best) best = ms;
    }
    printf("RESULT nc %d wps %d chain %d split %d pdl %d: %.4f ms/step best, %.4f mean; %.1f GB/s (%.3f of 6531.6), %.2f us/stage\n", nc, g_wps, chain,
           (int)split, (int)pdl, best, sum / 5, algo / best / 1e6, algo / best / 1e6 / 6531.6, best * 1e3 / (layers * 4));

    if (trace) {
        // one eager step with the marks on (first `launches` entries only need their own buffers)
        run_step(d_trace);
        CK(cudaStreamSynchronize(s));
        const size_t nl = launches.size() < 64 ? launches.size() : 64;
        std::vector<unsigned long long> h(trace_n * nl);
        CK(cudaMemcpy(h.data(), d_trace, h.size() * 8, cudaMemcpyDeviceToHost));
        const char* names[9] = {"start", "stage in", "x loaded", "x staged", "loop end", "epi end", "tma first", "tma last", "all warps"};
        // launches in the middle of the step
        for (size_t li = nl / 2; li < nl / 2 + 2 && li < nl; li++) {
            unsigned long long t0 = ~0ull;
            for (int st = 0; st < launches[li].n; st++)
                for (int b = 0; b < G; b++) {
                    const unsigned long long v = h[li * trace_n + ((size_t)st * G + b) * ring::kTraceSlots + 1];
                    if (v && v < t0) t0 = v;
                }
            printf("launch %zu (relative to its first 'stage in'), us: min / median / max over CTAs\n", li);
            for (int st = 0; st < launches[li].n; st++) {
                printf("  stage %d (%dx%d)\n", st, launches[li].st[st].rows, launches[li].st[st].K);
                for (int m = (st == 0 ? 0 : 1); m < 9; m++) {
                    std::vector<double> v;
                    for (int b = 0; b < G; b++) {
                        const unsigned long long t = h[li * trace_n + ((size_t)st * G + b) * ring::kTraceSlots + m];
                        if (t) v.push_back(((double)t - (double)t0) / 1e3);
                    }
                    if (v.empty()) continue;
                    std::sort(v.begin(), v.end());
                    printf("    %-9s %8.2f %8.2f %8.2f\n", names[m], v.front(), v[v.size() / 2], v.back());
                }
            }
            if (li == nl / 2 && launches[li].n > 2) {
                const int st = 2;
                printf("  per CTA, stage %d: cta smid slots | stage-in staged loop-end all-warps epi-end\n", st);
                for (int b = 0; b < G; b++) {
                    const unsigned long long* t = &h[li * trace_n + ((size_t)st * G + b) * ring::kTraceSlots];
                    auto us = [&](int m) { return t[m] ? ((double)t[m] - (double)t0) / 1e3 : 0.0; };
                    const ring::Stage& sg = launches[li].st[st];
                    const int nsl = b < sg.active ? (sg.per + (b < sg.rem ? 1 : 0)) * sg.gran : 0;
                    printf("   cta %3d sm %3llu slots %3d | %7.2f %7.2f %7.2f %7.2f %7.2f\n", b, t[9], nsl, us(1), us(3), us(4), us(8), us(5));
                }
            }
        }
        // is the lateness systematic?  per CTA index and per SM: mean of (all warps done - median over the CTAs) over every traced stage
        {
            std::vector<double> late_cta(G, 0.0), late_sm(1024, 0.0), skew_by_stage(ring::kMaxStages, 0.0);
            std::vector<int> n_cta(G, 0), n_sm(1024, 0), n_stage(ring::kMaxStages, 0);
            for (size_t li = 1; li < nl; li++)
                for (int st = 0; st < launches[li].n; st++) {
                    std::vector<double> v;
                    for (int b = 0; b < G; b++) {
                        const unsigned long long t = h[li * trace_n + ((size_t)st * G + b) * ring::kTraceSlots + 8];
                        if (t) v.push_back((double)t);
                    }
                    if ((int)v.size() < G / 2) continue;
                    std::vector<double> sv = v;
                    std::sort(sv.begin(), sv.end());
                    const double med = sv[sv.size() / 2];
                    skew_by_stage[st % 4] += (sv.back() - med) / 1e3;
                    n_stage[st % 4]++;
                    for (int b = 0; b < G; b++) {
                        const unsigned long long* t = &h[li * trace_n + ((size_t)st * G + b) * ring::kTraceSlots];
                        if (!t[8]) continue;
                        const double d = ((double)t[8] - med) / 1e3;
                        late_cta[b] += d; n_cta[b]++;
                        const int sm = (int)(t[9] & 1023);
                        late_sm[sm] += d; n_sm[sm]++;
                    }
                }
            printf("mean (slowest CTA - median CTA) of 'all warps done', us, by stage position in the layer:");
            for (int k = 0; k < 4; k++) printf(" %d: %.2f", k, n_stage[k] ? skew_by_stage[k] / n_stage[k] : 0.0);
            printf("\n");
            std::vector<std::pair<double, int>> oc, os;
            for (int b = 0; b < G; b++) if (n_cta[b]) oc.push_back({late_cta[b] / n_cta[b], b});
            for (int m = 0; m < 1024; m++) if (n_sm[m]) os.push_back({late_sm[m] / n_sm[m], m});
            std::sort(oc.rbegin(), oc.rend());
            std::sort(os.rbegin(), os.rend());
            printf("latest CTA indices (mean lateness us):");
            for (int k = 0; k < 12 && k < (int)oc.size(); k++) printf(" %d:%.2f", oc[k].second, oc[k].first);
            printf("\nearliest CTA indices:");
            for (int k = 0; k < 6 && k < (int)oc.size(); k++) printf(" %d:%.2f", oc[oc.size() - 1 - k].second, oc[oc.size() - 1 - k].first);
            printf("\nlatest SMs (mean lateness us):");
            for (int k = 0; k < 12 && k < (int)os.size(); k++) printf(" %d:%.2f", os[k].second, os[k].first);
            printf("\nearliest SMs:");
            for (int k = 0; k < 6 && k < (int)os.size(); k++) printf(" %d:%.2f", os[os.size() - 1 - k].second, os[os.size() - 1 - k].first);
            printf("\n");
        }
    }
    return 0;
}
