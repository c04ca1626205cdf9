// q4_gemv_ring.cu -- host side of the persistent ring GEMV (q4_gemv_ring.cuh): turns a chain of q4_gemv_fused_t stages
// (include/quantizations_b200.h) into one launch.  Replaces, per stage, what reference core.py:467-499 does with three launches.
//
// What "chain" means here: stage i + 1 consumes what stage i produced -- its x is stage i's out (or, for a SwiGLU pair, its
// x_gate / x are the two halves of a grouped gate/up stage's out), its bias may be the out of an earlier stage (the residual
// stream) -- and the kernel hands those values from CTA to CTA through tagged exchange words instead of ending the launch.  Any
// other aliasing between stages is refused (Q4_ERR_SHAPE): the caller then issues separate launches.
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>
#include <unordered_map>

#include "q4_gemv_ring.cuh"
#include "q4_launch.h"

namespace q4 {

extern int g_dyn_base_probed(cudaStream_t stream);  // q4_gemv.cu: where dynamic shared memory starts in a CTA's window (probed once)

namespace {

constexpr int kNC = 16, kWPS = 2;  // 16 consumer warps, two per ring slot (measured best: tools/micro/ring_bench.cu)

struct MapKey {
    const void* base;
    int64_t rows, K;
    int pair;
    bool operator==(const MapKey& o) const { return base == o.base && rows == o.rows && K == o.K && pair == o.pair; }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const
    {
        return std::hash<const void*>()(k.base) ^ (std::hash<int64_t>()(k.rows) * 1000003u) ^ (std::hash<int64_t>()(k.K) * 19349663u) ^ (size_t)k.pair;
    }
};

// tensor maps are pure functions of (pointer, shape, box): encoded once per weight (the driver call costs microseconds, a decode
// step issues 128 of them)
bool weight_map(CUtensorMap* out, const void* B, int64_t rows, int64_t K, bool pair)
{
    static std::mutex mu;
    static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{B, rows, K, pair ? 1 : 0};
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return true;
    }
    CUtensorMap m;
    if (!ring::make_weight_map(&m, B, rows, K, pair)) return false;
    if (cache.size() > 16384) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return true;
}

template <typename T, bool NESTED>
int launch_ring(const ring::Args& a, int grid, size_t smem, bool pdl, cudaStream_t stream)
{
    auto kern = ring::gemv_ring_kernel<T, NESTED, kNC, kWPS>;
    static bool attr_dev[kMaxDevices] = {};
    bool& attr = attr_dev[device_slot()];
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)ring::ring_threads(kNC));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = pdl ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return (int)e;
    }
    return finish_launch();
}

bool overlaps(const void* p, int64_t pbytes, const void* q, int64_t qbytes)
{
    if (!p || !q) return false;
    const uintptr_t a = reinterpret_cast<uintptr_t>(p), b = reinterpret_cast<uintptr_t>(q);
    return a < b + (uintptr_t)qbytes && b < a + (uintptr_t)pbytes;
}

}  // namespace

int gemv_4bit_ring(const q4_gemv_fused_t* stages, int n, void* workspace, int64_t workspace_bytes, cudaStream_t stream)
{
    if (n < 1) return 0;
    if (!stages || !workspace) return Q4_ERR_NULL;
    if (n > ring::kMaxStages || workspace_bytes < ring::kWsBytes || (reinterpret_cast<uintptr_t>(workspace) & 15)) return Q4_ERR_SHAPE;
    if (g_dyn_base_probed(stream) != kDynBase) return Q4_ERR_SHAPE;  // the table's place in shared memory is a compile-time constant of the kernel
    const int G = sm_count();
    if (G > ring::kWsMaxCtas) return Q4_ERR_SHAPE;
    const q4_gemv_fused_t& f0 = stages[0];
    if (f0.dtype != Q4_F16 && f0.dtype != Q4_BF16) return Q4_ERR_DTYPE;
    if (!f0.stats) return Q4_ERR_NULL;
    const bool nested = f0.stats->qabsmax != nullptr;
    const int64_t esz = 2;

    ring::Args c;
    memset(&c, 0, sizeof(c));
    c.n = n;
    c.lut = f0.lut;
    c.code = f0.code;
    c.ws = reinterpret_cast<unsigned*>(workspace);
    if (!c.lut || !c.code || (reinterpret_cast<uintptr_t>(c.lut) & 15)) return Q4_ERR_SHAPE;  // one table image per launch

    for (int i = 0; i < n; i++) {
        const q4_gemv_fused_t& f = stages[i];
        ring::Stage& st = c.st[i];
        if (!f.x || !f.B || !f.out || !f.stats || !f.code) return Q4_ERR_NULL;
        if (f.dtype != f0.dtype || f.lut != f0.lut || (f.stats->qabsmax != nullptr) != nested) return Q4_ERR_SHAPE;  // one table image = one (code, code2, dtype)
        if (f.blocksize != 64 || f.rows <= 0 || f.K <= 0 || (f.K % 256) != 0 || (f.rows % 32) != 0 || f.K > 16384 || f.rows > (1 << 20)) return Q4_ERR_SHAPE;
        if (f.flags & (Q4_GEMV_SWIGLU | Q4_GEMV_EXACT_F32)) return Q4_ERR_SHAPE;
        if (nested) {
            if (!f.stats->code2 || !f.stats->absmax2 || f.stats->blocksize2 < 128 ||
                (f.stats->blocksize2 & (f.stats->blocksize2 - 1)) || (reinterpret_cast<uintptr_t>(f.stats->qabsmax) & 1))
                return Q4_ERR_SHAPE;
        } else if (!f.stats->absmax || (reinterpret_cast<uintptr_t>(f.stats->absmax) & 7)) {
            return Q4_ERR_SHAPE;
        }
        if ((reinterpret_cast<uintptr_t>(f.B) & 15) || (reinterpret_cast<uintptr_t>(f.x) & 15) || (reinterpret_cast<uintptr_t>(f.x_gate) & 15) ||
            (reinterpret_cast<uintptr_t>(f.rms_weight) & 15) || (reinterpret_cast<uintptr_t>(f.out) & 3))
            return Q4_ERR_ALIGN;
        const int nmat = f.nmat < 1 ? 1 : f.nmat;
        if (nmat > kMaxMats || (nmat > 1 && (!f.row_end || (nested && !f.offsets)))) return Q4_ERR_SHAPE;

        // ---- how does this stage get its activation?
        st.x = f.x;
        st.x_gate = f.x_gate;
        st.x_tagged = 0;
        if (i > 0) {
            const q4_gemv_fused_t& p = stages[i - 1];
            ring::Stage& ps = c.st[i - 1];
            const int64_t half = p.rows / 2;
            if (!f.x_gate && f.x == p.out && f.K <= p.rows && p.rows <= ring::kXchMaxRows) {
                st.x_tagged = 1;
                ps.publish = 1;
            } else if (f.x_gate && f.x_gate == p.out && f.x == static_cast<const uint8_t*>(p.out) + half * esz && f.K == half &&
                       p.nmat == 2 && p.row_end && p.row_end[0] == half && (half % 16) == 0 && half <= ring::kXchMaxRows) {
                st.x_tagged = 1;
                st.x_gate = nullptr;  // the previous stage publishes silu(gate) * up
                ps.publish = 2;
                ps.pair = 1;
                ps.half = (int)half;
            }
        }
        if (!st.x_tagged) {  // plain staging is only ordered against the PREVIOUS kernel: nothing an earlier stage writes may be read
            for (int j = 0; j < i; j++)
                if (overlaps(f.x, f.K * esz, stages[j].out, stages[j].rows * esz) || overlaps(f.x_gate, f.K * esz, stages[j].out, stages[j].rows * esz))
                    return Q4_ERR_SHAPE;
        }
        st.rms_weight = f.rms_weight;
        st.rms_eps = f.rms_eps;
        st.s.absmax = f.stats->absmax;
        st.s.qabsmax = f.stats->qabsmax;
        st.s.code2 = f.stats->code2;
        st.s.absmax2 = f.stats->absmax2;
        st.s.offset = f.stats->offset;
        st.s.shift2 = nested ? ilog2(f.stats->blocksize2) : 0;
        for (int m = 0; m < kMaxMats; m++) {
            st.offsets[m] = nullptr;
            st.row_end[m] = 0x7fffffff;
        }
        st.offsets[0] = f.stats->offset;
        st.multi = (nmat > 1 && nested) ? 1 : 0;
        if (nmat > 1) {
            for (int m = 0; m < nmat; m++) {
                if (f.row_end[m] <= (m ? f.row_end[m - 1] : 0) || f.row_end[m] > f.rows || (f.row_end[m] % 32) != 0) return Q4_ERR_SHAPE;
                if (nested) st.offsets[m] = f.offsets[m];
                st.row_end[m] = f.row_end[m];
            }
            if (f.row_end[nmat - 1] != f.rows) return Q4_ERR_SHAPE;
            st.row_end[nmat - 1] = 0x7fffffff;
        }
        st.out = f.out;
        // ---- bias / residual: either memory the previous kernel left (plain loads), or the out of an EARLIER stage -- then through
        //      that stage's tagged copy (other CTAs wrote it; the tag says when it is there)
        st.bias = f.bias;
        st.bias_stage = -1;
        if (f.bias) {
            if (reinterpret_cast<uintptr_t>(f.bias) & 1) return Q4_ERR_ALIGN;
            for (int j = i - 1; j >= 0; j--) {
                if (!overlaps(f.bias, f.rows * esz, stages[j].out, stages[j].rows * esz)) continue;
                if (f.bias != stages[j].out || f.rows > stages[j].rows || c.st[j].pair || stages[j].rows > ring::kXchMaxRows) return Q4_ERR_SHAPE;
                if (i - j > ring::kXchBufs - 2) return Q4_ERR_SHAPE;  // its exchange buffer (stage % kXchBufs) would be in use again
                st.bias_stage = j;
                break;
            }
        }
        st.rows = (int)f.rows;
        st.K = (int)f.K;
        // ---- tensor-parallel row-parallel stage: the all-reduce over the ranks runs in the epilogue (q4_allreduce_t), before bias /
        //      residual and before the outputs are published to the next stage
        if (f.allreduce && f.allreduce->world > 1) {
            const q4_allreduce_t* ar = f.allreduce;
            if (!ar->peer_bases) return Q4_ERR_NULL;
            if (ar->world > kArMaxWorld || ar->rank < 0 || ar->rank >= ar->world || f.rows != ar->max_rows || G > kArMaxCtas || nmat > 1) return Q4_ERR_SHAPE;
            st.ar_peer_bases = ar->peer_bases;
            st.ar_world = ar->world;
            st.ar_rank = ar->rank;
            st.ar_max_rows = ar->max_rows;
        }
    }
    // publishing stages need their tagged copy even when only a later bias reads it
    for (int i = 0; i < n; i++)
        if (c.st[i].bias_stage >= 0) {
            ring::Stage& ps = c.st[c.st[i].bias_stage];
            if (ps.publish == 2) return Q4_ERR_SHAPE;
            ps.publish = 1;
        }
    // a stage whose out is overwritten by a later stage of the same launch (the residual stream, updated in place twice per layer)
    // keeps its values in the tagged copy only: two CTAs never store to the same address within a launch
    for (int j = 0; j < n; j++)
        for (int i = j + 1; i < n; i++)
            if (stages[i].out == stages[j].out && stages[i].rows == stages[j].rows && c.st[j].publish == 1) c.st[j].skip_out = 1;
    for (int j = 0; j < n; j++)
        for (int i = j + 1; i < n; i++)
            if (!c.st[j].skip_out && overlaps(stages[i].out, stages[i].rows * esz, stages[j].out, stages[j].rows * esz)) return Q4_ERR_SHAPE;
    for (int i = 0; i < n; i++) {
        ring::Stage& st = c.st[i];
        if (!weight_map(&st.map, stages[i].B, st.rows, st.K, st.pair != 0)) return Q4_ERR_SHAPE;
        ring::plan_stage(st, st.rows, st.K, G, true, kNC / kWPS);
        if (st.publish && (st.rows % 32) != 0) return Q4_ERR_SHAPE;
    }
    const size_t smem = ring::plan_launch(c, 227 * 1024);
    if (!smem) return Q4_ERR_SHAPE;
    const bool pdl = (f0.flags & Q4_GEMV_PDL) != 0;
    if (f0.dtype == Q4_F16) return nested ? launch_ring<__half, true>(c, G, smem, pdl, stream) : launch_ring<__half, false>(c, G, smem, pdl, stream);
    return nested ? launch_ring<__nv_bfloat16, true>(c, G, smem, pdl, stream) : launch_ring<__nv_bfloat16, false>(c, G, smem, pdl, stream);
}

}  // namespace q4

static_assert(q4::ring::kWsBytes == Q4_GEMV_RING_WS_BYTES, "header constant out of date");
static_assert(q4::ring::kMaxStages == Q4_GEMV_RING_MAX_STAGES, "header constant out of date");
static_assert(sizeof(q4::ring::Args) <= 32764, "kernel parameters exceed the 32-KB limit (CUDA >= 12.1, sm_70+)");
