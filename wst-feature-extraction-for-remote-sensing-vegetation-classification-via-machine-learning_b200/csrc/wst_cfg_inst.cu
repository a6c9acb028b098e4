// wst_cfg_inst.cu — one compiled cascade configuration: compile with -DWST_CFG_N=<padded side> -DWST_CFG_J=<J>
// (and -DWST_CFG_GLOBAL=1 -DWST_CFG_NT=<threads> for the global-workspace variant; -DWST_CFG_CL=<CTAs per cluster>
// -DWST_GLOBAL_BUDGET=<cfloats> to spread one signal over a thread-block cluster, WST_CFG_NT then counts the threads
// of the whole cluster).
#include "wst_ops.h"

#ifndef WST_CFG_GLOBAL
#define WST_CFG_GLOBAL 0
#endif
#ifndef WST_CFG_NT
#define WST_CFG_NT WST_NT
#endif
#ifndef WST_CFG_CL
#define WST_CFG_CL 1
#endif

using namespace wst;

namespace {

typedef Cfg<WST_CFG_N, WST_CFG_J, WST_CFG_NT, (WST_CFG_GLOBAL != 0), WST_CFG_CL> ThisCfg;

template <class C, bool SPLIT, class Exec>
__device__ __forceinline__ void run_cascade(Exec& ex, const PlanTables& pt, const InputDesc& in,
                                            long long nsig, cfloat* u0h_scratch, cfloat* workspace, float* maps_out,
                                            float* maps_scratch, float* feats, int split = 1, int* done = nullptr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // slot = the CTA (or cluster of C::CL CTAs) that owns one signal at a time; its scratch areas are indexed by it.
    // SPLIT (small batches, shared-memory variant): `split` consecutive CTAs share the signals of one slot, each
    // taking every split-th first-order group; U0^ scratch is then per CTA and the maps scratch per signal.
    const int cta = blockIdx.x / C::CL;
    const int slot = SPLIT ? cta / split : cta, nslots = SPLIT ? (gridDim.x / C::CL) / split : gridDim.x / C::CL;
    cfloat* sbase = reinterpret_cast<cfloat*>(smem_raw);
    cfloat* sm = C::WS_GLOBAL ? workspace + (size_t)slot * C::workspace_cfloats() : sbase;
    cfloat* twsm = C::WS_GLOBAL ? sbase : sbase + C::smem_cfloats();
    float* gsm = reinterpret_cast<float*>(twsm + C::tw_total);
    float* lpbuf = C::WS_GLOBAL ? reinterpret_cast<float*>(sm + C::smem_cfloats()) : gsm + C::g_total;
    // tile of the staged FFT passes (global-workspace variant), after the tables; 8-byte aligned: all counts are even
    cfloat* stage = reinterpret_cast<cfloat*>(gsm + C::g_total + (C::WS_GLOBAL ? 0 : C::lpbuf_floats()));
    const size_t map_elems = (size_t)pt.K * C::HOUT * C::HOUT;
    Cascade<C, Exec, SPLIT> prog{ex, pt, sm, twsm, gsm, lpbuf, stage,
                                 u0h_scratch + (size_t)cta * (C::N * (C::N / 2 + 1)), nullptr};
    if constexpr (SPLIT) { prog.part = cta % split; prog.nparts = split; }
    __shared__ int is_last;
    __shared__ unsigned long long input_mbar[2];         // [0] completion barrier of the input prefetch (TMA bulk copies), [1] its state
    if (threadIdx.x == 0) { mbar_init(&input_mbar[0], 1); input_mbar[1] = 0ull; }
    prog.mbar = &input_mbar[0];
    prog.load_twiddles();                                // its barrier also publishes the mbarrier initialisation
    if (slot < nsig) prog.prefetch_input(signal_source(in, slot, pt.H, pt.W));
    // Signals after a slot's first one are handed out by a device-wide ticket counter (`done`, zeroed by the host)
    // instead of a fixed stride: SMs do not run at exactly the same speed (distance to the L2 slices that hold the filter
    // bank, the other die), and over ~80 signals per CTA the static assignment waits for the slowest one.  The ticket of
    // the following signal is drawn at the top of an iteration, so the input prefetch still knows it one signal ahead.
    constexpr bool DYN = WST_OPT_DYNSCHED && !SPLIT && C::CL == 1;
    __shared__ long long next_sig;
    const bool dyn = DYN && done != nullptr;
    for (long long s = slot; s < nsig;) {
        if (dyn) {
            __syncthreads();                                 // everybody has read the previous ticket
            if (threadIdx.x == 0) next_sig = (long long)nslots + atomicAdd(done, 1);
        }                                                    // (published by the first barrier inside run())
        // maps go to the caller's buffer, or to this slot's own (L2-resident) scratch when only features are wanted
        prog.maps = maps_out ? maps_out + (size_t)s * map_elems
                             : maps_scratch + (size_t)(SPLIT ? s : slot) * map_elems;
        auto next = [&](SignalSrc& o) -> bool {
            const long long sn = dyn ? next_sig : s + nslots;
            if (sn >= nsig) return false;
            o = signal_source(in, sn, pt.H, pt.W);
            return true;
        };
        auto fptr = [&]() -> float* { return feats ? feats + (size_t)s * 2 * pt.K : nullptr; };
        if constexpr (!SPLIT) {
            prog.run(signal_source(in, s, pt.H, pt.W), fptr, next);
            s = dyn ? next_sig : s + nslots;
        } else {
            // shared signal: no pooling inside run(); the CTA that finishes last pools from the (L2) maps of all parts
            prog.run(signal_source(in, s, pt.H, pt.W), []() -> float* { return nullptr; }, next);
            if (feats) {
                __threadfence();                             // this thread's map stores are visible device-wide ...
                __syncthreads();
                if (threadIdx.x == 0) is_last = atomicAdd(&done[s], 1) == split - 1;     // ... before the CTA checks in
                __syncthreads();
                if (is_last) {
                    __threadfence();
                    prog.template pool<true>(fptr());
                }
            }
            s += nslots;
        }
    }
}

// One persistent CTA per SM; each CTA (or each cluster of C::CL CTAs, for sides whose arrays live in the L2-resident
// workspace) runs the whole scattering cascade of one (patch, channel) signal at a time (wst_cascade.h).
template <class C>
__global__ void __launch_bounds__(C::NTL, C::min_ctas())
cascade_kernel(const __grid_constant__ PlanTables pt, const InputDesc in, long long nsig, cfloat* u0h_scratch,
               cfloat* workspace, float* maps_out, float* maps_scratch, float* feats, int* ticket) {
    DevExec<C::CL> ex{C::CL > 1 ? cluster_cta_rank() * C::NTL : 0};
    run_cascade<C, false>(ex, pt, in, nsig, u0h_scratch, workspace, maps_out, maps_scratch, feats, 1, ticket);
}

// Small-batch twin (shared-memory variant only): `split` CTAs per signal, the last one to finish pools.
template <class C>
__global__ void __launch_bounds__(C::NTL, C::min_ctas())
cascade_split_kernel(const __grid_constant__ PlanTables pt, const InputDesc in, long long nsig, cfloat* u0h_scratch,
                     float* maps_out, float* maps_scratch, float* feats, int split, int* done) {
    if constexpr (!C::WS_GLOBAL && C::CL == 1) {
        DevExec<1> ex{0};
        run_cascade<C, true>(ex, pt, in, nsig, u0h_scratch, nullptr, maps_out, maps_scratch, feats, split, done);
    }
}

// Debug twin: same program, executor that accumulates clock64() per phase tag; CTA 0's totals -> cycles.
template <class C>
__global__ void __launch_bounds__(C::NTL, 1)
cascade_prof_kernel(const __grid_constant__ PlanTables pt, const InputDesc in, long long nsig, cfloat* u0h_scratch,
                    cfloat* workspace, float* maps_out, float* maps_scratch, float* feats, long long* cycles) {
    __shared__ long long acc[kNumPhaseTags];
    for (int i = threadIdx.x; i < kNumPhaseTags; i += C::NTL) acc[i] = 0;
    __syncthreads();
    ProfExec<C::CL> ex{C::CL > 1 ? cluster_cta_rank() * C::NTL : 0, acc};
    run_cascade<C, false>(ex, pt, in, nsig, u0h_scratch, workspace, maps_out, maps_scratch, feats);
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < kNumPhaseTags; i += C::NTL) cycles[i] = acc[i];
}

// `slots` = CTAs (clusters) to launch; the grid is slots * C::CL CTAs of C::NTL threads.
template <class C, class... Args>
cudaError_t launch_any(void (*kernel)(Args...), int slots, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(slots * C::CL), 1, 1);
    cfg.blockDim = dim3(C::NTL, 1, 1);
    cfg.dynamicSmemBytes = C::smem_bytes();
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = C::CL > 1 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <class C>
cudaError_t launch_cascade(const PlanTables& pt, const InputDesc& in, long long nsig, cfloat* u0h, cfloat* ws,
                           float* maps_out, float* maps_scratch, float* feats, int slots, cudaStream_t st, int split,
                           int* done) {
    // slots = CTAs (clusters) in the grid; with split > 1 consecutive groups of `split` CTAs share their signals
    if (split > 1)          // (its shared-memory attribute is set per device by max_slots at plan creation)
        return launch_any<C>(cascade_split_kernel<C>, slots, st, pt, in, nsig, u0h, maps_out, maps_scratch, feats, split, done);
    return launch_any<C>(cascade_kernel<C>, slots, st, pt, in, nsig, u0h, ws, maps_out, maps_scratch, feats, done);
}

template <class C>
cudaError_t launch_cascade_prof(const PlanTables& pt, const InputDesc& in, long long nsig, cfloat* u0h, cfloat* ws,
                                float* maps_out, float* maps_scratch, float* feats, long long* cycles, int slots,
                                cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(cascade_prof_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::smem_bytes());
    if (e == cudaSuccess && C::CL > 8)
        e = cudaFuncSetAttribute(cascade_prof_kernel<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    return launch_any<C>(cascade_prof_kernel<C>, slots, st, pt, in, nsig, u0h, ws, maps_out, maps_scratch, feats, cycles);
}

// Signals in flight: resident CTAs, or resident clusters (the hardware places a cluster inside one GPC).
template <class C>
cudaError_t max_slots(int device, int* slots) {
    cudaError_t e = cudaFuncSetAttribute(cascade_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
    if (e != cudaSuccess) return e;
    if constexpr (!C::WS_GLOBAL && C::CL == 1) {        // function attributes are per device: set the split twin's here too
        e = cudaFuncSetAttribute(cascade_split_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes());
        if (e != cudaSuccess) return e;
    }
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    if (C::CL == 1) {
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cascade_kernel<C>, C::NTL, C::smem_bytes());
        *slots = sms * per_sm;
        return e;
    }
    if (C::CL > 8) {
        e = cudaFuncSetAttribute(cascade_kernel<C>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(sms / C::CL * C::CL), 1, 1);
    cfg.blockDim = dim3(C::NTL, 1, 1);
    cfg.dynamicSmemBytes = C::smem_bytes();
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C::CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(slots, cascade_kernel<C>, &cfg);
}

}  // namespace

#define WST_CAT2(a, b, c, d) a##b##c##d
#define WST_CAT(a, b, c, d) WST_CAT2(a, b, c, d)

wst::CfgOps WST_CAT(wst_make_ops_, WST_CFG_N, _, WST_CFG_J)() {
    typedef ThisCfg C;
    static_assert(C::smem_bytes() + kNumPhaseTags * 8 + 16 <= 232448,
                  "configuration exceeds the 227 KB of shared memory a CTA may use");
    CfgOps o;
    o.N = C::N; o.J = C::J; o.NT = C::NT; o.hout = C::HOUT; o.cluster = C::CL;
    o.smem = C::smem_bytes();
    o.can_split = !C::WS_GLOBAL && C::CL == 1;
    o.num_units = &Cascade<C, HostExec<C::NT>>::num_units;
    o.workspace_cfloats = C::workspace_cfloats();
    o.kernel = (const void*)cascade_kernel<C>;
    o.build = &build_tables<C>;
    o.bind = &bind_tables<C>;
    o.launch = &launch_cascade<C>;
    o.launch_prof = &launch_cascade_prof<C>;
    o.max_slots = &max_slots<C>;
    return o;
}
