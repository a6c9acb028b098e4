// wst_cfg_inst.cu — one compiled cascade configuration: compile with -DWST_CFG_N=<padded side> -DWST_CFG_J=<J>
// (and -DWST_CFG_GLOBAL=1 -DWST_CFG_NT=<threads> for the global-workspace variant).
#include "wst_ops.h"

#ifndef WST_CFG_GLOBAL
#define WST_CFG_GLOBAL 0
#endif
#ifndef WST_CFG_NT
#define WST_CFG_NT WST_NT
#endif

using namespace wst;

namespace {

typedef Cfg<WST_CFG_N, WST_CFG_J, WST_CFG_NT, (WST_CFG_GLOBAL != 0)> ThisCfg;

template <class C, class Exec>
__device__ __forceinline__ void run_cascade(Exec& ex, const PlanTables& pt, const InputDesc& in,
                                            long long nsig, cfloat* u0h_scratch, cfloat* workspace, float* maps_out,
                                            float* maps_scratch, float* feats) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cfloat* sbase = reinterpret_cast<cfloat*>(smem_raw);
    cfloat* sm = C::WS_GLOBAL ? workspace + (size_t)blockIdx.x * C::workspace_cfloats() : sbase;
    cfloat* twsm = C::WS_GLOBAL ? sbase : sbase + C::smem_cfloats();
    float* gsm = reinterpret_cast<float*>(twsm + C::tw_total);
    float* lpbuf = gsm + C::g_total;
    const size_t map_elems = (size_t)pt.K * C::HOUT * C::HOUT;
    Cascade<C, Exec> prog{ex, pt, sm, twsm, gsm, lpbuf,
                          u0h_scratch + (size_t)blockIdx.x * (C::N * (C::N / 2 + 1)), nullptr};
    prog.load_twiddles();
    for (long long s = blockIdx.x; s < nsig; s += gridDim.x) {
        // maps go to the caller's buffer, or to this CTA's own (L2-resident) scratch when only features are wanted
        prog.maps = maps_out ? maps_out + (size_t)s * map_elems : maps_scratch + (size_t)blockIdx.x * map_elems;
        const SignalSrc src = signal_source(in, s, pt.H, pt.W);
        prog.run(src, feats ? feats + (size_t)s * 2 * pt.K : nullptr);
    }
}

// One persistent CTA per SM; each CTA runs the whole scattering cascade of one (patch, channel) signal at a
// time (wst_cascade.h).
template <class C>
__global__ void __launch_bounds__(C::NT, 1)
cascade_kernel(const PlanTables pt, const InputDesc in, long long nsig, cfloat* u0h_scratch,
               cfloat* workspace, float* maps_out, float* maps_scratch, float* feats) {
    DevExec ex;
    run_cascade<C>(ex, pt, in, nsig, u0h_scratch, workspace, maps_out, maps_scratch, feats);
}

// Debug twin: same program, executor that accumulates clock64() per phase tag; CTA 0's totals -> cycles.
template <class C>
__global__ void __launch_bounds__(C::NT, 1)
cascade_prof_kernel(const PlanTables pt, const InputDesc in, long long nsig, cfloat* u0h_scratch,
                    cfloat* workspace, float* maps_out, float* maps_scratch, float* feats, long long* cycles) {
    __shared__ long long acc[kNumPhaseTags];
    for (int i = threadIdx.x; i < kNumPhaseTags; i += C::NT) acc[i] = 0;
    __syncthreads();
    ProfExec ex{acc};
    run_cascade<C>(ex, pt, in, nsig, u0h_scratch, workspace, maps_out, maps_scratch, feats);
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < kNumPhaseTags; i += C::NT) cycles[i] = acc[i];
}

template <class C>
cudaError_t launch_cascade(const PlanTables& pt, const InputDesc& in, long long nsig, cfloat* u0h, cfloat* ws,
                           float* maps_out, float* maps_scratch, float* feats, int grid, cudaStream_t st) {
    cascade_kernel<C><<<grid, C::NT, C::smem_bytes(), st>>>(pt, in, nsig, u0h, ws, maps_out, maps_scratch, feats);
    return cudaGetLastError();
}

template <class C>
cudaError_t launch_cascade_prof(const PlanTables& pt, const InputDesc& in, long long nsig, cfloat* u0h, cfloat* ws,
                                float* maps_out, float* maps_scratch, float* feats, long long* cycles, int grid,
                                cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(cascade_prof_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::smem_bytes());
    if (e != cudaSuccess) return e;
    cascade_prof_kernel<C><<<grid, C::NT, C::smem_bytes(), st>>>(pt, in, nsig, u0h, ws, maps_out, maps_scratch, feats, cycles);
    return cudaGetLastError();
}

}  // namespace

#define WST_CAT2(a, b, c, d) a##b##c##d
#define WST_CAT(a, b, c, d) WST_CAT2(a, b, c, d)

wst::CfgOps WST_CAT(wst_make_ops_, WST_CFG_N, _, WST_CFG_J)() {
    typedef ThisCfg C;
    static_assert(C::smem_bytes() + kNumPhaseTags * 8 <= 232448,
                  "configuration exceeds the 227 KB of shared memory a CTA may use");
    CfgOps o;
    o.N = C::N; o.J = C::J; o.NT = C::NT; o.hout = C::HOUT;
    o.smem = C::smem_bytes();
    o.workspace_cfloats = C::workspace_cfloats();
    o.kernel = (const void*)cascade_kernel<C>;
    o.build = &build_tables<C>;
    o.bind = &bind_tables<C>;
    o.launch = &launch_cascade<C>;
    o.launch_prof = &launch_cascade_prof<C>;
    return o;
}
