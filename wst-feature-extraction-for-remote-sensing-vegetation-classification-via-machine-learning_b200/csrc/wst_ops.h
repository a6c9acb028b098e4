// wst_ops.h — the per-configuration entry points wst_lib.cu dispatches on.  Each compiled (N, J) lives in its
// own translation unit (wst_cfg_inst.cu compiled with -DWST_CFG_N=.. -DWST_CFG_J=..), so configurations build
// in parallel.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "wst_tables.h"

namespace wst {

struct CfgOps {
    int N, J, NT, hout;
    int cluster;                 // CTAs per signal (1: one CTA per signal; > 1: a thread-block cluster per signal)
    size_t smem;                 // dynamic shared memory per CTA
    size_t workspace_cfloats;    // per-slot global workspace (0: data region in shared memory)
    const void* kernel;
    bool (*build)(int, const float*, const float*, std::vector<float>&, TableOffsets&, std::string&);
    void (*bind)(PlanTables&, const float*, const TableOffsets&);
    bool can_split;              // small batches: several CTAs may share one signal (shared-memory variant)
    int (*num_units)(int L, int max_order);   // units of shared work per signal = the largest useful split
    // (tables, input descriptor, nsig, u0h scratch, workspace, maps_out | NULL, maps scratch | NULL, feats | NULL, grid, stream,
    //  split = CTAs per signal, per-signal completion counters (split > 1) | ticket counter or NULL (split == 1))
    cudaError_t (*launch)(const PlanTables&, const InputDesc&, long long, cfloat*, cfloat*, float*, float*, float*, int, cudaStream_t,
                          int, int*);
    cudaError_t (*max_slots)(int device, int* slots);   // signals in flight (resident CTAs or clusters); also sets kernel attributes
    cudaError_t (*launch_prof)(const PlanTables&, const InputDesc&, long long, cfloat*, cfloat*, float*, float*, float*, long long*, int, cudaStream_t);
};

}  // namespace wst

// the list of compiled cascades; tuning builds point WST_CONFIGS_FILE at a filtered copy
#ifndef WST_CONFIGS_FILE
#define WST_CONFIGS_FILE "wst_configs.inc"
#endif
#define CFG(n, j) wst::CfgOps wst_make_ops_##n##_##j();
#define CFGG(n, j) wst::CfgOps wst_make_ops_##n##_##j();
#include WST_CONFIGS_FILE
#undef CFG
#undef CFGG
