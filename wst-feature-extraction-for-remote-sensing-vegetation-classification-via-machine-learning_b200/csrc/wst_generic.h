// wst_generic.h — the shape-generic scattering engine (any H x W, any L): every Fourier-domain step of kymatio's
// cascade evaluated as dense DFT-matrix products.
//
// The fused cascades of wst_cascade.h are compiled per padded side and need square grids whose sides factor into
// the codelets' radices.  The reference builds its transform from whatever image it loads
// (train_and_save_model.py:355-359: Scattering2D(J, L, shape=(H, W))), so shapes outside that list run here:
//
//   U^        = Fr . x_pad . Fc^T                                        (forward DFT as two matrix products)
//   U1[t]     = | Ar . (U^ (.) psi^[t]) . Ac^T |                          Ar, Ac = *partial* inverse DFT matrices:
//                                                                        "multiply -> Fourier fold by 2^j -> inverse
//                                                                        FFT" is the inverse DFT evaluated at every
//                                                                        2^j-th sample only, so the fold disappears
//   S[t]      = Gr . U1[t] . Gc^T                                        (separable low-pass + subsample + unpad)
//
// for any padded size (prime factors included: no FFT factorisation is involved).  The products run as batched
// complex GEMMs with the operator shared by the whole batch, either on the fp32 SIMT pipe or on the tensor cores as
// 3xTF32 (mma.sync m16n8k8, hi/lo split of both operands, fp32 accumulation): the DFT-as-GEMM variant that
// BASELINE.json's north_star asks to be measured against the FFT path.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "wst_cascade.h"

namespace wst {

enum GenericEngine { kEngineSimt = 1, kEngineTf32x3 = 2 };

struct GenericPlan;

// psi_hat [J*L][Hp][Wp], phi_hat [Hp][Wp]: host copies of the full-resolution Fourier-domain filters.
int generic_create(GenericPlan** out, int device, int H, int W, int J, int L, int max_order, int engine,
                   const float* psi_hat, const float* phi_hat, cudaMemPool_t pool, std::string& err);
void generic_destroy(GenericPlan* p);
// feats [nsig][2][K] and/or maps [nsig][K][h][w] (either may be NULL).  Returns cudaSuccess or the failing call's error.
cudaError_t generic_forward(const GenericPlan* p, const InputDesc& in, long long nsig, float* feats, float* maps,
                            cudaStream_t st, std::string& err);
// kernels one forward call over nsig signals launches
long long generic_launch_count(const GenericPlan* p, long long nsig);
void generic_geometry(const GenericPlan* p, int* K, int* h, int* w, int* Hp, int* Wp);

}  // namespace wst
