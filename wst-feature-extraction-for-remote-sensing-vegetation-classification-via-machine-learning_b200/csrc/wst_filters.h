// wst_filters.h — the Morlet / Gabor filter bank of kymatio 0.3.0 (filter_bank.py::gabor_2d,
// morlet_2d, filter_bank), evaluated in double precision.
//
// The reference rebuilds this bank for every image (train_and_save_model.py:359,
// inference.py:242) and that is ~90% of its wall time; here it is built once per
// plan on the GPU: gabor_point() per pixel, a dense fp64 DFT per axis, then
//     psi^ = Re( W^ - K * Wmod^ ),  K = W^[0,0] / Wmod^[0,0]      (morlet_2d, zero mean)
//     phi^ = Re( G^ ).
// The per-point formula below is shared by the CUDA kernels (wst_lib.cu) and by the CPU
// emulation used in tests.
#pragma once
#include <cmath>
#include "wst_common.h"

namespace wst {

struct GaborParams {
    double c00, c01s, c11;     // quadratic form of the Gaussian envelope (c01s = curv[0,1] + curv[1,0])
    double wx, wy;             // carrier: xi*cos(theta), xi*sin(theta)
    double inv_norm;           // 1 / (2 * 3.1415 * sigma^2 / slant)   -- literal 3.1415 as in kymatio
};

// kymatio builds the rotation matrices in float32 and everything else in float64.
inline GaborParams make_gabor(double sigma, double theta, double xi, double slant) {
    const double c = (double)(float)std::cos(theta), s = (double)(float)std::sin(theta);
    const double ns = (double)(float)(-std::sin(theta));
    const double s2 = slant * slant, den = 2.0 * sigma * sigma;
    GaborParams p;
    // curv = R * diag(1, slant^2) * R_inv / (2 sigma^2), R = [[c, -s], [s, c]], R_inv = [[c, s], [-s, c]]
    p.c00 = (c * c + ns * s2 * ns) / den;
    double c01 = (c * s + ns * s2 * c) / den;
    double c10 = (s * c + c * s2 * ns) / den;
    p.c01s = c01 + c10;
    p.c11 = (s * s + c * s2 * c) / den;
    p.wx = xi * std::cos(theta);
    p.wy = xi * std::sin(theta);
    p.inv_norm = 1.0 / (2.0 * 3.1415 * sigma * sigma / slant);
    return p;
}

// gab[x][y] = inv_norm * sum_{ex,ey in -2..2} exp(-(c00 X^2 + c01s X Y + c11 Y^2) + i (wx X + wy Y)),
// X = x + ex*M, Y = y + ey*N  (x: row index, y: column index)
WST_HD void gabor_point(const GaborParams& p, int x, int y, int M, int N, double& re, double& im) {
    double sr = 0.0, si = 0.0;
    for (int ex = -2; ex <= 2; ++ex)
        for (int ey = -2; ey <= 2; ++ey) {
            double X = (double)(x + ex * M), Y = (double)(y + ey * N);
            double env = exp(-(p.c00 * X * X + p.c01s * X * Y + p.c11 * Y * Y));
            double ph = p.wx * X + p.wy * Y;
            double sn, cs;
#ifdef __CUDA_ARCH__
            sincos(ph, &sn, &cs);
#else
            sn = std::sin(ph); cs = std::cos(ph);
#endif
            sr += env * cs;
            si += env * sn;
        }
    re = sr * p.inv_norm;
    im = si * p.inv_norm;
}

// The 2*J*L + 1 Gabor functions a bank needs: for n = j*L + t the Morlet carrier wave (index 2n)
// and its envelope (index 2n+1), then the low-pass (index 2*J*L).
inline void bank_gabors(int J, int L, GaborParams* out) {
    const double pi = 3.14159265358979323846;   // np.pi
    for (int j = 0; j < J; ++j)
        for (int t = 0; t < L; ++t) {
            double sigma = 0.8 * std::pow(2.0, j);
            double theta = (double)((int)(L - L / 2.0 - 1) - t) * pi / L;
            double xi = 3.0 / 4.0 * pi / std::pow(2.0, j);
            double slant = 4.0 / L;
            out[2 * (j * L + t)] = make_gabor(sigma, theta, xi, slant);
            out[2 * (j * L + t) + 1] = make_gabor(sigma, theta, 0.0, slant);
        }
    out[2 * J * L] = make_gabor(0.8 * std::pow(2.0, J - 1), 0.0, 0.0, 1.0);
}

}  // namespace wst
