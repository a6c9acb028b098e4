// wst_emu.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's kernel headers (csrc/wst_cascade.h, the very code nvcc builds for
// sm_100a) with g++ and replays each CTA's barrier-separated phases thread by thread on the CPU.
// It exists so that the index arithmetic of the fused kernels (digit-swapped FFT orders,
// Hermitian half spectra, folds, shared-memory layout) can be checked against the oracle inside
// the GPU-less build container.  It is never imported by the product package and is not a
// fallback: wst_b200 fails loudly when libwst_b200.so / a GPU is missing.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "wst_tables.h"
#include "wst_filters.h"

#ifndef WST_EMU_CONFIG_FILE
#define WST_EMU_CONFIG_FILE "wst_configs.inc"     // tests may substitute a shorter list to compile faster
#endif

using namespace wst;

namespace {

template <class C>
int run_cfg(int L, int max_order, int H, int W, const float* psi_hat, const float* phi_hat,
            const float* x, int nsig, float* maps_out, float* feats_out, std::string& err) {
    std::vector<float> buf; TableOffsets off;
    if (!build_tables<C>(L, psi_hat, phi_hat, buf, off, err)) return -1;
    PlanTables pt{};
    pt.L = L; pt.max_order = max_order; pt.K = num_coefficients(C::J, L, max_order);
    pt.H = H; pt.W = W; pt.pad_top = (C::N - H) / 2; pt.pad_left = (C::N - W) / 2;
    bind_tables<C>(pt, buf.data(), off);
    std::vector<cfloat> sm(C::smem_cfloats() + 64), twsm(C::tw_total);
    std::vector<float> gsm(C::g_total), lpbuf(C::lpbuf_floats() + 1);
    // (the hybrid shared-memory region of the global-workspace variant aliases the stage tiles, as in the kernel)
    std::vector<cfloat> stage(cx_max(C::STAGE_BUFS * C::stage_cfloats(), C::hybrid_cfloats()) + 1);
    std::vector<cfloat> u0h((size_t)C::N * (C::N / 2 + 1));
    HostExec<C::NT> ex;
    const size_t map_sz = (size_t)pt.K * C::HOUT * C::HOUT;
    for (int s = 0; s < nsig; ++s) {
        // poison shared memory so that reads of never-written cells show up as NaN
        for (auto& v : sm) v = cmake(NAN, NAN);
        for (auto& v : lpbuf) v = NAN;
        Cascade<C, HostExec<C::NT>> prog{ex, pt, sm.data(), twsm.data(), gsm.data(), lpbuf.data(), stage.data(), u0h.data(), maps_out + s * map_sz};
        prog.load_twiddles();
        SignalSrc src{x + (size_t)s * H * W, nullptr, 1, W};
        prog.run(src, [&]() -> float* { return feats_out ? feats_out + (size_t)s * 2 * pt.K : nullptr; });
    }
    return 0;
}

std::string g_err;

}  // namespace

extern "C" {

const char* emu_last_error() { return g_err.c_str(); }

// 1-D/2-D FFT self-test hooks: forward then inverse through the digit-swapped layout.
// data: M x M complex (interleaved), in place; dir=-1 forward (natural -> natural, via explicit un-swap), +1 inverse.
int emu_fft2(int M, int dir, float* data) {
    auto run = [&](auto Mc) {
        constexpr int m = decltype(Mc)::value;
        constexpr int P = m + 1;
        constexpr int R1 = Fft1<m>::R1, R2 = Fft1<m>::R2;
        std::vector<cfloat> sm((size_t)m * P), tw(m);
        for (int k1 = 0; k1 < R1; ++k1) for (int i2 = 0; i2 < R2; ++i2) {
            double a = -6.283185307179586 * ((i2 * k1) % m) / m;
            tw[k1 * R2 + i2] = cmake((float)cos(a), (float)sin(a));
        }
        HostExec<256> ex;
        const cfloat* in = reinterpret_cast<const cfloat*>(data);
        cfloat* out = reinterpret_cast<cfloat*>(data);
        if (dir < 0) {
            for (int r = 0; r < m; ++r) for (int c = 0; c < m; ++c) sm[Fft1<m>::pos_s(r) * P + Fft1<m>::pos_s(c)] = in[r * m + c];
            fft_lines_fwd<m, m, P, 1, 256, 0, 0>(ex, sm.data(), 1, 0, tw.data());   // rows
            fft_lines_fwd<m, m, 1, P, 256, 0, 0>(ex, sm.data(), 1, 0, tw.data());   // columns
            for (int k = 0; k < m; ++k) for (int l = 0; l < m; ++l) out[k * m + l] = sm[Fft1<m>::pi(k) * P + Fft1<m>::pi(l)];
        } else {
            for (int k = 0; k < m; ++k) for (int l = 0; l < m; ++l) sm[Fft1<m>::pi(k) * P + Fft1<m>::pi(l)] = in[k * m + l];
            fft_lines_inv<m, m, 1, P, 256, 0, 0>(ex, sm.data(), 1, 0, tw.data());
            fft_lines_inv<m, m, P, 1, 256, 0, 0>(ex, sm.data(), 1, 0, tw.data());
            for (int r = 0; r < m; ++r) for (int c = 0; c < m; ++c) out[r * m + c] = sm[Fft1<m>::pos_s(r) * P + Fft1<m>::pos_s(c)];
        }
    };
    switch (M) {
#define CASE(m) case m: run(std::integral_constant<int, m>{}); return 0;
        CASE(6) CASE(10) CASE(12) CASE(18) CASE(20) CASE(24) CASE(34) CASE(36) CASE(40) CASE(48) CASE(64) CASE(68) CASE(72)
        CASE(80) CASE(96) CASE(136) CASE(144) CASE(160) CASE(288) CASE(576)
#undef CASE
    }
    g_err = "emu_fft2: unsupported size";
    return -2;
}

// Full cascade for nsig signals of H x W; writes maps [nsig][K][HOUT][HOUT] and (if non-NULL) feats [nsig][2][K].
int emu_forward(int N, int J, int L, int max_order, int H, int W, const float* psi_hat,
                const float* phi_hat, const float* x, int nsig, float* maps_out, float* feats_out) {
#define CFG(n, j) if (N == n && J == j) return run_cfg<Cfg<n, j>>(L, max_order, H, W, psi_hat, phi_hat, x, nsig, maps_out, feats_out, g_err);
#define CFGG(n, j) if (N == n && J == j) return run_cfg<Cfg<n, j, 256, true>>(L, max_order, H, W, psi_hat, phi_hat, x, nsig, maps_out, feats_out, g_err);
#include WST_EMU_CONFIG_FILE
#undef CFG
#undef CFGG
    g_err = "emu_forward: unsupported (N, J)";
    return -2;
}

// The plan's filter bank on the CPU: same per-point formula (wst_filters.h) and the same dense
// fp64 separable DFT + combine as the CUDA kernels in wst_lib.cu.
int emu_filter_bank(int N, int J, int L, float* psi_hat, float* phi_hat) {
    const int nwave = J * L, nf = 2 * nwave + 1;
    const size_t NN = (size_t)N * N;
    std::vector<GaborParams> gp(nf);
    bank_gabors(J, L, gp.data());
    std::vector<double> wr(N), wi(N);
    for (int t = 0; t < N; ++t) { double a = -2.0 * 3.14159265358979323846 * t / N; wr[t] = cos(a); wi[t] = sin(a); }
    std::vector<double> sr(NN), si(NN), tr(NN), ti(NN);
    std::vector<double> allr((size_t)nf * NN), alli((size_t)nf * NN);
    for (int f = 0; f < nf; ++f) {
        for (int x = 0; x < N; ++x) for (int y = 0; y < N; ++y) gabor_point(gp[f], x, y, N, N, sr[(size_t)x * N + y], si[(size_t)x * N + y]);
        for (int k = 0; k < N; ++k) for (int y = 0; y < N; ++y) {
            double ar = 0, ai = 0;
            for (int x = 0; x < N; ++x) { int t = (k * x) % N; double vr = sr[(size_t)x * N + y], vi = si[(size_t)x * N + y];
                ar += wr[t] * vr - wi[t] * vi; ai += wr[t] * vi + wi[t] * vr; }
            tr[(size_t)k * N + y] = ar; ti[(size_t)k * N + y] = ai;
        }
        for (int k = 0; k < N; ++k) for (int l = 0; l < N; ++l) {
            double ar = 0, ai = 0;
            for (int y = 0; y < N; ++y) { int t = (l * y) % N; double vr = tr[(size_t)k * N + y], vi = ti[(size_t)k * N + y];
                ar += wr[t] * vr - wi[t] * vi; ai += wr[t] * vi + wi[t] * vr; }
            allr[f * NN + (size_t)k * N + l] = ar; alli[f * NN + (size_t)k * N + l] = ai;
        }
    }
    for (int n = 0; n < nwave; ++n) {
        const double* wR = &allr[(size_t)(2 * n) * NN]; const double* wI = &alli[(size_t)(2 * n) * NN];
        const double* mR = &allr[(size_t)(2 * n + 1) * NN]; const double* mI = &alli[(size_t)(2 * n + 1) * NN];
        double den = mR[0] * mR[0] + mI[0] * mI[0];
        double kr = (wR[0] * mR[0] + wI[0] * mI[0]) / den, ki = (wI[0] * mR[0] - wR[0] * mI[0]) / den;
        for (size_t i = 0; i < NN; ++i) psi_hat[n * NN + i] = (float)(wR[i] - (kr * mR[i] - ki * mI[i]));
    }
    for (size_t i = 0; i < NN; ++i) phi_hat[i] = (float)allr[(size_t)(2 * nwave) * NN + i];
    return 0;
}

int emu_query(int N, int J, int* smem_bytes, int* gp, int* hout) {
#define CFGQ(C_) { using C = C_; *smem_bytes = (int)C::smem_bytes(); *hout = C::HOUT; \
        for (int i = 0; i < C::J; ++i) gp[i] = 0; \
        static_for<0, C::J>([&](auto Jc) { gp[decltype(Jc)::value] = C::GP(decltype(Jc)::value); }); return 0; }
#define CFG(n, j) if (N == n && J == j) CFGQ(Cfg<n COMMA j>)
#define CFGG(n, j) if (N == n && J == j) CFGQ(Cfg<n COMMA j COMMA 256 COMMA true>)
#define COMMA ,
#include WST_EMU_CONFIG_FILE
#undef CFG
#undef CFGG
    return -2;
}

}  // extern "C"
