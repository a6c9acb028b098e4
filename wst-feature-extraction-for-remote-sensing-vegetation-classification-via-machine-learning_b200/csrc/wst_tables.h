// wst_tables.h — host-side construction of the device tables of a plan (pure C++, no CUDA).
//
// Input: the full-resolution Fourier-domain filter bank (psi^[J*L][N][N], phi^[N][N], real fp32,
// laid out like kymatio's filters['psi'][n]['levels'][0] / filters['phi']['levels'][0],
// SURVEY.md Appendix A.2).  Output: one flat float buffer holding
//   * per-level FFT twiddles,
//   * per-level separable low-pass operators Gr, Gc (phi^ level -> spatial kernel -> kept outputs),
//   * psi^ periodised to every level it is used at (the corner crop of kymatio's
//     periodize_filter_fft), theta-interleaved in the group sizes the kernel consumes.
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>
#include "wst_cascade.h"

namespace wst {

struct TableOffsets {              // offsets in floats into the flat buffer
    size_t tw[kMaxJ], gr[kMaxJ], gc[kMaxJ], psi1[kMaxJ], psi2[kMaxJ][kMaxJ];
    size_t total;
    int bb1[kMaxJ][kMaxL][2], bb2[kMaxPairs][kMaxL][2];    // support bounding boxes (copied into PlanTables)
    float lpw[kMaxJ][kLpTaps];                             // banded low-pass taps (copied into PlanTables)
    double support_fraction;      // visited filter entries / all filter entries (diagnostic)
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <class C>
inline bool build_tables(int L, const float* psi_hat, const float* phi_hat,
                         std::vector<float>& buf, TableOffsets& off, std::string& err) {
    constexpr int N = C::N, J = C::J, HP = C::HP, HOUT = C::HOUT, NS = C::NS;
    const double kTwoPi = 6.283185307179586476925286766559;
    std::memset(&off, 0, sizeof(off));
    size_t cur = 0;
    auto reserve = [&](size_t nfloats) { size_t o = cur; cur = align_up(cur + nfloats, 64); return o; };

    for (int j = 0; j < J; ++j) {
        int m = N >> j;
        off.tw[j] = reserve(2 * (size_t)m);
        off.gr[j] = reserve((size_t)m * HP);
        off.gc[j] = reserve((size_t)m * HP);
    }
    auto groups = [](int Lv, int g) { return (Lv + g - 1) / g; };
    int gp[kMaxJ], g2[kMaxJ][kMaxJ];
    static_for<0, J>([&](auto Jc) {
        constexpr int j = decltype(Jc)::value;
        gp[j] = C::GP(j);
        static_for<j + 1, J>([&](auto J2c) { constexpr int j2 = decltype(J2c)::value; g2[j2][j] = C::G2(j, j2); });
    });
    int vw1[kMaxJ], vw2[kMaxJ][kMaxJ];          // interleave width = theta-group size
    for (int j = 0; j < J; ++j) { vw1[j] = gp[j]; for (int j2 = j + 1; j2 < J; ++j2) vw2[j2][j] = g2[j2][j]; }
    for (int j = 0; j < J; ++j) off.psi1[j] = reserve((size_t)groups(L, vw1[j]) * vw1[j] * N * N);
    for (int j2 = 1; j2 < J; ++j2)
        for (int j1 = 0; j1 < j2; ++j1) {
            int m = N >> j1;
            off.psi2[j2][j1] = reserve((size_t)groups(L, vw2[j2][j1]) * vw2[j2][j1] * m * m);
        }
    off.total = cur;
    buf.assign(cur, 0.0f);

    // ---- twiddles: tw[k1*R2 + i2] = exp(-2 pi i * i2 * k1 / m)
    static_for<0, J>([&](auto Jc) {
        constexpr int j = decltype(Jc)::value;
        constexpr int m = C::msize(j);
        constexpr int R1 = Fft1<m>::R1, R2 = Fft1<m>::R2;
        float* t = buf.data() + off.tw[j];
        for (int k1 = 0; k1 < R1; ++k1)
            for (int i2 = 0; i2 < R2; ++i2) {
                double a = -kTwoPi * (double)((i2 * k1) % m) / (double)m;
                t[2 * (k1 * R2 + i2)] = (float)std::cos(a);
                t[2 * (k1 * R2 + i2) + 1] = (float)std::sin(a);
            }
    });

    // ---- separable low-pass operators
    const double phi00 = (double)phi_hat[0];
    if (!(phi00 > 0.0)) { err = "phi_hat[0,0] must be positive"; return false; }
    const double rs = 1.0 / std::sqrt(phi00);
    {   // separability check: phi^[k][l] ~= a[k]*b[l]
        double worst = 0.0;
        for (int k = 0; k < N; ++k)
            for (int l = 0; l < N; ++l) {
                double d = std::fabs((double)phi_hat[(size_t)k * N + l]
                                     - (double)phi_hat[(size_t)k * N] * rs * (double)phi_hat[l] * rs);
                if (d > worst) worst = d;
            }
        if (worst > 2e-6 * phi00) { err = "low-pass filter is not separable (max dev " + std::to_string(worst) + ")"; return false; }
    }
    bool ok = true;
    static_for<0, J>([&](auto Jc) {
        constexpr int j = decltype(Jc)::value;
        constexpr int m = C::msize(j);
        const int s = m / NS;
        for (int dim = 0; dim < 2; ++dim) {
            std::vector<double> a(m), g(m);
            for (int k = 0; k < m; ++k) {
                int kk = (k < m / 2) ? k : N - m + k;
                // square grid: one operator serves rows and columns (mean of phi^'s first column and first row)
                a[k] = 0.5 * ((double)phi_hat[(size_t)kk * N] + (double)phi_hat[kk]) * rs;
            }
            for (int x = 0; x < m; ++x) {
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc += a[k] * std::cos(kTwoPi * (double)((k * x) % m) / (double)m);
                g[x] = acc / (double)m;
            }
            if constexpr (lp_banded(m, HOUT, j)) {                     // taps of the banded evaluation (same g)
                constexpr int R = lp_radius(m / NS);
                for (int d = 0; d <= R; ++d) off.lpw[j][d] = (float)g[d];
                for (int d = R + 1; d < m - R; ++d)
                    if (std::fabs(g[d]) > 1e-7 * std::fabs(g[0])) {
                        err = "low-pass kernel of level " + std::to_string(j) + " is wider than its compiled band";
                        ok = false;
                    }
            }
            float* G = buf.data() + (dim == 0 ? off.gr[j] : off.gc[j]);
            for (int x = 0; x < m; ++x)                            // rows of G follow the spatial storage order
                for (int i = 0; i < HOUT; ++i) {
                    int idx = (((i + 1) * s - x) % m + m) % m;
                    G[(size_t)Fft1<m>::pos_s(x) * HP + i] = (float)g[idx];
                }
        }
    });

    if (!ok) return false;

    // ---- wavelets, periodised per level (planar), with their supports
    // smallest cyclic interval of [0, n) covering all flagged positions, packed lo << 16 | len
    auto cyclic_span = [](const std::vector<char>& on) -> int {
        int n = (int)on.size(), first = -1, cnt = 0;
        for (int i = 0; i < n; ++i) if (on[i]) { if (first < 0) first = i; ++cnt; }
        if (cnt == 0) return 0;
        if (cnt == n) return n;                                   // lo = 0, len = n
        int best_gap = -1, best_end = 0, prev = -1, last = 0;     // largest run of unflagged positions (cyclic)
        for (int i = 0; i < n; ++i) if (on[i]) last = i;
        prev = last - n;
        for (int i = 0; i < n; ++i) if (on[i]) {
            int gap = i - prev - 1;
            if (gap > best_gap) { best_gap = gap; best_end = i; }
            prev = i;
        }
        int lo = best_end, len = n - best_gap;
        return (lo << 16) | len;
    };
    double visited = 0.0, all = 0.0;
    auto fill = [&](size_t ofilt, int (*bb)[2], int j, int res, int vw) {
        int m = N >> res;
        float* dst = buf.data() + ofilt;
        int nvg = groups(L, vw);
        for (int vg = 0; vg < nvg; ++vg) {
            std::vector<char> rowon(m, 0), colon(m, 0);
            for (int e = 0; e < vw; ++e) {
                int theta = vg * vw + e;
                if (theta >= L) continue;                          // padding orientations stay zero
                const float* src = psi_hat + (size_t)(j * L + theta) * N * N;
                float mx = 0.f;
                for (int k = 0; k < m; ++k) {
                    int kk = (k < m / 2) ? k : N - m + k;
                    for (int l = 0; l < m; ++l) {
                        int ll = (l < m / 2) ? l : N - m + l;
                        float v = src[(size_t)kk * N + ll];
                        dst[(((size_t)vg * m + k) * m + l) * vw + e] = v;
                        if (std::fabs(v) > mx) mx = std::fabs(v);
                    }
                }
                const float thr = kSupportEps * mx;
                for (int k = 0; k < m; ++k)
                    for (int l = 0; l < m; ++l)
                        if (std::fabs(dst[(((size_t)vg * m + k) * m + l) * vw + e]) > thr) { rowon[k] = 1; colon[l] = 1; }
            }
            bb[vg][0] = cyclic_span(rowon);
            bb[vg][1] = cyclic_span(colon);
            visited += (double)(bb[vg][0] & 0xffff) * (double)(bb[vg][1] & 0xffff) * vw;
            all += (double)m * m * vw;
        }
    };
    if (L > kMaxL) { err = "L > " + std::to_string(kMaxL) + " orientations is not supported by the compiled cascades"; return false; }
    std::memset(off.bb1, 0, sizeof(off.bb1));
    std::memset(off.bb2, 0, sizeof(off.bb2));
    for (int j = 0; j < J; ++j) fill(off.psi1[j], off.bb1[j], j, 0, vw1[j]);
    for (int j2 = 1; j2 < J; ++j2)
        for (int j1 = 0; j1 < j2; ++j1) fill(off.psi2[j2][j1], off.bb2[pair_index(j2, j1)], j2, j1, vw2[j2][j1]);
    off.support_fraction = all > 0 ? visited / all : 1.0;
    return true;
}

template <class C>
inline void bind_tables(PlanTables& pt, const float* base, const TableOffsets& off) {
    std::memcpy(pt.bb1, off.bb1, sizeof(pt.bb1));
    std::memcpy(pt.bb2, off.bb2, sizeof(pt.bb2));
    std::memcpy(pt.lpw, off.lpw, sizeof(pt.lpw));
    for (int j = 0; j < C::J; ++j) {
        pt.tw[j] = reinterpret_cast<const cfloat*>(base + off.tw[j]);
        pt.gr[j] = base + off.gr[j];
        pt.gc[j] = base + off.gc[j];
        pt.psi1[j] = base + off.psi1[j];
        for (int j1 = 0; j1 < j; ++j1) pt.psi2[j][j1] = base + off.psi2[j][j1];
    }
}

// kymatio geometry
inline int padded_size(int M, int J) { return ((M + (1 << J)) / (1 << J) + 1) * (1 << J); }
inline int num_coefficients(int J, int L, int max_order) {
    int K = 1 + L * J;
    if (max_order >= 2) K += L * L * J * (J - 1) / 2;
    return K;
}

}  // namespace wst
