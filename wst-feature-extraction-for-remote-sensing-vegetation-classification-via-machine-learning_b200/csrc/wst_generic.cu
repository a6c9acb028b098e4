// wst_generic.cu — shape-generic scattering engine: kymatio's cascade as batched
// DFT-matrix products (see wst_generic.h).  Kernels:
//
//   pad_kernel        reflect padding of the H x W signals (float32 planes, uint8 HWC pixels or tiles of a raster)
//   gemm_kernel<...>  C[b] = epilogue( (A[b] (.) filter[b]) . B[b] ), complex or real operands, operand shared by the
//                     batch or per batch item, fp32 SIMT (4x4 register tile per thread) or 3xTF32 tensor cores
//                     (mma.sync.m16n8k8, both operands split hi + lo, three products per term, fp32 accumulators)
//   pool_kernel       per-coefficient mean / population std (train_and_save_model.py:371-372), one warp per map
//
// All intermediates of a chunk of signals live in one stream-ordered workspace; nothing here is on the CPU.
#include "wst_generic.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "wst_tables.h"

namespace wst {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, BMP = 72, GEMM_THREADS = 256;   // BMP: padded tile pitch (floats)

struct GemmDesc {
    const float* A; const float* B; const float* F; float* C;
    int M, N, K;
    long long a_sm, a_sk, b_sk, b_sn, c_sm, c_sn, f_sm, f_sk;     // element strides (a complex element is one unit)
    int bd1, bd2;                                                // batch b -> (b / (bd1*bd2), (b / bd2) % bd1, b % bd2)
    long long a_b[3], b_b[3], c_b[3], f_b[3];                    // batch strides per index, in elements
};

__device__ __forceinline__ unsigned f2tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// d += a * b with a = ah + al, b = bh + bl split into TF32 halves: ah*bh + ah*bl + al*bh (al*bl is below fp32 rounding)
__device__ __forceinline__ void mma_3x(float (&d)[4], const unsigned (&ah)[4], const unsigned (&al)[4],
                                       const unsigned (&bh)[2], const unsigned (&bl)[2]) {
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bl);
    mma_tf32(d, ah, bh);
}

// AC / BC: operand is complex.  EPI: 0 store (complex if either operand is, else real), 1 modulus (real out).
// FILT: A elements are multiplied by a real filter on load.  BKFAST: B's K index is the contiguous one.
// TC: 3xTF32 tensor-core inner product instead of fp32 FMAs.
template <bool AC, bool BC, int EPI, bool FILT, bool BKFAST, bool TC>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_kernel(const GemmDesc d) {
    constexpr bool CC = AC || BC;
    __shared__ __align__(16) float As[2][BK][BMP];      // [re / im][k][m]
    __shared__ __align__(16) float Bs[2][BK][BMP];      // [re / im][k][n]
    const int tid = threadIdx.x;
    const long long b = blockIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.z * BN;
    const long long i0 = b / ((long long)d.bd1 * d.bd2), i1 = (b / d.bd2) % d.bd1, i2 = b % d.bd2;
    const float* A = d.A + (AC ? 2 : 1) * (i0 * d.a_b[0] + i1 * d.a_b[1] + i2 * d.a_b[2]);
    const float* B = d.B + (BC ? 2 : 1) * (i0 * d.b_b[0] + i1 * d.b_b[1] + i2 * d.b_b[2]);
    const float* F = FILT ? d.F + (i0 * d.f_b[0] + i1 * d.f_b[1] + i2 * d.f_b[2]) : nullptr;
    float* C = d.C + ((CC && EPI == 0) ? 2 : 1) * (i0 * d.c_b[0] + i1 * d.c_b[1] + i2 * d.c_b[2]);

    // SIMT: thread (ty, tx) owns the 4 x 4 outputs at rows ty*4.., columns tx*4..
    // TC:   warp (wm, wn) owns 32 x 16 outputs = 2 x 2 mma tiles of 16 x 8
    const int ty = tid >> 4, tx = tid & 15;
    const int lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3, g = lane >> 2, t = lane & 3;
    float accr[4][4], acci[4][4];                      // SIMT: [i][j]; TC: [mt*2+nt][c0..c3]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { accr[i][j] = 0.f; acci[i][j] = 0.f; }

    for (int k0 = 0; k0 < d.K; k0 += BK) {
        // ---- stage the A tile (k fastest in memory for every caller) and the B tile
#pragma unroll
        for (int it = 0; it < BM * BK / GEMM_THREADS; ++it) {
            const int e = tid + it * GEMM_THREADS, kk = e % BK, mm = e / BK;
            const int m = m0 + mm, k = k0 + kk;
            float re = 0.f, im = 0.f;
            if (m < d.M && k < d.K) {
                const long long o = (long long)m * d.a_sm + (long long)k * d.a_sk;
                if constexpr (AC) { const float2 v = *reinterpret_cast<const float2*>(A + 2 * o); re = v.x; im = v.y; }
                else re = A[o];
                if constexpr (FILT) { const float f = F[(long long)m * d.f_sm + (long long)k * d.f_sk]; re *= f; im *= f; }
            }
            As[0][kk][mm] = re;
            if constexpr (AC) As[1][kk][mm] = im;
        }
#pragma unroll
        for (int it = 0; it < BN * BK / GEMM_THREADS; ++it) {
            const int e = tid + it * GEMM_THREADS;
            const int kk = BKFAST ? e % BK : e / BN, nn = BKFAST ? e / BK : e % BN;
            const int n = n0 + nn, k = k0 + kk;
            float re = 0.f, im = 0.f;
            if (n < d.N && k < d.K) {
                const long long o = (long long)k * d.b_sk + (long long)n * d.b_sn;
                if constexpr (BC) { const float2 v = *reinterpret_cast<const float2*>(B + 2 * o); re = v.x; im = v.y; }
                else re = B[o];
            }
            Bs[0][kk][nn] = re;
            if constexpr (BC) Bs[1][kk][nn] = im;
        }
        __syncthreads();
        if constexpr (!TC) {
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                const float4 ar4 = *reinterpret_cast<const float4*>(&As[0][kk][ty * 4]);
                const float4 br4 = *reinterpret_cast<const float4*>(&Bs[0][kk][tx * 4]);
                const float ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w}, br[4] = {br4.x, br4.y, br4.z, br4.w};
                float ai[4] = {0.f, 0.f, 0.f, 0.f}, bi[4] = {0.f, 0.f, 0.f, 0.f};
                if constexpr (AC) { const float4 v = *reinterpret_cast<const float4*>(&As[1][kk][ty * 4]); ai[0] = v.x; ai[1] = v.y; ai[2] = v.z; ai[3] = v.w; }
                if constexpr (BC) { const float4 v = *reinterpret_cast<const float4*>(&Bs[1][kk][tx * 4]); bi[0] = v.x; bi[1] = v.y; bi[2] = v.z; bi[3] = v.w; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        accr[i][j] = fmaf(ar[i], br[j], accr[i][j]);
                        if constexpr (AC && BC) accr[i][j] = fmaf(-ai[i], bi[j], accr[i][j]);
                        if constexpr (BC) acci[i][j] = fmaf(ar[i], bi[j], acci[i][j]);
                        if constexpr (AC) acci[i][j] = fmaf(ai[i], br[j], acci[i][j]);
                    }
            }
        } else {
#pragma unroll
            for (int k8 = 0; k8 < BK; k8 += 8) {
                unsigned arh[2][4], arl[2][4], aih[2][4], ail[2][4], nih[2][4], nil_[2][4];
                unsigned brh[2][2], brl[2][2], bih[2][2], bil[2][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const int mb = wm * 32 + mt * 16;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int kk = k8 + t + (q >> 1) * 4, mm = mb + g + (q & 1) * 8;
                        const float v = As[0][kk][mm];
                        arh[mt][q] = f2tf32(v); arl[mt][q] = f2tf32(v - __uint_as_float(arh[mt][q]));
                        if constexpr (AC) {
                            const float w = As[1][kk][mm];
                            aih[mt][q] = f2tf32(w); ail[mt][q] = f2tf32(w - __uint_as_float(aih[mt][q]));
                            nih[mt][q] = aih[mt][q] ^ 0x80000000u; nil_[mt][q] = ail[mt][q] ^ 0x80000000u;
                        }
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 2; ++nt) {
                    const int nb = wn * 16 + nt * 8;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int kk = k8 + t + q * 4, nn = nb + g;
                        const float v = Bs[0][kk][nn];
                        brh[nt][q] = f2tf32(v); brl[nt][q] = f2tf32(v - __uint_as_float(brh[nt][q]));
                        if constexpr (BC) {
                            const float w = Bs[1][kk][nn];
                            bih[nt][q] = f2tf32(w); bil[nt][q] = f2tf32(w - __uint_as_float(bih[nt][q]));
                        }
                    }
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 2; ++nt) {
                        mma_3x(accr[mt * 2 + nt], arh[mt], arl[mt], brh[nt], brl[nt]);
                        if constexpr (AC && BC) mma_3x(accr[mt * 2 + nt], nih[mt], nil_[mt], bih[nt], bil[nt]);
                        if constexpr (BC) mma_3x(acci[mt * 2 + nt], arh[mt], arl[mt], bih[nt], bil[nt]);
                        if constexpr (AC) mma_3x(acci[mt * 2 + nt], aih[mt], ail[mt], brh[nt], brl[nt]);
                    }
            }
        }
        __syncthreads();
    }

    auto store = [&](int m, int n, float re, float im) {
        if (m >= d.M || n >= d.N) return;
        const long long o = (long long)m * d.c_sm + (long long)n * d.c_sn;
        if constexpr (EPI == 1) C[o] = sqrtf(re * re + im * im);
        else if constexpr (CC) *reinterpret_cast<float2*>(C + 2 * o) = make_float2(re, im);
        else C[o] = re;
    };
    if constexpr (!TC) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) store(m0 + ty * 4 + i, n0 + tx * 4 + j, accr[i][j], acci[i][j]);
    } else {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    store(m0 + wm * 32 + mt * 16 + g + (q >> 1) * 8, n0 + wn * 16 + nt * 8 + 2 * t + (q & 1),
                          accr[mt * 2 + nt][q], acci[mt * 2 + nt][q]);
    }
}

// z[s][r][c] = x_s reflect-padded (np.pad mode='reflect', kymatio Pad)
__global__ void pad_kernel(const InputDesc in, long long s0, long long nsig, int H, int W, int Hp, int Wp, int top, int left,
                           float* __restrict__ z) {
    const long long total = nsig * Hp * Wp;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(o % Wp), r = (int)((o / Wp) % Hp);
        const long long s = o / ((long long)Wp * Hp);
        int sc = c - left; sc = sc < 0 ? -sc : (sc >= W ? 2 * (W - 1) - sc : sc);
        int sr = r - top; sr = sr < 0 ? -sr : (sr >= H ? 2 * (H - 1) - sr : sr);
        z[o] = signal_source(in, s0 + s, H, W).at(sr, sc);
    }
}

// feats[s][0][k] = mean, feats[s][1][k] = population std of maps[s][k][:] — one warp per map, two passes
__global__ void pool_kernel(const float* __restrict__ maps, long long nmaps, int K, int npix, float* __restrict__ feats) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nmaps) return;
    const float* p = maps + w * npix;
    float s = 0.f;
    for (int i = lane; i < npix; i += 32) s += p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)npix;
    float v = 0.f;
    for (int i = lane; i < npix; i += 32) { const float dlt = p[i] - mean; v += dlt * dlt; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
        const long long sidx = w / K; const int k = (int)(w % K);
        feats[(sidx * 2 + 0) * K + k] = mean;
        feats[(sidx * 2 + 1) * K + k] = sqrtf(v / (float)npix);
    }
}

template <bool AC, bool BC, int EPI, bool FILT, bool BKFAST>
cudaError_t launch_gemm(const GemmDesc& d, long long nbatch, bool tc, cudaStream_t st) {
    if (nbatch <= 0 || d.M <= 0 || d.N <= 0) return cudaSuccess;
    dim3 grid((unsigned)nbatch, (unsigned)((d.M + BM - 1) / BM), (unsigned)((d.N + BN - 1) / BN));
    if (tc) gemm_kernel<AC, BC, EPI, FILT, BKFAST, true><<<grid, GEMM_THREADS, 0, st>>>(d);
    else gemm_kernel<AC, BC, EPI, FILT, BKFAST, false><<<grid, GEMM_THREADS, 0, st>>>(d);
    return cudaGetLastError();
}

}  // namespace

// ------------------------------------------------------------------------------------------------ plan
struct GenericPlan {
    int device = 0, H = 0, W = 0, J = 0, L = 0, max_order = 2, engine = kEngineSimt;
    int Hp = 0, Wp = 0, K = 0, h = 0, w = 0, top = 0, left = 0;
    float* d_tab = nullptr;
    // offsets (floats) into d_tab
    size_t fr[kMaxJ], fc[kMaxJ];                    // forward DFT matrices of level j: [Hj][Hj], [Wj][Wj] complex
    size_t ar[kMaxJ][kMaxJ], ac[kMaxJ][kMaxJ];      // [jc][jp] partial inverse DFT: [Hjc][Hjp], [Wjc][Wjp] complex
    size_t gr[kMaxJ], gc[kMaxJ];                    // low-pass operators of level j: [h][Hj], [w][Wj] real
    size_t psi1[kMaxJ];                             // scale j at level 0: [L][Hp][Wp] real
    size_t psi2[kMaxJ][kMaxJ];                      // [j2][j1]: scale j2 periodised to level j1: [L][Hj1][Wj1] real
    size_t floats_per_signal = 0;                   // workspace
    int sms = 148;
    cudaMemPool_t pool = nullptr;                   // stream-ordered pool the workspace comes from (NULL: the device's default)
};

int generic_create(GenericPlan** out, int device, int H, int W, int J, int L, int max_order, int engine,
                   const float* psi_hat, const float* phi_hat, cudaMemPool_t pool, std::string& err) {
    const double kTwoPi = 6.283185307179586476925286766559;
    GenericPlan* p = new GenericPlan();
    p->device = device; p->H = H; p->W = W; p->J = J; p->L = L; p->max_order = max_order; p->engine = engine;
    p->pool = pool;
    const int Hp = padded_size(H, J), Wp = padded_size(W, J);
    p->Hp = Hp; p->Wp = Wp; p->K = num_coefficients(J, L, max_order);
    p->h = (Hp >> J) - 2; p->w = (Wp >> J) - 2;
    p->top = (Hp - H) / 2; p->left = (Wp - W) / 2;
    if (p->h < 1 || p->w < 1) { err = "empty output"; delete p; return -1; }
    // kymatio's Pad reflects without repeating the edge sample: the pad must be smaller than the image
    if (Hp - H - p->top >= H || Wp - W - p->left >= W) { err = "padding is not smaller than the image (kymatio's special case is not supported)"; delete p; return -2; }
    cudaDeviceGetAttribute(&p->sms, cudaDevAttrMultiProcessorCount, device);

    std::vector<float> buf;
    auto reserve = [&](size_t n) { size_t o = buf.size(); buf.resize(o + ((n + 63) / 64) * 64, 0.f); return o; };
    auto dft = [&](size_t off, int rows, int cols, int period, double sign, double scale) {
        // M[i][k] = scale * exp(sign * 2 pi i * i*k / period)
        for (int i = 0; i < rows; ++i)
            for (int k = 0; k < cols; ++k) {
                const double a = sign * kTwoPi * (double)(((long long)i * k) % period) / (double)period;
                buf[off + 2 * ((size_t)i * cols + k)] = (float)(scale * std::cos(a));
                buf[off + 2 * ((size_t)i * cols + k) + 1] = (float)(scale * std::sin(a));
            }
    };
    const bool second = max_order >= 2;
    for (int j = 0; j < J; ++j) {
        const int Hj = Hp >> j, Wj = Wp >> j;
        if (j == 0 || (second && j < J - 1)) {
            p->fr[j] = reserve(2 * (size_t)Hj * Hj); dft(p->fr[j], Hj, Hj, Hj, -1.0, 1.0);
            p->fc[j] = reserve(2 * (size_t)Wj * Wj); dft(p->fc[j], Wj, Wj, Wj, -1.0, 1.0);
        }
    }
    // partial inverse DFT from the level-jp grid to the level-jc grid: x[i * 2^(jc-jp)] = (1/n_p) sum_k X[k] e^{+2 pi i k i / n_c}
    for (int jc = 0; jc < J; ++jc)
        for (int jp = 0; jp <= jc; ++jp) {
            if (jp > 0 && !(second && jp < jc)) continue;          // order 1 uses jp = 0; order 2 uses 0 < ... jp < jc
            const int Hc = Hp >> jc, Hq = Hp >> jp, Wc = Wp >> jc, Wq = Wp >> jp;
            p->ar[jc][jp] = reserve(2 * (size_t)Hc * Hq); dft(p->ar[jc][jp], Hc, Hq, Hc, +1.0, 1.0 / Hq);
            p->ac[jc][jp] = reserve(2 * (size_t)Wc * Wq); dft(p->ac[jc][jp], Wc, Wq, Wc, +1.0, 1.0 / Wq);
        }
    // separable low-pass: phi^[k][l] = a[k] b[l] (an isotropic Gaussian)
    const double phi00 = (double)phi_hat[0];
    if (!(phi00 > 0.0)) { err = "phi_hat[0,0] must be positive"; delete p; return -1; }
    const double rs = 1.0 / std::sqrt(phi00);
    for (int j = 0; j < J; ++j) {
        const int s = 1 << (J - j);
        for (int dim = 0; dim < 2; ++dim) {
            const int n0 = dim == 0 ? Hp : Wp, m = n0 >> j, nout = dim == 0 ? p->h : p->w;
            std::vector<double> a(m), gk(m);
            for (int k = 0; k < m; ++k) {
                const int kk = (k < m / 2) ? k : n0 - m + k;     // corner crop of kymatio's periodize_filter_fft
                a[k] = (double)(dim == 0 ? phi_hat[(size_t)kk * Wp] : phi_hat[kk]) * rs;
            }
            for (int x = 0; x < m; ++x) {
                double acc = 0.0;
                for (int k = 0; k < m; ++k) acc += a[k] * std::cos(kTwoPi * (double)(((long long)k * x) % m) / (double)m);
                gk[x] = acc / (double)m;
            }
            const size_t off = reserve((size_t)nout * m);
            (dim == 0 ? p->gr[j] : p->gc[j]) = off;
            for (int i = 0; i < nout; ++i)
                for (int x = 0; x < m; ++x) buf[off + (size_t)i * m + x] = (float)gk[(((i + 1) * s - x) % m + m) % m];
        }
    }
    auto crop = [&](size_t off, int j, int res) {
        const int mh = Hp >> res, mw = Wp >> res;
        for (int t = 0; t < L; ++t) {
            const float* src = psi_hat + (size_t)(j * L + t) * Hp * Wp;
            for (int k = 0; k < mh; ++k) {
                const int kk = (k < mh / 2) ? k : Hp - mh + k;
                for (int l = 0; l < mw; ++l) {
                    const int ll = (l < mw / 2) ? l : Wp - mw + l;
                    buf[off + ((size_t)t * mh + k) * mw + l] = src[(size_t)kk * Wp + ll];
                }
            }
        }
    };
    for (int j = 0; j < J; ++j) { p->psi1[j] = reserve((size_t)L * Hp * Wp); crop(p->psi1[j], j, 0); }
    if (second)
        for (int j2 = 1; j2 < J; ++j2)
            for (int j1 = 0; j1 < j2; ++j1) {
                p->psi2[j2][j1] = reserve((size_t)L * (Hp >> j1) * (Wp >> j1));
                crop(p->psi2[j2][j1], j2, j1);
            }
    if (cudaMalloc(&p->d_tab, buf.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(p->d_tab, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        err = std::string("generic table upload: ") + cudaGetErrorString(cudaGetLastError());
        cudaFree(p->d_tab); delete p; return -3;
    }
    // workspace per signal (floats): z0, two level-0 complex arrays, the order-1 arrays of one scale, the children of
    // one (j1, j2) pair, and the maps when the caller does not keep them
    const size_t HW = (size_t)Hp * Wp, LL = (size_t)L * L;
    size_t f = HW + 2 * HW + 2 * HW;                                   // z0, row-pass temp, U0^
    f += 2 * L * HW + L * HW + L * (size_t)Hp * p->w + 2 * L * HW + 2 * L * HW;      // V, U1, P, T1, U1^
    if (second && J > 1) f += 2 * LL * HW / 2 + LL * HW / 4 + LL * (size_t)(Hp / 2) * p->w;   // W, U2, P2
    f += (size_t)p->K * p->h * p->w;
    p->floats_per_signal = f + 64 * 16;
    *out = p;
    return 0;
}

void generic_destroy(GenericPlan* p) {
    if (!p) return;
    cudaFree(p->d_tab);
    delete p;
}

void generic_geometry(const GenericPlan* p, int* K, int* h, int* w, int* Hp, int* Wp) {
    if (K) *K = p->K;
    if (h) *h = p->h;
    if (w) *w = p->w;
    if (Hp) *Hp = p->Hp;
    if (Wp) *Wp = p->Wp;
}

namespace {
long long chunk_signals(const GenericPlan* p) {
    size_t budget = (size_t)3 << 29;                         // 1.5 GB of workspace
    if (const char* ev = getenv("WST_GENERIC_WORKSPACE_MB")) budget = (size_t)atoll(ev) << 20;
    long long s = (long long)(budget / (p->floats_per_signal * sizeof(float)));
    const long long cap = 60000 / ((long long)p->L * p->L);      // keeps every batch count far below 2^31 blocks
    if (s > cap) s = cap;
    return s < 1 ? 1 : s;
}
long long launches_per_chunk(const GenericPlan* p) {
    long long n = 1 + 2 + 2;                                 // pad, U0^, S0
    for (int j1 = 0; j1 < p->J; ++j1) {
        n += 2 + 2;                                          // order-1 product + inverse, low-pass
        if (p->max_order >= 2 && j1 < p->J - 1) n += 2 + 4 * (p->J - 1 - j1);   // U1^, then per j2: product, inverse, low-pass x2
    }
    return n;
}
}  // namespace

long long generic_launch_count(const GenericPlan* p, long long nsig) {
    if (nsig <= 0) return 0;
    const long long cs = chunk_signals(p), chunks = (nsig + cs - 1) / cs;
    return chunks * launches_per_chunk(p) + chunks;          // + pooling per chunk (counted even when only maps are wanted)
}

cudaError_t generic_forward(const GenericPlan* p, const InputDesc& in, long long nsig, float* feats, float* maps,
                            cudaStream_t st, std::string& err) {
    if (nsig <= 0) return cudaSuccess;
    const int J = p->J, L = p->L, Hp = p->Hp, Wp = p->Wp, h = p->h, w = p->w, K = p->K;
    const bool tc = p->engine == kEngineTf32x3, second = p->max_order >= 2;
    const long long CS = std::min<long long>(chunk_signals(p), nsig);
    float* ws = nullptr;
    // from the plan's private pool: its blocks stay cached across synchronisations (the default pool gives memory back
    // to the driver at every sync, which costs a 1.5 GB allocation per host-path call)
    const size_t ws_bytes = (size_t)CS * p->floats_per_signal * sizeof(float);
    cudaError_t e = p->pool ? cudaMallocFromPoolAsync((void**)&ws, ws_bytes, p->pool, st) : cudaMallocAsync(&ws, ws_bytes, st);
    if (e != cudaSuccess) { err = "cudaMallocAsync(generic workspace)"; return e; }
    const size_t HW = (size_t)Hp * Wp, LL = (size_t)L * L;
    auto al = [](size_t v) { return (v + 63) / 64 * 64; };
    size_t cur = 0;
    auto take = [&](size_t per_signal) { float* q = ws + cur; cur += al((size_t)CS * per_signal); return q; };
    float* z0 = take(HW);
    float* t0 = take(2 * HW);
    float* u0 = take(2 * HW);
    float* v = take(2 * L * HW);
    float* u1 = take(L * HW);
    float* pp = take(L * (size_t)Hp * w);
    float* t1 = take(2 * L * HW);
    float* u1h = take(2 * L * HW);
    float *wv = nullptr, *u2 = nullptr, *p2 = nullptr;
    if (second && J > 1) { wv = take(2 * LL * HW / 2); u2 = take(LL * HW / 4); p2 = take(LL * (size_t)(Hp / 2) * w); }
    float* mtmp = maps ? nullptr : take((size_t)K * h * w);
    const float* T = p->d_tab;
    const size_t map_sz = (size_t)h * w;

#define GEN_TRY(call) do { e = (call); if (e != cudaSuccess) { err = #call; cudaFreeAsync(ws, st); return e; } } while (0)
    auto zero = [](GemmDesc& d) { std::memset(&d, 0, sizeof d); d.bd1 = 1; d.bd2 = 1; };
    // low-pass of `count` real arrays [mh][mw] (contiguous) into maps: out index = i0*c0 + i1*c1 + i2*c2 (+ base), in
    // map units, with the batch split (bd1, bd2)
    auto lowpass = [&](const float* src, float* tmp, long long count, int j, float* mbase, long long c0, long long c1,
                       long long c2, int bd1, int bd2) -> cudaError_t {
        const int mh = Hp >> j, mw = Wp >> j;
        GemmDesc d; zero(d);                                   // P[b][r][i'] = sum_c U[b][r][c] Gc[i'][c]
        d.A = src; d.B = T + p->gc[j]; d.C = tmp; d.M = mh; d.N = w; d.K = mw;
        d.a_sm = mw; d.a_sk = 1; d.b_sk = 1; d.b_sn = mw; d.c_sm = w; d.c_sn = 1;
        d.a_b[2] = (long long)mh * mw; d.c_b[2] = (long long)mh * w; d.bd2 = 1 << 30;
        cudaError_t ee = launch_gemm<false, false, 0, false, true>(d, count, false, st);
        if (ee != cudaSuccess) return ee;
        zero(d);                                               // S[b][i][i'] = sum_r Gr[i][r] P[b][r][i']
        d.A = T + p->gr[j]; d.B = tmp; d.C = mbase; d.M = h; d.N = w; d.K = mh;
        d.a_sm = mh; d.a_sk = 1; d.b_sk = w; d.b_sn = 1; d.c_sm = w; d.c_sn = 1;
        d.bd1 = bd1; d.bd2 = bd2;
        d.b_b[0] = (long long)bd1 * bd2 * mh * w; d.b_b[1] = (long long)bd2 * mh * w; d.b_b[2] = (long long)mh * w;
        d.c_b[0] = c0 * (long long)map_sz; d.c_b[1] = c1 * (long long)map_sz; d.c_b[2] = c2 * (long long)map_sz;
        return launch_gemm<false, false, 0, false, false>(d, count, false, st);
    };

    for (long long s0 = 0; s0 < nsig; s0 += CS) {
        const long long S = std::min<long long>(CS, nsig - s0);
        float* mout = maps ? maps + (size_t)s0 * K * map_sz : mtmp;
        {
            const long long total = S * (long long)HW;
            const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)p->sms * 16);
            pad_kernel<<<blocks, 256, 0, st>>>(in, s0, S, p->H, p->W, Hp, Wp, p->top, p->left, z0);
            GEN_TRY(cudaGetLastError());
        }
        GemmDesc d;
        // U0^ = Fr . z0 . Fc^T : rows (real in), then columns
        zero(d); d.A = z0; d.B = T + p->fc[0]; d.C = t0; d.M = Hp; d.N = Wp; d.K = Wp;
        d.a_sm = Wp; d.a_sk = 1; d.b_sk = 1; d.b_sn = Wp; d.c_sm = Wp; d.c_sn = 1;
        d.a_b[2] = HW; d.c_b[2] = HW; d.bd2 = 1 << 30;
        GEN_TRY((launch_gemm<false, true, 0, false, true>(d, S, tc, st)));
        zero(d); d.A = T + p->fr[0]; d.B = t0; d.C = u0; d.M = Hp; d.N = Wp; d.K = Hp;
        d.a_sm = Hp; d.a_sk = 1; d.b_sk = Wp; d.b_sn = 1; d.c_sm = Wp; d.c_sn = 1;
        d.b_b[2] = HW; d.c_b[2] = HW; d.bd2 = 1 << 30;
        GEN_TRY((launch_gemm<true, true, 0, false, false>(d, S, tc, st)));
        // S0
        GEN_TRY(lowpass(z0, pp, S, 0, mout, 0, 0, K, 1, 1 << 30));

        for (int j1 = 0; j1 < J; ++j1) {
            const int H1 = Hp >> j1, W1 = Wp >> j1;
            const size_t HW1 = (size_t)H1 * W1;
            // V[s][t][k][c] = sum_l (U0^[s][k][l] psi1[t][k][l]) Ac[c][l]
            zero(d); d.A = u0; d.F = T + p->psi1[j1]; d.B = T + p->ac[j1][0]; d.C = v; d.M = Hp; d.N = W1; d.K = Wp;
            d.a_sm = Wp; d.a_sk = 1; d.f_sm = Wp; d.f_sk = 1; d.b_sk = 1; d.b_sn = Wp; d.c_sm = W1; d.c_sn = 1;
            d.bd1 = 1; d.bd2 = L; d.a_b[1] = HW; d.f_b[2] = HW;
            d.c_b[1] = (long long)L * Hp * W1; d.c_b[2] = (long long)Hp * W1;
            d.bd1 = 1 << 30;
            GEN_TRY((launch_gemm<true, true, 0, true, true>(d, S * L, tc, st)));
            // U1[s][t][r][c] = | sum_k Ar[r][k] V[s][t][k][c] |
            zero(d); d.A = T + p->ar[j1][0]; d.B = v; d.C = u1; d.M = H1; d.N = W1; d.K = Hp;
            d.a_sm = Hp; d.a_sk = 1; d.b_sk = W1; d.b_sn = 1; d.c_sm = W1; d.c_sn = 1;
            d.b_b[2] = (long long)Hp * W1; d.c_b[2] = HW1; d.bd2 = 1 << 30;
            GEN_TRY((launch_gemm<true, true, 1, false, false>(d, S * L, tc, st)));
            // S1 -> maps[s][1 + j1*L + t]
            GEN_TRY(lowpass(u1, pp, S * L, j1, mout + (size_t)(1 + j1 * L) * map_sz, 0, K, 1, 1 << 20, L));
            if (!(second && j1 < J - 1)) continue;
            // U1^ = Fr . U1 . Fc^T
            zero(d); d.A = u1; d.B = T + p->fc[j1]; d.C = t1; d.M = H1; d.N = W1; d.K = W1;
            d.a_sm = W1; d.a_sk = 1; d.b_sk = 1; d.b_sn = W1; d.c_sm = W1; d.c_sn = 1;
            d.a_b[2] = HW1; d.c_b[2] = HW1; d.bd2 = 1 << 30;
            GEN_TRY((launch_gemm<false, true, 0, false, true>(d, S * L, tc, st)));
            zero(d); d.A = T + p->fr[j1]; d.B = t1; d.C = u1h; d.M = H1; d.N = W1; d.K = H1;
            d.a_sm = H1; d.a_sk = 1; d.b_sk = W1; d.b_sn = 1; d.c_sm = W1; d.c_sn = 1;
            d.b_b[2] = HW1; d.c_b[2] = HW1; d.bd2 = 1 << 30;
            GEN_TRY((launch_gemm<true, true, 0, false, false>(d, S * L, tc, st)));
            int o2 = 1 + J * L;                                 // first order-2 coefficient of parents at scale j1
            for (int j = 0; j < j1; ++j) o2 += L * L * (J - 1 - j);
            for (int j2 = j1 + 1; j2 < J; ++j2) {
                const int H2 = Hp >> j2, W2 = Wp >> j2;
                // W[s][t1][t2][k][c] = sum_l (U1^[s][t1][k][l] psi2[t2][k][l]) Ac[c][l]
                zero(d); d.A = u1h; d.F = T + p->psi2[j2][j1]; d.B = T + p->ac[j2][j1]; d.C = wv;
                d.M = H1; d.N = W2; d.K = W1;
                d.a_sm = W1; d.a_sk = 1; d.f_sm = W1; d.f_sk = 1; d.b_sk = 1; d.b_sn = W1; d.c_sm = W2; d.c_sn = 1;
                d.bd1 = 1 << 30; d.bd2 = L; d.a_b[1] = HW1; d.f_b[2] = HW1;
                d.c_b[1] = (long long)L * H1 * W2; d.c_b[2] = (long long)H1 * W2;
                GEN_TRY((launch_gemm<true, true, 0, true, true>(d, S * L * L, tc, st)));
                // U2 = | Ar . W |
                zero(d); d.A = T + p->ar[j2][j1]; d.B = wv; d.C = u2; d.M = H2; d.N = W2; d.K = H1;
                d.a_sm = H1; d.a_sk = 1; d.b_sk = W2; d.b_sn = 1; d.c_sm = W2; d.c_sn = 1;
                d.b_b[2] = (long long)H1 * W2; d.c_b[2] = (long long)H2 * W2; d.bd2 = 1 << 30;
                GEN_TRY((launch_gemm<true, true, 1, false, false>(d, S * L * L, tc, st)));
                // S2 -> maps[s][o2 + t1*L*(J-1-j1) + (j2-j1-1)*L + t2]
                GEN_TRY(lowpass(u2, p2, S * L * L, j2, mout + (size_t)(o2 + (j2 - j1 - 1) * L) * map_sz,
                                K, (long long)L * (J - 1 - j1), 1, L, L));
            }
        }
        if (feats) {
            const long long nmaps = S * K;
            pool_kernel<<<(unsigned)((nmaps * 32 + 255) / 256), 256, 0, st>>>(mout, nmaps, K, h * w, feats + (size_t)s0 * 2 * K);
            GEN_TRY(cudaGetLastError());
        }
    }
#undef GEN_TRY
    cudaFreeAsync(ws, st);
    return cudaSuccess;
}

}  // namespace wst
