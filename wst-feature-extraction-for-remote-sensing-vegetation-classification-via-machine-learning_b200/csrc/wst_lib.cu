// wst_lib.cu — CUDA kernels (sm_100a) and the C ABI of libwst_b200.so (include/wst2d.h).
//
// Kernels:
//   cascade_kernel<Cfg>   (wst_cfg_inst.cu, one translation unit per configuration) one persistent CTA per
//                         SM; each CTA runs the whole scattering cascade of one (patch, channel) signal at a
//                         time out of shared memory and pools its maps (mean / population std,
//                         train_and_save_model.py:371-372) before moving on (wst_cascade.h)
//                         (uint8 HWC input is converted in its input stage: load_rgb_image, train...:51-56)
//   gabor_spatial_kernel, dft_axis{0,1}_kernel, combine_filters_kernel
//                         fp64 filter-bank construction, once per plan (wst_filters.h)
//
// There is no CPU fallback anywhere in this file: without a CUDA device plan creation fails.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "wst_tables.h"
#include "wst_filters.h"
#include "wst_ops.h"
#include "wst_generic.h"
#include "../../include/wst2d.h"

using namespace wst;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) { g_last_error = msg; return code; }

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess)                                                                  \
            return fail(WST2D_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));    \
    } while (0)

// ------------------------------------------------------------------------------------------------
// filter bank (fp64, once per plan)
// ------------------------------------------------------------------------------------------------
// M x N grid (x: row index < M, y: column index < N), like kymatio's gabor_2d(M, N, ...)
__global__ void gabor_spatial_kernel(const GaborParams* __restrict__ gp, int nf, int M, int N, double2* out) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    int x = blockIdx.y, f = blockIdx.z;
    if (y >= N || f >= nf) return;
    GaborParams p = gp[f];
    double re, im;
    gabor_point(p, x, y, M, N, re, im);
    out[((size_t)f * M + x) * N + y] = make_double2(re, im);
}

// out[f][k][y] = sum_x W[(k*x) % M] * in[f][x][y],   W[t] = exp(-2 pi i t / M)      (arrays are M x N)
__global__ void dft_axis0_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                 const double2* __restrict__ W, int M, int N) {
    int y = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y, f = blockIdx.z;
    if (y >= N) return;
    const double2* p = in + (size_t)f * M * N + y;
    double ar = 0.0, ai = 0.0;
    int t = 0;
    for (int x = 0; x < M; ++x) {
        double2 w = W[t], v = p[(size_t)x * N];
        ar += w.x * v.x - w.y * v.y;
        ai += w.x * v.y + w.y * v.x;
        t += k; if (t >= M) t -= M;
    }
    out[((size_t)f * M + k) * N + y] = make_double2(ar, ai);
}

// out[f][k][l] = sum_y in[f][k][y] * W[(l*y) % N]
__global__ void dft_axis1_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                 const double2* __restrict__ W, int M, int N) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y, f = blockIdx.z;
    if (l >= N) return;
    const double2* p = in + ((size_t)f * M + k) * N;
    double ar = 0.0, ai = 0.0;
    int t = 0;
    for (int y = 0; y < N; ++y) {
        double2 w = W[t], v = p[y];
        ar += w.x * v.x - w.y * v.y;
        ai += w.x * v.y + w.y * v.x;
        t += l; if (t >= N) t -= N;
    }
    out[((size_t)f * M + k) * N + l] = make_double2(ar, ai);
}

// psi^[n] = Re(W^ - K Wmod^), K = W^[0,0]/Wmod^[0,0];  phi^ = Re(G^)
__global__ void combine_filters_kernel(const double2* __restrict__ spec, int nwave, size_t NN,
                                       float* __restrict__ psi_hat, float* __restrict__ phi_hat) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int n = blockIdx.y;
    if (i >= NN) return;
    if (n < nwave) {
        const double2* w = spec + (size_t)(2 * n) * NN;
        const double2* m = spec + (size_t)(2 * n + 1) * NN;
        double2 a = w[0], b = m[0];
        double den = b.x * b.x + b.y * b.y;
        double kr = (a.x * b.x + a.y * b.y) / den, ki = (a.y * b.x - a.x * b.y) / den;
        psi_hat[(size_t)n * NN + i] = (float)(w[i].x - (kr * m[i].x - ki * m[i].y));
    } else {
        phi_hat[i] = (float)spec[(size_t)(2 * nwave) * NN + i].x;
    }
}

// ------------------------------------------------------------------------------------------------
// configuration dispatch
// ------------------------------------------------------------------------------------------------
const std::vector<CfgOps>& all_ops() {
    static const std::vector<CfgOps> ops = {
#define CFG(n, j) wst_make_ops_##n##_##j(),
#define CFGG(n, j) wst_make_ops_##n##_##j(),
#include WST_CONFIGS_FILE
#undef CFG
#undef CFGG
    };
    return ops;
}
}  // namespace

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct wst2d_plan {
    int device = 0;
    int H = 0, W = 0, J = 0, L = 0, max_order = 0;
    int Hp = 0, Wp = 0, K = 0, hout = 0, wout = 0;   // padded size, coefficients, output map size
    int N = 0;                        // padded side of the compiled (square) cascade; 0 for a generic plan
    int grid_max = 0;                 // persistent grid: SMs x resident CTAs per SM
    const CfgOps* ops = nullptr;      // compiled fused cascade (wst_cascade.h), or
    GenericPlan* gen = nullptr;       // the shape-generic DFT-matrix engine (wst_generic.cu)
    int engine = 0;                   // WST2D_ENGINE_* actually in use
    cudaMemPool_t pool = nullptr;     // private stream-ordered pool of this plan's device (scratch of forward calls)
    PlanTables pt{};
    float* d_tables = nullptr;
    std::vector<float> psi_hat, phi_hat;   // host copies (debug export)
    // optional per-kernel timing (wst2d_profile): event pairs recorded around each launch
    mutable std::mutex prof_mu;
    mutable bool profiling = false;
    mutable std::vector<cudaEvent_t> prof_cascade, prof_pool;   // (start, stop) pairs
    // resources of the host-buffer path (wst2d_forward_host), created on first use and reused:
    // two streams, each with its staging input / feature buffers and its cascade scratch
    struct HostPath {
        std::mutex mu;
        cudaStream_t st[2] = {nullptr, nullptr};
        float* x[2] = {nullptr, nullptr};
        float* f[2] = {nullptr, nullptr};
        cfloat* u0h[2] = {nullptr, nullptr};
        cfloat* ws[2] = {nullptr, nullptr};
        float* maps[2] = {nullptr, nullptr};
        long long cap_sig = 0;        // signals per chunk the buffers are sized for
    };
    mutable HostPath host;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int build_filter_bank_gpu(wst2d_plan* p) {
    const int M = p->Hp, N = p->Wp, J = p->J, L = p->L;
    const int nwave = J * L, nf = 2 * nwave + 1;
    const size_t NN = (size_t)M * N;
    std::vector<GaborParams> gp(nf);
    bank_gabors(J, L, gp.data());
    auto roots = [](int n) {
        std::vector<double2> W(n);
        for (int t = 0; t < n; ++t) {
            double a = -2.0 * 3.14159265358979323846 * (double)t / (double)n;
            W[t] = make_double2(std::cos(a), std::sin(a));
        }
        return W;
    };
    const std::vector<double2> W = roots(N), WM = roots(M);
    GaborParams* d_gp = nullptr; double2 *d_a = nullptr, *d_b = nullptr, *d_W = nullptr, *d_WM = nullptr;
    float *d_psi = nullptr, *d_phi = nullptr;
    auto cleanup = [&]() { cudaFree(d_gp); cudaFree(d_a); cudaFree(d_b); cudaFree(d_W); cudaFree(d_WM); cudaFree(d_psi); cudaFree(d_phi); };
#define TRYC(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { cleanup(); \
        return fail(WST2D_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)
    TRYC(cudaMalloc(&d_gp, nf * sizeof(GaborParams)));
    TRYC(cudaMalloc(&d_a, nf * NN * sizeof(double2)));
    TRYC(cudaMalloc(&d_b, nf * NN * sizeof(double2)));
    TRYC(cudaMalloc(&d_W, N * sizeof(double2)));
    TRYC(cudaMalloc(&d_WM, M * sizeof(double2)));
    TRYC(cudaMalloc(&d_psi, nwave * NN * sizeof(float)));
    TRYC(cudaMalloc(&d_phi, NN * sizeof(float)));
    TRYC(cudaMemcpy(d_gp, gp.data(), nf * sizeof(GaborParams), cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(d_W, W.data(), N * sizeof(double2), cudaMemcpyHostToDevice));
    TRYC(cudaMemcpy(d_WM, WM.data(), M * sizeof(double2), cudaMemcpyHostToDevice));
    dim3 blk(128), grd((N + 127) / 128, M, nf);
    gabor_spatial_kernel<<<grd, blk>>>(d_gp, nf, M, N, d_a);
    dft_axis0_kernel<<<grd, blk>>>(d_a, d_b, d_WM, M, N);
    dft_axis1_kernel<<<grd, blk>>>(d_b, d_a, d_W, M, N);
    dim3 cg((unsigned)((NN + 255) / 256), nwave + 1);
    combine_filters_kernel<<<cg, 256>>>(d_a, nwave, NN, d_psi, d_phi);
    TRYC(cudaGetLastError());
    p->psi_hat.resize(nwave * NN);
    p->phi_hat.resize(NN);
    TRYC(cudaMemcpy(p->psi_hat.data(), d_psi, nwave * NN * sizeof(float), cudaMemcpyDeviceToHost));
    TRYC(cudaMemcpy(p->phi_hat.data(), d_phi, NN * sizeof(float), cudaMemcpyDeviceToHost));
#undef TRYC
    cleanup();
    return WST2D_OK;
}

void prof_mark(const wst2d_plan* p, std::vector<cudaEvent_t>& v, cudaStream_t st) {
    if (!p->profiling) return;
    std::lock_guard<std::mutex> lk(p->prof_mu);
    cudaEvent_t e;
    if (cudaEventCreate(&e) == cudaSuccess) { cudaEventRecord(e, st); v.push_back(e); }
}

// peak-rate probe for the roofline denominator: 8 independent FMA chains per thread
__global__ void fma_peak_kernel(float* out, int iters) {
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float b = 0.999f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// One cascade launch over nsig signals.  The kernel pools each signal's maps itself; when the caller does not
// want the maps they live in a per-CTA scratch (grid x K*h*w floats, L2-resident) instead of HBM.
// own_*: caller-provided scratch (host path) sized for grid_max CTAs; when NULL the scratch is stream-ordered
// (cudaMallocAsync from the device's default pool).
// One private stream-ordered memory pool per device, shared by the plans of that device: forward calls take their
// scratch from it (cudaMallocFromPoolAsync) and it keeps freed blocks cached (release threshold = max) so that
// steady-state calls never reach cudaMalloc.  The device's default pool is left untouched.
cudaMemPool_t device_pool(int device) {
    static std::mutex mu;
    static std::vector<cudaMemPool_t> pools;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)pools.size() <= device) pools.resize(device + 1, nullptr);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        pools[device] = pool;
    }
    return pools[device];
}

template <class T>
cudaError_t pool_alloc(const wst2d_plan* p, T** ptr, size_t bytes, cudaStream_t st) {
    return p->pool ? cudaMallocFromPoolAsync((void**)ptr, bytes, p->pool, st) : cudaMallocAsync((void**)ptr, bytes, st);
}

InputDesc plain_input(const void* ptr, int u8_channels) {
    InputDesc in{};
    in.ptr = ptr; in.mode = u8_channels > 0 ? 1 : 0; in.C = u8_channels > 0 ? u8_channels : 1;
    return in;
}

// Signals of the ragged last wave that are better run as a split small batch (0: none, or not worth it).
long long tail_signals(const wst2d_plan* p, long long nsig) {
    if (p->gen || !p->ops->can_split || nsig <= p->grid_max || getenv("WST_NO_SPLIT") || getenv("WST_NO_TAIL_SPLIT")) return 0;
    // (over many waves the ticket scheduler already evens the CTAs out and a last wave is a small fraction of the run:
    // at 84 waves the second launch measured no gain, so it is kept for batches of up to 16 waves)
    if (nsig > 16ll * p->grid_max) return 0;
    const long long r = nsig % p->grid_max;
    return (r > 0 && r * 2 <= p->grid_max) ? r : 0;
}

int forward_impl(const wst2d_plan* p, const InputDesc& in, long long nsig, float* feats_dev,
                 float* maps_dev, cudaStream_t st, cfloat* own_u0h = nullptr, float* own_maps = nullptr,
                 cfloat* own_ws = nullptr, bool mark = true) {
    if (nsig == 0) return WST2D_OK;
    if (p->gen) {
        std::string err;
        prof_mark(p, p->prof_cascade, st);
        cudaError_t ge = generic_forward(p->gen, in, nsig, feats_dev, maps_dev, st, err);
        prof_mark(p, p->prof_cascade, st);
        if (ge != cudaSuccess) return fail(WST2D_ERR_CUDA, "generic engine: " + err + ": " + cudaGetErrorString(ge));
        return WST2D_OK;
    }
    const size_t map_elems = (size_t)p->K * p->hout * p->wout;
    const size_t u0h_elems = (size_t)p->N * (p->N / 2 + 1);
    const size_t ws_elems = p->ops->workspace_cfloats;
    // Ragged last wave: every signal costs the same and the grid is persistent, so nsig = q * grid + r with a small r
    // would keep grid - r CTAs idle for a whole signal time.  The r tail signals are run as a small batch instead — a
    // second launch on the same stream in which each of them is shared among grid / r CTAs (cascade_split_kernel).
    const long long tail = tail_signals(p, nsig);
    if (tail > 0) {
        const long long head = nsig - tail;
        if (mark) prof_mark(p, p->prof_cascade, st);       // one (start, stop) pair around both launches
        int rc2 = forward_impl(p, in, head, feats_dev, maps_dev, st, own_u0h, own_maps, own_ws, false);
        if (rc2 == WST2D_OK) {
            InputDesc rest = in;
            rest.sig0 += head;
            rc2 = forward_impl(p, rest, tail, feats_dev + (size_t)head * 2 * p->K, maps_dev ? maps_dev + (size_t)head * map_elems : nullptr,
                               st, own_u0h, own_maps, own_ws, false);
        }
        if (mark) prof_mark(p, p->prof_cascade, st);
        return rc2;
    }
    int grid = (int)(nsig < p->grid_max ? nsig : p->grid_max);
    // small batch: share each signal among `split` CTAs (first-order groups are independent), up to one CTA per group
    int split = 1;
    if (p->ops->can_split && nsig * 2 <= p->grid_max && !getenv("WST_NO_SPLIT")) {
        split = (int)(p->grid_max / nsig);
        const int units = p->ops->num_units(p->L, p->max_order);
        if (split > units) split = units;
        if (split < 1) split = 1;
        grid = (int)nsig * split;
    }
    cfloat* d_u0h = own_u0h; float* d_maps = own_maps; cfloat* d_ws = own_ws;
    int* d_done = nullptr;
    cudaError_t e = cudaSuccess;
    if (!d_u0h) e = pool_alloc(p, &d_u0h, (size_t)grid * u0h_elems * sizeof(cfloat), st);
    if (e == cudaSuccess && ws_elems && !d_ws) e = pool_alloc(p, &d_ws, (size_t)grid * ws_elems * sizeof(cfloat), st);
    if (e == cudaSuccess && !maps_dev && !d_maps)
        e = pool_alloc(p, &d_maps, (size_t)(split > 1 ? nsig : grid) * map_elems * sizeof(float), st);
    // split > 1: per-signal completion counters; otherwise the ticket counter that hands out the signals after each
    // CTA's first one (more than one wave only; WST_STATIC_SCHED keeps the fixed stride, for A/B runs)
    const bool tickets = split == 1 && nsig > grid && !getenv("WST_STATIC_SCHED");
    if (e == cudaSuccess && (split > 1 || tickets)) {
        const size_t n = split > 1 ? (size_t)nsig : 1;
        e = pool_alloc(p, &d_done, n * sizeof(int), st);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_done, 0, n * sizeof(int), st);
    }
    int rc = WST2D_OK;
    if (e != cudaSuccess) {
        rc = fail(WST2D_ERR_CUDA, std::string("stream-ordered scratch allocation: ") + cudaGetErrorString(e));
    } else {
        if (mark) prof_mark(p, p->prof_cascade, st);
        e = p->ops->launch(p->pt, in, nsig, d_u0h, d_ws, maps_dev, maps_dev ? nullptr : d_maps, feats_dev, grid, st, split, d_done);
        if (mark) prof_mark(p, p->prof_cascade, st);
        if (e != cudaSuccess) rc = fail(WST2D_ERR_CUDA, std::string("cascade launch: ") + cudaGetErrorString(e));
    }
    if (!own_u0h && d_u0h) cudaFreeAsync(d_u0h, st);
    if (!own_ws && d_ws) cudaFreeAsync(d_ws, st);
    if (!own_maps && d_maps) cudaFreeAsync(d_maps, st);
    if (d_done) cudaFreeAsync(d_done, st);
    return rc;
}

}  // namespace

extern "C" {

const char* wst2d_last_error(void) { return g_last_error.c_str(); }
const char* wst2d_version(void) { return "wst_b200 0.1 (sm_100a)"; }

int wst2d_plan_create_ex(wst2d_plan** out, int device, int H, int W, int J, int L, int max_order, int engine) {
    if (!out) return fail(WST2D_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (H <= 0 || W <= 0 || J < 1 || J >= kMaxJ || L < 1 || L > 64 || (max_order != 1 && max_order != 2))
        return fail(WST2D_ERR_ARG, "invalid plan arguments");
    if (engine < WST2D_ENGINE_AUTO || engine > WST2D_ENGINE_GEMM_TF32X3) return fail(WST2D_ERR_ARG, "unknown engine");
    if ((1 << J) > H || (1 << J) > W)
        return fail(WST2D_ERR_ARG, "The smallest dimension should be larger than 2^J.");
    if (const char* ev = getenv("WST_ENGINE")) {           // tuning / A-B runs: "fft", "gemm", "gemm_tf32x3"
        if (engine == WST2D_ENGINE_AUTO) {
            if (!strcmp(ev, "gemm")) engine = WST2D_ENGINE_GEMM_SIMT;
            else if (!strcmp(ev, "gemm_tf32x3")) engine = WST2D_ENGINE_GEMM_TF32X3;
            else if (!strcmp(ev, "fft")) engine = WST2D_ENGINE_FFT;
        }
    }
    const int Hp = padded_size(H, J), Wp = padded_size(W, J);
    const CfgOps* ops = nullptr;
    if (Hp == Wp && L <= kMaxL && (engine == WST2D_ENGINE_AUTO || engine == WST2D_ENGINE_FFT))
        for (const CfgOps& o : all_ops()) if (o.N == Hp && o.J == J) { ops = &o; break; }
    if (!ops && engine == WST2D_ENGINE_FFT) {
        char buf[160];
        snprintf(buf, sizeof buf, "no compiled cascade for padded size %dx%d, J=%d, L=%d", Hp, Wp, J, L);
        return fail(WST2D_ERR_UNSUPPORTED, buf);
    }
    if ((Hp - H + 1) / 2 >= H || (Wp - W + 1) / 2 >= W)
        return fail(WST2D_ERR_UNSUPPORTED, "reflect padding as wide as the image (H or W equal to 2^J) is not supported");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(WST2D_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(WST2D_ERR_ARG, "device index out of range");
    DeviceGuard guard(device);

    wst2d_plan* p = new wst2d_plan();
    p->device = device; p->H = H; p->W = W; p->J = J; p->L = L; p->max_order = max_order;
    p->Hp = Hp; p->Wp = Wp; p->K = num_coefficients(J, L, max_order);
    p->hout = (Hp >> J) - 2; p->wout = (Wp >> J) - 2;
    p->N = ops ? Hp : 0; p->ops = ops;
    p->engine = ops ? WST2D_ENGINE_FFT : (engine == WST2D_ENGINE_GEMM_TF32X3 ? WST2D_ENGINE_GEMM_TF32X3 : WST2D_ENGINE_GEMM_SIMT);
    p->pool = device_pool(device);

    int rc = build_filter_bank_gpu(p);
    if (rc != WST2D_OK) { delete p; return rc; }

    if (!ops) {
        std::string err;
        int grc = generic_create(&p->gen, device, H, W, J, L, max_order,
                                 p->engine == WST2D_ENGINE_GEMM_TF32X3 ? kEngineTf32x3 : kEngineSimt,
                                 p->psi_hat.data(), p->phi_hat.data(), p->pool, err);
        if (grc != 0) { delete p; return fail(grc == -2 ? WST2D_ERR_UNSUPPORTED : grc == -3 ? WST2D_ERR_CUDA : WST2D_ERR_ARG, err); }
        *out = p;
        return WST2D_OK;
    }

    std::vector<float> buf; TableOffsets off; std::string err;
    if (!ops->build(L, p->psi_hat.data(), p->phi_hat.data(), buf, off, err)) { delete p; return fail(WST2D_ERR_ARG, err); }
    if (cudaMalloc(&p->d_tables, buf.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(p->d_tables, buf.data(), buf.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        std::string m = cudaGetErrorString(cudaGetLastError());
        cudaFree(p->d_tables); delete p;
        return fail(WST2D_ERR_CUDA, "table upload: " + m);
    }
    p->pt.L = L; p->pt.max_order = max_order; p->pt.K = p->K;
    p->pt.H = H; p->pt.W = W; p->pt.pad_top = (Hp - H) / 2; p->pt.pad_left = (Wp - W) / 2;
    ops->bind(p->pt, p->d_tables, off);

    int slots = 0;
    e = ops->max_slots(device, &slots);
    if (e != cudaSuccess || slots < 1) {
        std::string m = e != cudaSuccess ? cudaGetErrorString(e) : "kernel does not fit on the device";
        cudaFree(p->d_tables); delete p;
        return fail(WST2D_ERR_CUDA, "cascade kernel setup: " + m);
    }
    p->grid_max = slots;
    *out = p;
    return WST2D_OK;
}

int wst2d_plan_create(wst2d_plan** out, int device, int H, int W, int J, int L, int max_order) {
    return wst2d_plan_create_ex(out, device, H, W, J, L, max_order, WST2D_ENGINE_AUTO);
}

int wst2d_plan_grid(const wst2d_plan* p) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    return p->grid_max;
}

int wst2d_plan_engine(const wst2d_plan* p) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    return p->engine;
}

int wst2d_plan_destroy(wst2d_plan* plan) {
    if (!plan) return WST2D_OK;
    DeviceGuard guard(plan->device);
    for (int i = 0; i < 2; ++i) {
        if (plan->host.st[i]) { cudaStreamSynchronize(plan->host.st[i]); cudaStreamDestroy(plan->host.st[i]); }
        cudaFree(plan->host.x[i]); cudaFree(plan->host.f[i]); cudaFree(plan->host.u0h[i]); cudaFree(plan->host.maps[i]);
        cudaFree(plan->host.ws[i]);
    }
    cudaFree(plan->d_tables);
    generic_destroy(plan->gen);
    // give the scratch this plan's forward calls left cached in the device's private pool back to the driver (other plans
    // of the device re-grow it on their next call)
    if (plan->pool) { cudaDeviceSynchronize(); cudaMemPoolTrimTo(plan->pool, 0); }
    delete plan;
    return WST2D_OK;
}

int wst2d_query(const wst2d_plan* p, int* K, int* h, int* w, int* Hp, int* Wp) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    if (K) *K = p->K;
    if (h) *h = p->hout;
    if (w) *w = p->wout;
    if (Hp) *Hp = p->Hp;
    if (Wp) *Wp = p->Wp;
    return WST2D_OK;
}

int wst2d_forward(const wst2d_plan* p, const float* x_dev, int64_t B, int C, float* feats_dev,
                  float* maps_dev, void* cuda_stream) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    if (B < 0 || C < 1) return fail(WST2D_ERR_ARG, "B must be >= 0 and C >= 1");
    if (B == 0) return WST2D_OK;
    if (!x_dev) return fail(WST2D_ERR_ARG, "x_dev is NULL");
    if (!feats_dev && !maps_dev) return fail(WST2D_ERR_ARG, "both outputs are NULL");
    DeviceGuard guard(p->device);
    return forward_impl(p, plain_input(x_dev, 0), (long long)B * C, feats_dev, maps_dev, (cudaStream_t)cuda_stream);
}

int wst2d_forward_u8(const wst2d_plan* p, const uint8_t* x_dev, int64_t B, int C, float* feats_dev,
                     float* maps_dev, void* cuda_stream) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    if (B < 0 || C < 1) return fail(WST2D_ERR_ARG, "B must be >= 0 and C >= 1");
    if (B == 0) return WST2D_OK;
    if (!x_dev) return fail(WST2D_ERR_ARG, "x_dev is NULL");
    if (!feats_dev && !maps_dev) return fail(WST2D_ERR_ARG, "both outputs are NULL");
    DeviceGuard guard(p->device);
    // the cascade's input stage reads the uint8 HWC pixels itself (value / 255, channel stride C)
    return forward_impl(p, plain_input(x_dev, C), (long long)B * C, feats_dev, maps_dev, (cudaStream_t)cuda_stream);
}

int wst2d_forward_scene(const wst2d_plan* p, const float* raster_dev, int C, int Himg, int Wimg, int stride_y,
                        int stride_x, int64_t tile_begin, int64_t tile_count, float* feats_dev, float* maps_dev,
                        void* cuda_stream) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    if (C < 1 || stride_y < 1 || stride_x < 1) return fail(WST2D_ERR_ARG, "C and the tile strides must be positive");
    if (Himg < p->H || Wimg < p->W) return fail(WST2D_ERR_ARG, "raster is smaller than one tile");
    const long long ny = (Himg - p->H) / stride_y + 1, nx = (Wimg - p->W) / stride_x + 1;
    if (tile_begin < 0 || tile_count < 0 || tile_begin + tile_count > ny * nx)
        return fail(WST2D_ERR_ARG, "tile range outside the " + std::to_string(ny) + " x " + std::to_string(nx) + " tile grid");
    if (tile_count == 0) return WST2D_OK;
    if (!raster_dev) return fail(WST2D_ERR_ARG, "raster_dev is NULL");
    if (!feats_dev && !maps_dev) return fail(WST2D_ERR_ARG, "both outputs are NULL");
    DeviceGuard guard(p->device);
    InputDesc in{};
    in.ptr = raster_dev; in.mode = 2; in.C = C; in.Himg = Himg; in.Wimg = Wimg; in.nx = (int)nx;
    in.sy = stride_y; in.sx = stride_x; in.tile0 = tile_begin;
    return forward_impl(p, in, (long long)tile_count * C, feats_dev, maps_dev, (cudaStream_t)cuda_stream);
}

static int forward_host_impl(const wst2d_plan* p, const void* x_host, bool u8, int64_t B, int C, float* feats_host);

int wst2d_forward_host(const wst2d_plan* p, const float* x_host, int64_t B, int C, float* feats_host) {
    return forward_host_impl(p, x_host, false, B, C, feats_host);
}

int wst2d_forward_host_u8(const wst2d_plan* p, const uint8_t* x_host, int64_t B, int C, float* feats_host) {
    return forward_host_impl(p, x_host, true, B, C, feats_host);
}

// x_host: float32 [B][C][H][W], or uint8 [B][H][W][C] (u8): the staging buffers hold either (sized for float32)
static int forward_host_impl(const wst2d_plan* p, const void* x_host, bool u8, int64_t B, int C, float* feats_host) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    if (B < 0 || C < 1) return fail(WST2D_ERR_ARG, "B must be >= 0 and C >= 1");
    if (B == 0) return WST2D_OK;
    if (!x_host || !feats_host) return fail(WST2D_ERR_ARG, "host buffer is NULL");
    DeviceGuard guard(p->device);
    wst2d_plan::HostPath& hp = p->host;
    std::lock_guard<std::mutex> lk(hp.mu);
    const size_t sig_in = (size_t)p->H * p->W, sig_out = (size_t)2 * p->K;
    const size_t map_elems = (size_t)p->K * p->hout * p->wout;
    // chunk: a few persistent-grid waves per copy so that H2D, compute and D2H of neighbouring chunks overlap
    long long chunk_sig = p->gen ? 1024 : (long long)p->grid_max * 6;
    if (const char* ev = getenv("WST_HOST_CHUNK_SIGNALS")) chunk_sig = atoll(ev);
    chunk_sig = chunk_sig / C * C;
    if (chunk_sig < C) chunk_sig = C;
    if (chunk_sig > B * C) chunk_sig = B * C;
    if (hp.cap_sig < chunk_sig) {           // (re)size the cached staging buffers; first call or larger C
        for (int i = 0; i < 2; ++i) {
            if (hp.st[i]) cudaStreamSynchronize(hp.st[i]);
            cudaFree(hp.x[i]); cudaFree(hp.f[i]); cudaFree(hp.u0h[i]); cudaFree(hp.maps[i]); cudaFree(hp.ws[i]);
            hp.x[i] = hp.f[i] = hp.maps[i] = nullptr; hp.u0h[i] = nullptr; hp.ws[i] = nullptr;
        }
        hp.cap_sig = 0;
        for (int i = 0; i < 2; ++i) {
            if (!hp.st[i]) CUDA_TRY(cudaStreamCreateWithFlags(&hp.st[i], cudaStreamNonBlocking));
            CUDA_TRY(cudaMalloc(&hp.x[i], chunk_sig * sig_in * sizeof(float)));
            CUDA_TRY(cudaMalloc(&hp.f[i], chunk_sig * sig_out * sizeof(float)));
            if (p->gen) continue;                  // the generic engine takes its workspace from the plan's pool
            CUDA_TRY(cudaMalloc(&hp.u0h[i], (size_t)p->grid_max * p->N * (p->N / 2 + 1) * sizeof(cfloat)));
            CUDA_TRY(cudaMalloc(&hp.maps[i], (size_t)p->grid_max * map_elems * sizeof(float)));
            if (p->ops->workspace_cfloats)
                CUDA_TRY(cudaMalloc(&hp.ws[i], (size_t)p->grid_max * p->ops->workspace_cfloats * sizeof(cfloat)));
        }
        hp.cap_sig = chunk_sig;
    }
    int rc = WST2D_OK, it = 0;
    const long long nsig = (long long)B * C;
    // The first copy has nothing to overlap with, so the chunks ramp up: one wave, two waves, then chunk_sig.
    long long n = 0;
    for (long long s0 = 0; s0 < nsig && rc == WST2D_OK; s0 += n, ++it) {
        long long want = chunk_sig;
        if (!p->gen && it < 2 && !getenv("WST_HOST_NO_RAMP")) {
            want = (long long)p->grid_max * (it + 1) / C * C;
            if (want < C) want = C;
            if (want > chunk_sig) want = chunk_sig;
        }
        n = nsig - s0 < want ? nsig - s0 : want;
        int i = it & 1;
        const size_t esz = u8 ? 1 : sizeof(float);      // a chunk is a whole number of patches, so uint8 pixels stay HWC-aligned
        cudaError_t e = cudaMemcpyAsync(hp.x[i], static_cast<const char*>(x_host) + (size_t)s0 * sig_in * esz, n * sig_in * esz,
                                        cudaMemcpyHostToDevice, hp.st[i]);
        if (e != cudaSuccess) { rc = fail(WST2D_ERR_CUDA, std::string("H2D: ") + cudaGetErrorString(e)); break; }
        rc = forward_impl(p, plain_input(hp.x[i], u8 ? C : 0), n, hp.f[i], nullptr, hp.st[i], hp.u0h[i], hp.maps[i], hp.ws[i]);
        if (rc != WST2D_OK) break;
        e = cudaMemcpyAsync(feats_host + (size_t)s0 * sig_out, hp.f[i], n * sig_out * sizeof(float),
                            cudaMemcpyDeviceToHost, hp.st[i]);
        if (e != cudaSuccess) { rc = fail(WST2D_ERR_CUDA, std::string("D2H: ") + cudaGetErrorString(e)); break; }
    }
    for (int i = 0; i < 2; ++i) {
        cudaError_t e = cudaStreamSynchronize(hp.st[i]);
        if (e != cudaSuccess && rc == WST2D_OK) rc = fail(WST2D_ERR_CUDA, std::string("forward_host sync: ") + cudaGetErrorString(e));
    }
    return rc;
}

int wst2d_plan_filters(const wst2d_plan* p, float* psi_hat, float* phi_hat) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    if (psi_hat) std::memcpy(psi_hat, p->psi_hat.data(), p->psi_hat.size() * sizeof(float));
    if (phi_hat) std::memcpy(phi_hat, p->phi_hat.data(), p->phi_hat.size() * sizeof(float));
    return WST2D_OK;
}

int wst2d_profile(wst2d_plan* p, int enable) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    std::lock_guard<std::mutex> lk(p->prof_mu);
    p->profiling = enable != 0;
    return WST2D_OK;
}

int wst2d_profile_read(wst2d_plan* p, double* cascade_ms, double* pool_ms, int* cascade_launches) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    DeviceGuard guard(p->device);
    CUDA_TRY(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(p->prof_mu);
    auto drain = [](std::vector<cudaEvent_t>& v, double& ms, int& n) {
        ms = 0.0; n = 0;
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, v[i], v[i + 1]) == cudaSuccess) { ms += t; ++n; }
        }
        for (cudaEvent_t e : v) cudaEventDestroy(e);
        v.clear();
    };
    double cm, pm; int cn, pn;
    drain(p->prof_cascade, cm, cn);
    drain(p->prof_pool, pm, pn);
    if (cascade_ms) *cascade_ms = cm;
    if (pool_ms) *pool_ms = pm;
    if (cascade_launches) *cascade_launches = cn;
    return WST2D_OK;
}

int wst2d_fma_peak(int device, double* tflops) {
    if (!tflops) return fail(WST2D_ERR_ARG, "tflops is NULL");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(WST2D_ERR_CUDA, "no such CUDA device");
    DeviceGuard guard(device);
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float* d = nullptr;
    CUDA_TRY(cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_peak_kernel<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) break;
        float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 128.0 * iters * (double)blocks * threads;   // 128 FMAs per iteration per thread
        if (rep > 0 && fl / (ms * 1e-3) / 1e12 > best) best = fl / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    CUDA_TRY(cudaGetLastError());
    *tflops = best;
    return WST2D_OK;
}

int wst2d_debug_phase_cycles(const wst2d_plan* p, const float* x_dev, int64_t nsig, int64_t* cycles_host,
                             int ntags) {
    if (!p || !x_dev || !cycles_host) return fail(WST2D_ERR_ARG, "NULL argument");
    if (ntags != kNumPhaseTags) return fail(WST2D_ERR_ARG, "ntags must be " + std::to_string(kNumPhaseTags));
    if (nsig <= 0) return fail(WST2D_ERR_ARG, "nsig must be positive");
    if (!p->ops) return fail(WST2D_ERR_UNSUPPORTED, "phase cycles exist for the fused cascades only");
    DeviceGuard guard(p->device);
    const int grid = (int)(nsig < p->grid_max ? nsig : p->grid_max);
    cfloat* d_u0h = nullptr; float* d_maps = nullptr; long long* d_cyc = nullptr; cfloat* d_ws = nullptr;
    CUDA_TRY(cudaMalloc(&d_u0h, (size_t)grid * p->N * (p->N / 2 + 1) * sizeof(cfloat)));
    if (p->ops->workspace_cfloats) CUDA_TRY(cudaMalloc(&d_ws, (size_t)grid * p->ops->workspace_cfloats * sizeof(cfloat)));
    CUDA_TRY(cudaMalloc(&d_maps, (size_t)grid * p->K * p->hout * p->wout * sizeof(float)));
    float* d_feats = nullptr;
    CUDA_TRY(cudaMalloc(&d_feats, (size_t)nsig * 2 * p->K * sizeof(float)));
    CUDA_TRY(cudaMalloc(&d_cyc, kNumPhaseTags * sizeof(long long)));
    cudaError_t e = p->ops->launch_prof(p->pt, plain_input(x_dev, 0), nsig, d_u0h, d_ws, nullptr, d_maps, d_feats, d_cyc, grid, nullptr);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(cycles_host, d_cyc, kNumPhaseTags * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d_u0h); cudaFree(d_maps); cudaFree(d_cyc); cudaFree(d_ws); cudaFree(d_feats);
    if (e != cudaSuccess) return fail(WST2D_ERR_CUDA, std::string("phase profile: ") + cudaGetErrorString(e));
    return WST2D_OK;
}

int wst2d_launch_count(const wst2d_plan* p, int64_t B, int C) {
    if (!p) return fail(WST2D_ERR_ARG, "plan is NULL");
    if (p->gen) return (int)generic_launch_count(p->gen, (long long)B * C);
    if ((long long)B * C <= 0) return 0;
    return tail_signals(p, (long long)B * C) > 0 ? 2 : 1;   // one fused cascade + pooling kernel per forward call (+ one for a ragged last wave)
}

int wst2d_debug_num_phase_tags(void) { return kNumPhaseTags; }

}  // extern "C"
