// wst_noise.cu — the five noise models of the robustness sweep on the GPU.
//
// Replaces add_gaussian_noise / add_salt_and_pepper_noise / add_speckle_noise / add_poisson_noise /
// add_uniform_noise (src/preprocessing/add_noise.py:14-72) for batches of uint8 [B][H][W][C] images that are
// already resident on the device, so that BASELINE configs[3] (noise -> WST features -> Random Forest) needs no
// host pass: the output feeds wst2d_forward_u8 directly.
//
// Two entry points share one arithmetic core:
//   wst2d_add_noise_draws  the caller supplies the random draws (what numpy's global RNG produced for the
//                          reference); the result is then the reference's, bit for bit — float64 arithmetic
//                          with the reference's operation order and no FMA contraction, clip, truncating cast;
//   wst2d_add_noise        draws come from a counter-based generator (Philox4x32-10 keyed by the seed, counter =
//                          element index), same distributions as the reference's numpy calls; host RNG streams
//                          cannot be matched, so this one is checked statistically.
// The salt-and-pepper model keeps the reference's quirks: ceil(amount * H*W*C * 0.5) coordinates per colour
// (the channel count is in the product), coordinates drawn from [0, H-1) x [0, W-1) (numpy's exclusive upper
// bound: the last row and column are never touched), all channels of a hit pixel set, pepper applied after salt.
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/wst2d.h"

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11)
struct Philox {
    uint32_t k0, k1;
    uint32_t c0, c1, c2, c3;
    uint32_t out[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t index, uint32_t stream)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), c0((uint32_t)index), c1((uint32_t)(index >> 32)), c2(0),
          c3(stream), have(0) {}
    __device__ void block() {
        uint32_t a0 = c0, a1 = c1, a2 = c2, a3 = c3, q0 = k0, q1 = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, a0), lo0 = 0xD2511F53u * a0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, a2), lo1 = 0xCD9E8D57u * a2;
            const uint32_t n0 = hi1 ^ a1 ^ q0, n2 = hi0 ^ a3 ^ q1;
            a0 = n0; a1 = lo1; a2 = n2; a3 = lo0;
            q0 += 0x9E3779B9u; q1 += 0xBB67AE85u;
        }
        out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
        ++c2;                                   // next block of this (index, stream)
        have = 4;
    }
    __device__ uint32_t next32() {
        if (have == 0) block();
        return out[4 - have--];
    }
    // 53-bit uniform in [0, 1)
    __device__ double uniform() {
        const uint32_t hi = next32(), lo = next32();
        return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6)) * (1.0 / 9007199254740992.0);
    }
    __device__ double normal() {                // Box-Muller, one value per pair of uniforms
        const double u1 = 1.0 - uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    }
    // 24-bit uniform in (0, 1) for the rejection tests below
    __device__ float uniformf() { return ((float)(next32() >> 8) + 0.5f) * (1.0f / 16777216.0f); }
    // Poisson(lam): sequential search below 10, Hörmann's transformed rejection (PTRS, 1993) from 10 up — the
    // same two regimes numpy's generator uses.  The draw is an integer decided by accept/reject tests, for which
    // single precision is ample (and twice as fast as the double-precision logarithms and lgamma).
    __device__ long long poisson(double lam_d) {
        if (lam_d <= 0.0) return 0;
        const float lam = (float)lam_d;
        if (lam < 10.0f) {
            const float enlam = __expf(-lam);
            long long k = 0;
            float prod = uniformf();
            while (prod > enlam) { ++k; prod *= uniformf(); }
            return k;
        }
        const float slam = sqrtf(lam), loglam = __logf(lam);
        const float b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
        const float invalpha = 1.1239f + 1.1328f / (b - 3.4f), vr = 0.9277f - 3.6224f / (b - 2.0f);
        for (;;) {
            const float U = uniformf() - 0.5f, V = uniformf();
            const float us = 0.5f - fabsf(U);
            const float kf = floorf((2.0f * a / us + b) * U + lam + 0.43f);
            if (us >= 0.07f && V <= vr) return (long long)kf;
            if (kf < 0.0f || (us < 0.013f && V > us)) continue;
            if (__logf(V) + __logf(invalpha) - __logf(a / (us * us) + b) <= -lam + kf * loglam - lgammaf(kf + 1.0f))
                return (long long)kf;
        }
    }
};

// ---------------------------------------------------------------- the reference's arithmetic (float64, unfused)
__device__ __forceinline__ unsigned char clip_u8(double v) {   // np.clip(v, 0, 255).astype(np.uint8): truncation
    v = v < 0.0 ? 0.0 : (v > 255.0 ? 255.0 : v);
    return (unsigned char)(int)v;
}
__device__ __forceinline__ unsigned char apply_additive(unsigned char p, double d) {        // add_noise.py:20-21, :72-73
    return clip_u8(__dadd_rn((double)p, d));
}
__device__ __forceinline__ unsigned char apply_speckle(unsigned char p, double g, double factor) {   // :52-54
    const double a = (double)p;
    return clip_u8(__dadd_rn(a, __dmul_rn(__dmul_rn(a, g), factor)));
}
__device__ __forceinline__ double poisson_lambda(unsigned char p, double scale) {            // :61
    return __ddiv_rn(__dmul_rn((double)p, scale), 255.0);
}
__device__ __forceinline__ unsigned char apply_poisson(long long k, double scale) {          // :64-65
    return clip_u8(__ddiv_rn(__dmul_rn((double)k, 255.0), scale));
}

struct Params {
    int kind;
    double sigma;        // gaussian: intensity * 255 / 100
    double factor;       // speckle: intensity / 100
    double scale;        // poisson: 10 + (intensity / 100) * 90
    double lo, width;    // uniform: low = -range/2, width = high - low
};

template <bool DRAWS>
__global__ void __launch_bounds__(kThreads)
noise_elementwise(Params pr, const unsigned char* __restrict__ img, long long n, const void* __restrict__ draws,
                  unsigned long long seed, unsigned char* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        const unsigned char p = img[i];
        unsigned char r;
        if (DRAWS) {
            if (pr.kind == WST2D_NOISE_POISSON) r = apply_poisson(static_cast<const long long*>(draws)[i], pr.scale);
            else if (pr.kind == WST2D_NOISE_SPECKLE) r = apply_speckle(p, static_cast<const double*>(draws)[i], pr.factor);
            else r = apply_additive(p, static_cast<const double*>(draws)[i]);
        } else {
            Philox g(seed, (uint64_t)i, (uint32_t)pr.kind);
            if (pr.kind == WST2D_NOISE_GAUSSIAN) r = apply_additive(p, __dmul_rn(pr.sigma, g.normal()));
            else if (pr.kind == WST2D_NOISE_SPECKLE) r = apply_speckle(p, g.normal(), pr.factor);
            else if (pr.kind == WST2D_NOISE_POISSON) r = apply_poisson(g.poisson(poisson_lambda(p, pr.scale)), pr.scale);
            else r = apply_additive(p, __dadd_rn(pr.lo, __dmul_rn(pr.width, g.uniform())));
        }
        out[i] = r;
    }
}

// one thread per coordinate: sets all C channels of pixel (row, col) of image b to `value`
template <bool DRAWS>
__global__ void __launch_bounds__(kThreads)
noise_scatter(unsigned char* __restrict__ out, long long B, int H, int W, int C, long long ncoord, int colour,
              unsigned char value, const long long* __restrict__ coords, unsigned long long seed) {
    const long long total = B * ncoord;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        const long long b = t / ncoord, i = t - b * ncoord;
        long long row, col;
        if (DRAWS) {                                   // coords: [B][2 colours][2 (row, col)][ncoord]
            const long long* cb = coords + ((b * 2 + colour) * 2) * ncoord;
            row = cb[i]; col = cb[ncoord + i];
            if (row < 0 || row >= H || col < 0 || col >= W) continue;
        } else {                                       // np.random.randint(0, H - 1), np.random.randint(0, W - 1)
            Philox g(seed, (uint64_t)t, (uint32_t)(16 + colour));
            row = (long long)(((unsigned long long)g.next32() * (unsigned long long)(H - 1)) >> 32);
            col = (long long)(((unsigned long long)g.next32() * (unsigned long long)(W - 1)) >> 32);
        }
        unsigned char* p = out + ((b * H + row) * W + col) * C;
        for (int c = 0; c < C; ++c) p[c] = value;
    }
}

thread_local std::string g_noise_error;

int run(int device, int kind, double intensity, const uint8_t* img, int64_t B, int H, int W, int C, bool with_draws,
        const void* draws, int64_t ncoord_in, uint64_t seed, uint8_t* out, void* cuda_stream) {
    auto fail = [](int code, const std::string& m) { g_noise_error = m; return code; };
    if (kind < WST2D_NOISE_GAUSSIAN || kind > WST2D_NOISE_UNIFORM) return fail(WST2D_ERR_ARG, "Unknown noise type");
    if (B < 0 || H < 1 || W < 1 || C < 1) return fail(WST2D_ERR_ARG, "invalid shape");
    if (!(intensity >= 0.0 && intensity <= 100.0)) return fail(WST2D_ERR_ARG, "Intensity must be between 0 and 100");
    if (kind == WST2D_NOISE_SALT_AND_PEPPER && (H < 2 || W < 2 || C < 2))
        return fail(WST2D_ERR_ARG, "salt_and_pepper: low >= high (the reference draws coordinates from [0, dim - 1) for every axis)");
    if (B == 0) return WST2D_OK;
    if (!img || !out || (with_draws && !draws)) return fail(WST2D_ERR_ARG, "NULL buffer");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(WST2D_ERR_CUDA, "no such CUDA device");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const long long n = (long long)B * H * W * C;
    auto grid_for = [&](long long work) {
        long long g = (work + kThreads - 1) / kThreads, cap = 8LL * sms;
        return (int)(g < cap ? (g > 0 ? g : 1) : cap);
    };
    cudaError_t e = cudaSuccess;
    if (kind == WST2D_NOISE_SALT_AND_PEPPER) {
        // add_noise.py:33,38: ceil(amount * image.size * 0.5) coordinates for salt, the same count for pepper
        const double amount = intensity / 100.0;
        long long ncoord = (long long)ceil(amount * (double)((long long)H * W * C) * 0.5);
        if (with_draws) {
            if (ncoord_in < 0) { if (prev >= 0 && prev != device) cudaSetDevice(prev); return fail(WST2D_ERR_ARG, "negative coordinate count"); }
            ncoord = ncoord_in;
        }
        if (out != img) e = cudaMemcpyAsync(out, img, (size_t)n, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess && ncoord > 0) {
            const int g = grid_for(B * ncoord);
            if (with_draws) {
                noise_scatter<true><<<g, kThreads, 0, st>>>(out, B, H, W, C, ncoord, 0, 255, static_cast<const long long*>(draws), 0ULL);
                noise_scatter<true><<<g, kThreads, 0, st>>>(out, B, H, W, C, ncoord, 1, 0, static_cast<const long long*>(draws), 0ULL);
            } else {
                noise_scatter<false><<<g, kThreads, 0, st>>>(out, B, H, W, C, ncoord, 0, 255, nullptr, seed);
                noise_scatter<false><<<g, kThreads, 0, st>>>(out, B, H, W, C, ncoord, 1, 0, nullptr, seed);
            }
            e = cudaGetLastError();
        }
    } else {
        Params pr;
        pr.kind = kind;
        pr.sigma = intensity * 255.0 / 100.0;
        pr.factor = intensity / 100.0;
        pr.scale = 10.0 + (intensity / 100.0) * 90.0;
        const double range = intensity * 255.0 / 100.0;
        pr.lo = -range / 2.0;
        pr.width = range / 2.0 - pr.lo;
        const int g = grid_for(n);
        if (with_draws) noise_elementwise<true><<<g, kThreads, 0, st>>>(pr, img, n, draws, 0ULL, out);
        else noise_elementwise<false><<<g, kThreads, 0, st>>>(pr, img, n, nullptr, seed, out);
        e = cudaGetLastError();
    }
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(WST2D_ERR_CUDA, std::string("add_noise: ") + cudaGetErrorString(e));
    return WST2D_OK;
}

}  // namespace

extern "C" {

const char* wst2d_noise_last_error(void) { return g_noise_error.c_str(); }

int wst2d_add_noise(int device, int kind, double intensity, const uint8_t* img_dev, int64_t B, int H, int W, int C,
                    uint64_t seed, uint8_t* out_dev, void* cuda_stream) {
    return run(device, kind, intensity, img_dev, B, H, W, C, false, nullptr, 0, seed, out_dev, cuda_stream);
}

int wst2d_add_noise_draws(int device, int kind, double intensity, const uint8_t* img_dev, int64_t B, int H, int W,
                          int C, const void* draws_dev, int64_t n_coords, uint8_t* out_dev, void* cuda_stream) {
    return run(device, kind, intensity, img_dev, B, H, W, C, true, draws_dev, n_coords, 0, out_dev, cuda_stream);
}

}  // extern "C"
