// wst_common.h — shared scalar/complex helpers, compile-time trigonometry and loop unrolling.
//
// Everything in csrc/*.h is written once and compiled twice:
//   * by nvcc for sm_100a (the product: wst_lib.cu -> libwst_b200.so), and
//   * by g++ as a single-threaded "one CTA at a time" emulation (tests/emu only), so the
//     index arithmetic of the fused kernels can be checked against the oracle in the
//     CPU-only container.  The emulation is test infrastructure, never a product path.
#pragma once
#include <cstdint>
#include <cstddef>
#include <type_traits>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WST_HD __host__ __device__ __forceinline__
#define WST_D __device__ __forceinline__
#define WST_D_NOINLINE __device__ __noinline__
#define WST_CX __host__ __device__ constexpr
#else
#include <cmath>
#define WST_HD inline
#define WST_D inline
#define WST_D_NOINLINE
#define WST_CX constexpr
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#endif

namespace wst {

typedef float2 cfloat;

WST_HD cfloat cmake(float re, float im) { return make_float2(re, im); }
WST_HD cfloat cadd(cfloat a, cfloat b) { return make_float2(a.x + b.x, a.y + b.y); }
WST_HD cfloat csub(cfloat a, cfloat b) { return make_float2(a.x - b.x, a.y - b.y); }
WST_HD cfloat cmul(cfloat a, cfloat b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
WST_HD cfloat cmulc(cfloat a, cfloat b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
WST_HD cfloat cscale(cfloat a, float s) { return make_float2(a.x * s, a.y * s); }
WST_HD cfloat cconj(cfloat a) { return make_float2(a.x, -a.y); }
WST_HD cfloat cmul_i(cfloat a) { return make_float2(-a.y, a.x); }    // a * (+i)
WST_HD cfloat cmul_ni(cfloat a) { return make_float2(a.y, -a.x); }   // a * (-i)
WST_HD float cabs_(cfloat a) {
#ifdef __CUDA_ARCH__
    float s = fmaf(a.x, a.x, a.y * a.y), r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(s));      // MUFU.SQRT, ~1 ulp, sqrt(0) = 0
    return r;
#else
    return sqrtf(a.x * a.x + a.y * a.y);
#endif
}

// ---------------------------------------------------------------- compile-time loops
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <int B, int E, class F>
WST_HD void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// ---------------------------------------------------------------- compile-time trigonometry
// cos/sin(2*pi*m/R) evaluated by the compiler in double precision (Taylor series on [-pi, pi],
// exact values on the axes) so that every in-register butterfly constant is an immediate.
constexpr double kPi = 3.14159265358979323846264338327950288;

WST_CX double cx_sin_series(double x) {
    double term = x, sum = x;
    for (int k = 1; k < 24; ++k) {
        term *= -x * x / double((2 * k) * (2 * k + 1));
        sum += term;
    }
    return sum;
}
WST_CX double cx_cos_series(double x) {
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 24; ++k) {
        term *= -x * x / double((2 * k - 1) * (2 * k));
        sum += term;
    }
    return sum;
}
WST_CX int cx_mod(int a, int n) { return ((a % n) + n) % n; }

WST_CX double cx_cos2pi(int m, int R) {
    m = cx_mod(m, R);
    if (m == 0) return 1.0;
    if (2 * m == R) return -1.0;
    if (4 * m == R || 4 * m == 3 * R) return 0.0;
    int mm = (2 * m > R) ? m - R : m;          // angle in (-pi, pi)
    return cx_cos_series(2.0 * kPi * double(mm) / double(R));
}
WST_CX double cx_sin2pi(int m, int R) {
    m = cx_mod(m, R);
    if (m == 0 || 2 * m == R) return 0.0;
    if (4 * m == R) return 1.0;
    if (4 * m == 3 * R) return -1.0;
    int mm = (2 * m > R) ? m - R : m;
    return cx_sin_series(2.0 * kPi * double(mm) / double(R));
}

WST_CX int cx_gcd(int a, int b) { return b == 0 ? a : cx_gcd(b, a % b); }
WST_CX int cx_modinv(int a, int n) {   // a^-1 mod n, gcd(a,n)=1
    a = cx_mod(a, n);
    for (int x = 1; x < n; ++x) if ((a * x) % n == 1) return x;
    return 0;
}
WST_CX int cx_smallest_prime_factor(int n) {
    for (int p = 2; p * p <= n; ++p) if (n % p == 0) return p;
    return n;
}
WST_CX bool cx_is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
WST_CX int cx_max(int a, int b) { return a > b ? a : b; }
WST_CX int cx_min(int a, int b) { return a < b ? a : b; }

}  // namespace wst
