// wst_advstats.cu — the reference's "advanced statistics" extractor on the GPU.
//
// Replaces extract_advanced_features (src/training/train_and_save_model.py:58-112, duplicated at
// src/inference/inference.py:181-235): 18 statistics per channel
//   mean std var min max range skew kurt cv p10 p25 p50 p75 p90 iqr mad grad_mean edge_density
// so that the `hybrid` feature vector (advanced stats + WST, train...:380-387) can be produced on the device.
//
// One CTA per (patch, channel) signal.  The signal lives in shared memory as the image (for the Sobel / Laplace
// stencils) next to a buffer of |laplace| values.  np.percentile needs ten order statistics of the pixels (two
// neighbouring ranks for each of the five percentiles) and two of |laplace| (the 90th-percentile edge threshold):
// they are found exactly by a most-significant-digit radix select — four passes of 8 bits over order-preserving
// integer keys, all ranks at once, one 256-bin histogram per distinct prefix, warp-aggregated shared-memory atomics
// (remote-sensing patches sit on the uint8 grid, so thousands of pixels share a bin) — instead of sorting 16 384
// values twice (the sort was 85 % of this kernel).  Moments are accumulated in double.  The stencil
// arithmetic reproduces scipy.ndimage's rounding points (each 1-D correlate pass is evaluated in double and
// rounded to float32) because edge_density counts `edges > threshold` on heavily tied data, where one ulp
// moves whole groups of pixels across the threshold.
#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>
#include <string>

#include "../../include/wst2d.h"

namespace {

constexpr int kThreads = 1024;
constexpr int kFeatures = 18;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum of up to NV doubles per thread; result broadcast to all threads
template <int NV>
__device__ void block_sum(double (&v)[NV], double* red /* [NV][kThreads/32] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) red[i * (kThreads / 32) + warp] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[i * (kThreads / 32) + w];
        v[i] = s;
    }
}

constexpr int kMaxRanks = 10;

// order-preserving map float -> unsigned (and back); -0.0 sorts just below +0.0, which compares equal like in numpy
__device__ __forceinline__ unsigned to_key(float v) {
    const unsigned u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_key(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

struct SelectState {
    unsigned prefix[kMaxRanks];        // digits resolved so far (the top 8 * pass bits of the key)
    int rank[kMaxRanks];               // rank of the target among the elements sharing its prefix
    int owner[kMaxRanks];              // first target with the same prefix: the one whose histogram is filled
    unsigned hist[kMaxRanks][256];
};

// Exact order statistics: on return st.prefix[t] is the key of the element of rank st.rank[t] (0-based, ascending)
// of buf[0..n), for t < T.  st.rank[t] must be set by the caller (thread 0) before a barrier; all threads call.
__device__ void radix_select(const float* buf, int n, int T, SelectState& st) {
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < T) st.prefix[tid] = 0u;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        __syncthreads();
        if (tid < T) {
            int o = tid;
            for (int t = 0; t < tid; ++t) if (st.prefix[t] == st.prefix[tid]) { o = t; break; }
            st.owner[tid] = o;
        }
        for (int i = tid; i < T * 256; i += kThreads) (&st.hist[0][0])[i] = 0u;
        __syncthreads();
        const int nround = (n + kThreads - 1) / kThreads * kThreads;
        for (int i = tid; i < nround; i += kThreads) {
            int slot = -1;
            if (i < n) {
                const unsigned key = to_key(buf[i]);
                const unsigned hi = pass == 0 ? 0u : key >> (shift + 8);
                if (pass == 0) slot = (int)(key >> 24);
                else
                    for (int t = 0; t < T; ++t)
                        if (st.owner[t] == t && st.prefix[t] == hi) { slot = t * 256 + (int)((key >> shift) & 255u); break; }
            }
            const unsigned act = __ballot_sync(0xffffffffu, slot >= 0);
            if (slot >= 0) {
                const unsigned same = __match_any_sync(act, slot);
                if (lane == __ffs(same) - 1) atomicAdd(&(&st.hist[0][0])[slot], (unsigned)__popc(same));
            }
        }
        __syncthreads();
        if (tid < T * 32) {                       // one warp per target: bin whose cumulative count passes the rank
            const int t = tid >> 5;
            const unsigned* h = st.hist[st.owner[t]] + lane * 8;
            int c[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { c[k] = (int)h[k]; sum += c[k]; }
            int incl = sum;                          // inclusive prefix over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int r = st.rank[t];
            const unsigned before = __ballot_sync(0xffffffffu, incl <= r);     // lanes wholly below the rank
            const int src = __popc(before) < 31 ? __popc(before) : 31;
            if (lane == src) {
                int rr = r - (incl - sum), bin = 0;
                while (bin < 7 && rr >= c[bin]) { rr -= c[bin]; ++bin; }
                st.rank[t] = rr;
                st.prefix[t] = (st.prefix[t] << 8) | (unsigned)(lane * 8 + bin);
            }
        }
    }
    __syncthreads();
}

// np.percentile(a, q) (method 'linear') for a float32 array.  numpy 2.x evaluates the whole thing in the array's
// dtype: q / float32(100), (n - 1) * q, the fractional weight t and the _lerp
//   a + (b - a) * t,   or   b - (b - a) * (1 - t) when t >= 0.5
// are all float32 operations (unfused), and the result is a float32.  Reproduced operation by operation, because
// the edge-density threshold is compared against heavily tied float32 data.  percentile_ranks gives the two
// neighbouring ranks the interpolation needs, percentile_lerp combines the two order statistics.
__device__ __forceinline__ void percentile_ranks(int n, float q_percent, int& lo, int& hi, float& t) {
    const float q = __fdiv_rn(q_percent, 100.0f);
    const float virt = __fmul_rn((float)(n - 1), q);
    const float fl = floorf(virt);
    lo = (int)fl;
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    hi = lo + 1 < n ? lo + 1 : n - 1;
    t = __fsub_rn(virt, fl);
}
__device__ __forceinline__ float percentile_lerp(float a, float b, float t) {
    const float diff = __fsub_rn(b, a);
    float r = __fadd_rn(a, __fmul_rn(diff, t));
    if (t >= 0.5f) r = __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, t)));
    return r;
}

__global__ void __launch_bounds__(kThreads, 1)
advstats_kernel(const void* __restrict__ x, int u8_channels, long long nsig, int H, int W,
                float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* img = reinterpret_cast<float*>(smem_raw);
    float* edg = img + H * W;                                   // |laplace|
    __shared__ double red[5 * (kThreads / 32)];
    __shared__ float redf[2 * (kThreads / 32)];
    __shared__ SelectState st;
    __shared__ float pct[6];                                    // p10 p25 p50 p75 p90 of the pixels, p90 of |laplace|
    __shared__ int cnt;
    const int n = H * W, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float qs[5] = {10.0f, 25.0f, 50.0f, 75.0f, 90.0f};

    for (long long s = blockIdx.x; s < nsig; s += gridDim.x) {
        __syncthreads();
        // ---- load (float32 plane, or one channel of uint8 HWC pixels / 255 like load_rgb_image)
        float mn = FLT_MAX, mx = -FLT_MAX;
        double acc1[1] = {0.0};
        for (int i = tid; i < n; i += kThreads) {
            float v;
            if (u8_channels > 0) {
                long long b = s / u8_channels; int c = (int)(s - b * u8_channels);
                v = __fdiv_rn((float)static_cast<const unsigned char*>(x)[((size_t)b * n + i) * u8_channels + c], 255.0f);
            } else {
                v = static_cast<const float*>(x)[(size_t)s * n + i];
            }
            img[i] = v;
            acc1[0] += (double)v;
            mn = fminf(mn, v); mx = fmaxf(mx, v);
        }
        mn = warp_min(mn); mx = warp_max(mx);
        if (lane == 0) { redf[warp] = mn; redf[kThreads / 32 + warp] = mx; }
        if (tid == 0) {                                          // the ten ranks np.percentile interpolates between
            for (int k = 0; k < 5; ++k) { float t; percentile_ranks(n, qs[k], st.rank[2 * k], st.rank[2 * k + 1], t); }
            cnt = 0;
        }
        block_sum<1>(acc1, red);
        mn = FLT_MAX; mx = -FLT_MAX;
        for (int w = 0; w < kThreads / 32; ++w) { mn = fminf(mn, redf[w]); mx = fmaxf(mx, redf[kThreads / 32 + w]); }
        const double mean = acc1[0] / (double)n;
        const double mean_val = (double)(float)mean;            // features[base+0] is a float32 mean in the reference

        // ---- central moments, mean absolute deviation, Sobel magnitude, |laplace|
        double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};              // m2, m3, m4, sum|x - mean|, sum grad_mag
        for (int i = tid; i < n; i += kThreads) {
            const int r = i / W, c = i - r * W;
            const int rm = r > 0 ? r - 1 : 0, rp = r < H - 1 ? r + 1 : H - 1;      // scipy 'reflect': d c b a | a b c d
            const int cm = c > 0 ? c - 1 : 0, cp = c < W - 1 ? c + 1 : W - 1;
            const float v = img[i];
            const float v_u = img[rm * W + c], v_d = img[rp * W + c], v_l = img[r * W + cm], v_r = img[r * W + cp];
            const double d = (double)v - mean;
            const double d2 = d * d;
            acc[0] += d2; acc[1] += d2 * d; acc[2] += d2 * d2;
            acc[3] += fabs((double)v - mean_val);
            // sobel(axis=0): [-1,0,1] along rows (rounded to float32), then [1,2,1] along columns
            const float t_l = __fsub_rn(img[rp * W + cm], img[rm * W + cm]);
            const float t_c = __fsub_rn(v_d, v_u);
            const float t_r = __fsub_rn(img[rp * W + cp], img[rm * W + cp]);
            const float gx = (float)((double)t_l + 2.0 * (double)t_c + (double)t_r);
            // sobel(axis=1): [-1,0,1] along columns, then [1,2,1] along rows
            const float s_u = __fsub_rn(img[rm * W + cp], img[rm * W + cm]);
            const float s_c = __fsub_rn(v_r, v_l);
            const float s_d = __fsub_rn(img[rp * W + cp], img[rp * W + cm]);
            const float gy = (float)((double)s_u + 2.0 * (double)s_c + (double)s_d);
            acc[4] += (double)__fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)));
            // laplace: [1,-2,1] along rows rounded to float32, plus [1,-2,1] along columns rounded, float32 add
            const double v2 = 2.0 * (double)v;
            const float l0 = (float)((double)v_u - v2 + (double)v_d);
            const float l1 = (float)((double)v_l - v2 + (double)v_r);
            edg[i] = fabsf(__fadd_rn(l0, l1));
        }
        block_sum<5>(acc, red);

        // ---- percentiles of the pixels
        radix_select(img, n, 10, st);
        if (tid < 5) {
            int lo, hi; float t;
            percentile_ranks(n, qs[tid], lo, hi, t);
            pct[tid] = percentile_lerp(from_key(st.prefix[2 * tid]), from_key(st.prefix[2 * tid + 1]), t);
        }
        __syncthreads();
        // ---- edge density: share of |laplace| above its own 90th percentile
        if (tid == 0) { float t; percentile_ranks(n, 90.0f, st.rank[0], st.rank[1], t); }
        radix_select(edg, n, 2, st);
        if (tid == 0) {
            int lo, hi; float t;
            percentile_ranks(n, 90.0f, lo, hi, t);
            pct[5] = percentile_lerp(from_key(st.prefix[0]), from_key(st.prefix[1]), t);
        }
        __syncthreads();
        const float thr = pct[5];
        int local = 0;
        for (int i = tid; i < n; i += kThreads) local += (edg[i] > thr) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0 && local) atomicAdd(&cnt, local);
        __syncthreads();

        if (tid == 0) {
            const double var = acc[0] / n, sd = sqrt(var);
            const double m3 = acc[1] / n, m4 = acc[2] / n;
            float* o = out + (size_t)s * kFeatures;
            const double stdf = (double)(float)sd;                 // features[base+1] is float32 in the reference
            const float p25 = pct[1], p75 = pct[3];
            o[0] = (float)mean; o[1] = (float)sd; o[2] = (float)var;
            o[3] = mn; o[4] = mx; o[5] = __fsub_rn(mx, mn);
            o[6] = (float)(m3 / (var * sd));                       // scipy.stats.skew (biased)
            o[7] = (float)(m4 / (var * var) - 3.0);                // scipy.stats.kurtosis (Fisher, biased)
            o[8] = (float)(stdf / fmax(mean_val, 1e-8));
            o[9] = pct[0]; o[10] = p25; o[11] = pct[2]; o[12] = p75; o[13] = pct[4];
            o[14] = (float)((double)p75 - (double)p25);        // features[] is float64: the difference is taken there
            o[15] = (float)(acc[3] / n);
            o[16] = (float)(acc[4] / n);
            o[17] = (float)((double)cnt / (double)n);
        }
    }
}

thread_local std::string g_adv_error;

}  // namespace

extern "C" {

const char* wst2d_advanced_stats_last_error(void) { return g_adv_error.c_str(); }

int wst2d_advanced_stats(int device, const void* x_dev, int is_u8, int64_t B, int C, int H, int W, float* out_dev,
                         void* cuda_stream) {
    auto fail = [](int code, const std::string& m) { g_adv_error = m; return code; };
    if (B < 0 || C < 1 || H < 2 || W < 2) return fail(WST2D_ERR_ARG, "invalid shape");
    if (B == 0) return WST2D_OK;
    if (!x_dev || !out_dev) return fail(WST2D_ERR_ARG, "NULL buffer");
    const int n = H * W;
    const size_t smem = (size_t)2 * n * sizeof(float);
    if (smem > 200 * 1024) return fail(WST2D_ERR_UNSUPPORTED, "patch too large for the shared-memory statistics kernel (the image and its |laplace| map must fit in 200 KB)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(WST2D_ERR_CUDA, "no such CUDA device");
    int prev = -1;
    cudaGetDevice(&prev);
    if (prev != device) cudaSetDevice(device);
    cudaError_t e = cudaFuncSetAttribute(advstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const long long nsig = (long long)B * C;
    const int grid = (int)(nsig < 2LL * sms ? nsig : 2LL * sms);
    if (e == cudaSuccess) {
        advstats_kernel<<<grid, kThreads, smem, (cudaStream_t)cuda_stream>>>(x_dev, is_u8 ? C : 0, nsig, H, W, out_dev);
        e = cudaGetLastError();
    }
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
    if (e != cudaSuccess) return fail(WST2D_ERR_CUDA, std::string("advanced_stats: ") + cudaGetErrorString(e));
    return WST2D_OK;
}

}  // extern "C"
