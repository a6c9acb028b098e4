// wst_cascade.h — the fused scattering cascade for ONE (patch, channel) signal, executed by ONE CTA.
//
// Replaces, for a batch, the per-image / per-channel Python loop around kymatio's
// core.scattering2d (reference call sites: src/training/train_and_save_model.py:364-376,
// src/inference/inference.py:246-268; algorithm: SURVEY.md Appendix A.3).
//
// Data flow per signal (all intermediates stay in shared memory; only the half spectrum U0^ of
// the padded input is parked in a per-CTA, L2-resident global scratch because every one of the
// J*L first-order filters re-reads it):
//
//   x --reflect pad--> z0 --real 2-D FFT--> U0^                                    (once)
//   S0      = lowpass(z0)
//   U1[n1]  = | ifft2( fold( U0^ . psi^[n1] ) ) |            n1 = (j1, theta1)
//   S1[n1]  = lowpass(U1)
//   U1^     = real 2-D FFT(U1)                               (only when the parent has children)
//   U2      = | ifft2( fold( U1^ . psi^[n2] ) ) |            j2 > j1, U1^ resident in shared memory
//   S2      = lowpass(U2)
//
// Differences from the reference dataflow that keep the numbers (within fp32 rounding) but
// remove work:
//   * "lowpass" = phi^ multiply + Fourier fold + small inverse FFT + unpad in kymatio.  phi is an
//     isotropic Gaussian, so phi^ is separable; the same result is a pair of small dense matrix
//     products with Gr[hout x m], Gc[hout x m] (built on the host from phi^'s first row/column),
//     evaluated only at the kept (un-padded) outputs.  No forward FFT of U2 is ever needed.
//   * forward FFTs act on real data (|.|): two rows are packed into one complex row, and only the
//     Hermitian half spectrum is kept.
//   * the FFTs are in-place two-pass transforms n = R1*R2 whose frequency-domain side is left in
//     digit-swapped order pi(k) = (k % R1)*R2 + k / R1; forward = natural->swapped (DIF),
//     inverse = swapped->natural (DIT), so no pass needs a reorder or a second barrier.
//
// The code is written as a sequence of barrier-separated phases `ex.template phase<TAG>([&](int tid){...})`.
// On the GPU a phase is the lambda + __syncthreads(); in tests/emu the same phases are replayed
// thread by thread on the CPU to check the index arithmetic against the oracle.
#pragma once
#include "wst_dft.h"

namespace wst {

#ifndef WST_NT
#define WST_NT 640
#endif
constexpr int kMaxJ = 6;
constexpr int kMaxL = 8;                 // orientations per scale supported by the compiled cascades
constexpr int kMaxPairs = kMaxJ * (kMaxJ - 1) / 2;
WST_CX int pair_index(int j2, int j1) { return j2 * (j2 - 1) / 2 + j1; }
#ifndef WST_SMEM_BUDGET
#define WST_SMEM_BUDGET 27000
#endif
constexpr int kSmemCfloats = WST_SMEM_BUDGET;   // data region budget: 216,000 B of the 227 KB a CTA may use by default;
                                                // smaller for configurations that run several CTAs per SM
#ifndef WST_GLOBAL_BUDGET
#define WST_GLOBAL_BUDGET (1 << 24)
#endif
// Data region budget of the global-workspace variant, in cfloats.  It sets how many same-scale arrays are processed
// together (GP, G2) and therefore the hot working set of one signal; the cluster variant picks it so that the working
// sets of all signals in flight stay resident in L2.
constexpr int kGlobalCfloats = WST_GLOBAL_BUDGET;
// Hybrid form of the global-workspace variant: levels whose arrays (with their children) fit in this many cfloats of
// shared memory are processed there, exactly like the shared-memory cascade; only the levels too large for one SM keep
// their arrays in the workspace.  0: every level in the workspace.
#ifndef WST_HYBRID_BUDGET
#define WST_HYBRID_BUDGET 0
#endif
constexpr int kHybridCfloats = WST_HYBRID_BUDGET;
#ifndef WST_GLOBAL_THREADS_PER_SM
#define WST_GLOBAL_THREADS_PER_SM 512      // resident threads per SM the global-workspace variant is compiled for (register cap)
#endif

// ------------------------------------------------------------------ 1-D factorisation n = R1*R2
// Prime-factor (Good-Thomas) split n = P * Q with P the power-of-two part and Q the odd part: the two passes
// need no twiddles at all.  Used whenever both radices fit in registers (2 <= P <= 16, 3 <= Q <= 24).
WST_CX int fft_pfa_Q(int n) {
    if (n <= 18) return 0;
    int odd = n; while (odd % 2 == 0) odd /= 2;
    int p2 = n / odd;
    return (odd >= 3 && odd <= 24 && p2 >= 2 && p2 <= 16) ? odd : 0;
}
WST_CX int fft_R2(int n) {               // contiguous radix
    if (n <= 18) return n;                 // one in-register pass
    if (fft_pfa_Q(n) > 0) return fft_pfa_Q(n);
    int odd = n; while (odd % 2 == 0) odd /= 2;
    int best = 0; long best_score = 1L << 60;
    for (int r2 = odd; r2 <= 24 && r2 <= n; r2 *= 2) {
        int r1 = n / r2;
        if (r1 > 16 || r2 < 2) continue;
        long d = (long)r2 * r2 - n;      // distance of r2 from sqrt(n), in squared units
        if (d < 0) d = -d;
        if (d < best_score) { best_score = d; best = r2; }
    }
    if (best == 0) {                     // large sides (e.g. 576 = 24 x 24): any divisor pair with both radices <= 24
        for (int r2 = 2; r2 <= 24; ++r2) {
            if (n % r2 != 0 || n / r2 > 24) continue;
            long d = (long)r2 * r2 - n;
            if (d < 0) d = -d;
            if (d < best_score) { best_score = d; best = r2; }
        }
    }
    return best;
}
WST_CX int fft_R1(int n) { return fft_R2(n) > 0 ? n / fft_R2(n) : 0; }
WST_CX bool fft_supported(int n) { return n >= 2 && n % 2 == 0 && fft_R2(n) > 0; }

// In-place two-pass transform of length N = R1 * R2 (strided radix R1, contiguous radix R2).
//   Cooley-Tukey (PFA == false): frequency k is stored at (k % R1)*R2 + k / R1, space is in natural order, and
//     one twiddle multiply sits between the passes.
//   Good-Thomas (PFA == true, R1 and R2 coprime): frequency k is stored at (k*u % R1)*R2 + (k*v % R2) with
//     u = R2^-1 mod R1, v = R1^-1 mod R2; sample i of the spatial side at (i % R1)*R2 + (i % R2); no twiddles.
//     Because R1 is even and R2 odd, samples i and i + N/2 sit exactly N/2 storage slots apart, which is what
//     the row-pairing of the real-input transforms relies on.
template <int N> struct Fft1 {
    static constexpr int R2 = fft_R2(N);
    static constexpr int R1 = N / R2;
    static_assert(R2 > 0 && R1 * R2 == N, "unsupported FFT length");
    static constexpr bool PFA = (R1 > 1) && fft_pfa_Q(N) == R2;
    static constexpr int U = PFA ? cx_modinv(R2, R1) : 0;
    static constexpr int V = PFA ? cx_modinv(R1, R2) : 0;
    // storage position of frequency k
    static WST_HD int pi(int k) {
        if constexpr (R1 == 1) return k;
        else if constexpr (PFA) return ((k * U) % R1) * R2 + (k * V) % R2;
        else return (k % R1) * R2 + k / R1;
    }
    // storage position of spatial sample i, and its inverse
    static WST_CX int pos_s(int i) {
        if constexpr (PFA) return (i % R1) * R2 + (i % R2); else return i;
    }
    static WST_HD int inv_pos_s(int p) {
        if constexpr (PFA) return ((p / R2) * R2 * U + (p % R2) * R1 * V) % N; else return p;
    }
};

// ------------------------------------------------------------------ banded low-pass
// The low-pass kernel at a level whose output stride is S = 2^(J-j) samples is a Gaussian of sigma 0.4*S samples
// (phi has sigma 0.8 * 2^(J-1) pixels): beyond R = ceil(5.68 * 0.4 * S) samples it is below
// 1e-7 of its peak (the same threshold the wavelet supports use).  When that band is narrow against the row length
// (large output maps: 128x128 at J=2 keeps 32x32 outputs) the separable low-pass is evaluated as a strided
// (2R+1)-tap convolution instead of a dense [m x HOUT] operator.
constexpr int kLpTaps = 40;              // taps w[0..R] kept per level in the kernel parameter block
WST_CX int lp_radius(int stride) { return stride == 2 ? 5 : stride == 4 ? 10 : stride == 8 ? 19 : stride == 16 ? 37 : 0; }
// Level 0 holds phi as sampled in space (an exact periodised Gaussian).  Deeper levels use the corner crop of phi^,
// whose kink at the level's Nyquist frequency gives the spatial kernel 1/d^2 tails: 4e-2 of the peak spectrum at
// stride 2, 3e-6 at stride 4, negligible from stride 8 on.  Hence: level 0, or stride >= 8.
WST_CX bool lp_banded(int m, int hout, int level) {
    int s = m / (hout + 2), r = lp_radius(s);
    return r > 0 && s * (hout + 2) == m && (2 * r + 1) * 3 <= m && (level == 0 || s >= 8);
}

#ifndef WST_STAGE_BUFS
#define WST_STAGE_BUFS 1
#endif
// Tuning knobs of the shared-memory cascade (A/B builds: tools/tune_variants.sh):
//   WST_OPT_HOIST  task loops whose thread stride is a multiple of the line count keep the line / column index of a
//                  thread fixed across iterations (no per-iteration div / mod, row-independent addressing hoisted)
//   WST_OPT_BATCH  radix passes with short butterflies take two tasks per iteration (loads of both, butterflies of
//                  both, stores of both): twice the independent work per thread between shared-memory round trips
//   WST_OPT_TMA    the H x W input of the NEXT signal is fetched by bulk asynchronous copies (TMA, cp.async.bulk +
//                  mbarrier) into a free part of shared memory while the last level of the current signal runs
#ifndef WST_OPT_HOIST
#define WST_OPT_HOIST 0
#endif
//   WST_OPT_HOIST_PROD  the same for the filter products only
//   WST_OPT_PREFETCH    dense two-orientation products issue the filter loads of a thread's next output pair before
//                       the arithmetic of the current one (L2 latency overlaps the shared-memory reads and FMAs)
#ifndef WST_OPT_HOIST_PROD
#define WST_OPT_HOIST_PROD WST_OPT_HOIST
#endif
#ifndef WST_OPT_PREFETCH
#define WST_OPT_PREFETCH 0
#endif
#ifndef WST_OPT_BATCH
#define WST_OPT_BATCH 0
#endif
#ifndef WST_OPT_TMA
#define WST_OPT_TMA 1
#endif
//   WST_OPT_ROWSEQ  the fused last pass transforms its two rows one after the other (fewer live registers: what lets
//                   the cascade run with 800 threads per CTA)
#ifndef WST_OPT_ROWSEQ
#define WST_OPT_ROWSEQ 0
#endif
//   WST_OPT_PAIRHALF  dense few-orientation products pair the outputs (l, l + MC/2) instead of (l, l + 1): consecutive
//                     lanes then read consecutive spectrum samples (no two-way bank conflicts: these loads and stores
//                     were 35 % of the kernel's conflict replays) and whether an alias comes from the stored half
//                     spectrum or from its Hermitian mirror is known at compile time (no branches)
#ifndef WST_OPT_PAIRHALF
#define WST_OPT_PAIRHALF 1
#endif
//   WST_OPT_HERMSEL   Hermitian lookups select row, column and sign arithmetically instead of branching
#ifndef WST_OPT_HERMSEL
#define WST_OPT_HERMSEL 0
#endif
//   WST_OPT_SPARSEROW bit mask.  1: sparse products that visit all F column aliases (F = 4) compute the two spectrum row
//                     pointers (row k and its mirror -k) once per alias row and know the side of the half spectrum of
//                     each alias at compile time; 2: the other sparse products (F = 8) select row, column and sign
//                     without a branch; 4: dense many-orientation products at fold 2 know the side at compile time
#ifndef WST_OPT_SPARSEROW
#define WST_OPT_SPARSEROW 0
#endif
//   WST_OPT_LPPAIR    the first contraction of the dense low-pass takes two columns (y, y + M/2) per thread: every
//                     (warp-uniform) vector load of the operator feeds twice the FMAs; the phase is bound by
//                     shared-memory wavefronts, not by arithmetic
#ifndef WST_OPT_LPPAIR
#define WST_OPT_LPPAIR 1
#endif
//   WST_OPT_L2HINT   bit mask, global-workspace variant only.  1: filter loads of the product phases carry an L2
//                     evict_last policy (the bank is re-read by every CTA for every signal; without it the workspaces
//                     streaming through L2 push it out to HBM).  2: tile loads of the staged passes carry evict_first.
//                     4: tile stores carry evict_first.  8: product outputs written to the workspace carry evict_first.
#ifndef WST_OPT_L2HINT
#define WST_OPT_L2HINT 9
#endif
//   WST_OPT_DYNSCHED  signals after a CTA's first one come from a device-wide ticket counter instead of a fixed stride
//                     (run_cascade in wst_cfg_inst.cu)
#ifndef WST_OPT_DYNSCHED
#define WST_OPT_DYNSCHED 1
#endif
//   WST_OPT_PRODTILE  global-workspace levels with fold factor <= 2: the filter product is computed straight into the
//                     shared-memory column tile that the inverse column transforms run on (product_ifft_cols_staged)
//                     instead of being written to the workspace and loaded back tile by tile: one write and one read
//                     of every such array less.  The workspace then holds the array with its columns in natural
//                     frequency order between the column and the row passes (the row-tile loads permute).
#ifndef WST_OPT_PRODTILE
#define WST_OPT_PRODTILE 0
#endif
#ifndef WST_STAGE_TC_MAX
#define WST_STAGE_TC_MAX 16
#endif
// columns per staged tile of the global-workspace variant: the largest power of two <= WST_STAGE_TC_MAX dividing m
WST_CX int stage_tc(int m) { int t = WST_STAGE_TC_MAX; while (m % t) t /= 2; return t; }
// columns per tile of the fused product + column transform: gs arrays of mc rows, pitch t + 1, in `avail` cfloats
// (0: does not fit with at least four columns — a row segment of a tile store is then a whole 32-byte sector)
WST_CX int prodtile_cols(int mc, int gs, int avail) {
    for (int t = 16; t >= 4; t /= 2) if (mc % t == 0 && gs * mc * (t + 1) <= avail) return t;
    return 0;
}

// ------------------------------------------------------------------ geometry shared by host and device
// WS_GLOBAL = false: the data region lives in shared memory (the fast path, N <= 160).
// WS_GLOBAL = true : same program, data region in a per-CTA global-memory workspace — correct for sides whose
//                    spectra exceed one SM's shared memory (512x512 J=5 -> 576), not yet tuned.
// CL > 1 (global-workspace variant only): one signal is processed by a thread-block cluster of CL CTAs.  NT is the
//                    number of threads of the whole cluster (the stride of every phase loop), NTL = NT / CL the
//                    threads of one CTA; a phase ends with a cluster barrier instead of __syncthreads(), every CTA
//                    keeps its own copy of the twiddle / low-pass tables in shared memory, and the low-pass reduction
//                    buffer moves from shared memory to the workspace.
template <int N_, int J_, int NT_ = WST_NT, bool WS_GLOBAL_ = false, int CL_ = 1>
struct Cfg {
    static constexpr int N = N_, J = J_, NT = NT_, CL = CL_, NTL = NT_ / CL_;
    static constexpr bool WS_GLOBAL = WS_GLOBAL_;
    static_assert(CL_ >= 1 && NT_ % CL_ == 0 && (CL_ == 1 || WS_GLOBAL_), "clusters need the global-workspace variant");
    static constexpr int BUDGET = WS_GLOBAL_ ? kGlobalCfloats : kSmemCfloats;
    static constexpr int NS = N >> J;           // side of the subsampled (still padded) output grid
    static constexpr int HOUT = NS - 2;         // kept outputs per side after unpad [1:-1]
    // row length of the low-pass operator tables: padded to float4, and to two float4s from 8 outputs on (the dense
    // low-pass takes its outputs eight at a time)
    static constexpr int HP = HOUT >= 8 ? (HOUT + 7) & ~7 : (HOUT + 3) & ~3;
    static_assert((N >> J) << J == N, "padded size must be a multiple of 2^J");
    static_assert(HOUT >= 1, "empty output");

    static WST_CX int msize(int j) { return N >> j; }
    static WST_CX int vsz(int m) { return m * (m + 1); }           // complex m x m array, odd pitch m+1
    static WST_CX int uhsz(int m) { return m * (m / 2 + 1); }       // half spectrum, pitch m/2+1
    static WST_CX int zend(int m, int gp) { return (gp - 1) * vsz(m) + (m / 2) * (m + 1); }

    // children group size for child side mc when `room` cfloats are free below the parents' U^ arrays
    static WST_CX int pick_group(int mc, int room) {
        for (int g = 8; g > 1; g /= 2) if (g * vsz(mc) <= room) return g;
        return 1;
    }
    static WST_CX bool has_children(int j) { return j < J - 1; }
    // Memory space of a level's arrays.  Shared-memory cascade: everything in shared memory (budget kSmemCfloats).
    // Global-workspace variant: the per-CTA workspace (budget kGlobalCfloats), except — hybrid form, SB > 0 — the levels
    // that fit, with all their children, in SB cfloats of shared memory.  Children of a workspace-level parent that are
    // themselves small enough also go to shared memory (their parent's spectrum is then read from the workspace).
    static constexpr int SB = (WS_GLOBAL_ && CL_ == 1) ? kHybridCfloats : 0;
    static WST_CX int level_total_in(int j, int gp, int budget, bool hybrid_parent) {
        int m = msize(j);
        if (!has_children(j)) return gp * vsz(m);
        int room = budget - gp * uhsz(m);
        if (room < zend(m, gp)) return 1 << 30;
        int offb = zend(m, gp);
        for (int j2 = j + 1; j2 < J; ++j2) {
            if (!hybrid_parent && WS_GLOBAL_ && SB > 0 && fits_shared(j2)) continue;   // those children live in shared memory
            int ch = pick_group(msize(j2), room) * vsz(msize(j2));
            if (ch > room) return 1 << 30;
            if (ch > offb) offb = ch;
        }
        int t = offb + gp * uhsz(m);
        return t > gp * vsz(m) ? t : gp * vsz(m);
    }
    static WST_CX bool fits_shared(int j) { return SB > 0 && level_total_in(j, 1, SB, true) <= SB; }
    // true: level j's arrays are in shared memory
    static WST_CX bool in_smem(int j) { return !WS_GLOBAL_ || (j > 0 && fits_shared(j)); }
    static WST_CX int budget_of(int j) { return WS_GLOBAL_ ? (in_smem(j) ? SB : kGlobalCfloats) : kSmemCfloats; }
    static WST_CX int level_total(int j, int gp) { return level_total_in(j, gp, budget_of(j), WS_GLOBAL_ && in_smem(j)); }
    // number of same-scale parents processed together at level j
    static WST_CX int GP(int j) {
        for (int g = 8; g > 1; g /= 2) if (level_total(j, g) <= budget_of(j)) return g;
        return 1;
    }
    static WST_CX int G2(int j1, int j2) {   // children group size
        if (in_smem(j2) && !in_smem(j1)) return pick_group(msize(j2), SB);      // children alone in the shared region
        return pick_group(msize(j2), budget_of(j1) - GP(j1) * uhsz(msize(j1)));
    }
    static WST_CX int OFFB(int j) {          // offset of the parents' half spectra
        int gp = GP(j), m = msize(j);
        int offb = zend(m, gp);
        for (int j2 = j + 1; j2 < J; ++j2)
            if (in_smem(j2) == in_smem(j)) offb = cx_max(offb, G2(j, j2) * vsz(msize(j2)));
        return offb;
    }
    // cfloats of the data region: shared memory (shared-memory cascade) or the per-CTA workspace (levels kept there)
    static WST_CX int smem_cfloats() {
        int t = 0;
        for (int j = 0; j < J; ++j) if (!WS_GLOBAL_ || !in_smem(j)) t = cx_max(t, level_total(j, GP(j)));
        // input stage uses level-0 layout with one array
        t = cx_max(t, cx_max(vsz(N) / 2 + 1, OFFB(0) + uhsz(N)));
        return t;
    }
    // cfloats of the hybrid shared-memory region (global-workspace variant): its levels, and the small children of
    // workspace-level parents
    static WST_CX int hybrid_cfloats() {
        if (!WS_GLOBAL_ || SB == 0) return 0;
        int t = 0;
        for (int j = 1; j < J; ++j) {
            if (!in_smem(j)) continue;
            t = cx_max(t, level_total(j, GP(j)));
            for (int j1 = 0; j1 < j; ++j1) if (!in_smem(j1)) t = cx_max(t, G2(j1, j) * vsz(msize(j)));
        }
        return t;
    }
    static WST_CX int tw_offset(int j) {     // twiddle tables, one per level, appended after the data
        int o = 0;
        for (int i = 0; i < j; ++i) o += msize(i);
        return o;
    }
    static WST_CX int tw_sum() { int o = 0; for (int i = 0; i < J; ++i) o += msize(i); return o; }
    static constexpr int tw_total = (2 * N - (N >> (J - 1)) + 7) & ~7;   // sum_j N>>j, padded to 8
    // low-pass operator tables G[j][m_j][HP] (floats), one per level, resident in shared memory
    static WST_CX int g_offset(int j) { return tw_offset(j) * HP; }
    // (the global-workspace variant reads them from global memory through L1 instead: without their 74 KB two CTAs fit
    // on an SM, and one CTA's tile loads then overlap the other's transforms — 914 -> 971 tiles/s at 512 x 512 J=5)
    static constexpr int g_total = WS_GLOBAL_ ? 0 : tw_total * HP;
    // The low-pass of an array can be fused into the last pass of its inverse FFT when the 2*HOUT
    // partial sums of a thread fit in the shared-memory slots that thread owns (see pass_rows_final).
    static WST_CX bool lp_fused(int m, bool write_z) {
        int r1 = fft_R1(m);
        if (r1 == 1) return !write_z;
        return write_z ? (HOUT <= r1) : (HOUT / 2 <= r1);
    }
    static WST_CX bool any_fused() {
        for (int j = 0; j < J; ++j) if (lp_fused(msize(j), has_children(j)) || (j > 0 && lp_fused(msize(j), false))) return true;
        return false;
    }
    // (array, row-chunk) partial maps held between the two reduce phases: ~10 KB, at least one slot per array
    // (the global-workspace variant keeps them in its workspace and takes enough chunks for every thread to have work:
    // with one chunk per array the reduction of eight 288-row arrays would be 128 work items for 512 threads)
    static constexpr int LP_SLOTS = WS_GLOBAL_ ? 256 : cx_max(8, cx_min(40, 2560 / (HOUT * HOUT)));
    static WST_CX int lpbuf_floats() { return any_fused() ? LP_SLOTS * HOUT * HOUT : 0; }
    // bytes of dynamic shared memory, and of per-CTA global workspace (0 for the shared-memory variant)
    static WST_CX size_t smem_bytes() {
        // (the hybrid region and the stage tiles share their shared memory: a staged pass and a shared-memory level
        // never run at the same time)
        return (size_t)((WS_GLOBAL ? 0 : smem_cfloats()) + tw_total + cx_max(STAGE_BUFS * stage_cfloats(), hybrid_cfloats())) * sizeof(cfloat)
               + (size_t)(g_total + (WS_GLOBAL ? 0 : lpbuf_floats())) * sizeof(float);
    }
    // Global-workspace variant: the passes of an inverse FFT run on shared-memory tiles (a batch of columns, then a
    // batch of row pairs) so that each array crosses HBM once per dimension instead of once per pass.
    static WST_CX int stage_cols(int m) { return stage_tc(m); }                                 // columns per tile
    static WST_CX int stage_cfloats() { return WS_GLOBAL ? N * (stage_cols(N) + 1) : 0; }      // one level-0 column tile
    // shared memory behind `stage`: the stage tiles and the hybrid region alias each other
    static WST_CX int stage_avail() { return WS_GLOBAL ? cx_max(WST_STAGE_BUFS * stage_cfloats(), hybrid_cfloats()) : 0; }
    static constexpr int STAGE_BUFS = WST_STAGE_BUFS;   // 2: tile t+1 is loaded (asynchronously) while tile t is transformed
    static WST_CX int stage_rows(int m) {                                                       // row pairs per tile
        int t = 16;
        while (t > 1 && ((m / 2) % t != 0 || 2 * t * (m + 1) > stage_cfloats())) t /= 2;
        return t;
    }
    // CTAs meant to share an SM (small configurations run several narrower CTAs, so that one CTA's barrier waits are
    // filled by another's work): bounded by shared memory and by keeping 640 threads' worth of registers per SM.
    static WST_CX int min_ctas() {
        if (CL > 1) return 1;
        int by_smem = (int)(232448 / (smem_bytes() + 1024)), by_threads = (WS_GLOBAL ? WST_GLOBAL_THREADS_PER_SM : 640) / NTL;
        int m = by_smem < by_threads ? by_smem : by_threads;
        return m < 1 ? 1 : m;
    }
    static WST_CX size_t workspace_cfloats() {
        return WS_GLOBAL ? (size_t)smem_cfloats() + (size_t)(lpbuf_floats() + 1) / 2 : 0;
    }
    static_assert(level_total(0, 1) <= BUDGET, "padded size too large for the shared-memory cascade");
};

// ------------------------------------------------------------------ device-side plan tables
struct PlanTables {
    int L, max_order, K;
    int H, W, pad_top, pad_left;           // un-padded input size and reflect-pad offsets
    const cfloat* tw[kMaxJ];               // tw[j][k1*R2+i2] = exp(-2 pi i * i2*k1 / m_j)
    const float* gr[kMaxJ];                // row   low-pass operator, [m_j][HP]  (x-major)
    const float* gc[kMaxJ];                // column low-pass operator, [m_j][HP]
    const float* psi1[kMaxJ];              // order-1 filters of scale j at level 0: [ceil(L/GP)][N][N][GP], GP = GP(j)
    const float* psi2[kMaxJ][kMaxJ];       // [j2][j1]: scale j2 periodised to level j1: [ceil(L/G)][m][m][G], G = G2(j1, j2)
    // Bounding box of the support of every theta-group of filters (entries above kSupportEps * max|psi^|):
    // one cyclic row span and one cyclic column span, each packed lo << 16 | len.  They travel in the kernel parameter
    // block (constant bank), so the product loops have no dependent global loads.
    int bb1[kMaxJ][kMaxL][2];              // order-1: scale j at level 0
    int bb2[kMaxPairs][kMaxL][2];          // order-2: pair_index(j2, j1): scale j2 at level j1
    // Banded low-pass taps of level j (zero where the level uses the dense operator): lpw[j][d] = weight of the
    // sample d positions away from the output's centre, d <= lp_radius.  Read with compile-time indices, i.e. as
    // constant-bank operands of the FMAs.
    float lpw[kMaxJ][kLpTaps];
};
constexpr float kSupportEps = 1e-7f;

// One H x W input signal: float32 (plane of a CHW patch) or uint8 (channel of an HWC patch, stride C;
// value / 255 in IEEE float32 like load_rgb_image, train_and_save_model.py:54).
struct SignalSrc {
    const float* f32;
    const unsigned char* u8;
    int stride;          // element stride (C for interleaved uint8 pixels, 1 otherwise)
    int pitch;           // elements per image row (W for a patch; the raster width for a tile of a scene)
    WST_D float at(int r, int c) const {
        const size_t idx = (size_t)r * pitch + c;
        return u8 ? (float)u8[idx * stride] / 255.0f : f32[idx];
    }
};

// Where the signals of one launch come from (kernel parameter).
//   mode 0: float32 [nsig][H][W]                      (planes of CHW patches)
//   mode 1: uint8   [nsig / C][H][W][C]               (HWC patches, load_rgb_image order)
//   mode 2: float32 raster [C][Himg][Wimg], tiles of H x W taken every (sy, sx) pixels, nx tiles per tile row;
//           signal s = (tile tile0 + s / C, channel s % C) — the whole-scene tiler of BASELINE configs[4]
struct InputDesc {
    const void* ptr;
    int mode, C;
    int Himg, Wimg, nx, sy, sx;
    long long tile0;
    long long sig0;    // signal 0 of the launch is signal sig0 of the described input (a launch over the tail of a batch)
};

WST_D SignalSrc signal_source(const InputDesc& in, long long s, int H, int W) {
    SignalSrc src;
    s += in.sig0;
    src.f32 = nullptr; src.u8 = nullptr; src.stride = 1; src.pitch = W;
    if (in.mode == 1) {
        const long long b = s / in.C;
        src.u8 = static_cast<const unsigned char*>(in.ptr) + (size_t)b * H * W * in.C + (s - b * in.C);
        src.stride = in.C;
    } else if (in.mode == 2) {
        const long long t = in.tile0 + s / in.C;
        const int c = (int)(s % in.C), ty = (int)(t / in.nx), tx = (int)(t % in.nx);
        src.f32 = static_cast<const float*>(in.ptr) + ((size_t)c * in.Himg + (size_t)ty * in.sy) * in.Wimg + (size_t)tx * in.sx;
        src.pitch = in.Wimg;
    } else {
        src.f32 = static_cast<const float*>(in.ptr) + (size_t)s * H * W;
    }
    return src;
}

// ------------------------------------------------------------------ executors
// Phase tags (kind * 8 + level of the array side being processed) — only used by the cycle-profiling
// executor of the debug entry point; the production executor ignores them.
enum PhaseKind { PK_TWIDDLE = 0, PK_INPUT, PK_LP1, PK_LP2, PK_RFFT_ROW_S, PK_RFFT_ROW_C, PK_RFFT_SPLIT,
                 PK_RFFT_COL_S, PK_RFFT_COL_C, PK_U0_STORE, PK_PROD1, PK_PROD2, PK_IFFT_COL_C, PK_IFFT_COL_S,
                 PK_IFFT_ROW_C, PK_IFFT_FINAL, PK_LPR, PK_LPS, PK_POOL, PK_STAGE_LD, PK_STAGE_ST, PK_COUNT };
constexpr int kNumPhaseTags = PK_COUNT * 8;

#ifdef __CUDACC__
// Barrier closing a phase: the CTA's, or the cluster's (release/acquire at cluster scope, so workspace writes of
// one CTA are visible to the others afterwards).
template <int CL> WST_D void phase_barrier() {
    if constexpr (CL > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
}
WST_D int cluster_cta_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return (int)r; }
// `base` = rank of this CTA in its cluster * threads per CTA (0 without clusters): phase lambdas see the thread's
// index within the whole group that works on the signal.
template <int CL> struct DevExec {
    int base;
    template <int TAG, class F> WST_D void phase(F&& f) { f(base + (int)threadIdx.x); phase_barrier<CL>(); }
};
// Accumulates, per tag, the cycles from phase entry to barrier release as seen by thread 0.
template <int CL> struct ProfExec {
    int base;
    long long* acc;    // shared memory, kNumPhaseTags entries
    template <int TAG, class F> WST_D void phase(F&& f) {
        long long t0 = clock64();
        f(base + (int)threadIdx.x);
        phase_barrier<CL>();
        if (threadIdx.x == 0) acc[TAG] += clock64() - t0;
    }
};
#endif
// ------------------------------------------------------------------ bulk asynchronous copies (TMA) + mbarrier
// cp.async.bulk moves a contiguous run of bytes global -> shared without passing through registers and signals an
// mbarrier with the byte count; the waiting threads spin on the barrier's phase parity.  Used to fetch the next
// signal's pixels while the current signal is still being transformed (Cascade::prefetch_input).
#ifdef __CUDACC__
WST_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
WST_D void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
WST_D void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
WST_D void tma_bulk_g2s(void* dst_shared, const void* src_global, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_shared)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
WST_D void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
WST_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

template <int NT> struct HostExec {
    template <int TAG, class F> void phase(F&& f) { for (int t = 0; t < NT; ++t) f(t); }
};

// ------------------------------------------------------------------ FFT passes over shared memory
// A "line set": narr arrays (stride AS), NL lines per array (stride LS), elements of a line ES apart.
// Lanes run across lines, so both column transforms (LS = 1) and row transforms (LS = odd pitch)
// are bank-conflict free for 64-bit accesses.

// radix-R1 butterflies over elements {k1*R2 + i2}, optional twiddle after the butterfly.
// SUBFAST: consecutive threads take consecutive butterflies of the same line instead of the same butterfly of
// consecutive lines.  Used by the global-workspace variant for row transforms, where it is what coalesces.
template <int M, int DIR, bool TW, int NL, int LS, int ES, int NT, bool SUBFAST = false>
WST_D void pass_strided(int tid, cfloat* base, int narr, int AS, const cfloat* tw) {
    constexpr int R1 = Fft1<M>::R1, R2 = Fft1<M>::R2;
    const int total = narr * R2 * NL;
    auto load = [&](cfloat* p, cfloat (&a)[R1]) {
        static_for<0, R1>([&](auto K) { constexpr int k = decltype(K)::value; a[k] = p[k * R2 * ES]; });
    };
    auto store = [&](cfloat* p, int i2, cfloat (&a)[R1]) {
        static_for<0, R1>([&](auto K) {
            constexpr int k = decltype(K)::value;
            cfloat v = a[k];
            if constexpr (TW && !Fft1<M>::PFA && k > 0) {
                cfloat w = tw[k * R2 + i2];
                v = (DIR < 0) ? cmul(v, w) : cmulc(v, w);
            }
            p[k * R2 * ES] = v;
        });
    };
    if constexpr (WST_OPT_HOIST && !SUBFAST && NT % NL == 0) {
        // the thread's line is the same in every iteration; r = (array, sub-butterfly) advances by NT / NL
        constexpr int STEP = NT / NL;
        cfloat* pl = base + (tid % NL) * LS;
        const int rtot = narr * R2;
        if constexpr (WST_OPT_BATCH && R1 <= 8) {
            for (int r = tid / NL; r < rtot; r += 2 * STEP) {
                const bool two = r + STEP < rtot;
                const int rb = two ? r + STEP : r;
                const int i2a = r % R2, i2b = rb % R2;
                cfloat* pa = pl + (r / R2) * AS + i2a * ES;
                cfloat* pb = pl + (rb / R2) * AS + i2b * ES;
                cfloat a[R1], c[R1];
                load(pa, a); load(pb, c);
                dft<R1, DIR>(a); dft<R1, DIR>(c);
                store(pa, i2a, a);
                if (two) store(pb, i2b, c);
            }
        } else {
            for (int r = tid / NL; r < rtot; r += STEP) {
                const int i2 = r % R2;
                cfloat* p = pl + (r / R2) * AS + i2 * ES;
                cfloat a[R1];
                load(p, a);
                dft<R1, DIR>(a);
                store(p, i2, a);
            }
        }
        return;
    }
    for (int b = tid; b < total; b += NT) {
        int line, i2, g;
        if constexpr (SUBFAST) { i2 = b % R2; int r = b / R2; line = r % NL; g = r / NL; }
        else { line = b % NL; int r = b / NL; i2 = r % R2; g = r / R2; }
        cfloat* p = base + g * AS + line * LS + i2 * ES;
        cfloat a[R1];
        load(p, a);
        dft<R1, DIR>(a);
        store(p, i2, a);
    }
}

// radix-R2 butterflies over contiguous elements {k1*R2 + i2}, optional twiddle after.
template <int M, int DIR, bool TW, int NL, int LS, int ES, int NT, bool SUBFAST = false>
WST_D void pass_contig(int tid, cfloat* base, int narr, int AS, const cfloat* tw) {
    constexpr int R1 = Fft1<M>::R1, R2 = Fft1<M>::R2;
    const int total = narr * R1 * NL;
    auto load = [&](cfloat* p, cfloat (&v)[R2]) {
        static_for<0, R2>([&](auto I) { constexpr int i = decltype(I)::value; v[i] = p[i * ES]; });
    };
    auto store = [&](cfloat* p, int k1, cfloat (&v)[R2]) {
        static_for<0, R2>([&](auto I) {
            constexpr int i = decltype(I)::value;
            cfloat o = v[i];
            if constexpr (TW && !Fft1<M>::PFA && R1 > 1 && i > 0) {
                cfloat w = tw[k1 * R2 + i];
                o = (DIR < 0) ? cmul(o, w) : cmulc(o, w);
            }
            p[i * ES] = o;
        });
    };
    if constexpr (WST_OPT_HOIST && !SUBFAST && NT % NL == 0) {
        constexpr int STEP = NT / NL;
        cfloat* pl = base + (tid % NL) * LS;
        const int rtot = narr * R1;
        if constexpr (WST_OPT_BATCH && R2 <= 10) {
            for (int r = tid / NL; r < rtot; r += 2 * STEP) {
                const bool two = r + STEP < rtot;
                const int rb = two ? r + STEP : r;
                const int k1a = r % R1, k1b = rb % R1;
                cfloat* pa = pl + (r / R1) * AS + k1a * R2 * ES;
                cfloat* pb = pl + (rb / R1) * AS + k1b * R2 * ES;
                cfloat v[R2], u[R2];
                load(pa, v); load(pb, u);
                dft<R2, DIR>(v); dft<R2, DIR>(u);
                store(pa, k1a, v);
                if (two) store(pb, k1b, u);
            }
        } else {
            for (int r = tid / NL; r < rtot; r += STEP) {
                const int k1 = r % R1;
                cfloat* p = pl + (r / R1) * AS + k1 * R2 * ES;
                cfloat v[R2];
                load(p, v);
                dft<R2, DIR>(v);
                store(p, k1, v);
            }
        }
        return;
    }
    for (int b = tid; b < total; b += NT) {
        int line, k1, g;
        if constexpr (SUBFAST) { k1 = b % R1; int r = b / R1; line = r % NL; g = r / NL; }
        else { line = b % NL; int r = b / NL; k1 = r % R1; g = r / R1; }
        cfloat* p = base + g * AS + line * LS + k1 * R2 * ES;
        cfloat v[R2];
        load(p, v);
        dft<R2, DIR>(v);
        store(p, k1, v);
    }
}

// forward 1-D transforms of all lines: natural -> digit-swapped
template <int M, int NL, int LS, int ES, int NT, int TAG_S, int TAG_C, bool GLOB = false, class Exec>
WST_D void fft_lines_fwd(Exec& ex, cfloat* base, int narr, int AS, const cfloat* tw) {
    constexpr bool SF = GLOB && ES == 1;
    if constexpr (Fft1<M>::R1 > 1) {
        ex.template phase<TAG_S>([&](int tid) { pass_strided<M, -1, true, NL, LS, ES, NT, SF>(tid, base, narr, AS, tw); });
    }
    ex.template phase<TAG_C>([&](int tid) { pass_contig<M, -1, false, NL, LS, ES, NT, SF>(tid, base, narr, AS, tw); });
}

// inverse 1-D transforms of all lines: digit-swapped -> natural (unnormalised)
template <int M, int NL, int LS, int ES, int NT, int TAG_C, int TAG_S, bool GLOB = false, class Exec>
WST_D void fft_lines_inv(Exec& ex, cfloat* base, int narr, int AS, const cfloat* tw) {
    constexpr bool SF = GLOB && ES == 1;
    ex.template phase<TAG_C>([&](int tid) { pass_contig<M, +1, true, NL, LS, ES, NT, SF>(tid, base, narr, AS, tw); });
    if constexpr (Fft1<M>::R1 > 1) {
        ex.template phase<TAG_S>([&](int tid) { pass_strided<M, +1, false, NL, LS, ES, NT, SF>(tid, base, narr, AS, tw); });
    }
}

// Last pass of the inverse row transform, fused with modulus, row pairing and (optionally) the first
// half of the separable low-pass.
//   WRITE_Z: z[x][y] = ( |u[x][y]|, |u[x + M/2][y]| ), x < M/2, written over row x (input of rfft2_from_pairs)
//   LPF:     each thread contracts the |u| values it holds with Gc over y and parks the 2*HOUT partial sums
//            in shared-memory slots it owns (they were its own inputs):
//              no z :  row r, column slot (q*R2 + i2) <- (acc_r[2q], acc_r[2q+1]),  q < HOUT/2
//              z    :  row x+M/2, slot (q*R2 + i2) <- acc_x[..],  slot ((HOUT/2+q)*R2 + i2) <- acc_{x+M/2}[..]
//            (single-pass lengths hold whole rows: the slot is column q and the sum over y is complete).
//   XN, POFF: rows x < XN are paired with the rows POFF elements further on (the whole array: M/2 and M/2 rows; a
//            staged tile of the global-workspace variant: its row pairs, stored top half then bottom half).
template <int M, int NT, bool WRITE_Z, bool LPF, int HOUT, int HP, bool SUBFAST = false, int XN = M / 2,
          int POFF = (M / 2) * (M + 1)>
WST_D void pass_rows_final(int tid, cfloat* base, int narr, int AS, const float* gc) {
    constexpr int R1 = Fft1<M>::R1, R2 = Fft1<M>::R2, P = M + 1, HALF = XN;
    constexpr int NV = (R1 > 1) ? R1 : R2;            // values per row held by one thread
    constexpr int YS = (R1 > 1) ? R2 : 1;             // their stride along y
    const int nsub = (R1 > 1) ? R2 : 1;
    const int total = narr * nsub * HALF;
    // HOIST: the thread's row pair is the same in every iteration, b = (array, sub-butterfly) advances by NT / HALF
    constexpr bool HOIST = WST_OPT_HOIST && !SUBFAST && NT % HALF == 0;
    for (int b = HOIST ? tid / HALF : tid; b < (HOIST ? narr * nsub : total); b += (HOIST ? NT / HALF : NT)) {
        int x, i2, g;
        if constexpr (HOIST) { x = tid % HALF; i2 = b % nsub; g = b / nsub; }
        else if constexpr (SUBFAST) { i2 = b % nsub; int r = b / nsub; x = r % HALF; g = r / HALF; }
        else { x = b % HALF; int r = b / HALF; i2 = r % nsub; g = r / nsub; }
        cfloat* p0 = base + g * AS + x * P + i2;
        cfloat* p1 = p0 + POFF;
        cfloat a[NV];
        if constexpr (NV > 16 || (WST_OPT_ROWSEQ && NV >= 8)) {
            // long butterflies (the 24-point lines of 576): one row at a time, so that only one butterfly's 2*NV
            // registers are live together with the NV moduli of the other
            float mc[NV];
            {
                cfloat c[NV];
                static_for<0, NV>([&](auto K) { constexpr int k = decltype(K)::value; c[k] = p1[k * YS]; });
                dft<NV, +1>(c);
                static_for<0, NV>([&](auto K) { constexpr int k = decltype(K)::value; mc[k] = cabs_(c[k]); });
            }
            static_for<0, NV>([&](auto K) { constexpr int k = decltype(K)::value; a[k] = p0[k * YS]; });
            dft<NV, +1>(a);
            static_for<0, NV>([&](auto K) {
                constexpr int k = decltype(K)::value;
                a[k].x = cabs_(a[k]);
                a[k].y = mc[k];
            });
        } else {
            cfloat c[NV];
            static_for<0, NV>([&](auto K) { constexpr int k = decltype(K)::value; a[k] = p0[k * YS]; c[k] = p1[k * YS]; });
            dft<NV, +1>(a);
            dft<NV, +1>(c);
            static_for<0, NV>([&](auto K) {
                constexpr int k = decltype(K)::value;
                a[k].x = cabs_(a[k]);
                a[k].y = cabs_(c[k]);
            });
        }
        if constexpr (LPF) {
            float acc0[HOUT], acc1[HOUT];
            static_for<0, HOUT>([&](auto I) { acc0[decltype(I)::value] = 0.f; acc1[decltype(I)::value] = 0.f; });
            static_for<0, NV>([&](auto K) {
                constexpr int k = decltype(K)::value;
                const float* gy = gc + (k * YS + i2) * HP;
                static_for<0, HP / 4>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    float4 w = *reinterpret_cast<const float4*>(gy + 4 * q);
                    if constexpr (4 * q + 0 < HOUT) { acc0[4 * q + 0] += w.x * a[k].x; acc1[4 * q + 0] += w.x * a[k].y; }
                    if constexpr (4 * q + 1 < HOUT) { acc0[4 * q + 1] += w.y * a[k].x; acc1[4 * q + 1] += w.y * a[k].y; }
                    if constexpr (4 * q + 2 < HOUT) { acc0[4 * q + 2] += w.z * a[k].x; acc1[4 * q + 2] += w.z * a[k].y; }
                    if constexpr (4 * q + 3 < HOUT) { acc0[4 * q + 3] += w.w * a[k].x; acc1[4 * q + 3] += w.w * a[k].y; }
                });
            });
            if constexpr (WRITE_Z) {
                static_for<0, NV>([&](auto K) { constexpr int k = decltype(K)::value; p0[k * YS] = a[k]; });
                static_for<0, HOUT / 2>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    p1[q * YS] = cmake(acc0[2 * q], acc0[2 * q + 1]);
                    p1[(HOUT / 2 + q) * YS] = cmake(acc1[2 * q], acc1[2 * q + 1]);
                });
            } else {
                static_for<0, HOUT / 2>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    p0[q * YS] = cmake(acc0[2 * q], acc0[2 * q + 1]);
                    p1[q * YS] = cmake(acc1[2 * q], acc1[2 * q + 1]);
                });
            }
        } else {
            static_for<0, NV>([&](auto K) { constexpr int k = decltype(K)::value; p0[k * YS] = a[k]; });
        }
    }
}

// Second half of the fused low-pass: reduce the parked partial sums over the threads of a row, contract
// with Gr over chunks of rows (phase LPR), then sum the chunks and store the map (phase LPS).
template <int M, int NT, bool WRITE_Z, int HOUT, int HP, int LPSLOTS, int LV, class Exec, class CoefFn>
WST_D void lowpass_reduce(Exec& ex, cfloat* base, int narr, int AS, const float* gr, float* lpbuf, float* maps,
                          CoefFn coef) {
    constexpr int R1 = Fft1<M>::R1, R2 = Fft1<M>::R2, P = M + 1, HALF = M / 2;
    constexpr int NSUB = (R1 > 1) ? R2 : 1, YS = (R1 > 1) ? R2 : 1;
    int nchunks = LPSLOTS / narr;
    if (nchunks > (M + 3) / 4) nchunks = (M + 3) / 4;
    if (nchunks < 1) nchunks = 1;
    const int rc = (M + nchunks - 1) / nchunks;
    ex.template phase<PK_LPR * 8 + LV>([&](int tid) {
        const int total = narr * nchunks * HOUT;
        for (int b = tid; b < total; b += NT) {
            int ic = b % HOUT, r = b / HOUT;
            int ch = r % nchunks, g = r / nchunks;
            const float* fb = reinterpret_cast<const float*>(base + g * AS);
            float acc[HOUT];
            static_for<0, HOUT>([&](auto I) { acc[decltype(I)::value] = 0.f; });
            int r0 = ch * rc, r1 = r0 + rc < M ? r0 + rc : M;
            for (int row = r0; row < r1; ++row) {
                int srow, scol;
                if constexpr (WRITE_Z) { srow = row < HALF ? row + HALF : row; scol = (row < HALF ? 0 : HOUT / 2) + ic / 2; }
                else { srow = row; scol = ic / 2; }
                const float* sp = fb + (size_t)(srow * P + scol * YS) * 2 + (ic & 1);
                float t = 0.f;
                for (int i2 = 0; i2 < NSUB; ++i2) t += sp[2 * i2];
                const float* gx = gr + row * HP;
                static_for<0, HOUT>([&](auto I) { acc[decltype(I)::value] += gx[decltype(I)::value] * t; });
            }
            float* o = lpbuf + (size_t)(g * nchunks + ch) * HOUT * HOUT + ic;
            static_for<0, HOUT>([&](auto I) { o[decltype(I)::value * HOUT] = acc[decltype(I)::value]; });
        }
    });
    ex.template phase<PK_LPS * 8 + LV>([&](int tid) {
        const int total = narr * HOUT * HOUT;
        for (int b = tid; b < total; b += NT) {
            int e = b % (HOUT * HOUT), g = b / (HOUT * HOUT);
            int cidx = coef(g);
            if (cidx < 0) continue;
            const float* o = lpbuf + (size_t)g * nchunks * HOUT * HOUT + e;
            float sacc = 0.f;
            for (int ch = 0; ch < nchunks; ++ch) sacc += o[ch * HOUT * HOUT];
            maps[(size_t)cidx * (HOUT * HOUT) + e] = sacc;
        }
    });
}

// L2 eviction-priority hints (createpolicy + .L2::cache_hint); host emulation: plain accesses.
template <bool LAST> WST_D unsigned long long l2_policy() {
#ifdef __CUDA_ARCH__
    unsigned long long pol;
    if constexpr (LAST) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
#else
    return 0;
#endif
}
// Filter-bank loads (read-only for the kernel's lifetime).  FH: keep the line in L2 (evict_last).
template <bool FH> WST_D float ldf1(const float* p) {
#ifdef __CUDA_ARCH__
    if constexpr (FH) { float v; asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(l2_policy<true>())); return v; }
#endif
    return *p;
}
template <bool FH> WST_D float2 ldf2(const float* p) {
#ifdef __CUDA_ARCH__
    if constexpr (FH) { float2 v; asm volatile("ld.global.nc.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(l2_policy<true>())); return v; }
#endif
    return *reinterpret_cast<const float2*>(p);
}
template <bool FH> WST_D float4 ldf4(const float* p) {
#ifdef __CUDA_ARCH__
    if constexpr (FH) { float4 v; asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(l2_policy<true>())); return v; }
#endif
    return *reinterpret_cast<const float4*>(p);
}
// Product outputs written to the workspace (OH: evict_first — a group of arrays is written in full before its first tile
// is read back).
template <bool OH> WST_D void prod_store(cfloat* dst, cfloat v) {
#ifdef __CUDA_ARCH__
    if constexpr (OH) {
        asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(dst), "f"(v.x), "f"(v.y), "l"(l2_policy<false>()) : "memory");
        return;
    }
#endif
    *dst = v;
}
// Workspace tile stores of the staged passes (the tile is not read again before the rest of the array has streamed by).
WST_D void stage_store(cfloat* dst_global, cfloat v) {
#ifdef __CUDA_ARCH__
    if constexpr ((WST_OPT_L2HINT & 4) != 0) {
        asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(dst_global), "f"(v.x), "f"(v.y), "l"(l2_policy<false>()) : "memory");
        return;
    }
#endif
    *dst_global = v;
}

// Tile loads of the staged passes: asynchronous 8-byte global -> shared copies (LDGSTS), so that a thread's whole
// share of the tile is in flight at once instead of one register round trip after another; stage_copy_wait() closes the
// phase that issued them.  (Host emulation: plain copies.)
WST_D void stage_copy(cfloat* dst_shared, const cfloat* src_global) {
#ifdef __CUDA_ARCH__
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_shared);
    if constexpr ((WST_OPT_L2HINT & 2) != 0)
        asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(d), "l"(src_global), "l"(l2_policy<false>()) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src_global) : "memory");
#else
    *dst_shared = *src_global;
#endif
}
WST_D void stage_copy_wait() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}

// Tile loop of the staged passes.  load(tid, t, buf) issues the asynchronous copies of tile t, first(tid, t, buf) is the
// first transform pass, mid(t, buf) runs the remaining transform phases, store(tid, t, buf) writes the tile back.
// With WST_STAGE_BUFS = 2 the loop is software-pipelined: tile t lives in stage buffer t & 1 and the copies of tile
// t + 1 are issued at the start of tile t's first pass and awaited in its store phase.  Measured slower at 512 x 512 J=5
// (786 against 914 tiles/s: the second 78 KB buffer has to be paid for with shared memory that the L1 of the product
// phases was using, and issuing the copies lengthens the transform phases by most of what the load phase took), so the
// default is one buffer.
template <int TAG_FIRST, int TAG_ST, int LV, class Exec, class Load, class First, class Mid, class Store>
WST_D void staged_tiles(Exec& ex, int ntiles, cfloat* stage, int stage_cf, Load load, First first, Mid mid, Store store) {
    constexpr bool PIPE = WST_STAGE_BUFS > 1;
    if constexpr (PIPE) ex.template phase<PK_STAGE_LD * 8 + LV>([&](int tid) { load(tid, 0, stage); stage_copy_wait(); });
    for (int t = 0; t < ntiles; ++t) {
        cfloat* cur = PIPE ? stage + (t & 1) * stage_cf : stage;
        cfloat* nxt = stage + ((t + 1) & 1) * stage_cf;
        if constexpr (!PIPE) ex.template phase<PK_STAGE_LD * 8 + LV>([&](int tid) { load(tid, t, cur); stage_copy_wait(); });
        ex.template phase<TAG_FIRST * 8 + LV>([&](int tid) {
            if (PIPE && t + 1 < ntiles) load(tid, t + 1, nxt);
            first(tid, t, cur);
        });
        mid(t, cur);
        ex.template phase<TAG_ST * 8 + LV>([&](int tid) { store(tid, t, cur); if (PIPE) stage_copy_wait(); });
    }
}

// Global-workspace variant of rfft2_from_pairs: the two row passes and the split run on shared-memory tiles of TZ rows
// of z (read once, coalesced; the split writes whole rows of U^), the two column passes on tiles of 16 columns of U^.
template <int M, int NT, int STAGE, int LV, class Exec>
WST_D void rfft2_from_pairs_staged(Exec& ex, cfloat* z, int ZS, cfloat* uh, int narr, const cfloat* tw, cfloat* stage) {
    constexpr int P = M + 1, HALF = M / 2, PH = M / 2 + 1, UHS = M * PH;
    constexpr bool TWO = Fft1<M>::R1 > 1;
    {   // rows
        constexpr int TZ = (HALF % 16 == 0 && 16 * P <= STAGE) ? 16 : (HALF % 8 == 0 && 8 * P <= STAGE) ? 8
                           : (HALF % 4 == 0 && 4 * P <= STAGE) ? 4 : (HALF % 2 == 0 && 2 * P <= STAGE) ? 2 : 1;
        constexpr int TILE = TZ * P, NAB = STAGE / TILE, NX = HALF / TZ;
        static_assert(NAB >= 1, "stage too small for a row tile");
        const int nbatch = (narr + NAB - 1) / NAB;
        auto geom = [&](int t, int& g0, int& na, int& x0) { g0 = (t / NX) * NAB; na = narr - g0 < NAB ? narr - g0 : NAB; x0 = (t % NX) * TZ; };
        staged_tiles<TWO ? PK_RFFT_ROW_S : PK_RFFT_ROW_C, PK_RFFT_SPLIT, LV>(ex, nbatch * NX, stage, STAGE,
            [&](int tid, int t, cfloat* buf) {
                int g0, na, x0; geom(t, g0, na, x0);
                for (int rq = tid >> 5; rq < na * TZ; rq += NT / 32) {       // one warp per tile row, lanes along it
                    const int q = rq % TZ, a = rq / TZ;
                    const cfloat* src = z + (g0 + a) * ZS + (x0 + q) * P;
                    cfloat* dst = buf + a * TILE + q * P;
                    for (int c = tid & 31; c < M; c += 32) stage_copy(dst + c, src + c);
                }
            },
            [&](int tid, int t, cfloat* buf) {
                int g0, na, x0; geom(t, g0, na, x0);
                if constexpr (TWO) pass_strided<M, -1, true, TZ, P, 1, NT>(tid, buf, na, TILE, tw);
                else pass_contig<M, -1, false, TZ, P, 1, NT>(tid, buf, na, TILE, tw);
            },
            [&](int t, cfloat* buf) {
                if constexpr (TWO) {
                    int g0, na, x0; geom(t, g0, na, x0);
                    ex.template phase<PK_RFFT_ROW_C * 8 + LV>([&](int tid) { pass_contig<M, -1, false, TZ, P, 1, NT>(tid, buf, na, TILE, tw); });
                }
            },
            // split the packed rows:  A = FFT(row x), B = FFT(row x + M/2); consecutive threads write consecutive l
            [&](int tid, int t, cfloat* buf) {
                int g0, na, x0; geom(t, g0, na, x0);
                for (int rq = tid >> 5; rq < na * TZ; rq += NT / 32) {
                    const int q = rq % TZ, a = rq / TZ;
                    const cfloat* zr = buf + a * TILE + q * P;
                    cfloat* ua = uh + (g0 + a) * UHS + (x0 + q) * PH;
                    cfloat* ub = ua + HALF * PH;
                    for (int l = tid & 31; l < PH; l += 32) {
                        const cfloat zl = zr[Fft1<M>::pi(l)];
                        const cfloat zm = zr[Fft1<M>::pi(l == 0 ? 0 : M - l)];
                        ua[l] = cmake(0.5f * (zl.x + zm.x), 0.5f * (zl.y - zm.y));
                        ub[l] = cmake(0.5f * (zl.y + zm.y), -0.5f * (zl.x - zm.x));
                    }
                }
            });
    }
    {   // columns of U^ (PH of them, not a multiple of the tile width: the last tile is partly idle)
        constexpr int TC = M * 17 <= STAGE ? 16 : M * 9 <= STAGE ? 8 : M * 5 <= STAGE ? 4 : 2;
        constexpr int TCP = TC + 1, TILE = M * TCP, NAB = STAGE / TILE, NCT = (PH + TC - 1) / TC;
        static_assert(NAB >= 1, "stage too small for a column tile");
        const int nbatch = (narr + NAB - 1) / NAB;
        auto geom = [&](int t, int& g0, int& na, int& c0, int& tc) {
            g0 = (t / NCT) * NAB; na = narr - g0 < NAB ? narr - g0 : NAB; c0 = (t % NCT) * TC; tc = PH - c0 < TC ? PH - c0 : TC;
        };
        staged_tiles<TWO ? PK_RFFT_COL_S : PK_RFFT_COL_C, PK_STAGE_ST, LV>(ex, nbatch * NCT, stage, STAGE,
            [&](int tid, int t, cfloat* buf) {
                int g0, na, c0, tc; geom(t, g0, na, c0, tc);
                const int c = tid % TC;                                         // TC consecutive threads per tile row
                for (int a = 0; a < na; ++a) {
                    const cfloat* src = uh + (g0 + a) * UHS + c0 + c;
                    cfloat* dst = buf + a * TILE + c;
                    for (int r = tid / TC; r < M; r += NT / TC) {
                        if (c < tc) stage_copy(dst + r * TCP, src + r * PH);
                        else dst[r * TCP] = cmake(0.f, 0.f);
                    }
                }
            },
            [&](int tid, int t, cfloat* buf) {
                int g0, na, c0, tc; geom(t, g0, na, c0, tc);
                if constexpr (TWO) pass_strided<M, -1, true, TC, 1, TCP, NT>(tid, buf, na, TILE, tw);
                else pass_contig<M, -1, false, TC, 1, TCP, NT>(tid, buf, na, TILE, tw);
            },
            [&](int t, cfloat* buf) {
                if constexpr (TWO) {
                    int g0, na, c0, tc; geom(t, g0, na, c0, tc);
                    ex.template phase<PK_RFFT_COL_C * 8 + LV>([&](int tid) { pass_contig<M, -1, false, TC, 1, TCP, NT>(tid, buf, na, TILE, tw); });
                }
            },
            [&](int tid, int t, cfloat* buf) {
                int g0, na, c0, tc; geom(t, g0, na, c0, tc);
                const int c = tid % TC;
                if (c < tc)
                    for (int a = 0; a < na; ++a) {
                        cfloat* dst = uh + (g0 + a) * UHS + c0 + c;
                        const cfloat* src = buf + a * TILE + c;
                        for (int r = tid / TC; r < M; r += NT / TC) dst[r * PH] = src[r * TCP];
                    }
            });
    }
}

// Real 2-D forward FFT of narr paired-row arrays z (stride ZS, pitch M+1, M/2 rows) into
// half spectra U^[pi(k)][l], l = 0..M/2 (stride UHS = M*(M/2+1), pitch M/2+1).
template <int M, int NT, int LV, bool GLOB = false, int STAGE = 0, class Exec>
WST_D void rfft2_from_pairs(Exec& ex, cfloat* z, int ZS, cfloat* uh, int narr, const cfloat* tw, cfloat* stage = nullptr) {
    constexpr int P = M + 1, HALF = M / 2, PH = M / 2 + 1, UHS = M * PH;
    if constexpr (GLOB && STAGE > 0) { rfft2_from_pairs_staged<M, NT, STAGE, LV>(ex, z, ZS, uh, narr, tw, stage); return; }
    // rows of z (along y): natural -> swapped
    fft_lines_fwd<M, HALF, P, 1, NT, PK_RFFT_ROW_S * 8 + LV, PK_RFFT_ROW_C * 8 + LV, GLOB>(ex, z, narr, ZS, tw);
    // split the packed rows:  A = FFT(row x), B = FFT(row x + M/2)
    ex.template phase<PK_RFFT_SPLIT * 8 + LV>([&](int tid) {
        const int total = narr * PH * HALF;
        for (int b = tid; b < total; b += NT) {
            int x = b % HALF, r = b / HALF;
            int l = r % PH, g = r / PH;
            const cfloat* zr = z + g * ZS + x * P;
            cfloat zl = zr[Fft1<M>::pi(l)];
            cfloat zm = zr[Fft1<M>::pi(l == 0 ? 0 : M - l)];
            cfloat A = cmake(0.5f * (zl.x + zm.x), 0.5f * (zl.y - zm.y));
            cfloat B = cmake(0.5f * (zl.y + zm.y), -0.5f * (zl.x - zm.x));
            cfloat* u = uh + g * UHS + l;
            u[x * PH] = A;
            u[(x + HALF) * PH] = B;
        }
    });
    // columns of U^ (along rows index): natural -> swapped
    fft_lines_fwd<M, PH, 1, PH, NT, PK_RFFT_COL_S * 8 + LV, PK_RFFT_COL_C * 8 + LV>(ex, uh, narr, UHS, tw);
}

// Hermitian lookup U(k, l) from a half spectrum stored as [pi(k)][l], l <= M/2.
template <int M>
WST_D cfloat herm_get(const cfloat* uh, int k, int l) {
    constexpr int PH = M / 2 + 1;
    if constexpr (WST_OPT_HERMSEL) {
        const bool dir = l <= M / 2;
        const int kk = dir ? k : (k == 0 ? 0 : M - k), ll = dir ? l : M - l;
        cfloat v = uh[Fft1<M>::pi(kk) * PH + ll];
        v.y = dir ? v.y : -v.y;
        return v;
    }
    if (l <= M / 2) return uh[Fft1<M>::pi(k) * PH + l];
    int kk = (k == 0) ? 0 : M - k;
    cfloat v = uh[Fft1<M>::pi(kk) * PH + (M - l)];
    return cmake(v.x, -v.y);
}

// The aliases a in [0, F) for which index base + a*MC lies inside the cyclic span (lo << 16 | len) of a
// length-MP axis are consecutive modulo F: {first, first+1, ...} (count of them).  Branch-free.
struct AliasRun { int first, count; };
template <int MP, int MC>
WST_D AliasRun alias_run(int base, int packed) {
    constexpr int F = MP / MC;
    const int lo = packed >> 16, len = packed & 0xffff;
    int t = lo - base;
    t += t < 0 ? MP : 0;                              // distance from base up to the span start, cyclic
    const int a0 = (t + MC - 1) / MC;                 // first alias at or after the span start (may equal F)
    const int d0 = a0 * MC - t;                       // its offset inside the span, in [0, MC)
    int cnt = d0 < len ? (len - 1 - d0) / MC + 1 : 0;
    AliasRun r;
    r.first = a0 >= F ? a0 - F : a0;
    r.count = cnt > F ? F : cnt;
    return r;
}

// V_g = fold( U . psi_g ) for the GS filters of one theta-group, written digit-swapped and pre-scaled by
// 1/(F^2 * MC^2) (fold mean and inverse-FFT normalisation).
//   uh   : parent half spectrum, side MP (shared or global memory)
//   filt : [MP][MP][GS] real filters of this group, orientations interleaved so that one vector load feeds
//          GS accumulators and the index arithmetic is shared
//   rows, cols : cyclic bounding box of the group's support.  psi^ of scale j is a Gaussian bump around the
//          origin, so at fold factors >= 4 most aliases fall outside it.  The aliases inside are consecutive
//          modulo F: rows are a short runtime loop, columns a fixed-length straight-line batch of NB loads
//          starting at the first alias inside (extra ones only add exact zeros' worth of tail), so the loads
//          of a row are independent and in flight together — these phases are L2-latency bound.
//   out  : GS arrays of MC x MC (pitch MC+1, stride MC*(MC+1))
template <int MP, int GS, bool FH = false>
WST_D void load_filter_vec(const float* fp, float (&w)[GS]) {
    if constexpr (GS % 4 == 0) {
        static_for<0, GS / 4>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            float4 t = ldf4<FH>(fp + 4 * q);
            w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
        });
    } else if constexpr (GS == 2) {
        float2 t = ldf2<FH>(fp);
        w[0] = t.x; w[1] = t.y;
    } else {
        static_for<0, GS>([&](auto G) { w[decltype(G)::value] = ldf1<FH>(fp + decltype(G)::value); });
    }
}

template <int MP, int MC, int GS, int NT, bool FH = false, bool OH = false>
WST_D void product_fold(int tid, const cfloat* uh, const float* filt, int rows, int cols, cfloat* out) {
    constexpr int F = MP / MC, PC = MC + 1, AS = MC * (MC + 1);
    constexpr bool SPARSE = F >= 4;
    constexpr float scale = 1.0f / ((float)F * (float)F * (float)MC * (float)MC);
    if constexpr (WST_OPT_PAIRHALF && !SPARSE && GS <= 2 && MC % 2 == 0 && MP % 4 == 0) {
        constexpr int HM = MP / 2, PH = MP / 2 + 1, HC = MC / 2;
        static_assert(HM % HC == 0, "alias columns must fall wholly on one side of the half spectrum");
        for (int o = tid; o < MC * HC; o += NT) {
            const int lc = o % HC, kc = o / HC;
            float ar[2][GS], ai[2][GS];
            static_for<0, 2 * GS>([&](auto E) { ar[decltype(E)::value / GS][decltype(E)::value % GS] = 0.f;
                                                ai[decltype(E)::value / GS][decltype(E)::value % GS] = 0.f; });
            static_for<0, F>([&](auto A) {
                constexpr int a = decltype(A)::value;
                const int k = kc + a * MC;
                const cfloat* rd = uh + Fft1<MP>::pi(k) * PH;                       // row k, columns 0..MP/2
                const cfloat* rm = uh + Fft1<MP>::pi(k == 0 ? 0 : MP - k) * PH;       // row -k for the mirrored half
                const float* frow = filt + (size_t)k * MP * GS;
                static_for<0, 2 * F>([&](auto Q) {
                    constexpr int q = decltype(Q)::value, p = q % 2, off = (q / 2) * MC + p * HC;   // column l = lc + off
                    constexpr bool MIR = off >= HM;            // every lc < HC lands on the same side of column MP/2
                    const int l = lc + off;
                    const cfloat u = MIR ? rm[MP - l] : rd[l];                  // U(k, l) = conj(U(-k, -l))
                    float w[GS];
                    if constexpr (GS == 2) { const float2 t = ldf2<FH>(frow + (size_t)l * 2); w[0] = t.x; w[1] = t.y; }
                    else w[0] = ldf1<FH>(frow + l);
                    static_for<0, GS>([&](auto G) {
                        constexpr int g = decltype(G)::value;
                        ar[p][g] += u.x * w[g];
                        if constexpr (MIR) ai[p][g] -= u.y * w[g]; else ai[p][g] += u.y * w[g];
                    });
                });
            });
            cfloat* orow = out + Fft1<MC>::pi(kc) * PC;
            static_for<0, 2>([&](auto Pp) {
                constexpr int p = decltype(Pp)::value;
                cfloat* op = orow + Fft1<MC>::pi(lc + p * HC);
                static_for<0, GS>([&](auto G) {
                    constexpr int g = decltype(G)::value;
                    prod_store<OH>(op + g * AS, cmake(ar[p][g] * scale, ai[p][g] * scale));
                });
            });
        }
    } else if constexpr (!SPARSE && GS <= 2 && MC % 2 == 0 && MP % 4 == 0) {
        // dense, few orientations per pass (the 80 x 80 children and the full-resolution parents, where only
        // one or two arrays fit): these phases are issue-bound and the index arithmetic is shared by only GS
        // accumulators, so each thread takes two adjacent outputs (l even, l+1): one 8/16-byte filter load and
        // one row/mirror computation serve both.
        constexpr int HM = MP / 2, PH = MP / 2 + 1, HC = MC / 2;
        // HOIST: a thread keeps its output column pair; everything that depends on the column only (mirror branch,
        // column offsets, output column positions) leaves the loop
        constexpr bool HOIST = (WST_OPT_HOIST_PROD || WST_OPT_PREFETCH) && NT % HC == 0;
        constexpr bool PREF = WST_OPT_PREFETCH && HOIST;
        // filter values of one output pair: alias (a, b) -> [pair member][orientation]
        auto load_w = [&](const float* fb, float (&w)[F * F][2][GS]) {
            static_for<0, F * F>([&](auto S) {
                constexpr int sl = decltype(S)::value, a = sl / F, b = sl % F;
                const float* fp = fb + ((size_t)a * MC * MP + b * MC) * GS;
                if constexpr (GS == 2) {
                    float4 t = ldf4<FH>(fp);
                    w[sl][0][0] = t.x; w[sl][0][1] = t.y; w[sl][1][0] = t.z; w[sl][1][1] = t.w;
                } else {
                    float2 t = ldf2<FH>(fp);
                    w[sl][0][0] = t.x; w[sl][1][0] = t.y;
                }
            });
        };
        float wn[F * F][2][GS];                       // PREF: the next pair's filter values, in flight during this pair
        if constexpr (PREF) {
            const int k0 = tid / HC;
            load_w(filt + ((size_t)(k0 < MC ? k0 : 0) * MP + (tid % HC) * 2) * GS, wn);
        }
        for (int o = HOIST ? tid / HC : tid; o < (HOIST ? MC : MC * HC); o += (HOIST ? NT / HC : NT)) {
            const int lc = HOIST ? (tid % HC) * 2 : (o % HC) * 2, kc = HOIST ? o : o / HC;
            float ar[2][GS], ai[2][GS];
            static_for<0, 2 * GS>([&](auto E) { ar[decltype(E)::value / GS][decltype(E)::value % GS] = 0.f;
                                                ai[decltype(E)::value / GS][decltype(E)::value % GS] = 0.f; });
            float w[F * F][2][GS];
            cfloat u[F * F][2];
            if constexpr (PREF) {
                static_for<0, F * F * 2 * GS>([&](auto E) {
                    constexpr int e = decltype(E)::value;
                    w[e / (2 * GS)][(e / GS) % 2][e % GS] = wn[e / (2 * GS)][(e / GS) % 2][e % GS];
                });
                const int kn = kc + NT / HC;           // rows past the end re-read row kc (never used)
                load_w(filt + ((size_t)(kn < MC ? kn : kc) * MP + lc) * GS, wn);
            } else {
                load_w(filt + ((size_t)kc * MP + lc) * GS, w);
            }
            static_for<0, F>([&](auto A) {
                constexpr int a = decltype(A)::value;
                const int k = kc + a * MC;
                const cfloat* rd = uh + Fft1<MP>::pi(k) * PH;                       // row k, columns 0..MP/2
                const cfloat* rm = uh + Fft1<MP>::pi(k == 0 ? 0 : MP - k) * PH;       // row -k for the mirrored half
                static_for<0, F>([&](auto B) {
                    constexpr int b = decltype(B)::value, sl = a * F + b;
                    const int l = lc + b * MC;                                      // even
                    // U(k, l) and U(k, l+1) from the Hermitian half spectrum
                    if (l + 1 <= HM) { u[sl][0] = rd[l]; u[sl][1] = rd[l + 1]; }
                    else if (l >= HM + 1) { cfloat p = rm[MP - l], q = rm[MP - l - 1];
                                            u[sl][0] = cmake(p.x, -p.y); u[sl][1] = cmake(q.x, -q.y); }
                    else { cfloat q = rm[MP - l - 1]; u[sl][0] = rd[l]; u[sl][1] = cmake(q.x, -q.y); }   // l == MP/2
                });
            });
            static_for<0, F * F>([&](auto S) {
                constexpr int sl = decltype(S)::value;
                static_for<0, 2>([&](auto Pp) {
                    constexpr int p = decltype(Pp)::value;
                    static_for<0, GS>([&](auto G) {
                        constexpr int g = decltype(G)::value;
                        ar[p][g] += u[sl][p].x * w[sl][p][g];
                        ai[p][g] += u[sl][p].y * w[sl][p][g];
                    });
                });
            });
            cfloat* orow = out + Fft1<MC>::pi(kc) * PC;
            static_for<0, 2>([&](auto Pp) {
                constexpr int p = decltype(Pp)::value;
                cfloat* op = orow + Fft1<MC>::pi(lc + p);
                static_for<0, GS>([&](auto G) {
                    constexpr int g = decltype(G)::value;
                    prod_store<OH>(op + g * AS, cmake(ar[p][g] * scale, ai[p][g] * scale));
                });
            });
        }
    } else if constexpr (!SPARSE) {
        // dense: all F*F aliases of U outputs per thread are loaded as one straight-line batch
        constexpr int U = cx_max(1, cx_min(8, 32 / (F * F * (GS + 2))));
        constexpr bool HOIST = WST_OPT_HOIST_PROD && NT % MC == 0;       // the thread's output column is fixed
        for (int o0 = HOIST ? tid / MC : tid; o0 < (HOIST ? MC : MC * MC); o0 += (HOIST ? NT / MC : NT) * U) {
            float w[U][F * F][GS];
            cfloat u[U][F * F];
            int kc[U], lc[U];
            static_for<0, U>([&](auto Uc) {
                constexpr int ui = decltype(Uc)::value;
                if constexpr (HOIST) {
                    int k = o0 + ui * (NT / MC);
                    kc[ui] = k < MC ? k : MC - 1;           // tail items recompute the last row, never stored
                    lc[ui] = tid % MC;
                } else {
                    int o = o0 + ui * NT;
                    o = o < MC * MC ? o : MC * MC - 1;      // tail items recompute the last output, never stored
                    lc[ui] = o % MC; kc[ui] = o / MC;
                }
                static_for<0, F * F>([&](auto S) {
                    constexpr int sl = decltype(S)::value;
                    const int k = kc[ui] + (sl / F) * MC, l = lc[ui] + (sl % F) * MC;
                    load_filter_vec<MP, GS, FH>(filt + ((size_t)k * MP + l) * GS, w[ui][sl]);
                    if constexpr ((WST_OPT_SPARSEROW & 4) && F == 2) {
                        // fold by two: alias column b = 0 lies in the stored half, b = 1 in the mirrored half, for every output
                        if constexpr (sl % F == 0) u[ui][sl] = uh[Fft1<MP>::pi(k) * (MP / 2 + 1) + l];
                        else { const cfloat v = uh[Fft1<MP>::pi(k == 0 ? 0 : MP - k) * (MP / 2 + 1) + (MP - l)]; u[ui][sl] = cmake(v.x, -v.y); }
                    } else {
                        u[ui][sl] = herm_get<MP>(uh, k, l);
                    }
                });
            });
            static_for<0, U>([&](auto Uc) {
                constexpr int ui = decltype(Uc)::value;
                float ar[GS], ai[GS];
                static_for<0, GS>([&](auto G) { ar[decltype(G)::value] = 0.f; ai[decltype(G)::value] = 0.f; });
                static_for<0, F * F>([&](auto S) {
                    constexpr int sl = decltype(S)::value;
                    static_for<0, GS>([&](auto G) {
                        constexpr int g = decltype(G)::value;
                        ar[g] += u[ui][sl].x * w[ui][sl][g];
                        ai[g] += u[ui][sl].y * w[ui][sl][g];
                    });
                });
                if (HOIST ? o0 + ui * (NT / MC) < MC : o0 + ui * NT < MC * MC) {
                    cfloat* op = out + Fft1<MC>::pi(kc[ui]) * PC + Fft1<MC>::pi(lc[ui]);
                    static_for<0, GS>([&](auto G) {
                        constexpr int g = decltype(G)::value;
                        prod_store<OH>(op + g * AS, cmake(ar[g] * scale, ai[g] * scale));
                    });
                }
            });
        }
    } else {
        constexpr int NB = F >= 8 ? 4 : F;                // columns visited per row and batch
        constexpr bool HOIST = WST_OPT_HOIST_PROD && NT % MC == 0;       // fixed output column: its alias run leaves the loop
        for (int o = HOIST ? tid / MC : tid; o < (HOIST ? MC : MC * MC); o += (HOIST ? NT / MC : NT)) {
            const int lc = HOIST ? tid % MC : o % MC, kc = HOIST ? o : o / MC;
            float ar[GS], ai[GS];
            static_for<0, GS>([&](auto G) { ar[decltype(G)::value] = 0.f; ai[decltype(G)::value] = 0.f; });
            AliasRun ra = alias_run<MP, MC>(kc, rows), rb = alias_run<MP, MC>(lc, cols);
            if constexpr (NB == F) { rb.first = 0; rb.count = rb.count == 0 ? 0 : F; }
            for (int ia = 0; ia < ra.count; ++ia) {
                int a = ra.first + ia; a -= a >= F ? F : 0;
                const int k = kc + a * MC;
                [[maybe_unused]] const cfloat* rd = uh + Fft1<MP>::pi(k) * (MP / 2 + 1);
                [[maybe_unused]] const cfloat* rm = uh + Fft1<MP>::pi(k == 0 ? 0 : MP - k) * (MP / 2 + 1);
                for (int ib0 = 0; ib0 < rb.count; ib0 += NB) {
                    float w[NB][GS];
                    cfloat u[NB];
                    static_for<0, NB>([&](auto B) {
                        constexpr int bi = decltype(B)::value;
                        if constexpr ((WST_OPT_SPARSEROW & 1) && NB == F) {
                            // every alias is visited, in order: the side of column MP/2 is known at compile time
                            constexpr bool MIR = bi * MC >= MP / 2;
                            const int l = lc + bi * MC;
                            load_filter_vec<MP, GS, FH>(filt + ((size_t)k * MP + l) * GS, w[bi]);
                            if constexpr (MIR) { const cfloat v = rm[MP - l]; u[bi] = cmake(v.x, -v.y); }
                            else u[bi] = rd[l];
                        } else {
                            int b = rb.first + ib0 + bi; b -= b >= F ? F : 0;
                            const int l = lc + b * MC;
                            load_filter_vec<MP, GS, FH>(filt + ((size_t)k * MP + l) * GS, w[bi]);
                            if constexpr ((WST_OPT_SPARSEROW & 2) != 0) {
                                const bool dir = l <= MP / 2;
                                const cfloat v = dir ? rd[l] : rm[MP - l];
                                u[bi] = cmake(v.x, dir ? v.y : -v.y);
                            } else {
                                u[bi] = herm_get<MP>(uh, k, l);
                            }
                        }
                    });
                    static_for<0, NB>([&](auto B) {
                        constexpr int bi = decltype(B)::value;
                        static_for<0, GS>([&](auto G) {
                            constexpr int g = decltype(G)::value;
                            ar[g] += u[bi].x * w[bi][g];
                            ai[g] += u[bi].y * w[bi][g];
                        });
                    });
                }
            }
            cfloat* op = out + Fft1<MC>::pi(kc) * PC + Fft1<MC>::pi(lc);
            static_for<0, GS>([&](auto G) {
                constexpr int g = decltype(G)::value;
                prod_store<OH>(op + g * AS, cmake(ar[g] * scale, ai[g] * scale));
            });
        }
    }
}

// Separable low-pass + subsample + unpad of narr paired-row arrays z (stride ZS, pitch M+1):
//   S[i][i'] = sum_{x,y} Gr[x][i] * U[x][y] * Gc[y][i'],  i, i' < HOUT
// written to maps + coef(g)*HOUT*HOUT for arrays with coef(g) >= 0.
// Scratch: the dead second half of each array (rows >= M/2).
template <int M, int HOUT, int HP, int NT, int LV, class Exec, class CoefFn>
WST_D void lowpass_maps_dense(Exec& ex, cfloat* z, int ZS, int narr, const float* gr, const float* gc,
                              float* maps, CoefFn coef) {
    constexpr int P = M + 1, HALF = M / 2, PT = HOUT | 1;
    constexpr int GO1 = HOUT >= 8 ? 8 : HP, NG1 = (HOUT + GO1 - 1) / GO1;      // phase 1: outputs per thread (float4s of G)
    constexpr int NG2 = HP / 4;                                                // phase 2: one float4 of outputs per thread
    static_assert(PT * M <= 2 * HALF * P && GO1 % 4 == 0, "low-pass scratch must fit in the dead half of the array");
    static_assert(HOUT < 8 || HP % 8 == 0, "operator rows must hold whole groups of eight outputs");
    // phase LP1: T[y][i] = sum_x Gr[x][i] * U[x][y]; thread = (array, group of GO1 outputs i, column slot y), the
    // G rows are warp-uniform (broadcast) vector loads, the U column is read as (row x, row x + M/2) pairs.
    if constexpr (WST_OPT_LPPAIR && M % 2 == 0) {
        ex.template phase<PK_LP1 * 8 + LV>([&](int tid) {
            constexpr int HM2 = M / 2;
            const int total = narr * NG1 * HM2;
            for (int b = tid; b < total; b += NT) {
                const int y = b % HM2, r = b / HM2;
                const int q = r % NG1, g = r / NG1;
                const cfloat* zp = z + g * ZS + y;
                const float* g0 = gr + q * GO1;
                float acc[2][GO1];
                static_for<0, 2 * GO1>([&](auto I) { acc[decltype(I)::value / GO1][decltype(I)::value % GO1] = 0.f; });
#pragma unroll 2
                for (int x = 0; x < HALF; ++x) {
                    const cfloat v0 = zp[x * P], v1 = zp[x * P + HM2];
                    static_for<0, GO1 / 4>([&](auto Q) {
                        constexpr int k = decltype(Q)::value;
                        const float4 a = *reinterpret_cast<const float4*>(g0 + x * HP + 4 * k);
                        const float4 d = *reinterpret_cast<const float4*>(g0 + (x + HALF) * HP + 4 * k);
                        acc[0][4 * k + 0] += a.x * v0.x; acc[0][4 * k + 0] += d.x * v0.y;
                        acc[0][4 * k + 1] += a.y * v0.x; acc[0][4 * k + 1] += d.y * v0.y;
                        acc[0][4 * k + 2] += a.z * v0.x; acc[0][4 * k + 2] += d.z * v0.y;
                        acc[0][4 * k + 3] += a.w * v0.x; acc[0][4 * k + 3] += d.w * v0.y;
                        acc[1][4 * k + 0] += a.x * v1.x; acc[1][4 * k + 0] += d.x * v1.y;
                        acc[1][4 * k + 1] += a.y * v1.x; acc[1][4 * k + 1] += d.y * v1.y;
                        acc[1][4 * k + 2] += a.z * v1.x; acc[1][4 * k + 2] += d.z * v1.y;
                        acc[1][4 * k + 3] += a.w * v1.x; acc[1][4 * k + 3] += d.w * v1.y;
                    });
                }
                float* tp = reinterpret_cast<float*>(z + g * ZS + HALF * P) + y * PT + q * GO1;
                static_for<0, GO1>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    if (q * GO1 + i < HOUT) { tp[i] = acc[0][i]; tp[HM2 * PT + i] = acc[1][i]; }
                });
            }
        });
    } else
    ex.template phase<PK_LP1 * 8 + LV>([&](int tid) {
        const int total = narr * NG1 * M;
        for (int b = tid; b < total; b += NT) {
            const int y = b % M, r = b / M;
            const int q = r % NG1, g = r / NG1;
            const cfloat* zp = z + g * ZS + y;
            const float* g0 = gr + q * GO1;
            float acc[GO1];
            static_for<0, GO1>([&](auto I) { acc[decltype(I)::value] = 0.f; });
#pragma unroll 2
            for (int x = 0; x < HALF; ++x) {
                const cfloat v = zp[x * P];
                static_for<0, GO1 / 4>([&](auto Q) {
                    constexpr int k = decltype(Q)::value;
                    const float4 a = *reinterpret_cast<const float4*>(g0 + x * HP + 4 * k);
                    const float4 d = *reinterpret_cast<const float4*>(g0 + (x + HALF) * HP + 4 * k);
                    acc[4 * k + 0] += a.x * v.x; acc[4 * k + 0] += d.x * v.y;      // two FMAs (a sum of products first
                    acc[4 * k + 1] += a.y * v.x; acc[4 * k + 1] += d.y * v.y;      // would cost a multiply, an FMA and an add)
                    acc[4 * k + 2] += a.z * v.x; acc[4 * k + 2] += d.z * v.y;
                    acc[4 * k + 3] += a.w * v.x; acc[4 * k + 3] += d.w * v.y;
                });
            }
            float* tp = reinterpret_cast<float*>(z + g * ZS + HALF * P) + y * PT + q * GO1;
            static_for<0, GO1>([&](auto I) {
                constexpr int i = decltype(I)::value;
                if (q * GO1 + i < HOUT) tp[i] = acc[i];
            });
        }
    });
    // phase LP2: S[i][i'] = sum_y T[y][i] * Gc[y][i']; thread = (array, float4 of outputs i', row i)
    ex.template phase<PK_LP2 * 8 + LV>([&](int tid) {
        const int total = narr * NG2 * HOUT;
        for (int b = tid; b < total; b += NT) {
            const int ir = b % HOUT, r = b / HOUT;
            const int q = r % NG2, g = r / NG2;
            const int cidx = coef(g);
            if (cidx < 0) continue;
            const float* tp = reinterpret_cast<const float*>(z + g * ZS + HALF * P) + ir;
            const float* gq = gc + 4 * q;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
            for (int y = 0; y < M; ++y) {
                const float t = tp[y * PT];
                const float4 w = *reinterpret_cast<const float4*>(gq + y * HP);
                a0 += w.x * t; a1 += w.y * t; a2 += w.z * t; a3 += w.w * t;
            }
            float* mp = maps + (size_t)cidx * (HOUT * HOUT) + ir * HOUT + 4 * q;
            if (4 * q + 0 < HOUT) mp[0] = a0;
            if (4 * q + 1 < HOUT) mp[1] = a1;
            if (4 * q + 2 < HOUT) mp[2] = a2;
            if (4 * q + 3 < HOUT) mp[3] = a3;
        }
    });
}

// Banded form of lowpass_maps for levels where lp_banded(M, HOUT): the operator is circulant, G[x][i] = w[|(i+1)*S - x|],
// and w vanishes beyond R samples, so every output is a (2R+1)-tap strided convolution.  All positions are
// compile-time constants (the spatial side may be stored in Good-Thomas order, Fft1<M>::pos_s), the taps are
// constant-bank operands, and a thread produces GO consecutive outputs from one sliding window of loads.
//   phase LP1: T[y][ir]  = sum_d w[|d|] * U[(ir+1)*S + d][y]      thread = (array, output group, y slot)
//   phase LP2: S[ir][ic] = sum_d w[|d|] * T[(ic+1)*S + d][ir]     thread = (array, output group, ir)
// T (pitch HOUT|1 floats) lives in the dead second half of each array, like lowpass_maps' scratch.
template <int M, int HOUT, int NT, int LV, class Exec, class CoefFn>
WST_D void lowpass_maps_banded(Exec& ex, cfloat* z, int ZS, int narr, const float (&w)[kLpTaps], float* maps,
                               CoefFn coef) {
    constexpr int P = M + 1, HALF = M / 2, S = M / (HOUT + 2), R = lp_radius(S), PT = HOUT | 1;
    constexpr int GO = HOUT >= 8 ? 8 : HOUT, NG = (HOUT + GO - 1) / GO;        // outputs per thread, groups per line
    static_assert(R > 0 && R < kLpTaps && PT * M <= 2 * HALF * P, "banded low-pass geometry");
    ex.template phase<PK_LP1 * 8 + LV>([&](int tid) {
        const int total = narr * NG * M;
        for (int b = tid; b < total; b += NT) {
            const int ys = b % M, r = b / M;
            const int grp = r % NG, g = r / NG;
            const float* zp = reinterpret_cast<const float*>(z + g * ZS + ys);
            float* tp = reinterpret_cast<float*>(z + g * ZS + HALF * P) + ys * PT;
            static_for<0, NG>([&](auto Gc) {
                constexpr int G = decltype(Gc)::value;
                if (grp != G) return;
                constexpr int I0 = G * GO, I1 = (I0 + GO < HOUT ? I0 + GO : HOUT);      // outputs [I0, I1)
                float acc[GO];
                static_for<0, GO>([&](auto I) { acc[decltype(I)::value] = 0.f; });
                static_for<(I0 + 1) * S - R, I1 * S + R + 1>([&](auto Xc) {
                    constexpr int xx = decltype(Xc)::value;                         // window position (may wrap)
                    constexpr int slot = Fft1<M>::pos_s(((xx % M) + M) % M);
                    constexpr int row = slot < HALF ? slot : slot - HALF, comp = slot < HALF ? 0 : 1;
                    const float v = zp[row * P * 2 + comp];
                    static_for<0, GO>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        constexpr int d = xx - (I0 + i + 1) * S;
                        if constexpr (I0 + i < I1 && d >= -R && d <= R) acc[i] += w[d < 0 ? -d : d] * v;
                    });
                });
                static_for<0, GO>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    if constexpr (I0 + i < I1) tp[I0 + i] = acc[i];
                });
            });
        }
    });
    ex.template phase<PK_LP2 * 8 + LV>([&](int tid) {
        const int total = narr * NG * HOUT;
        for (int b = tid; b < total; b += NT) {
            const int ir = b % HOUT, r = b / HOUT;
            const int grp = r % NG, g = r / NG;
            const int cidx = coef(g);
            if (cidx < 0) continue;
            const float* tp = reinterpret_cast<const float*>(z + g * ZS + HALF * P) + ir;
            float* mp = maps + (size_t)cidx * (HOUT * HOUT) + ir * HOUT;
            static_for<0, NG>([&](auto Gc) {
                constexpr int G = decltype(Gc)::value;
                if (grp != G) return;
                constexpr int I0 = G * GO, I1 = (I0 + GO < HOUT ? I0 + GO : HOUT);
                float acc[GO];
                static_for<0, GO>([&](auto I) { acc[decltype(I)::value] = 0.f; });
                static_for<(I0 + 1) * S - R, I1 * S + R + 1>([&](auto Yc) {
                    constexpr int yy = decltype(Yc)::value;
                    constexpr int slot = Fft1<M>::pos_s(((yy % M) + M) % M);
                    const float v = tp[slot * PT];
                    static_for<0, GO>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        constexpr int d = yy - (I0 + i + 1) * S;
                        if constexpr (I0 + i < I1 && d >= -R && d <= R) acc[i] += w[d < 0 ? -d : d] * v;
                    });
                });
                static_for<0, GO>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    if constexpr (I0 + i < I1) mp[I0 + i] = acc[i];
                });
            });
        }
    });
}

template <int M, int HOUT, int HP, int NT, int LV, bool BAND, class Exec, class CoefFn>
WST_D void lowpass_maps(Exec& ex, cfloat* z, int ZS, int narr, const float* gr, const float* gc,
                        const float (&w)[kLpTaps], float* maps, CoefFn coef) {
    if constexpr (BAND) lowpass_maps_banded<M, HOUT, NT, LV>(ex, z, ZS, narr, w, maps, coef);
    else lowpass_maps_dense<M, HOUT, HP, NT, LV>(ex, z, ZS, narr, gr, gc, maps, coef);
}

// Global-workspace variant: both passes of the inverse column transforms on shared-memory tiles of TC columns (all M
// rows; as many arrays per tile batch as the stage holds), then the row pass and the final pass (modulus, pairing,
// first low-pass contraction) on tiles of TR row pairs.  An array is read and written once per dimension instead of
// once per pass; tile loads and stores move 8*TC-byte / whole-row segments.
template <int M, int NT, int STAGE, int LV, class Exec>
WST_D void ifft_cols_staged(Exec& ex, cfloat* base, int narr, int AS, const cfloat* tw, cfloat* stage) {
    constexpr int P = M + 1, TC = stage_tc(M), TCP = TC + 1, TILE = M * TCP, NCT = M / TC;
    constexpr int NAB = STAGE / TILE;
    constexpr bool TWO = Fft1<M>::R1 > 1;
    static_assert(NAB >= 1, "stage too small for a column tile");
    const int nbatch = (narr + NAB - 1) / NAB;
    auto geom = [&](int t, int& g0, int& na, int& c0) { g0 = (t / NCT) * NAB; na = narr - g0 < NAB ? narr - g0 : NAB; c0 = (t % NCT) * TC; };
    staged_tiles<PK_IFFT_COL_C, PK_STAGE_ST, LV>(ex, nbatch * NCT, stage, STAGE,
        [&](int tid, int t, cfloat* buf) {
            int g0, na, c0; geom(t, g0, na, c0);
            const int c = tid % TC;                                             // TC consecutive threads per tile row
            for (int a = 0; a < na; ++a) {
                const cfloat* src = base + (g0 + a) * AS + c0 + c;
                cfloat* dst = buf + a * TILE + c;
                for (int r = tid / TC; r < M; r += NT / TC) stage_copy(dst + r * TCP, src + r * P);
            }
        },
        [&](int tid, int t, cfloat* buf) {
            int g0, na, c0; geom(t, g0, na, c0);
            pass_contig<M, +1, true, TC, 1, TCP, NT>(tid, buf, na, TILE, tw);
        },
        [&](int t, cfloat* buf) {
            if constexpr (TWO) {
                int g0, na, c0; geom(t, g0, na, c0);
                ex.template phase<PK_IFFT_COL_S * 8 + LV>([&](int tid) { pass_strided<M, +1, false, TC, 1, TCP, NT>(tid, buf, na, TILE, tw); });
            }
        },
        [&](int tid, int t, cfloat* buf) {
            int g0, na, c0; geom(t, g0, na, c0);
            const int c = tid % TC;
            for (int a = 0; a < na; ++a) {
                cfloat* dst = base + (g0 + a) * AS + c0 + c;
                const cfloat* src = buf + a * TILE + c;
                for (int r = tid / TC; r < M; r += NT / TC) stage_store(dst + r * P, src[r * TCP]);
            }
        });
}

// NATCOL: the workspace holds the columns in natural frequency order (written by product_ifft_cols_staged); the tile
// loads put column l at its transform position pi(l).
template <int M, int NT, int STAGE, int TR, bool WRITE_Z, bool LPF, int HOUT, int HP, int LV, bool NATCOL = false, class Exec>
WST_D void ifft_rows_final_staged(Exec& ex, cfloat* base, int narr, int AS, const cfloat* tw, const float* g,
                                  cfloat* stage) {
    constexpr int P = M + 1, HALF = M / 2, TILE = 2 * TR * P, NX = HALF / TR;
    constexpr int NAB = STAGE / TILE;
    constexpr bool TWO = Fft1<M>::R1 > 1;
    static_assert(NAB >= 1 && HALF % TR == 0, "stage too small for a row-pair tile");
    const int nbatch = (narr + NAB - 1) / NAB;
    auto geom = [&](int t, int& g0, int& na, int& x0) { g0 = (t / NX) * NAB; na = narr - g0 < NAB ? narr - g0 : NAB; x0 = (t % NX) * TR; };
    // tile row q < TR holds array row x0 + q, tile row TR + q holds array row x0 + q + M/2
    auto final_pass = [&](int tid, int na, cfloat* buf) {
        pass_rows_final<M, NT, WRITE_Z, LPF, HOUT, HP, false, TR, TR * P>(tid, buf, na, TILE, g);
    };
    staged_tiles<TWO ? PK_IFFT_ROW_C : PK_IFFT_FINAL, PK_STAGE_ST, LV>(ex, nbatch * NX, stage, STAGE,
        [&](int tid, int t, cfloat* buf) {
            int g0, na, x0; geom(t, g0, na, x0);
            for (int rq = tid >> 5; rq < na * 2 * TR; rq += NT / 32) {          // one warp per tile row, lanes along it
                const int q = rq % (2 * TR), a = rq / (2 * TR);
                const int row = x0 + (q < TR ? q : q - TR + HALF);
                const cfloat* src = base + (g0 + a) * AS + row * P;
                cfloat* dst = buf + a * TILE + q * P;
                if constexpr (NATCOL) { for (int c = tid & 31; c < M; c += 32) stage_copy(dst + Fft1<M>::pi(c), src + c); }
                else
                for (int c = tid & 31; c < M; c += 32) stage_copy(dst + c, src + c);
            }
        },
        [&](int tid, int t, cfloat* buf) {
            int g0, na, x0; geom(t, g0, na, x0);
            if constexpr (TWO) pass_contig<M, +1, true, 2 * TR, P, 1, NT>(tid, buf, na, TILE, tw);
            else final_pass(tid, na, buf);
        },
        [&](int t, cfloat* buf) {
            if constexpr (TWO) {
                int g0, na, x0; geom(t, g0, na, x0);
                ex.template phase<PK_IFFT_FINAL * 8 + LV>([&](int tid) { final_pass(tid, na, buf); });
            }
        },
        [&](int tid, int t, cfloat* buf) {
            int g0, na, x0; geom(t, g0, na, x0);
            for (int rq = tid >> 5; rq < na * 2 * TR; rq += NT / 32) {
                const int q = rq % (2 * TR), a = rq / (2 * TR);
                const int row = x0 + (q < TR ? q : q - TR + HALF);
                cfloat* dst = base + (g0 + a) * AS + row * P;
                const cfloat* src = buf + a * TILE + q * P;
                for (int c = tid & 31; c < M; c += 32) stage_store(dst + c, src[c]);
            }
        });
}

// Global-workspace levels, fold factor F <= 2: V_g = fold(U . psi_g) for the GS filters of a group computed straight into
// a shared-memory tile of TCF natural-frequency columns (all MC rows, all GS arrays), transformed along the columns there
// and stored to the workspace once — the array skips one write and one read (the product_fold output and the column-tile
// load of ifft_cols_staged).  Rows of the tile are in transform order pi(k); the columns go to the workspace in natural
// order l0 .. l0 + TCF - 1 (whole 32-byte sectors), which the row-tile loads of ifft_rows_final_staged<NATCOL> undo.
// Same accumulation order as product_fold's dense branches (aliases a-major).
template <int MP, int MC, int GS, int NT, int AVAIL, int LV, int TAG_PROD, bool FH, class Exec>
WST_D void product_ifft_cols_staged(Exec& ex, const cfloat* uh, const float* filt, cfloat* base, const cfloat* tw,
                                    cfloat* stage) {
    constexpr int F = MP / MC, P = MC + 1, AS = MC * (MC + 1);
    constexpr int TCF = prodtile_cols(MC, GS, AVAIL), TCP = TCF + 1, TILE = MC * TCP, NCT = MC / TCF;
    constexpr bool TWO = Fft1<MC>::R1 > 1;
    constexpr float scale = 1.0f / ((float)F * (float)F * (float)MC * (float)MC);
    static_assert(F <= 2 && TCF >= 4, "fused product tile: fold factor <= 2 and at least four columns");
    for (int t = 0; t < NCT; ++t) {
        const int l0 = t * TCF;
        ex.template phase<TAG_PROD * 8 + LV>([&](int tid) {
            for (int o = tid; o < MC * TCF; o += NT) {
                const int c = o % TCF, kc = o / TCF, lc = l0 + c;
                float w[F * F][GS];
                cfloat u[F * F];
                static_for<0, F * F>([&](auto S) {
                    constexpr int sl = decltype(S)::value;
                    const int k = kc + (sl / F) * MC, l = lc + (sl % F) * MC;
                    load_filter_vec<MP, GS, FH>(filt + ((size_t)k * MP + l) * GS, w[sl]);
                    u[sl] = herm_get<MP>(uh, k, l);
                });
                float ar[GS], ai[GS];
                static_for<0, GS>([&](auto G) { ar[decltype(G)::value] = 0.f; ai[decltype(G)::value] = 0.f; });
                static_for<0, F * F>([&](auto S) {
                    constexpr int sl = decltype(S)::value;
                    static_for<0, GS>([&](auto G) {
                        constexpr int g = decltype(G)::value;
                        ar[g] += u[sl].x * w[sl][g];
                        ai[g] += u[sl].y * w[sl][g];
                    });
                });
                cfloat* op = stage + Fft1<MC>::pi(kc) * TCP + c;
                static_for<0, GS>([&](auto G) {
                    constexpr int g = decltype(G)::value;
                    op[g * TILE] = cmake(ar[g] * scale, ai[g] * scale);
                });
            }
        });
        ex.template phase<PK_IFFT_COL_C * 8 + LV>([&](int tid) { pass_contig<MC, +1, true, TCF, 1, TCP, NT>(tid, stage, GS, TILE, tw); });
        if constexpr (TWO)
            ex.template phase<PK_IFFT_COL_S * 8 + LV>([&](int tid) { pass_strided<MC, +1, false, TCF, 1, TCP, NT>(tid, stage, GS, TILE, tw); });
        ex.template phase<PK_STAGE_ST * 8 + LV>([&](int tid) {
            for (int o = tid; o < GS * MC * TCF; o += NT) {
                const int c = o % TCF, r = (o / TCF) % MC, g = o / (TCF * MC);
                stage_store(base + g * AS + r * P + l0 + c, stage[g * TILE + r * TCP + c]);
            }
        });
    }
}

// Inverse 2-D FFT of narr digit-swapped M x M spectra (pitch M+1) + modulus (+ row-paired z for a following
// rfft2_from_pairs) + low-pass map of every array.
// COLS_DONE: the column transforms have already run (product_ifft_cols_staged) and the columns are in natural order.
template <int M, int NT, int LV, bool WRITE_Z, int HOUT, int HP, int LPSLOTS, bool GLOB = false, int STAGE = 0, int TR = 1,
          bool COLS_DONE = false, class Exec, class CoefFn>
WST_D void ifft2_modulus_lowpass(Exec& ex, cfloat* base, int narr, const cfloat* tw, const float* g,
                                 const float (&w)[kLpTaps], float* lpbuf, float* maps, cfloat* stage, CoefFn coef) {
    constexpr int P = M + 1, AS = M * (M + 1);
    constexpr bool FUSED = (Fft1<M>::R1 == 1) ? !WRITE_Z
                                              : (WRITE_Z ? (HOUT <= Fft1<M>::R1) : (HOUT / 2 <= Fft1<M>::R1));
    if constexpr (GLOB && STAGE > 0) {
        if constexpr (!COLS_DONE) ifft_cols_staged<M, NT, STAGE, LV>(ex, base, narr, AS, tw, stage);
        ifft_rows_final_staged<M, NT, STAGE, TR, FUSED ? WRITE_Z : true, FUSED, HOUT, HP, LV, COLS_DONE>(ex, base, narr, AS, tw, g, stage);
        if constexpr (FUSED) lowpass_reduce<M, NT, WRITE_Z, HOUT, HP, LPSLOTS, LV>(ex, base, narr, AS, g, lpbuf, maps, coef);
        else lowpass_maps<M, HOUT, HP, NT, LV, false>(ex, base, AS, narr, g, g, w, maps, coef);
        return;
    }
    // (COLS_DONE only comes with GLOB && STAGE > 0: the branch above)
    fft_lines_inv<M, M, 1, P, NT, PK_IFFT_COL_C * 8 + LV, PK_IFFT_COL_S * 8 + LV>(ex, base, narr, AS, tw);   // columns
    if constexpr (Fft1<M>::R1 > 1) {
        ex.template phase<PK_IFFT_ROW_C * 8 + LV>([&](int tid) { pass_contig<M, +1, true, M, P, 1, NT, GLOB>(tid, base, narr, AS, tw); });
    }
    if constexpr (FUSED) {
        ex.template phase<PK_IFFT_FINAL * 8 + LV>([&](int tid) { pass_rows_final<M, NT, WRITE_Z, true, HOUT, HP, GLOB>(tid, base, narr, AS, g); });
        lowpass_reduce<M, NT, WRITE_Z, HOUT, HP, LPSLOTS, LV>(ex, base, narr, AS, g, lpbuf, maps, coef);
    } else {
        ex.template phase<PK_IFFT_FINAL * 8 + LV>([&](int tid) { pass_rows_final<M, NT, true, false, HOUT, HP, GLOB>(tid, base, narr, AS, g); });
        lowpass_maps<M, HOUT, HP, NT, LV, lp_banded(M, HOUT, LV) && !GLOB>(ex, base, AS, narr, g, g, w, maps, coef);
    }
}

// ------------------------------------------------------------------ the per-signal program
// SPLIT: the signal is shared by several CTAs (small batches, see `part` below); a separate instantiation so that the
// throughput kernel carries none of its state in registers — the 128 x 128 J=4 cascade sits exactly at its 96-register
// budget and three more live values cost it 27 % in spills.
template <class C, class Exec, bool SPLIT = false>
struct Cascade {
    static constexpr int N = C::N, J = C::J, NT = C::NT, HOUT = C::HOUT, HP = C::HP;

    Exec& ex;
    const PlanTables& pt;
    cfloat* sm;          // shared data region (C::smem_cfloats())
    cfloat* twsm;        // shared twiddles   (C::tw_total)
    float* gsm;          // shared low-pass operators (C::g_total)
    float* lpbuf;        // shared scratch of the fused low-pass reduction (C::lpbuf_floats())
    cfloat* stage;       // shared tile of the staged passes (C::stage_cfloats(); global-workspace variant only)
    cfloat* u0h;         // per-CTA global scratch: N * (N/2+1)
    float* maps;         // this signal's output maps [K][HOUT][HOUT]
    // input prefetch (TMA): raw H x W pixels of the next signal at sm + PF_OFF, completion on *mbar
    // (the prefetch state lives in shared memory, not in registers that would stay live through the whole cascade:
    // word 0-1 the mbarrier, word 2 its phase parity, word 3 "a prefetch is pending")
    unsigned long long* mbar = nullptr;
    // Small batches (the reference calls the extractor one image at a time, train_and_save_model.py:486-488): the
    // first-order groups of a signal — a group of same-scale parents with all their children — are independent once
    // U0^ exists, so `nparts` CTAs share one signal: every CTA runs the input stage, then takes the groups whose
    // running index is congruent to `part`.  The last CTA to finish pools the signal (wst_cfg_inst.cu).
    // A unit of shared work is (first-order group, one group of its children): every CTA that owns a unit of a group
    // rebuilds that group's parents (product, inverse FFT, forward FFT), the owner of the group's first unit also
    // writes the parents' own maps, and each CTA transforms only its child groups.  Groups without children are one unit.
    int part = 0, nparts = 1, unit = 0;
    // child groups of one first-order group at level j (all its parents, all deeper scales)
    static WST_CX int child_tasks(int j, int L) {
        if (!C::has_children(j)) return 1;
        int n = 0;
        for (int j2 = j + 1; j2 < J; ++j2) n += (L + C::G2(j, j2) - 1) / C::G2(j, j2);
        return n * C::GP(j);
    }
    // number of units of a plan with L orientations and max_order (host side: how far a signal can be split)
    static WST_CX int num_units(int L, int max_order = 2) {
        int n = 0;
        for (int j = 0; j < J; ++j) n += ((L + C::GP(j) - 1) / C::GP(j)) * (max_order >= 2 ? child_tasks(j, L) : 1);
        return n;
    }
    // does unit u belong to this CTA?
    WST_D bool mine(int u) const { return !SPLIT || u % nparts == part; }
    // does any unit of [first, first + n) belong to this CTA?
    WST_D bool any_mine(int first, int n) const {
        if constexpr (!SPLIT) return true;
        else {
            if (n >= nparts) return true;
            const int r = first % nparts, d = part >= r ? part - r : part + nparts - r;
            return d < n;
        }
    }
    WST_D volatile unsigned* pf_state() const { return reinterpret_cast<volatile unsigned*>(mbar) + 2; }
    // first cfloat of the raw-pixel area: right after the paired rows z0 of the input stage (16-byte aligned)
    static constexpr int PF_OFF = ((N / 2) * (N + 1) + 1) & ~1;
    static constexpr bool PF_COMPILED = WST_OPT_TMA && !C::WS_GLOBAL && C::CL == 1;
    // the area is free while the last level runs if that level's arrays end below it
    static constexpr bool PF_EARLY = C::level_total(J - 1, C::GP(J - 1)) <= PF_OFF;

    // Can this signal's pixels be fetched by bulk copies?  float32 rows, 16-byte aligned, sizes in multiples of 16 bytes,
    // and the raw area inside the data region.
    WST_D bool pf_usable(const SignalSrc& x) const {
        if (!PF_COMPILED || !x.f32 || x.stride != 1) return false;
        const int H = pt.H, W = pt.W;
        if (PF_OFF + (H * W + 1) / 2 > C::smem_cfloats()) return false;
        if ((reinterpret_cast<size_t>(x.f32) & 15) != 0) return false;
        return x.pitch == W ? (H * W) % 4 == 0 : (W % 4 == 0 && x.pitch % 4 == 0);
    }

    // Issue the bulk copies of signal x into the raw area (one phase; warp 0 issues, everybody passes the barrier).
    WST_D void prefetch_input(const SignalSrc& x) {
#ifdef __CUDA_ARCH__
        if constexpr (PF_COMPILED) {
            if (!pf_usable(x)) return;
            ex.template phase<PK_INPUT * 8 + 1>([&](int tid) {
                if (tid == 0) pf_state()[1] = 1u;
                if (tid >= 32) return;
                const int H = pt.H, W = pt.W;
                float* raw = reinterpret_cast<float*>(sm + PF_OFF);
                fence_proxy_async();                  // earlier generic-proxy writes to the area are ordered before the copies
                if (tid == 0) mbar_expect_tx(mbar, (unsigned)(H * W * 4));
                __syncwarp();
                if (x.pitch == W) {                   // one contiguous patch: a few large copies
                    const unsigned total = (unsigned)(H * W * 4), CH = 16384;
                    for (unsigned o = tid * CH; o < total; o += 32 * CH)
                        tma_bulk_g2s(reinterpret_cast<char*>(raw) + o, reinterpret_cast<const char*>(x.f32) + o,
                                     total - o < CH ? total - o : CH, mbar);
                } else {                              // a tile of a larger raster: one copy per row
                    for (int r = tid; r < H; r += 32)
                        tma_bulk_g2s(raw + r * W, x.f32 + (size_t)r * x.pitch, (unsigned)(W * 4), mbar);
                }
            });
        }
#else
        (void)x;
#endif
    }

    // data region of level Jl's arrays: the workspace / the shared-memory region, or — hybrid form of the
    // global-workspace variant — the shared-memory region that aliases the stage tiles
    template <int Jl> WST_D cfloat* base() const { return (C::WS_GLOBAL && C::in_smem(Jl)) ? stage : sm; }
    template <int Jl> static WST_CX bool glob() { return C::WS_GLOBAL && !C::in_smem(Jl); }
    WST_D const cfloat* tw(int j) const { return twsm + C::tw_offset(j); }
    WST_D const float* g(int j) const { return C::g_total == 0 ? pt.gr[j] : gsm + C::g_offset(j); }

    // index of the first order-2 coefficient of parent (j1, t1)
    WST_D int order2_base(int j1, int t1) const {
        int L = pt.L, idx = 1 + J * L;
        for (int j = 0; j < j1; ++j) idx += L * L * (J - 1 - j);
        return idx + t1 * L * (J - 1 - j1);
    }

    // once per CTA
    WST_D void load_twiddles() {
        ex.template phase<PK_TWIDDLE * 8>([&](int tid) {
            static_for<0, J>([&](auto Jj) {
                constexpr int j = decltype(Jj)::value;
                constexpr int m = C::msize(j);
                const int lt = tid % C::NTL;      // every CTA of a cluster fills its own copy of the tables
                for (int i = lt; i < m; i += C::NTL) twsm[C::tw_offset(j) + i] = pt.tw[j][i];
                if (C::g_total > 0)
                    for (int i = lt; i < m * HP; i += C::NTL) gsm[C::g_offset(j) + i] = pt.gr[j][i];
            });
        });
    }

    // reflect-pad the H x W input into paired rows z0, S0, and U0^ -> global scratch
    WST_D void input_stage(const SignalSrc& x) {
        constexpr int P = N + 1, HALF = N / 2, PH = N / 2 + 1;
        const int H = pt.H, W = pt.W, pt_top = pt.pad_top, pt_left = pt.pad_left;
#ifdef __CUDA_ARCH__
        if (PF_COMPILED && pf_state()[1]) {
            // the pixels are (being) delivered to shared memory by the bulk copies issued during the previous signal
            const float* raw = reinterpret_cast<const float*>(sm + PF_OFF);
            const unsigned parity = pf_state()[0];
            ex.template phase<PK_INPUT * 8>([&](int tid) {
                mbar_wait(mbar, parity);
                for (int o = tid; o < HALF * N; o += NT) {
                    const int cs = o % N, rs = o / N;
                    const int c = Fft1<N>::inv_pos_s(cs), r = Fft1<N>::inv_pos_s(rs);
                    const int rp = r + HALF < N ? r + HALF : r + HALF - N;
                    int sc = c - pt_left; sc = sc < 0 ? -sc : (sc >= W ? 2 * (W - 1) - sc : sc);
                    int r0 = r - pt_top; r0 = r0 < 0 ? -r0 : (r0 >= H ? 2 * (H - 1) - r0 : r0);
                    int r1 = rp - pt_top; r1 = r1 < 0 ? -r1 : (r1 >= H ? 2 * (H - 1) - r1 : r1);
                    sm[rs * P + cs] = cmake(raw[r0 * W + sc], raw[r1 * W + sc]);
                }
            });
            // (every thread has read the state before the barrier that closed the phase; the next write to it is in a
            // later phase)
            if (ex.base + (int)threadIdx.x == 0) { pf_state()[0] = parity ^ 1u; pf_state()[1] = 0u; }
        } else
#endif
        ex.template phase<PK_INPUT * 8>([&](int tid) {
            for (int o = tid; o < HALF * N; o += NT) {
                const int cs = o % N, rs = o / N;                      // storage column / row (row < N/2)
                const int c = Fft1<N>::inv_pos_s(cs), r = Fft1<N>::inv_pos_s(rs);   // padded-image coordinates
                const int rp = r + HALF < N ? r + HALF : r + HALF - N;              // the row stored N/2 slots below
                int sc = c - pt_left; sc = sc < 0 ? -sc : (sc >= W ? 2 * (W - 1) - sc : sc);
                int r0 = r - pt_top; r0 = r0 < 0 ? -r0 : (r0 >= H ? 2 * (H - 1) - r0 : r0);
                int r1 = rp - pt_top; r1 = r1 < 0 ? -r1 : (r1 >= H ? 2 * (H - 1) - r1 : r1);
                sm[rs * P + cs] = cmake(x.at(r0, sc), x.at(r1, sc));
            }
        });
        lowpass_maps<N, HOUT, HP, NT, 0, lp_banded(N, HOUT, 0) && !C::WS_GLOBAL>(ex, sm, 0, 1, g(0), g(0), pt.lpw[0], maps,
                                                                              [&](int) { return (!SPLIT || part == 0) ? 0 : -1; });
        cfloat* uh = sm + C::OFFB(0);
        rfft2_from_pairs<N, NT, 0, C::WS_GLOBAL, C::stage_cfloats()>(ex, sm, 0, uh, 1, tw(0), stage);
        ex.template phase<PK_U0_STORE * 8>([&](int tid) {
            for (int o = tid; o < N * PH; o += NT) u0h[o] = uh[o];
        });
    }

    // ubase: index of the first unit of this parent's child groups at scale J2 (SPLIT: only the groups this CTA owns run)
    template <int J1, int J2>
    WST_D void children(const cfloat* uh_parent, int t1, int ubase) {
        constexpr int MP = C::msize(J1), MC = C::msize(J2), G = C::G2(J1, J2);
        const int L = pt.L;
        const int ngroups = (L + G - 1) / G;
        const int cbase = order2_base(J1, t1) + (J2 - J1 - 1) * L;
        cfloat* arr = base<J2>();
        for (int grp = 0; grp < ngroups; ++grp) {
            if (!mine(ubase + grp)) continue;
            // fused product + column transforms on a shared-memory tile (workspace-level children, fold factor <= 2)
            constexpr bool PT = WST_OPT_PRODTILE && C::CL == 1 && glob<J2>() && C::stage_cfloats() > 0 && MP / MC <= 2 &&
                                prodtile_cols(MC, G, C::stage_avail()) > 0;
            if constexpr (PT) {
                product_ifft_cols_staged<MP, MC, G, NT, C::stage_avail(), J2, PK_PROD2, (WST_OPT_L2HINT & 1) != 0>(
                    ex, uh_parent, pt.psi2[J2][J1] + (size_t)grp * MP * MP * G, arr, tw(J2), stage);
            } else
            ex.template phase<PK_PROD2 * 8 + J2>([&](int tid) {
                product_fold<MP, MC, G, NT, C::WS_GLOBAL && (WST_OPT_L2HINT & 1), glob<J2>() && (WST_OPT_L2HINT & 8)>(tid, uh_parent, pt.psi2[J2][J1] + (size_t)grp * MP * MP * G,
                                            pt.bb2[pair_index(J2, J1)][grp][0], pt.bb2[pair_index(J2, J1)][grp][1], arr);
            });
            ifft2_modulus_lowpass<MC, NT, J2, false, HOUT, HP, C::LP_SLOTS, glob<J2>(), C::stage_cfloats(), C::stage_rows(MC), PT>(
                ex, arr, G, tw(J2), g(J2), pt.lpw[J2], lpbuf, maps, stage,
                [&](int a) { int t2 = grp * G + a; return t2 < L ? cbase + t2 : -1; });
        }
    }

    template <int J1>
    WST_D void level() {
        constexpr int M = C::msize(J1), GPn = C::GP(J1);
        const int L = pt.L;
        const int ngroups = (L + GPn - 1) / GPn;
        const bool second = C::has_children(J1) && pt.max_order >= 2;
        const int ntasks = second ? child_tasks(J1, L) : 1;          // units of one group
        cfloat* arr = base<J1>();
        for (int grp = 0; grp < ngroups; ++grp) {
            int first = 0;
            if constexpr (SPLIT) { first = unit; unit += ntasks; if (!any_mine(first, ntasks)) continue; }
            const bool own_maps = mine(first);                        // the owner of the group's first unit writes S1
            constexpr bool PT = WST_OPT_PRODTILE && C::CL == 1 && glob<J1>() && C::stage_cfloats() > 0 && N / M <= 2 &&
                                prodtile_cols(M, GPn, C::stage_avail()) > 0;
            if constexpr (PT) {
                product_ifft_cols_staged<N, M, GPn, NT, C::stage_avail(), J1, PK_PROD1, (WST_OPT_L2HINT & 1) != 0>(
                    ex, u0h, pt.psi1[J1] + (size_t)grp * N * N * GPn, arr, tw(J1), stage);
            } else
            ex.template phase<PK_PROD1 * 8 + J1>([&](int tid) {
                product_fold<N, M, GPn, NT, C::WS_GLOBAL && (WST_OPT_L2HINT & 1), glob<J1>() && (WST_OPT_L2HINT & 8)>(tid, u0h, pt.psi1[J1] + (size_t)grp * N * N * GPn,
                                            pt.bb1[J1][grp][0], pt.bb1[J1][grp][1], arr);
            });
            ifft2_modulus_lowpass<M, NT, J1, C::has_children(J1), HOUT, HP, C::LP_SLOTS, glob<J1>(), C::stage_cfloats(), C::stage_rows(M), PT>(
                ex, arr, GPn, tw(J1), g(J1), pt.lpw[J1], lpbuf, maps, stage,
                [&](int a) { int t1 = grp * GPn + a; return (t1 < L && own_maps) ? 1 + J1 * L + t1 : -1; });
            if constexpr (C::has_children(J1)) {
                if (pt.max_order >= 2) {
                    cfloat* uh = arr + C::OFFB(J1);
                    rfft2_from_pairs<M, NT, J1, glob<J1>(), C::stage_cfloats()>(ex, arr, C::vsz(M), uh, GPn, tw(J1), stage);
                    int ubase = first;
                    for (int g = 0; g < GPn; ++g) {
                        int t1 = grp * GPn + g;
                        const cfloat* uhp = uh + g * C::uhsz(M);
                        static_for<J1 + 1, J>([&](auto J2c) {
                            constexpr int j2 = decltype(J2c)::value;
                            if (t1 < L) this->template children<J1, j2>(uhp, t1, ubase);
                            ubase += (L + C::G2(J1, j2) - 1) / C::G2(J1, j2);
                        });
                    }
                }
            }
        }
    }

    // Per-coefficient mean and population std over the h x w map (np.mean / np.std of
    // train_and_save_model.py:371-372), two-pass, from the maps this CTA has just written:
    //   feats[0][k] = mean, feats[1][k] = std.
    // CG: the maps were written by other CTAs (split signals): read them through L2 (ld.global.cg), never from L1
    template <bool CG = false>
    WST_D void pool(float* feats) {
        constexpr int NPIX = HOUT * HOUT;
        auto ld4 = [](const float4* q) -> float4 {
#ifdef __CUDA_ARCH__
            if constexpr (CG) return __ldcg(q);
#endif
            return *q;
        };
        const int K = pt.K;
        int parts = 1;
        while (parts < 8 && K * parts * 2 <= NT && NPIX / (parts * 2) >= 16 && NPIX % (parts * 8) == 0) parts *= 2;
        const int slice = NPIX / parts;                      // a multiple of 4 (the slices are read as float4)
        float* scr = reinterpret_cast<float*>(sm);           // the data region is free at the end of a signal
        ex.template phase<PK_POOL * 8>([&](int tid) {
            for (int o = tid; o < K * parts; o += NT) {
                const float4* p = reinterpret_cast<const float4*>(maps + (size_t)(o / parts) * NPIX + (o % parts) * slice);
                float sacc = 0.f;
                for (int i = 0; i < slice / 4; ++i) { float4 v = ld4(p + i); sacc += (v.x + v.y) + (v.z + v.w); }
                scr[o] = sacc;
            }
        });
        ex.template phase<PK_POOL * 8 + 1>([&](int tid) {
            for (int o = tid; o < K * parts; o += NT) {
                const int k = o / parts;
                float tot = 0.f;
                for (int q = 0; q < parts; ++q) tot += scr[k * parts + q];
                const float mean = tot * (1.0f / NPIX);
                const float4* p = reinterpret_cast<const float4*>(maps + (size_t)k * NPIX + (o % parts) * slice);
                float vacc = 0.f;
                for (int i = 0; i < slice / 4; ++i) {
                    float4 v = ld4(p + i);
                    float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
                    vacc += (a * a + b * b) + (c * c + d * d);
                }
                scr[K * parts + o] = vacc;
            }
        });
        ex.template phase<PK_POOL * 8 + 2>([&](int tid) {
            for (int k = tid; k < K; k += NT) {
                float tot = 0.f, var = 0.f;
                for (int q = 0; q < parts; ++q) { tot += scr[k * parts + q]; var += scr[K * parts + k * parts + q]; }
                feats[k] = tot * (1.0f / NPIX);
                feats[K + k] = sqrtf(var * (1.0f / NPIX));
            }
        });
    }

    // next(SignalSrc&): fills in the signal this CTA will process after x and returns true, or returns false — its
    // pixels are prefetched while the last level of x runs (or right before pooling when the last level's arrays reach
    // into the raw-pixel area).  feats(): where the pooled features of x go (nullptr: no pooling here).  Both are
    // callables evaluated where needed, so that nothing of them stays live in registers across the cascade.
    struct NoNext { WST_D bool operator()(SignalSrc&) const { return false; } };
    template <class FeatsFn, class NextFn = NoNext>
    WST_D void run(const SignalSrc& x, FeatsFn feats, NextFn next = NextFn()) {
        if constexpr (SPLIT) unit = 0;
        input_stage(x);
        static_for<0, J>([&](auto Jc) {
            constexpr int j = decltype(Jc)::value;
            if constexpr (j == J - 1 && PF_EARLY && PF_COMPILED) { SignalSrc nx; if (next(nx)) prefetch_input(nx); }
            this->template level<j>();
        });
        if constexpr (!PF_EARLY && PF_COMPILED) {
            // pooling keeps its partial sums at the start of the data region: 2 * K * parts floats
            SignalSrc nx;
            if (8 * pt.K <= PF_OFF && next(nx)) prefetch_input(nx);
        }
        float* f = feats();
        if (f) pool(f);
    }
};

}  // namespace wst
