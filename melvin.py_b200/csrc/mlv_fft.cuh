// melvin-b200: register-resident fp64 Stockham FFT building block.
//
// One transform line of N = 2^LOG2N complex128 points is owned by T = N/16
// threads; every thread keeps 16 points in registers for the whole transform:
// thread tau holds x[tau + T*j], j = 0..15 on entry and X[tau + T*j] on exit
// (natural order both ways, so global loads/stores are unit-stride across
// lanes).  The transform is  [radix R0] -> 16 -> 16 ...  with
// R0 = 2^(LOG2N mod 4) (or 16): between passes the 16 points are exchanged
// through shared memory (the only shared-memory traffic of the transform),
// every butterfly runs in registers.
//
// Derivation (self-sorting Cooley-Tukey, decimation in time over the top
// digit):  N = R*N3, n = N3*d + n'',  X[k1 + R*k'] needs
//   Y[k1][n''] = W_N^{n'' k1} * sum_d x[N3 d + n''] W_R^{d k1}
// followed by R independent length-N3 transforms over n''.  With "butterfly
// beta = n'' + N3*K" (K = already produced low output digits) assigned to
// thread tau = beta (beta = tau + T*u when R < 16), the results of every pass
// land at linear slot  beta + (N/R)*d' = tau + T*j : the scatter side of every
// exchange is the identity layout, only the gather side is strided.  The last
// gather (stride 16) is made bank-conflict free by padding one slot per 16.
#pragma once

#include "mlv_common.cuh"

namespace mlv {

#define MLV_MAX_PASS 4

// Per-pass twiddle tables (device pointers).  Table p holds log2(R_p) planes of N3_p entries
//   tw[b*N3 + n] = exp(-2 pi i n 2^b / (R_p N3_p));   unused when N3_p == 1.
// (Plane-major: the lanes of a warp read consecutive n of one plane, i.e. consecutive 16-byte
// entries; with the entries of one n side by side every lane touched its own 64-byte stretch and
// a table load cost 16 LSU wavefronts instead of 4 -- more than the data loads of the z stage.)
struct FftTw {
    const cplx* p[MLV_MAX_PASS];
    // three-pass lengths only (N = R0*256): tables of the mirrored pass order 16 -> 16 -> R0
    // (fft_grp2nat): g[0] = (R 16, N3 N/16), g[1] = (R 16, N3 R0)
    const cplx* g[2];
};

template <int LOG2N>
struct FftCfg {
    static_assert(LOG2N >= 4 && LOG2N <= 16, "line length 16 .. 65536");
    static constexpr int N = 1 << LOG2N;
    static constexpr int T = N / 16;                     // threads per line
    static constexpr int Q = LOG2N & 3;
    static constexpr int R0 = Q ? (1 << Q) : 16;         // first-pass radix
    static constexpr int NPASS = (LOG2N >> 2) + (Q ? 1 : 0);
    static constexpr int XSLOTS = (N + N / 16 + 1) & ~1; // exchange slots per line (even: keeps
                                                         // per-line buffers 16-byte aligned)
    // sub-problem length left after pass p
    __host__ __device__ static constexpr int n3(int p) {
        int n = N / R0;
        for (int i = 0; i < p; ++i) n /= 16;
        return n;
    }
    __host__ __device__ static constexpr int radix(int p) { return p == 0 ? R0 : 16; }
};

// ------------------------------------------------------------- butterflies
// Forward sign: exp(-2 pi i jk/R).  INV conjugates every twiddle.
template <bool INV>
MLV_HD cplx rot_w4(cplx a) { return INV ? cmuli(a) : cmulni(a); }        // * W4^1

template <bool INV>
MLV_HD cplx mul_w8_1(cplx a) {                                          // * W8^1
    const double c = 0.70710678118654752440;
    return INV ? mk(c * (a.x - a.y), c * (a.x + a.y))
               : mk(c * (a.x + a.y), c * (a.y - a.x));
}
template <bool INV>
MLV_HD cplx mul_w8_3(cplx a) {                                          // * W8^3
    const double c = 0.70710678118654752440;
    return INV ? mk(-c * (a.x + a.y), c * (a.x - a.y))
               : mk(c * (a.y - a.x), -c * (a.x + a.y));
}
// * (wr + i wi) forward, * conj for inverse
template <bool INV>
MLV_HD cplx mul_w(cplx a, double wr, double wi) {
    return INV ? cmulc(a, mk(wr, wi)) : cmul(a, mk(wr, wi));
}

MLV_HD void bfly2(cplx& a, cplx& b) {
    cplx t = a;
    a = cadd(t, b);
    b = csub(t, b);
}

// natural order in, natural order out
template <bool INV>
MLV_HD void bfly4(cplx& a0, cplx& a1, cplx& a2, cplx& a3) {
    cplx t0 = cadd(a0, a2), t1 = csub(a0, a2);
    cplx t2 = cadd(a1, a3), t3 = rot_w4<INV>(csub(a1, a3));
    a0 = cadd(t0, t2);
    a1 = cadd(t1, t3);
    a2 = csub(t0, t2);
    a3 = csub(t1, t3);
}

template <bool INV>
MLV_HD void bfly8(cplx (&a)[8]) {
    bfly2(a[0], a[4]);
    bfly2(a[1], a[5]);
    bfly2(a[2], a[6]);
    bfly2(a[3], a[7]);
    a[5] = mul_w8_1<INV>(a[5]);
    a[6] = rot_w4<INV>(a[6]);
    a[7] = mul_w8_3<INV>(a[7]);
    bfly4<INV>(a[0], a[1], a[2], a[3]);      // X0 X2 X4 X6
    bfly4<INV>(a[4], a[5], a[6], a[7]);      // X1 X3 X5 X7
    cplx x1 = a[4], x2 = a[1], x3 = a[5], x4 = a[2], x5 = a[6], x6 = a[3];
    a[1] = x1; a[2] = x2; a[3] = x3; a[4] = x4; a[5] = x5; a[6] = x6;
}

template <bool INV>
MLV_HD void bfly16(cplx (&a)[16]) {
    const double c1 = 0.92387953251128675613;   // cos(pi/8)
    const double s1 = 0.38268343236508977173;   // sin(pi/8)
    // n = 4 n1 + n2 : four radix-4 over n1 (stride 4) -> a[n2 + 4 k1]
    MLV_UNROLL
    for (int n2 = 0; n2 < 4; ++n2) bfly4<INV>(a[n2], a[n2 + 4], a[n2 + 8], a[n2 + 12]);
    // twiddle a[n2 + 4 k1] *= W16^(n2 k1)
    a[5] = mul_w<INV>(a[5], c1, -s1);           // e = 1
    a[6] = mul_w8_1<INV>(a[6]);                 // e = 2
    a[7] = mul_w<INV>(a[7], s1, -c1);           // e = 3
    a[9] = mul_w8_1<INV>(a[9]);                 // e = 2
    a[10] = rot_w4<INV>(a[10]);                 // e = 4
    a[11] = mul_w8_3<INV>(a[11]);               // e = 6
    a[13] = mul_w<INV>(a[13], s1, -c1);         // e = 3
    a[14] = mul_w8_3<INV>(a[14]);               // e = 6
    a[15] = mul_w<INV>(a[15], -c1, s1);         // e = 9
    // four radix-4 over n2 -> X[k1 + 4 k2] at a[4 k1 + k2]
    MLV_UNROLL
    for (int k1 = 0; k1 < 4; ++k1)
        bfly4<INV>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
    // transpose 4x4 to natural order
    cplx t;
    t = a[1];  a[1] = a[4];   a[4] = t;
    t = a[2];  a[2] = a[8];   a[8] = t;
    t = a[3];  a[3] = a[12];  a[12] = t;
    t = a[6];  a[6] = a[9];   a[9] = t;
    t = a[7];  a[7] = a[13];  a[13] = t;
    t = a[11]; a[11] = a[14]; a[14] = t;
}

// radix-R butterflies over the 16 registers: U = 16/R interleaved butterflies,
// butterfly u uses registers u + U*d.
template <int R, bool INV>
MLV_HD void pass_butterflies(cplx (&v)[16]) {
    if constexpr (R == 16) {
        bfly16<INV>(v);
    } else if constexpr (R == 8) {
        MLV_UNROLL
        for (int u = 0; u < 2; ++u) {
            cplx a[8];
            MLV_UNROLL
            for (int d = 0; d < 8; ++d) a[d] = v[u + 2 * d];
            bfly8<INV>(a);
            MLV_UNROLL
            for (int d = 0; d < 8; ++d) v[u + 2 * d] = a[d];
        }
    } else if constexpr (R == 4) {
        MLV_UNROLL
        for (int u = 0; u < 4; ++u) bfly4<INV>(v[u], v[u + 4], v[u + 8], v[u + 12]);
    } else {
        MLV_UNROLL
        for (int u = 0; u < 8; ++u) bfly2(v[u], v[u + 8]);
    }
}

// ---------------------------------------------------------------- twiddles
// After a radix-R pass:  v[u + U d] *= W_{R N3}^{n'' d},  n'' = (tau + T u) mod N3.
// The kernels leave almost no L1 (shared memory takes ~all of the 228 KB), so a
// table lookup per factor costs an L2 round trip each.  Only the log2(R) "binary"
// powers  w^1, w^2, w^4, w^8  are tabulated (table row n'' = log2(R) entries); the
// other powers are products of those (depth <= 3, error a few ulp).  The loads are
// issued before the butterfly of the pass so that their latency hides behind it.
template <int R> struct Log2R { static constexpr int v = R == 16 ? 4 : R == 8 ? 3 : R == 4 ? 2 : 1; };

template <int R, int N3, int T>
MLV_DEV void tw_fetch(cplx (&wb)[4], int tau, int u, const cplx* tw) {
    constexpr int LR = Log2R<R>::v;
    const int n = (tau + T * u) & (N3 - 1);
    MLV_UNROLL
    for (int b = 0; b < LR; ++b) wb[b] = tw[b * N3 + n];
}

template <bool INV>
MLV_HD cplx tw_mul(cplx a, cplx w) { return INV ? cmulc(a, w) : cmul(a, w); }

template <int R, bool INV>
MLV_HD void tw_apply(cplx (&v)[16], int u, const cplx (&wb)[4]) {
    constexpr int U = 16 / R;
#define MLV_V(d) v[u + U * (d)]
    const cplx w1 = wb[0];
    MLV_V(1) = tw_mul<INV>(MLV_V(1), w1);
    if constexpr (R >= 4) {
        const cplx w2 = wb[1];
        const cplx w3 = cmul(w1, w2);
        MLV_V(2) = tw_mul<INV>(MLV_V(2), w2);
        MLV_V(3) = tw_mul<INV>(MLV_V(3), w3);
        if constexpr (R >= 8) {
            const cplx w4 = wb[2];
            const cplx w5 = cmul(w1, w4), w6 = cmul(w2, w4), w7 = cmul(w3, w4);
            MLV_V(4) = tw_mul<INV>(MLV_V(4), w4);
            MLV_V(5) = tw_mul<INV>(MLV_V(5), w5);
            MLV_V(6) = tw_mul<INV>(MLV_V(6), w6);
            MLV_V(7) = tw_mul<INV>(MLV_V(7), w7);
            if constexpr (R >= 16) {
                const cplx w8 = wb[3];
                MLV_V(8) = tw_mul<INV>(MLV_V(8), w8);
                MLV_V(9) = tw_mul<INV>(MLV_V(9), cmul(w1, w8));
                MLV_V(10) = tw_mul<INV>(MLV_V(10), cmul(w2, w8));
                MLV_V(11) = tw_mul<INV>(MLV_V(11), cmul(w3, w8));
                MLV_V(12) = tw_mul<INV>(MLV_V(12), cmul(w4, w8));
                MLV_V(13) = tw_mul<INV>(MLV_V(13), cmul(w5, w8));
                MLV_V(14) = tw_mul<INV>(MLV_V(14), cmul(w6, w8));
                MLV_V(15) = tw_mul<INV>(MLV_V(15), cmul(w7, w8));
            }
        }
    }
#undef MLV_V
}

// butterflies + twiddles of one pass (R, N3); N3 == 1: no twiddles
template <int R, int N3, int T, bool INV>
MLV_DEV void fft_pass(cplx (&v)[16], int tau, const cplx* tw) {
    constexpr int U = 16 / R;
    if constexpr (N3 == 1) {
        pass_butterflies<R, INV>(v);
    } else if constexpr (U == 1) {
        cplx wb[4];
        MLV_SCHED_FENCE();            // keep the table loads inside this pass
        tw_fetch<R, N3, T>(wb, tau, 0, tw);
        pass_butterflies<R, INV>(v);
        tw_apply<R, INV>(v, 0, wb);
    } else {
        pass_butterflies<R, INV>(v);
        MLV_SCHED_FENCE();
        MLV_UNROLL
        for (int u = 0; u < U; ++u) {
            cplx wb[4];
            tw_fetch<R, N3, T>(wb, tau, u, tw);
            tw_apply<R, INV>(v, u, wb);
        }
    }
}

// ------------------------------------------------------ exchange policies
// Order in which the 16 registers of a thread go through an exchange: 0,4,8,12, 1,5,9,13, ...
// The radix-16 butterfly starts with radix-4 butterflies over registers {n, n+4, n+8, n+12} and
// ends with results that land in the same groups, so the first butterflies can start while the
// other loads are still in flight, and the first stores can leave before the last results exist.
#ifdef MLV_USE_XORDER
#define MLV_XORDER(i) ((((i) & 3) << 2) | ((i) >> 2))
#else
#define MLV_XORDER(i) (i)          // measured: no gain (profiles/r02_experiments.md)
#endif

// Full complex128 exchange buffer shared by C interleaved lines (x passes):
// slot L of line c lives at buf[L*C + c].  Two barriers per exchange.
template <int C>
struct XchgFull {
    static constexpr bool SWIZZLE = false;
    cplx* buf;
    int c;
    template <class WI, class RI>
    MLV_DEV void exchange(cplx (&v)[16], WI wi, RI ri) {
        __syncthreads();
        MLV_UNROLL
        for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); buf[wi(j) * C + c] = v[j]; }
        __syncthreads();
        MLV_UNROLL
        for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); v[j] = buf[ri(j) * C + c]; }
    }
};

// Half-size exchange buffer (N+N/16 doubles per line): real parts, then
// imaginary parts.  Four barriers per exchange, half the shared memory.
struct XchgSplit {
    static constexpr bool SWIZZLE = false;
    double* buf;
    template <class WI, class RI>
    MLV_DEV void exchange(cplx (&v)[16], WI wi, RI ri) {
        __syncthreads();
        MLV_UNROLL
        for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); buf[wi(j)] = v[j].x; }
        __syncthreads();
        MLV_UNROLL
        for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); v[j].x = buf[ri(j)]; }
        __syncthreads();
        MLV_UNROLL
        for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); buf[wi(j)] = v[j].y; }
        __syncthreads();
        MLV_UNROLL
        for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); v[j].y = buf[ri(j)]; }
    }
};

// One line, full complex128 slots, exactly N of them: the stride-16 gather of the last exchange is
// made conflict-free by XOR-ing the slot with bits of its own row (L ^ ((L >> 4) & 7)) instead of
// padding one slot per 16 -- 64 KB for a 4096-point line, which lets two such CTAs (plus a
// 44 KB column stash each) share an SM.
struct XchgLineSwz {
    static constexpr bool SWIZZLE = true;
    cplx* buf;
    template <class WI, class RI>
    MLV_DEV void exchange(cplx (&v)[16], WI wi, RI ri) {
        __syncthreads();
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) buf[wi(j)] = v[j];
        __syncthreads();
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) v[j] = buf[ri(j)];
    }
};

// ------------------------------------------------------------ the transform
template <int LOG2N, int P, bool INV, class X>
MLV_DEV void fft_later_passes(cplx (&v)[16], const int tau, const FftTw& tw, X& xc) {
    typedef FftCfg<LOG2N> C;
    constexpr int N3 = C::n3(P);            // length left after this pass
    constexpr bool PAD = (N3 == 1);         // stride-16 gather: pad 1 per 16
    const int lo = tau & (N3 - 1);
    const int hi = tau / N3;
    xc.exchange(
        v,
        [&](int j) { const int L = tau + C::T * j; return PAD ? (X::SWIZZLE ? (L ^ ((L >> 4) & 7)) : L + (L >> 4)) : L; },
        [&](int j) {
            const int L = lo + N3 * j + 16 * N3 * hi;
            return PAD ? (X::SWIZZLE ? (L ^ ((L >> 4) & 7)) : L + (L >> 4)) : L;
        });
    fft_pass<16, N3, C::T, INV>(v, tau, tw.p[P]);
    if constexpr (N3 > 1) fft_later_passes<LOG2N, P + 1, INV, X>(v, tau, tw, xc);
}

// In-register transform of one line.  All T threads of the line (and all
// other lines sharing the CTA) must call this together (it contains CTA
// barriers).  INV = unnormalised inverse (sum with exp(+...)).
template <int LOG2N, bool INV, class X>
MLV_DEV void fft_line(cplx (&v)[16], const int tau_, const FftTw& tw, X& xc) {
    typedef FftCfg<LOG2N> C;
    const int tau = opaque_int(tau_);    // per-transform index arithmetic (see opaque_int)
    fft_pass<C::R0, C::n3(0), C::T, INV>(v, tau, tw.p[0]);
    if constexpr (C::NPASS > 1) fft_later_passes<LOG2N, 1, INV, X>(v, tau, tw, xc);
}

// ------------------------------------------------ grouped three-pass transforms
// N = R0 * 256 (three passes).  The last two passes of the standard order R0 -> 16 -> 16 are 256-point
// sub-transforms that live inside groups of 16 consecutive threads (thread tau = 16 k0 + c): if
// the result may stay in *grouped order*
//     thread tau, register j  <->  index (tau >> 4) + R0 (tau & 15) + (N/16) j
// the exchange before the last pass is a 16x16 transpose inside the group -- half a warp, no
// CTA barrier, and the warps of a line drift apart instead of meeting at every exchange.
// Point-wise work (products, maxima, sums) does not care about the order, and the mirrored
// pass order 16 -> 16 -> R0 takes grouped order back to natural order with its FIRST exchange
// inside the group.  A physical-space stage  inverse -> product -> forward  therefore needs one
// CTA-wide exchange per transform instead of two.
MLV_DEV void warp_sync() {
#ifdef MLV_EMU
    const int w = (int)threadIdx.x >> 5;
    const int left = (int)blockDim.x - 32 * w;
    emu_bar_sync(16 + w, left < 32 ? left : 32);
#else
    __syncwarp();
#endif
}

// 16x16 transpose inside a group of 16 threads through the group's 16*17 doubles of the (half-size)
// exchange buffer: register j of lane c <-> register c of lane j.  Real parts, then imaginary
// parts.  The region belongs to the group (one half-warp), so warp-level ordering is enough
// against the group's own previous use; the caller orders it against CTA-wide users of the
// same buffer with a CTA barrier.
MLV_DEV void group_transpose(cplx (&v)[16], double* gbuf, const int lane) {
    warp_sync();
    MLV_UNROLL
    for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); gbuf[j * 17 + lane] = v[j].x; }
    warp_sync();
    MLV_UNROLL
    for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); v[j].x = gbuf[lane * 17 + j]; }
    warp_sync();
    MLV_UNROLL
    for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); gbuf[j * 17 + lane] = v[j].y; }
    warp_sync();
    MLV_UNROLL
    for (int i = 0; i < 16; ++i) { const int j = MLV_XORDER(i); v[j].y = gbuf[lane * 17 + j]; }
}

// natural order in -> grouped order out (passes R0, 16, 16).  `buf`: the line's XSLOTS doubles.
template <int LOG2N, bool INV>
MLV_DEV void fft_nat2grp(cplx (&v)[16], const int tau_, const FftTw& tw, double* buf) {
    typedef FftCfg<LOG2N> C;
    static_assert(C::NPASS == 3, "grouped transforms: N = R0 * 256");
    const int tau = opaque_int(tau_);
    fft_pass<C::R0, C::n3(0), C::T, INV>(v, tau, tw.p[0]);
    XchgSplit xc;
    xc.buf = buf;
    const int lo = tau & 15, hi = tau >> 4;
    xc.exchange(v, [&](int j) { return tau + C::T * j; }, [&](int j) { return lo + 16 * j + 256 * hi; });
    fft_pass<16, 16, C::T, INV>(v, tau, tw.p[1]);
    __syncthreads();                       // every CTA-wide read done before the groups reuse the buffer
    group_transpose(v, buf + 272 * hi, lo);
    fft_pass<16, 1, C::T, INV>(v, tau, nullptr);
}

// grouped order in -> natural order out (passes 16, 16, R0).  The caller guarantees that no
// CTA-wide exchange is still reading `buf` when this is called (a preceding fft_nat2grp is fine).
template <int LOG2N, bool INV>
MLV_DEV void fft_grp2nat(cplx (&v)[16], const int tau_, const FftTw& tw, double* buf) {
    typedef FftCfg<LOG2N> C;
    static_assert(C::NPASS == 3, "grouped transforms: N = R0 * 256");
    constexpr int U = 16 / C::R0;
    const int tau = opaque_int(tau_);
    const int s = tau & 15, p = tau >> 4;
    // x[n_lo + T j], n_lo = p + R0 s: radix 16 over j, twiddle W_N^(n_lo k0)
    fft_pass<16, C::T, C::T, INV>(v, p + C::R0 * s, tw.g[0]);
    group_transpose(v, buf + 272 * p, s);
    // thread (p, k0), register s: radix 16 over s, twiddle W_T^(p k1)
    fft_pass<16, C::R0, C::T, INV>(v, p, tw.g[1]);
    XchgSplit xc;
    xc.buf = buf;
    xc.exchange(v, [&](int j) { return tau + C::T * j; },
                [&](int j) {
                    const int m = tau + C::T * (j % U);          // k0 + 16 k1 of the output this register feeds
                    return 16 * (j / U) + (m & 15) + 16 * C::R0 * (m >> 4);
                });
    fft_pass<C::R0, 1, C::T, INV>(v, tau, nullptr);
}

}  // namespace mlv
