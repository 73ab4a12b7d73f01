// melvin-b200: kernels for transform lines longer than one SM can hold in registers
// (16384 points; BASELINE config 5).  A complex128 line of 16384 points is 256 KB -- the
// whole register file of an SM -- so these kernels wrap the register-resident transform of
// mlv_fft.cuh (at most 8192 points) in one exact radix-2 step:
//
//   x passes (complex columns): decimation around the half-length transform, N = NF/2,
//       inverse:  y[2p+s] = IDFT_N( (X[k] + (-1)^s X[k+N]) e^{+2 pi i s k/NF} )[p]
//       forward:  X[2q+s] =  DFT_N( (x[p] + (-1)^s x[p+N]) e^{-2 pi i s p/NF} )[q]
//     the two halves s = 0, 1 run one after the other in the same CTA (the second read of
//     the operands hits the L2); the forward pass is k_xfwd<.., SPLIT = 2> in
//     mlv_kernels_fft.cuh, the inverse pass is k_xinv_split below.
//   z stage (real rows): ONE real row of nz points per line as a complex transform of
//     H = nz/2 points, z[p] = x[2p] + i x[2p+1], with the usual untangling step
//       c2r:  Z[k] = (X[k] + conj X[H-k]) + i e^{+2 pi i k/nz} (X[k] - conj X[H-k])
//       r2c:  X[k] = (Z[k] + conj Z[H-k])/2 - (i/2) e^{-2 pi i k/nz} (Z[k] - conj Z[H-k])
//     (the pair-packed form of mlv_kernels_fft.cuh would need a 16384-point complex line).
//
// Semantics are those of the kernels they stand in for (same reference citations):
// SpectralTransformer.py:90-199, Variable.py:119-128, utility.py:62-79, Integrator.py:35-44.
#pragma once

#include "mlv_kernels_fft.cuh"

namespace mlv {

// ===================================================================== x inverse, split
// spectral (2nn+1, nm) -> I (nx, ipitch), nx = 2N; no column stash (it would not fit next
// to the exchange buffer): every field reads its source column from global memory / L2.
template <int LOG2N, int C>
__global__ void __launch_bounds__(C * FftCfg<LOG2N>::T, (C * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_xinv_split(const XInvArgs a) {
    typedef FftCfg<LOG2N> F;
    constexpr int NF = 2 * F::N;
    const int c = threadIdx.x % C, tau = threadIdx.x / C;
    const int mreal = blockIdx.x * C + c;
    const bool valid = mreal < a.nm;
    const int m = valid ? mreal : a.nm - 1;          // clamp: loads stay unpredicated
    XchgFull<C> xc;
    xc.buf = reinterpret_cast<cplx*>(MLV_SMEM_BASE());
    xc.c = c;
    const int mg = m + a.sh.m_off;                   // global column (spectral symbols)
    const int rmask = (1 << a.sh.rpc_shift) - 1;
    for (int f = 0; f < a.nf; ++f) {
        const cplx* __restrict__ src = a.src[f];
        const int op = a.op[f];
        for (int s = 0; s < 2; ++s) {
            cplx v[16];
            const double sg = s ? -1.0 : 1.0;
            MLV_UNROLL
            for (int j0 = 0; j0 < 16; j0 += 4) {     // groups of 4 points: 8 loads in flight
                MLV_SCHED_FENCE();
                const int tq = opaque_int(tau);
                cplx X0[4], X1[4];
                MLV_UNROLL
                for (int u = 0; u < 4; ++u) {
                    const int k = tq + F::T * (j0 + u);
                    int r0 = 0, n0 = 0, r1 = 0, n1 = 0;
                    const bool ok0 = xrow_of(k, NF, a.nn, r0, n0);
                    const bool ok1 = xrow_of(k + F::N, NF, a.nn, r1, n1);
                    X0[u] = ldg_pred(src + (size_t)(ok0 ? r0 : 0) * a.spitch + m, ok0);
                    X1[u] = ldg_pred(src + (size_t)(ok1 ? r1 : 0) * a.spitch + m, ok1);
                }
                MLV_UNROLL
                for (int u = 0; u < 4; ++u) {
                    const int k = tq + F::T * (j0 + u);
                    int r0 = 0, n0 = 0, r1 = 0, n1 = 0;
                    const bool ok0 = xrow_of(k, NF, a.nn, r0, n0);
                    const bool ok1 = xrow_of(k + F::N, NF, a.nn, r1, n1);
                    cplx A = X0[u], B = X1[u];
                    if (op != XOP_IDENT) {               // truncated rows hold 0 and stay 0
                        A = spectral_op(op, A, ok0 ? n0 : 0, mg, a.k);
                        B = spectral_op(op, B, ok1 ? n1 : 0, mg, a.k);
                    }
                    cplx t = mk(A.x + sg * B.x, A.y + sg * B.y);
                    if (s) t = cmulc(t, a.tws[k]);       // * e^{+2 pi i k/NF}
                    v[j0 + u] = t;
                }
            }
            MLV_SCHED_FENCE();
            fft_line<LOG2N, true>(v, tau, a.tw, xc);
            if (valid) {
                const size_t off = (size_t)a.dstoff[f] + m;
                MLV_UNROLL
                for (int j = 0; j < 16; ++j) {
                    const int x = 2 * (tau + F::T * j) + s;   // global row -> block of its owner
                    a.out.blk[x >> a.sh.rpc_shift][off + (size_t)(x & rmask) * a.ipitch] = v[j];
                }
            }
        }
    }
}

// ===================================================================== z stage, real rows
// Untangled input of the half-length inverse transform of one real row (one-sided spectrum
// `row`, retained columns m < nm, F4: Im of the m = 0 bin dropped).  tws[k] = e^{-2 pi i k/nz}.
template <int LOG2H, bool SHARDED>
MLV_DEV void zreal_load_line_(cplx (&v)[16], const cplx* __restrict__ row,
                              const cplx* __restrict__ tws, int tau_, int nm, const Shard& sh) {
    typedef FftCfg<LOG2H> F;
    MLV_UNROLL
    for (int j0 = 0; j0 < 16; j0 += 4) {             // groups of 4 points: 12 loads in flight
        MLV_SCHED_FENCE();
        const int tau = opaque_int(tau_);
        cplx X[4], P[4], W[4];
        MLV_UNROLL
        for (int u = 0; u < 4; ++u) {
            const int k = tau + F::T * (j0 + u), kp = F::N - k;
            const bool ok0 = k < nm, ok1 = kp < nm;
            const int i0 = ok0 ? k : 0, i1 = ok1 ? kp : 0;
            X[u] = ldg_pred(row + (SHARDED ? inv_col_off(i0, sh) : (size_t)i0), ok0);
            P[u] = ldg_pred(row + (SHARDED ? inv_col_off(i1, sh) : (size_t)i1), ok1);
            W[u] = tws[k];
        }
        MLV_UNROLL
        for (int u = 0; u < 4; ++u) {
            const int k = tau + F::T * (j0 + u);
            const double xy = k == 0 ? 0.0 : X[u].y;                 // F4
            const double sx = X[u].x + P[u].x, sy = xy - P[u].y;     // X + conj(P)
            const double dx = X[u].x - P[u].x, dy = xy + P[u].y;     // X - conj(P)
            const double tx = W[u].x * dx + W[u].y * dy;             // conj(w) * D
            const double ty = W[u].x * dy - W[u].y * dx;
            v[j0 + u] = mk(sx - ty, sy + tx);                        // S + i t
        }
    }
    MLV_SCHED_FENCE();
}
template <int LOG2H>
MLV_DEV void zreal_load_line(cplx (&v)[16], const cplx* __restrict__ row,
                             const cplx* __restrict__ tws, int tau, int nm, const Shard& sh) {
    if (sh.inv_chunk == 0) zreal_load_line_<LOG2H, false>(v, row, tws, tau, nm, sh);
    else zreal_load_line_<LOG2H, true>(v, row, tws, tau, nm, sh);
}

// One-sided spectrum X[k], k < nm, of the real row whose half-length forward transform
// every thread holds as Z[tau + T j].  Partners Z[H-k] travel through `pbuf` (nm doubles,
// real parts then imaginary parts: four CTA barriers).  `emit(k, X)` is called for every
// retained k owned by the thread.
template <int LOG2H, class Emit>
MLV_DEV void zreal_unpack(const cplx (&v)[16], int tau, int nm, const cplx* __restrict__ tws,
                          double* pbuf, Emit emit) {
    typedef FftCfg<LOG2H> F;
    // retained k < nm <= (2H-1)/3 < 11 T: j <= 10; partners kk > H - nm > 5 T: j >= 5
    double px[11], py[11];
    __syncthreads();
    MLV_UNROLL
    for (int j = 5; j < 16; ++j) {
        const int kk = tau + F::T * j;
        if (kk > F::N - nm) pbuf[F::N - kk] = v[j].x;
    }
    __syncthreads();
    MLV_UNROLL
    for (int j = 0; j < 11; ++j) {
        const int k = tau + F::T * j;
        px[j] = pbuf[(k < nm && k > 0) ? k : 1];
    }
    __syncthreads();
    MLV_UNROLL
    for (int j = 5; j < 16; ++j) {
        const int kk = tau + F::T * j;
        if (kk > F::N - nm) pbuf[F::N - kk] = v[j].y;
    }
    __syncthreads();
    MLV_UNROLL
    for (int j = 0; j < 11; ++j) {
        const int k = tau + F::T * j;
        py[j] = pbuf[(k < nm && k > 0) ? k : 1];
    }
    MLV_UNROLL
    for (int j0 = 0; j0 < 11; j0 += 4) {
        cplx W[4];
        MLV_UNROLL
        for (int u = 0; u < 4; ++u) {
            const int k = tau + F::T * (j0 + u);
            if (j0 + u < 11) W[u] = tws[k < nm ? k : 0];
        }
        MLV_UNROLL
        for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            if (j >= 11) continue;
            const int k = tau + F::T * j;
            if (k >= nm) continue;
            const cplx Z = v[j];
            const cplx P = k == 0 ? Z : mk(px[j], py[j]);            // Z[H] == Z[0]
            const double ex = 0.5 * (Z.x + P.x), ey = 0.5 * (Z.y - P.y);   // (Z + conj P)/2
            const double ox = 0.5 * (Z.x - P.x), oy = 0.5 * (Z.y + P.y);   // (Z - conj P)/2
            const double ax = W[u].x * ox - W[u].y * oy;             // w * O
            const double ay = W[u].x * oy + W[u].y * ox;
            emit(k, mk(ex + ay, ey - ax));                           // E - i w O
        }
    }
}

struct ZRealArgs {
    int nx, nm, ipitch, ct;      // nx = local rows, nm = global retained columns
    Shard sh;
    const cplx* I;       // c2r input
    cplx* Iout;          // r2c output (tile layout)
    const double* Pin;   // r2c input
    double* P;           // c2r output (nx, nz)
    const cplx* tws;     // e^{-2 pi i k/nz}, k < nz/2
    FftTw tw;
};

// I (nx, ipitch) -> P (nx, nz): one real row per CTA
template <int LOG2H>
__global__ void __launch_bounds__(FftCfg<LOG2H>::T, (FftCfg<LOG2H>::T <= 256) ? 2 : 1)
k_zr_c2r(const ZRealArgs a) {
    typedef FftCfg<LOG2H> F;
    const int tau = threadIdx.x;
    const int x = blockIdx.x;
    XchgSplit xc;
    xc.buf = reinterpret_cast<double*>(MLV_SMEM_BASE());
    cplx v[16];
    zreal_load_line<LOG2H>(v, a.I + (size_t)x * a.ipitch, a.tws, tau, a.nm, a.sh);
    fft_line<LOG2H, true>(v, tau, a.tw, xc);
    cplx* out = reinterpret_cast<cplx*>(a.P + (size_t)x * (2 * F::N));
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) out[tau + F::T * j] = v[j];     // (x[2p], x[2p+1])
}

// P (nx, nz) -> I (tile layout), truncated to m < nm, unnormalised
template <int LOG2H>
__global__ void __launch_bounds__(FftCfg<LOG2H>::T, (FftCfg<LOG2H>::T <= 256) ? 2 : 1)
k_zr_r2c(const ZRealArgs a) {
    typedef FftCfg<LOG2H> F;
    const int tau = threadIdx.x;
    const int x = blockIdx.x;
    XchgSplit xc;
    xc.buf = reinterpret_cast<double*>(MLV_SMEM_BASE());
    cplx v[16];
    const cplx* in = reinterpret_cast<const cplx*>(a.Pin + (size_t)x * (2 * F::N));
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) v[j] = in[tau + F::T * j];
    fft_line<LOG2H, false>(v, tau, a.tw, xc);
    const int cts = log2_pow2(a.ct);
    const bool sharded = a.sh.fwd_chunk != 0;
    zreal_unpack<LOG2H>(v, tau, a.nm, a.tws, xc.buf, [&](int k, cplx X) {
        const int t = k >> cts;
        int h = 0, tl = t;
        if (sharded) { h = t / a.sh.tpr; tl = t - h * a.sh.tpr; }
        a.Iout[(size_t)h * a.sh.fwd_peer + fwd_store_off(x, tl, k & (a.ct - 1), a.ct, a.sh)] = X;
    });
}

// fused physical-space stage of Variable.vec_dot_nabla (see k_z_advect), one real row per CTA.
// Shared memory: [ XSLOTS doubles exchange | H cplx thread-private stash | 4*T doubles ].
template <int LOG2H, bool RED, bool SHARDED>
__global__ void __launch_bounds__(FftCfg<LOG2H>::T, (FftCfg<LOG2H>::T <= 256) ? 2 : 1)
k_zr_advect(const ZAdvArgs a) {
    typedef FftCfg<LOG2H> F;
    constexpr int NT = F::T;
    const int tau = threadIdx.x;
    const int x = a.row0 + blockIdx.x;
    unsigned char* base = MLV_SMEM_BASE();
    XchgSplit xc;
    xc.buf = reinterpret_cast<double*>(base);
    cplx* stash = reinterpret_cast<cplx*>(base + (size_t)F::XSLOTS * sizeof(double)) + tau;
    double* rbuf = reinterpret_cast<double*>(base + (size_t)F::XSLOTS * sizeof(double) +
                                             (size_t)F::N * sizeof(cplx));
    const size_t rowoff = (size_t)x * a.ipitch;
    const int cts = log2_pow2(a.ct);
    if constexpr (SHARDED) wait_arrivals(a.wait_counter, a.wait_expect);

    cplx v[16];
    zreal_load_line_<LOG2H, SHARDED>(v, a.Iq + rowoff, a.tws, tau, a.nm, a.sh);
    fft_line<LOG2H, true>(v, tau, a.tw, xc);
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) stash[j * F::T] = v[j];
    for (int pass = 0; pass < 2; ++pass) {        // pass 0: A = ux q, pass 1: B = uz q
        zreal_load_line_<LOG2H, SHARDED>(v, (pass == 0 ? a.Iux : a.Iuz) + rowoff, a.tws, tau, a.nm, a.sh);
        fft_line<LOG2H, true>(v, tau, a.tw, xc);
        if constexpr (RED) {
            double mx = -INFINITY, ss = 0.0;
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                mx = fmax(mx, fmax(v[j].x, v[j].y));
                ss += v[j].x * v[j].x + v[j].y * v[j].y;
                const cplx q = stash[j * F::T];
                v[j] = mk(v[j].x * q.x, v[j].y * q.y);
            }
            rbuf[pass * NT + tau] = ss != ss ? NAN : mx;      // NaN-propagating, see k_z_advect
            rbuf[(2 + pass) * NT + tau] = ss;
        } else {
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                const cplx q = stash[j * F::T];
                v[j] = mk(v[j].x * q.x, v[j].y * q.y);
            }
        }
        fft_line<LOG2H, false>(v, tau, a.tw, xc);
        const size_t foff = (size_t)a.outoff[pass];
        zreal_unpack<LOG2H>(v, tau, a.nm, a.tws, xc.buf, [&](int k, cplx X) {
            const int t = k >> cts;
            if constexpr (SHARDED) {
                const int h = t / a.sh.tpr, tl = t - h * a.sh.tpr;
                a.out.blk[h][foff + fwd_store_off(x, tl, k & (a.ct - 1), a.ct, a.sh)] = X;
            } else {                                                     // [tile][nx][ct]
                a.IA[foff + ((((size_t)t) << a.sh.fwd_rshift) + (size_t)x) * a.ct + (k & (a.ct - 1))] = X;
            }
        });
    }
    // ---- reductions: per-CTA partials (deterministic two-stage reduction)
    if constexpr (RED) {
    __syncthreads();
    {
        constexpr int G = NT / 4 > 0 ? NT / 4 : 1;          // threads per quantity
        const int w = threadIdx.x / G, g = threadIdx.x % G;
        if (w < 4) {
            double r = rbuf[w * NT + g];
            for (int i = g + G; i < NT; i += G) r = (w < 2) ? nan_max(r, rbuf[w * NT + i]) : r + rbuf[w * NT + i];
            rbuf[w * NT + g] = r;
        }
        for (int s2 = G / 2; s2 > 0; s2 >>= 1) {
            __syncthreads();
            if (w < 4 && g < s2) {
                const double p = rbuf[w * NT + g], q = rbuf[w * NT + g + s2];
                rbuf[w * NT + g] = (w < 2) ? nan_max(p, q) : p + q;
            }
        }
        if (w < 4 && g == 0) a.red[(size_t)blockIdx.x * 4 + w] = rbuf[w * NT];
    }
    }
}

}  // namespace mlv
