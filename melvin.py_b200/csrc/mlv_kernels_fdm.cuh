// melvin-b200: kernels of the Fourier-x / finite-difference-z step (BASELINE config 3;
// reference examples/rayleigh_benard_convection.py:95-145).  Spectral arrays are S_f =
// (nn, nz) complex128, rows n = 0..nn-1 (x modes), z contiguous.  The step is three kernels
// (SURVEY 8(d), config-3 model):
//   K1  k_fdm_solve     row-wise: psi = solve(-w) (LaplacianSolver.py:22-56,70-79) and, from the
//                       row still in shared memory, ux^ = -pddz(psi), uz^ = i kx n psi
//                       (utility.py:62-79; the z stencil commutes with the x transform)
//   K2  k_x1d_advect    column-wise: c2r of ux^, uz^, q^ -> products -> r2c of ux q, uz q
//                       (Variable.py:119-128), CFL / energy reductions on the way
//   K3  k_integrate     row-wise (mlv_kernels_pw.cuh): right-hand side from the stencil ops
//                       below + AB predictor + explicit update (Integrator.py:5-18,53-56)
#pragma once

#include "mlv_kernels_fft.cuh"

namespace mlv {

// ------------------------------------------------------------ row-wise stencil operators
// Linear-term op codes that need the neighbours of an element in its row (FDM-z mode only).
enum {
    XOP_FDM_D2DZ2 = 10,   // pd2dz2(src): central second difference, interior only (SpatialDifferentiator.py:106-128)
    XOP_FDM_DDZ = 11,     // pddz(src): central first difference, edge columns zero (:91-104, :157-185)
    XOP_FDM_NABLA2 = 12,  // d2x n^2 src + pd2dz2(src)   (Variable.snabla2, Variable.py:111-113)
    XOP_FDX_SYM = 13      // i symx[n] src: Fourier symbol of the periodic central x stencil
};

struct FdmConsts {
    int nz, order;
    double dz;
    const double* symx;   // [nn] imaginary part of the x stencil symbol
};

MLV_DEV cplx fdm_ddz(const cplx* __restrict__ s, size_t idx, int z, const FdmConsts& f) {
    const int w = f.order == 2 ? 1 : 2;
    if (z < w || z >= f.nz - w) return mk(0.0, 0.0);
    if (f.order == 2) {
        const cplx p = s[idx + 1], q = s[idx - 1];
        return mk((p.x - q.x) / (2 * f.dz), (p.y - q.y) / (2 * f.dz));
    }
    const cplx p2 = s[idx + 2], p1 = s[idx + 1], q1 = s[idx - 1], q2 = s[idx - 2];
    return mk((-0.25 * p2.x + 2 * p1.x - 2 * q1.x + 0.25 * q2.x) / (3 * f.dz),
              (-0.25 * p2.y + 2 * p1.y - 2 * q1.y + 0.25 * q2.y) / (3 * f.dz));
}

MLV_DEV cplx fdm_d2dz2(const cplx* __restrict__ s, size_t idx, int z, const FdmConsts& f) {
    const int w = f.order == 2 ? 1 : 2;
    if (z < w || z >= f.nz - w) return mk(0.0, 0.0);
    const double h2 = f.dz * f.dz;
    if (f.order == 2) {
        const cplx p = s[idx + 1], c = s[idx], q = s[idx - 1];
        return mk((p.x - 2 * c.x + q.x) / h2, (p.y - 2 * c.y + q.y) / h2);
    }
    const cplx p2 = s[idx + 2], p1 = s[idx + 1], c = s[idx], q1 = s[idx - 1], q2 = s[idx - 2];
    return mk((-1.0 / 12 * p2.x + 4.0 / 3 * p1.x - 5.0 / 2 * c.x + 4.0 / 3 * q1.x - 1.0 / 12 * q2.x) / h2,
              (-1.0 / 12 * p2.y + 4.0 / 3 * p1.y - 5.0 / 2 * c.y + 4.0 / 3 * q1.y - 1.0 / 12 * q2.y) / h2);
}

// value of one linear term in FDM-z mode (row n, column z)
MLV_DEV cplx fdm_term(int op, const cplx* __restrict__ src, size_t idx, int n, int z,
                      const SpecConsts& k, const FdmConsts& f) {
    switch (op) {
        case XOP_FDM_D2DZ2: return fdm_d2dz2(src, idx, z, f);
        case XOP_FDM_DDZ: return fdm_ddz(src, idx, z, f);
        case XOP_FDM_NABLA2: {
            const cplx s = src[idx], d = fdm_d2dz2(src, idx, z, f);
            const double b = k.d2x * ((double)n * (double)n);
            return mk(b * s.x + d.x, b * s.y + d.y);
        }
        case XOP_FDX_SYM: { const cplx s = src[idx]; const double b = f.symx[n]; return mk(-b * s.y, b * s.x); }
        default: return spectral_op(op, src[idx], n, 0, k);
    }
}

MLV_DEV cplx fdm_lin_terms_at(const LinTerms& lt, size_t idx, int n, int z, const SpecConsts& k,
                              const FdmConsts& f) {
    cplx acc = mk(0.0, 0.0);
    for (int i = 0; i < lt.n; ++i)
        acc = cadd(acc, cmul(mk(lt.cre[i], lt.cim[i]), fdm_term(lt.op[i], lt.src[i], idx, n, z, k, f)));
    return acc;
}

// ===================================================================== K1: batched solve
// nn independent systems (LaplacianSolver.py:22-56), one per CTA:
//   rows 1..nz-2:  x[i-1]/dz^2 - (kx_n^2 + 2/dz^2) x[i] + x[i+1]/dz^2 = r[i];  x[0] = r[0], x[nz-1] = r[nz-1].
// With the (real, right-hand-side independent) pivots inv[i] = 1/(b - a c'[i-1]) tabulated at
// context creation both Thomas sweeps are first-order linear recurrences
//   forward   d'[i] = r[i] inv[i] - (a inv[i]) d'[i-1]        backward  x[i] = d'[i] - (a inv[i]) x[i+1]
// i.e. compositions of affine maps y -> g + m y, which are associative: every thread composes
// the maps of its PER consecutive unknowns, the CTA runs a parallel scan over the per-thread
// maps (warp shuffles inside a warp, shared memory across warps), then every thread replays
// its unknowns from its carry-in.  Depth 2 PER + log2(NT) instead of nz, all lanes busy, one
// pass over HBM: the row comes in and the results leave with unit-stride 16-byte lanes through a
// padded shared-memory row (8 consecutive unknowns per thread would otherwise be a 32-way
// bank conflict).
struct FdmSolveArgs {
    const cplx* rhs;
    double sign;          // solves A x = sign * rhs
    cplx* out;            // x (may be null when only the velocities are wanted)
    cplx* uxh;            // optional: -pddz(x)
    cplx* uzh;            // optional: i kx0 n x
    const double* inv;    // (nn, nz) pivots
    int nn, nz;
    double off;           // a = 1/dz^2
    double kx0;
    FdmConsts f;
};

#define MLV_FDM_PER 8

// carry-in of every thread: exclusive scan of the affine maps (M, G) in thread order
// (REV: in reverse thread order).  wbuf: 3 * 32 doubles of shared memory.
template <bool REV>
MLV_DEV cplx affine_scan_carry(double M, cplx G, double* wbuf) {
#ifdef MLV_EMU
    // emulation build (fibers, no warp intrinsics): serial scan by thread 0 through shared memory
    double* sm = wbuf + 96;                       // [3][blockDim.x]
    const int nt = blockDim.x, t = threadIdx.x;
    sm[t] = M; sm[nt + t] = G.x; sm[2 * nt + t] = G.y;
    __syncthreads();
    if (t == 0) {
        cplx c = mk(0.0, 0.0);
        for (int q = 0; q < nt; ++q) {
            const int u = REV ? nt - 1 - q : q;
            const double m = sm[u];
            const cplx g = mk(sm[nt + u], sm[2 * nt + u]);
            sm[nt + u] = c.x; sm[2 * nt + u] = c.y;
            c = mk(g.x + m * c.x, g.y + m * c.y);
        }
    }
    __syncthreads();
    const cplx r = mk(sm[nt + t], sm[2 * nt + t]);
    __syncthreads();
    return r;
#else
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // inclusive scan inside the warp: (M, G) <- (M, G) o (M_prev, G_prev)
    MLV_UNROLL
    for (int d = 1; d < 32; d <<= 1) {
        const double Mo = REV ? __shfl_down_sync(full, M, d) : __shfl_up_sync(full, M, d);
        const double gx = REV ? __shfl_down_sync(full, G.x, d) : __shfl_up_sync(full, G.x, d);
        const double gy = REV ? __shfl_down_sync(full, G.y, d) : __shfl_up_sync(full, G.y, d);
        const bool has = REV ? (lane + d < 32) : (lane >= d);
        if (has) { G = mk(G.x + M * gx, G.y + M * gy); M *= Mo; }
    }
    const int last = REV ? 0 : 31;
    if (lane == last) { wbuf[warp] = M; wbuf[32 + warp] = G.x; wbuf[64 + warp] = G.y; }
    // exclusive value of the previous lane
    double Mp = REV ? __shfl_down_sync(full, M, 1) : __shfl_up_sync(full, M, 1);
    double px = REV ? __shfl_down_sync(full, G.x, 1) : __shfl_up_sync(full, G.x, 1);
    double py = REV ? __shfl_down_sync(full, G.y, 1) : __shfl_up_sync(full, G.y, 1);
    if (lane == (REV ? 31 : 0)) { Mp = 1.0; px = 0.0; py = 0.0; }
    __syncthreads();
    // carry into this warp: totals of the warps before it (at most 32: replayed by every thread)
    cplx c = mk(0.0, 0.0);
    if (REV) {
        for (int w = nw - 1; w > warp; --w) c = mk(wbuf[32 + w] + wbuf[w] * c.x, wbuf[64 + w] + wbuf[w] * c.y);
    } else {
        for (int w = 0; w < warp; ++w) c = mk(wbuf[32 + w] + wbuf[w] * c.x, wbuf[64 + w] + wbuf[w] * c.y);
    }
    __syncthreads();                               // wbuf is reused by the next scan
    return mk(px + Mp * c.x, py + Mp * c.y);
#endif
}

__global__ void __launch_bounds__(512) k_fdm_solve(const FdmSolveArgs a) {
    constexpr int PER = MLV_FDM_PER;
    cplx* row = reinterpret_cast<cplx*>(MLV_SMEM_BASE());            // padded: slot e + (e >> 3)
    const int padded = a.nz + (a.nz >> 3) + 1;
    double* wbuf = reinterpret_cast<double*>(row + padded);
    const int n = blockIdx.x, nt = blockDim.x, t = threadIdx.x;
    const size_t base = (size_t)n * a.nz;
    for (int e = t; e < a.nz; e += nt) row[e + (e >> 3)] = a.rhs[base + e];
    __syncthreads();
    const int e0 = t * PER;
    cplx g[PER];
    double m[PER];
    MLV_UNROLL
    for (int k = 0; k < PER; ++k) {
        const int i = e0 + k;
        g[k] = mk(0.0, 0.0);
        m[k] = 0.0;
        if (i < a.nz) {
            const double iv = a.inv[base + i];
            const double lower = (i == 0 || i == a.nz - 1) ? 0.0 : a.off;
            const cplx r = row[i + (i >> 3)];
            g[k] = mk(a.sign * r.x * iv, a.sign * r.y * iv);
            m[k] = -lower * iv;                    // also -c'[i] of the back substitution
        }
    }
    // ---- forward sweep
    {
        double M = 1.0;
        cplx G = mk(0.0, 0.0);
        MLV_UNROLL
        for (int k = 0; k < PER; ++k)
            if (e0 + k < a.nz) { G = mk(g[k].x + m[k] * G.x, g[k].y + m[k] * G.y); M *= m[k]; }
        cplx c = affine_scan_carry<false>(M, G, wbuf);
        MLV_UNROLL
        for (int k = 0; k < PER; ++k)
            if (e0 + k < a.nz) { c = mk(g[k].x + m[k] * c.x, g[k].y + m[k] * c.y); g[k] = c; }
    }
    // ---- back substitution (unknowns and threads in reverse order)
    {
        double M = 1.0;
        cplx G = mk(0.0, 0.0);
        MLV_UNROLL
        for (int k = PER - 1; k >= 0; --k)
            if (e0 + k < a.nz) { G = mk(g[k].x + m[k] * G.x, g[k].y + m[k] * G.y); M *= m[k]; }
        cplx c = affine_scan_carry<true>(M, G, wbuf);
        MLV_UNROLL
        for (int k = PER - 1; k >= 0; --k)
            if (e0 + k < a.nz) { c = mk(g[k].x + m[k] * c.x, g[k].y + m[k] * c.y); g[k] = c; }
    }
    MLV_UNROLL
    for (int k = 0; k < PER; ++k) {
        const int i = e0 + k;
        if (i < a.nz) row[i + (i >> 3)] = g[k];
    }
    __syncthreads();
    // ---- results leave with unit stride; the velocities come from the row in shared memory
    const double kxn = a.kx0 * n;
    const int w = a.f.order == 2 ? 1 : 2;
    for (int e = t; e < a.nz; e += nt) {
        const cplx x = row[e + (e >> 3)];
        if (a.out) a.out[base + e] = x;
        if (a.uzh) a.uzh[base + e] = mk(-kxn * x.y, kxn * x.x);
        if (a.uxh) {
            cplx d = mk(0.0, 0.0);
            if (e >= w && e < a.nz - w) {
#define MLV_ROW(o) row[(e + (o)) + ((e + (o)) >> 3)]
                if (a.f.order == 2) {
                    const cplx p = MLV_ROW(1), q = MLV_ROW(-1);
                    d = mk((p.x - q.x) / (2 * a.f.dz), (p.y - q.y) / (2 * a.f.dz));
                } else {
                    const cplx p2 = MLV_ROW(2), p1 = MLV_ROW(1), q1 = MLV_ROW(-1), q2 = MLV_ROW(-2);
                    d = mk((-0.25 * p2.x + 2 * p1.x - 2 * q1.x + 0.25 * q2.x) / (3 * a.f.dz),
                           (-0.25 * p2.y + 2 * p1.y - 2 * q1.y + 0.25 * q2.y) / (3 * a.f.dz));
                }
#undef MLV_ROW
            }
            a.uxh[base + e] = mk(-d.x, -d.y);
        }
    }
}

// ===================================================================== 4th-order (pentadiagonal) solve
// An extension with no counterpart in the reference (its finite-difference Laplacian is always the
// 2nd-order tridiagonal one, LaplacianSolver.py:22-45; BASELINE's north star names the pentadiagonal
// case): per x mode n the system  (D4 - (kx n)^2) psi = rhs  with D4 the 4th-order central second
// difference the reference uses for derivatives (SpatialDifferentiator.py:121-130) on rows
// 2..nz-3, the 2nd-order one on rows 1 and nz-2, and identity rows 0 and nz-1 (the reference's
// "solution matches the right-hand side on the boundary" convention, LaplacianSolver.py:46-51).
// The context holds the LU factors (no pivoting: the operator is negative definite); both
// substitutions are SECOND-order linear recurrences
//     z_i = g_i + a_i z_(i-1) + b_i z_(i-2)
// evaluated as a parallel scan of affine maps on the pair (z_i, z_(i-1)): PER consecutive unknowns
// per thread composed in registers, warp shuffles inside a warp, one exchange between warps.
struct Aff2 {              // (z_i, z_(i-1)) = M (z_s, z_(s-1)) + G,  M = [[m00, m01], [m10, m11]] real
    double m00, m01, m10, m11;
    cplx g0, g1;
};
MLV_DEV Aff2 aff2_identity() { Aff2 r; r.m00 = 1; r.m01 = 0; r.m10 = 0; r.m11 = 1; r.g0 = mk(0, 0); r.g1 = mk(0, 0); return r; }
// one more element on top of `p`:  z_new = g + a z_i + b z_(i-1)
MLV_DEV Aff2 aff2_push(const Aff2& p, double a, double b, cplx g) {
    Aff2 r;
    r.m00 = a * p.m00 + b * p.m10; r.m01 = a * p.m01 + b * p.m11;
    r.m10 = p.m00; r.m11 = p.m01;
    r.g0 = mk(g.x + a * p.g0.x + b * p.g1.x, g.y + a * p.g0.y + b * p.g1.y);
    r.g1 = p.g0;
    return r;
}
// later o earlier
MLV_DEV Aff2 aff2_compose(const Aff2& l, const Aff2& e) {
    Aff2 r;
    r.m00 = l.m00 * e.m00 + l.m01 * e.m10; r.m01 = l.m00 * e.m01 + l.m01 * e.m11;
    r.m10 = l.m10 * e.m00 + l.m11 * e.m10; r.m11 = l.m10 * e.m01 + l.m11 * e.m11;
    r.g0 = mk(l.g0.x + l.m00 * e.g0.x + l.m01 * e.g1.x, l.g0.y + l.m00 * e.g0.y + l.m01 * e.g1.y);
    r.g1 = mk(l.g1.x + l.m10 * e.g0.x + l.m11 * e.g1.x, l.g1.y + l.m10 * e.g0.y + l.m11 * e.g1.y);
    return r;
}

// carry-in pair of every thread: exclusive scan of the maps in thread order (REV: reverse order),
// applied to the zero start pair.  wbuf: 8 * 32 doubles of shared memory (+ 8 * blockDim.x in the
// emulation build).
template <bool REV>
MLV_DEV void aff2_scan_carry(Aff2 t, double* wbuf, cplx& c0, cplx& c1) {
#ifdef MLV_EMU
    double* sm = wbuf + 256;                      // [8][blockDim.x]
    const int nt = blockDim.x, id = threadIdx.x;
    sm[id] = t.m00; sm[nt + id] = t.m01; sm[2 * nt + id] = t.m10; sm[3 * nt + id] = t.m11;
    sm[4 * nt + id] = t.g0.x; sm[5 * nt + id] = t.g0.y; sm[6 * nt + id] = t.g1.x; sm[7 * nt + id] = t.g1.y;
    __syncthreads();
    if (id == 0) {
        cplx a = mk(0.0, 0.0), b = mk(0.0, 0.0);
        for (int q = 0; q < nt; ++q) {
            const int u = REV ? nt - 1 - q : q;
            const double m00 = sm[u], m01 = sm[nt + u], m10 = sm[2 * nt + u], m11 = sm[3 * nt + u];
            const cplx g0 = mk(sm[4 * nt + u], sm[5 * nt + u]), g1 = mk(sm[6 * nt + u], sm[7 * nt + u]);
            sm[4 * nt + u] = a.x; sm[5 * nt + u] = a.y; sm[6 * nt + u] = b.x; sm[7 * nt + u] = b.y;
            const cplx na = mk(g0.x + m00 * a.x + m01 * b.x, g0.y + m00 * a.y + m01 * b.y);
            const cplx nb = mk(g1.x + m10 * a.x + m11 * b.x, g1.y + m10 * a.y + m11 * b.y);
            a = na; b = nb;
        }
    }
    __syncthreads();
    c0 = mk(sm[4 * nt + id], sm[5 * nt + id]);
    c1 = mk(sm[6 * nt + id], sm[7 * nt + id]);
    __syncthreads();
#else
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    auto shfl = [&](double v, int d) { return REV ? __shfl_down_sync(full, v, d) : __shfl_up_sync(full, v, d); };
    auto fetch = [&](const Aff2& s, int d) {
        Aff2 o;
        o.m00 = shfl(s.m00, d); o.m01 = shfl(s.m01, d); o.m10 = shfl(s.m10, d); o.m11 = shfl(s.m11, d);
        o.g0 = mk(shfl(s.g0.x, d), shfl(s.g0.y, d)); o.g1 = mk(shfl(s.g1.x, d), shfl(s.g1.y, d));
        return o;
    };
    MLV_UNROLL
    for (int d = 1; d < 32; d <<= 1) {             // inclusive scan inside the warp
        const Aff2 o = fetch(t, d);
        const bool has = REV ? (lane + d < 32) : (lane >= d);
        if (has) t = aff2_compose(t, o);
    }
    if (lane == (REV ? 0 : 31)) {
        double* w = wbuf + 8 * warp;
        w[0] = t.m00; w[1] = t.m01; w[2] = t.m10; w[3] = t.m11;
        w[4] = t.g0.x; w[5] = t.g0.y; w[6] = t.g1.x; w[7] = t.g1.y;
    }
    Aff2 prev = fetch(t, 1);                       // exclusive value inside the warp
    if (lane == (REV ? 31 : 0)) prev = aff2_identity();
    __syncthreads();
    cplx a = mk(0.0, 0.0), b = mk(0.0, 0.0);       // carry into this warp (at most 32 warps: replayed)
    auto step = [&](int w) {
        const double* q = wbuf + 8 * w;
        const cplx na = mk(q[4] + q[0] * a.x + q[1] * b.x, q[5] + q[0] * a.y + q[1] * b.y);
        const cplx nb = mk(q[6] + q[2] * a.x + q[3] * b.x, q[7] + q[2] * a.y + q[3] * b.y);
        a = na; b = nb;
    };
    if (REV) { for (int w = nw - 1; w > warp; --w) step(w); }
    else { for (int w = 0; w < warp; ++w) step(w); }
    __syncthreads();                               // wbuf is reused by the next scan
    c0 = mk(prev.g0.x + prev.m00 * a.x + prev.m01 * b.x, prev.g0.y + prev.m00 * a.y + prev.m01 * b.y);
    c1 = mk(prev.g1.x + prev.m10 * a.x + prev.m11 * b.x, prev.g1.y + prev.m10 * a.y + prev.m11 * b.y);
#endif
}

struct FdmSolve5Args {
    const cplx* rhs;      // (nn, nz)
    cplx* out;
    const double* fa;     // forward:  y_i = r_i + fa_i y_(i-1) + fb_i y_(i-2)
    const double* fb;
    const double* dinv;   // backward: x_i = dinv_i y_i + ba_i x_(i+1) + bb_i x_(i+2)
    const double* ba;
    const double* bb;
    int nn, nz;
};

__global__ void __launch_bounds__(512) k_fdm_solve5(const FdmSolve5Args a) {
    constexpr int PER = MLV_FDM_PER;
    cplx* row = reinterpret_cast<cplx*>(MLV_SMEM_BASE());            // padded: slot e + (e >> 3)
    const int padded = a.nz + (a.nz >> 3) + 1;
    double* wbuf = reinterpret_cast<double*>(row + padded);
    const int n = blockIdx.x, nt = blockDim.x, t = threadIdx.x;
    const size_t base = (size_t)n * a.nz;
    for (int e = t; e < a.nz; e += nt) row[e + (e >> 3)] = a.rhs[base + e];
    __syncthreads();
    const int e0 = t * PER;
    cplx g[PER];
    // ---- forward substitution
    {
        double ca[PER], cb[PER];
        Aff2 m = aff2_identity();
        MLV_UNROLL
        for (int k = 0; k < PER; ++k) {
            const int i = e0 + k;
            ca[k] = 0.0; cb[k] = 0.0; g[k] = mk(0.0, 0.0);
            if (i < a.nz) {
                ca[k] = a.fa[base + i]; cb[k] = a.fb[base + i];
                g[k] = row[i + (i >> 3)];
                m = aff2_push(m, ca[k], cb[k], g[k]);
            }
        }
        cplx z1, z2;                                 // y_(e0-1), y_(e0-2)
        aff2_scan_carry<false>(m, wbuf, z1, z2);
        MLV_UNROLL
        for (int k = 0; k < PER; ++k)
            if (e0 + k < a.nz) {
                const cplx y = mk(g[k].x + ca[k] * z1.x + cb[k] * z2.x, g[k].y + ca[k] * z1.y + cb[k] * z2.y);
                z2 = z1; z1 = y; g[k] = y;
            }
    }
    // ---- back substitution (unknowns and threads in reverse order)
    {
        double ca[PER], cb[PER];
        Aff2 m = aff2_identity();
        MLV_UNROLL
        for (int k = PER - 1; k >= 0; --k) {
            const int i = e0 + k;
            ca[k] = 0.0; cb[k] = 0.0;
            if (i < a.nz) {
                const double dv = a.dinv[base + i];
                ca[k] = a.ba[base + i]; cb[k] = a.bb[base + i];
                g[k] = mk(dv * g[k].x, dv * g[k].y);
                m = aff2_push(m, ca[k], cb[k], g[k]);
            }
        }
        cplx z1, z2;                                 // x_(e0+PER), x_(e0+PER+1)
        aff2_scan_carry<true>(m, wbuf, z1, z2);
        MLV_UNROLL
        for (int k = PER - 1; k >= 0; --k)
            if (e0 + k < a.nz) {
                const cplx x = mk(g[k].x + ca[k] * z1.x + cb[k] * z2.x, g[k].y + ca[k] * z1.y + cb[k] * z2.y);
                z2 = z1; z1 = x; g[k] = x;
            }
    }
    MLV_UNROLL
    for (int k = 0; k < PER; ++k) {
        const int i = e0 + k;
        if (i < a.nz) row[i + (i >> 3)] = g[k];
    }
    __syncthreads();
    for (int e = t; e < a.nz; e += nt) a.out[base + e] = row[e + (e >> 3)];     // unit stride
}

// ===================================================================== K2: fused 1-D advection
// Half-size exchange buffer shared by C interleaved lines: real parts, then imaginary parts
// (four barriers per exchange; leaves room for the physical-space stash of q).
template <int C>
struct XchgSplitC {
    static constexpr bool SWIZZLE = false;
    double* buf;
    int c;
    template <class WI, class RI>
    MLV_DEV void exchange(cplx (&v)[16], WI wi, RI ri) {
        __syncthreads();
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) buf[wi(j) * C + c] = v[j].x;
        __syncthreads();
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) v[j].x = buf[ri(j) * C + c];
        __syncthreads();
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) buf[wi(j) * C + c] = v[j].y;
        __syncthreads();
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) v[j].y = buf[ri(j) * C + c];
    }
};

struct X1dAdvArgs {
    int nn, nz;
    const cplx* uxh;     // (nn, nz) x spectra of ux, uz and the advected scalar
    const cplx* uzh;
    const cplx* q;
    cplx* A;             // out: x spectrum of ux q / nx
    cplx* B;             // out: x spectrum of uz q / nx
    double scale;        // 1/nx (SpectralTransformer.py:85)
    double* red;         // [gridDim.x][4] partials: max ux, max uz, sum ux^2, sum uz^2
    FftTw tw;
};

// packed line of the two adjacent real columns z0, z0+1 (Hermitian extension in x, 2/3-rule
// truncation, SpectralTransformer.py:33-61); branch-free loads
template <int LOG2N>
MLV_DEV void x1d_load_line(cplx (&v)[16], const cplx* __restrict__ S, int tau_, int nn, int nz,
                           int z0, bool valid, bool has2) {
    typedef FftCfg<LOG2N> F;
    const int tau = opaque_int(tau_);
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) {
        if (MLV_MID(j)) { v[j] = mk(0.0, 0.0); continue; }
        const int kk = tau + F::T * j;
        const bool lo = kk < nn, hi = kk > F::N - nn;
        const int mm = lo ? kk : (hi ? F::N - kk : 0);
        const cplx A = ldg_pred(S + (size_t)mm * nz + z0, valid && (lo || hi));
        const cplx B = ldg_pred(S + (size_t)mm * nz + z0 + 1, has2 && (lo || hi));
        const double s = lo ? 1.0 : -1.0;            // hi: conj(A) + i conj(B)
        const double ay = kk == 0 ? 0.0 : A.y, by = kk == 0 ? 0.0 : B.y;
        v[j] = mk(A.x - s * by, s * ay + B.x);
    }
}

template <int LOG2N, int C, bool RED>
__global__ void __launch_bounds__(C * FftCfg<LOG2N>::T, (C * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_x1d_advect(const X1dAdvArgs a) {
    typedef FftCfg<LOG2N> F;
    constexpr int NT = C * F::T;
    const int c = threadIdx.x % C, tau = threadIdx.x / C;
    const int z0 = 2 * (blockIdx.x * C + c);
    const bool valid = z0 < a.nz, has2 = z0 + 1 < a.nz;
    const int zl = valid ? z0 : 0;                   // clamp: addresses stay in range
    // shared memory: [ XSLOTS*C doubles exchange | N*C cplx thread-private stash | 4*NT doubles ]
    unsigned char* base = MLV_SMEM_BASE();
    XchgSplitC<C> xc;
    xc.buf = reinterpret_cast<double*>(base);
    xc.c = c;
    cplx* stash = reinterpret_cast<cplx*>(base + (size_t)F::XSLOTS * C * sizeof(double)) + threadIdx.x;
    double* rbuf = reinterpret_cast<double*>(base + (size_t)F::XSLOTS * C * sizeof(double) +
                                             (size_t)F::N * C * sizeof(cplx));
    cplx* pbuf = reinterpret_cast<cplx*>(xc.buf);    // partners of the r2c unpacking: [nn][C]
    cplx v[16];
    x1d_load_line<LOG2N>(v, a.q, tau, a.nn, a.nz, zl, valid, has2);
    fft_line<LOG2N, true>(v, tau, a.tw, xc);
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) stash[j * NT] = v[j];
    for (int pass = 0; pass < 2; ++pass) {           // pass 0: A = ux q, pass 1: B = uz q
        MLV_SCHED_FENCE();
        x1d_load_line<LOG2N>(v, pass == 0 ? a.uxh : a.uzh, tau, a.nn, a.nz, zl, valid, has2);
        MLV_SCHED_FENCE();
        fft_line<LOG2N, true>(v, tau, a.tw, xc);
        if constexpr (RED) {
            double mx = -INFINITY, ss = 0.0;
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                mx = fmax(mx, has2 ? fmax(v[j].x, v[j].y) : v[j].x);
                ss += v[j].x * v[j].x + (has2 ? v[j].y * v[j].y : 0.0);
                const cplx q = stash[j * NT];
                v[j] = mk(v[j].x * q.x, v[j].y * q.y);
            }
            rbuf[pass * NT + threadIdx.x] = valid ? (ss != ss ? NAN : mx) : -INFINITY;
            rbuf[(2 + pass) * NT + threadIdx.x] = valid ? ss : 0.0;
        } else {                                 // no ticker reads the reductions of this step
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                const cplx q = stash[j * NT];
                v[j] = mk(v[j].x * q.x, v[j].y * q.y);
            }
        }
        fft_line<LOG2N, false>(v, tau, a.tw, xc);
        __syncthreads();
        MLV_UNROLL
        for (int j = 10; j < 16; ++j) {              // retained high modes: j >= 10
            const int kk = tau + F::T * j;
            if (kk > F::N - a.nn) pbuf[(size_t)(F::N - kk) * C + c] = v[j];
        }
        __syncthreads();
        if (valid) {
            cplx* out = pass == 0 ? a.A : a.B;
            MLV_UNROLL
            for (int j = 0; j < 6; ++j) {            // retained low modes: j <= 5
                const int kk = tau + F::T * j;
                if (kk < a.nn) {
                    const cplx P = (kk == 0) ? v[j] : pbuf[(size_t)kk * C + c];
                    cplx X, Y;
                    zpair_unpack(v[j], P, X, Y);
                    cplx* o = out + (size_t)kk * a.nz + z0;
                    o[0] = cscale(X, a.scale);
                    if (has2) o[1] = cscale(Y, a.scale);
                }
            }
        }
    }
    // ---- reductions: per-CTA partials
    if constexpr (RED) {
    __syncthreads();
    {
        constexpr int G = NT / 4 > 0 ? NT / 4 : 1;
        const int w = threadIdx.x / G, g = threadIdx.x % G;
        if (w < 4) {
            double r = rbuf[w * NT + g];
            for (int i = g + G; i < NT; i += G) r = (w < 2) ? nan_max(r, rbuf[w * NT + i]) : r + rbuf[w * NT + i];
            rbuf[w * NT + g] = r;
        }
        for (int s2 = G / 2; s2 > 0; s2 >>= 1) {
            __syncthreads();
            if (w < 4 && g < s2) {
                const double x = rbuf[w * NT + g], y = rbuf[w * NT + g + s2];
                rbuf[w * NT + g] = (w < 2) ? nan_max(x, y) : x + y;
            }
        }
        if (w < 4 && g == 0) a.red[(size_t)blockIdx.x * 4 + w] = rbuf[w * NT];
    }
    }
}

}  // namespace mlv
