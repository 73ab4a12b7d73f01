// melvin-b200: pointwise / stencil / reduction / banded-solve kernels.
// All are HBM-bound streaming kernels: one pass over their operands, unit-stride
// across lanes (16-byte complex128 or 8-byte float64 per lane).
#pragma once

#include "mlv_kernels_fdm.cuh"

namespace mlv {

// A 2-D strided view (element strides).  cplx views use 16-byte elements.
struct View2 {
    void* p;
    long long rs, cs;   // row / column stride in elements
};

// ------------------------------------------------------------------ spectral
// out = sum_i coef_i * op_i(src_i) on a spectral-shaped contiguous array.
// Fully spectral: rows r=0..2nn (n = r <= nn ? r : r-2nn-1), cols m.
// FDM-z: rows n = 0..nn-1, cols are z points (ops that need m are rejected by host).
struct SpecLinArgs {
    int rows, cols, nn, fdm;
    int m_off, nm_glob;          // slab decomposition: global column = m_off + local column
    LinTerms lin;
    cplx* out;
    SpecConsts k;
    FdmConsts f;                 // FDM-z: row-wise stencil operators
};

__global__ void __launch_bounds__(256) k_spec_lincomb(const SpecLinArgs a) {
    const size_t total = (size_t)a.rows * a.cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / a.cols), m = (int)(i % a.cols);
        const int n = a.fdm ? r : (r <= a.nn ? r : r - 2 * a.nn - 1);
        const int mg = a.fdm ? 0 : m + a.m_off;
        if (a.fdm) a.out[i] = fdm_lin_terms_at(a.lin, i, n, m, a.k, a.f);
        else a.out[i] = mg < a.nm_glob ? lin_terms_at(a.lin, i, n, mg, a.k) : mk(0.0, 0.0);
    }
}

// real array of the Laplacian symbol (SpatialDifferentiator.py:70-74)
__global__ void __launch_bounds__(256)
k_lap_array(double* out, int rows, int cols, int nn, SpecConsts k, double coef, int m_off) {
    const size_t total = (size_t)rows * cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), m = (int)(i % cols);
        const int n = (r <= nn ? r : r - 2 * nn - 1);
        out[i] = coef * lap_symbol(n, m + m_off, k);
    }
}

// f0 += lin terms (optional); q_out = integrate(q_in, history)   (Integrator.py:53-63)
struct IntegKArgs {
    int rows, cols, nn, fdm;
    int m_off;
    int f0_set;                  // 1: f0 := linear terms (the whole right-hand side); 0: f0 += linear terms
    LinTerms lin;
    IntegArgs integ;
    SpecConsts k;
    FdmConsts f;
};

__global__ void __launch_bounds__(256) k_integrate(const IntegKArgs a) {
    const size_t total = (size_t)a.rows * a.cols;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / a.cols), m = (int)(i % a.cols);
        const int n = a.fdm ? r : (r <= a.nn ? r : r - 2 * a.nn - 1);
        const int mg = a.fdm ? 0 : m + a.m_off;
        cplx f0 = a.f0_set ? mk(0.0, 0.0) : a.integ.f0[i];
        if (a.lin.n > 0 || a.f0_set) {
            f0 = cadd(f0, a.fdm ? fdm_lin_terms_at(a.lin, i, n, m, a.k, a.f) : lin_terms_at(a.lin, i, n, mg, a.k));
            a.integ.f0[i] = f0;
        }
        integrate_point(a.integ, f0, i, n, mg, a.k);
    }
}

// ------------------------------------------------------------------ stencils
// Central first derivative along x (axis 0) or z (axis 1) of a (rows, cols) array
// of `ncomp`-component elements (1: float64, 2: complex128 viewed as 2 doubles).
// SpatialDifferentiator.py:76-104 (order 2), :130-185 (order 4).
MLV_DEV int wrapi(int i, int n) { return i < 0 ? i + n : (i >= n ? i - n : i); }

struct StencilArgs {
    const double* in;
    double* out;
    int rows, cols, ncomp;
    int axis, order, periodic, second;   // second=1: d2/dz2 (interior only, :106-128)
    double h;
};

MLV_DEV double stencil_at(const StencilArgs& a, int i, int j, int comp) {
    const int n = a.axis == 0 ? a.rows : a.cols;
    const int p = a.axis == 0 ? i : j;
    const int w = a.order == 2 ? 1 : 2;
    const bool edge = (p < w) || (p >= n - w);
    if (edge && !(a.periodic && !a.second)) return 0.0;
    const size_t cs = a.ncomp;
    const size_t rs = (size_t)a.cols * a.ncomp;
#define MLV_AT(off)                                                                      \
    (a.axis == 0 ? a.in[(size_t)wrapi(i + (off), n) * rs + (size_t)j * cs + comp]        \
                 : a.in[(size_t)i * rs + (size_t)wrapi(j + (off), n) * cs + comp])
    if (a.second) {
        if (a.order == 2) return (MLV_AT(1) - 2 * MLV_AT(0) + MLV_AT(-1)) / (a.h * a.h);
        return (-1.0 / 12 * MLV_AT(2) + 4.0 / 3 * MLV_AT(1) - 5.0 / 2 * MLV_AT(0) +
                4.0 / 3 * MLV_AT(-1) - 1.0 / 12 * MLV_AT(-2)) / (a.h * a.h);
    }
    if (a.order == 2) return (MLV_AT(1) - MLV_AT(-1)) / (2 * a.h);
    return (-0.25 * MLV_AT(2) + 2 * MLV_AT(1) - 2 * MLV_AT(-1) + 0.25 * MLV_AT(-2)) / (3 * a.h);
#undef MLV_AT
}

__global__ void __launch_bounds__(256) k_stencil(const StencilArgs a) {
    const size_t total = (size_t)a.rows * a.cols * a.ncomp;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const int comp = (int)(t % a.ncomp);
        const size_t e = t / a.ncomp;
        const int i = (int)(e / a.cols), j = (int)(e % a.cols);
        a.out[t] = stencil_at(a, i, j, comp);
    }
}

// out = d/dx(ux*q) + d/dz(uz*q) in physical space (Variable.py:125-127), used when
// the operands are materialised physical arrays (FDM-z mode, user-supplied velocities).
struct AdvectArgs {
    const double* ux;
    const double* uz;
    const double* q;
    double* out;
    int nx, nz, order, z_periodic;
    double dx, dz;
};

__global__ void __launch_bounds__(256) k_advect_phys(const AdvectArgs a) {
    const size_t total = (size_t)a.nx * a.nz;
    const int w = a.order == 2 ? 1 : 2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / a.nz), j = (int)(t % a.nz);
#define MLV_FX(off) (a.ux[(size_t)wrapi(i + (off), a.nx) * a.nz + j] * a.q[(size_t)wrapi(i + (off), a.nx) * a.nz + j])
#define MLV_FZ(off) (a.uz[(size_t)i * a.nz + wrapi(j + (off), a.nz)] * a.q[(size_t)i * a.nz + wrapi(j + (off), a.nz)])
        double ddx, ddz = 0.0;
        const bool zedge = (j < w) || (j >= a.nz - w);
        if (a.order == 2) {
            ddx = (MLV_FX(1) - MLV_FX(-1)) / (2 * a.dx);
            if (a.z_periodic || !zedge) ddz = (MLV_FZ(1) - MLV_FZ(-1)) / (2 * a.dz);
        } else {
            ddx = (-0.25 * MLV_FX(2) + 2 * MLV_FX(1) - 2 * MLV_FX(-1) + 0.25 * MLV_FX(-2)) / (3 * a.dx);
            if (a.z_periodic || !zedge)
                ddz = (-0.25 * MLV_FZ(2) + 2 * MLV_FZ(1) - 2 * MLV_FZ(-1) + 0.25 * MLV_FZ(-2)) / (3 * a.dz);
        }
#undef MLV_FX
#undef MLV_FZ
        a.out[t] = ddx + ddz;
    }
}

// ------------------------------------------------------ generic elementwise
enum { EW_ADD = 0, EW_SUB = 1, EW_MUL = 2, EW_DIV = 3, EW_COPY = 4, EW_POW = 5 };
// operand kinds
enum { EK_REAL = 0, EK_CPLX = 1, EK_SCALAR = 2 };

struct EwArgs {
    int op;
    int rows, cols;
    View2 out;  int out_kind;          // EK_REAL / EK_CPLX
    View2 a;    int a_kind;            // EK_REAL / EK_CPLX / EK_SCALAR
    View2 b;    int b_kind;
    double a_re, a_im, b_re, b_im;     // scalar operands
};

MLV_DEV cplx ew_load(const View2& v, int kind, double sre, double sim, int i, int j) {
    if (kind == EK_SCALAR) return mk(sre, sim);
    const long long off = (long long)i * v.rs + (long long)j * v.cs;
    if (kind == EK_REAL) return mk(reinterpret_cast<const double*>(v.p)[off], 0.0);
    return reinterpret_cast<const cplx*>(v.p)[off];
}

MLV_DEV cplx cdiv(cplx a, cplx b, bool b_real) {
    if (b_real) return mk(a.x / b.x, a.y / b.x);
    // Smith's algorithm, as NumPy's complex division
    if (fabs(b.x) >= fabs(b.y)) {
        const double r = b.y / b.x, d = b.x + b.y * r;
        return mk((a.x + a.y * r) / d, (a.y - a.x * r) / d);
    }
    const double r = b.x / b.y, d = b.x * r + b.y;
    return mk((a.x * r + a.y) / d, (a.y * r - a.x) / d);
}

__global__ void __launch_bounds__(256) k_elementwise(const EwArgs g) {
    const size_t total = (size_t)g.rows * g.cols;
    const bool a_real = g.a_kind == EK_REAL || (g.a_kind == EK_SCALAR && g.a_im == 0.0);
    const bool b_real = g.b_kind == EK_REAL || (g.b_kind == EK_SCALAR && g.b_im == 0.0);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / g.cols), j = (int)(t % g.cols);
        const cplx x = ew_load(g.a, g.a_kind, g.a_re, g.a_im, i, j);
        cplx r = x;
        if (g.op != EW_COPY) {
            const cplx y = ew_load(g.b, g.b_kind, g.b_re, g.b_im, i, j);
            switch (g.op) {
                case EW_ADD: r = cadd(x, y); break;
                case EW_SUB: r = csub(x, y); break;
                case EW_MUL:
                    r = a_real ? mk(x.x * y.x, x.x * y.y) : (b_real ? mk(x.x * y.x, x.y * y.x) : cmul(x, y));
                    break;
                case EW_DIV: r = cdiv(x, y, b_real); break;
                case EW_POW: {       // real base, small non-negative integer exponent (x**2)
                    const int e = (int)y.x;
                    double p = 1.0;
                    for (int q = 0; q < e; ++q) p *= x.x;
                    r = mk(p, 0.0);
                } break;
                default: break;
            }
        }
        const long long off = (long long)i * g.out.rs + (long long)j * g.out.cs;
        if (g.out_kind == EK_REAL) reinterpret_cast<double*>(g.out.p)[off] = r.x;
        else reinterpret_cast<cplx*>(g.out.p)[off] = r;
    }
}

// ---------------------------------------------------------------- reductions
enum { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2, RED_SUMSQ = 3, RED_SUMPROD = 4 };

struct RedArgs {
    int op, rows, cols;
    View2 a, b;               // real views (b only for RED_SUMPROD)
    double* partial;          // [gridDim.x]
};

MLV_DEV double red_combine(int op, double x, double y) {
    if (op == RED_MAX) return fmax(x, y);
    if (op == RED_MIN) return fmin(x, y);
    return x + y;
}
MLV_DEV double red_identity(int op) {
    if (op == RED_MAX) return -INFINITY;
    if (op == RED_MIN) return INFINITY;
    return 0.0;
}

// max/min propagate NaN like NumPy (any NaN -> NaN), needed by the CFL check
// (Integrator.py:41 `np.isnan(cfl_dt)`).
__global__ void __launch_bounds__(256) k_reduce(const RedArgs g) {
    double* sm = reinterpret_cast<double*>(MLV_SMEM_BASE());
    const size_t total = (size_t)g.rows * g.cols;
    double acc = red_identity(g.op);
    bool nan = false;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t / g.cols), j = (int)(t % g.cols);
        double x = reinterpret_cast<const double*>(g.a.p)[(long long)i * g.a.rs + (long long)j * g.a.cs];
        if (g.op == RED_SUMSQ) x = x * x;
        if (g.op == RED_SUMPROD)
            x *= reinterpret_cast<const double*>(g.b.p)[(long long)i * g.b.rs + (long long)j * g.b.cs];
        nan = nan || (x != x);
        acc = red_combine(g.op, acc, x);
    }
    if (nan) acc = NAN;
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            const double x = sm[threadIdx.x], y = sm[threadIdx.x + s];
            sm[threadIdx.x] = (x != x || y != y) ? NAN : red_combine(g.op, x, y);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) g.partial[blockIdx.x] = sm[0];
}

// final stage: combine `n` partials (stride `stride`, offset `off`) into out[0]
__global__ void __launch_bounds__(256)
k_reduce_final(const double* partial, int n, int stride, int off, int op, double* out) {
    double* sm = reinterpret_cast<double*>(MLV_SMEM_BASE());
    double acc = red_identity(op);
    bool nan = false;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = partial[(size_t)i * stride + off];
        nan = nan || (x != x);
        acc = red_combine(op, acc, x);
    }
    sm[threadIdx.x] = nan ? NAN : acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            const double x = sm[threadIdx.x], y = sm[threadIdx.x + s];
            sm[threadIdx.x] = (x != x || y != y) ? NAN : red_combine(op, x, y);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sm[0];
}

// final stage of the fused z stage's reductions: block w combines column w of the
// [n][4] partials (w < 2: max, else sum) into out[w]
__global__ void __launch_bounds__(256) k_reduce_final4(const double* partial, int n, double* out) {
    double* sm = reinterpret_cast<double*>(MLV_SMEM_BASE());
    const int w = blockIdx.x;
    const int op = w < 2 ? RED_MAX : RED_SUM;
    double acc = red_identity(op);
    bool nan = false;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = partial[(size_t)i * 4 + w];
        nan = nan || (x != x);
        acc = red_combine(op, acc, x);
    }
    sm[threadIdx.x] = nan ? NAN : acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            const double x = sm[threadIdx.x], y = sm[threadIdx.x + s];
            sm[threadIdx.x] = (x != x || y != y) ? NAN : red_combine(op, x, y);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[w] = sm[0];
}

// contiguous block -> (peer-mapped) destination with coalesced 16-byte accesses, four in flight per thread
__global__ void __launch_bounds__(512) k_peer_copy(cplx* __restrict__ dst, const cplx* __restrict__ src, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const cplx a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

}  // namespace mlv
