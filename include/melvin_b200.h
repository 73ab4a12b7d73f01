/* melvin_b200.h -- C ABI of libmelvin_b200.so
 *
 * B200-native (sm_100a) implementation of the per-timestep pseudo-spectral hot
 * path of Melvin.py.  The reference has no FFI: its extension point is the `xp`
 * array module injected into every class plus the operator slots bound in the
 * constructors.  Each entry point below therefore cites the
 * reference *method* it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (MLV_ERR_*);
 *     mlv_last_error() returns a thread-local description of the last failure.
 *   - all array arguments are DEVICE pointers owned by the caller (the Python
 *     layer allocates them through torch); the library never frees or retains
 *     them beyond the stream work it enqueues.  The context owns only plans:
 *     twiddle tables, stencil-symbol tables, banded-solve factors and a small
 *     reduction scratch.
 *   - all work is enqueued on the context's stream (mlv_set_stream) and the
 *     calls return immediately; nothing synchronises.
 *   - complex128 = interleaved (re, im) doubles.  Layouts:
 *       S  spectral  (2nn+1, nm), rows n = 0..nn,-nn..-1   (melvin/ArrayFactory.py:8-45)
 *       Sf spectral, FDM-z mode (nn, nz)                   (melvin/Parameters.py:76-78)
 *       P  physical  (nx, nz) float64                      (melvin/ArrayFactory.py:47-62)
 *       I  private x-transformed intermediate (nx, ipitch) complex128
 *   - one host thread per context; distinct contexts are independent.
 */
#ifndef MELVIN_B200_H
#define MELVIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLV_OK 0
#define MLV_ERR_INVALID (-1)
#define MLV_ERR_UNSUPPORTED (-2)
#define MLV_ERR_CUDA (-3)
#define MLV_ERR_NOMEM (-4)

/* diagonal spectral operators g(n, m) applied to a spectral array */
#define MLV_OP_IDENT 0
#define MLV_OP_PSI 1    /* (-w)/lap, lap(0,0):=1    LaplacianSolver.py:58-68 + utility.py:65 */
#define MLV_OP_UX 2     /* -(i kz m) psi(w)          utility.py:71                            */
#define MLV_OP_UZ 3     /*  (i kx n) psi(w)          utility.py:78                            */
#define MLV_OP_DDX 4    /* SpatialDifferentiator.py:50  */
#define MLV_OP_DDZ 5    /* SpatialDifferentiator.py:55  */
#define MLV_OP_D2DX2 6  /* SpatialDifferentiator.py:60  */
#define MLV_OP_D2DZ2 7  /* SpatialDifferentiator.py:65  */
#define MLV_OP_LAP 8    /* Variable.py:111-113 (spectral snabla2) */
#define MLV_OP_INVLAP 9 /* LaplacianSolver.py:58-68 */
/* FDM-z mode only: operators that act along a spectral row (z is the finite-difference axis) */
#define MLV_OP_FDM_D2DZ2 10  /* pd2dz2, interior only         SpatialDifferentiator.py:106-128     */
#define MLV_OP_FDM_DDZ 11    /* pddz, edge columns zero       SpatialDifferentiator.py:91-104,157-185 */
#define MLV_OP_FDM_NABLA2 12 /* sd2dx2 + pd2dz2               Variable.py:111-113 (FDM snabla2)    */
#define MLV_OP_FDX_SYM 13    /* i * Fourier symbol of the periodic central x stencil   */
#define MLV_MAXLIN 6

/* multiplier symbols of the forward x pass */
#define MLV_SYM_ONE 0
#define MLV_SYM_FDX 1   /* Fourier symbol of the central x stencil, SpatialDifferentiator.py:76,130 */
#define MLV_SYM_FDZ 2   /* Fourier symbol of the central z stencil, SpatialDifferentiator.py:91,157 */

/* integrator schemes */
#define MLV_SCHEME_SEMI_IMPLICIT_LAP 0   /* Integrator.py:58-63, L = lcoef * lap symbol */
#define MLV_SCHEME_SEMI_IMPLICIT_ARR 1   /* Integrator.py:58-63, L = caller's real array */
#define MLV_SCHEME_EXPLICIT 2            /* Integrator.py:53-56 */

/* elementwise / reduction op codes (array-namespace surface) */
#define MLV_EW_ADD 0
#define MLV_EW_SUB 1
#define MLV_EW_MUL 2
#define MLV_EW_DIV 3
#define MLV_EW_COPY 4
#define MLV_EW_POW 5
#define MLV_KIND_REAL 0
#define MLV_KIND_CPLX 1
#define MLV_KIND_SCALAR 2
#define MLV_RED_SUM 0
#define MLV_RED_MAX 1
#define MLV_RED_MIN 2
#define MLV_RED_SUMSQ 3
#define MLV_RED_SUMPROD 4

typedef struct mlv_ctx mlv_ctx;

/* melvin/Parameters.py:63-92 + melvin/BasisFunctions.py:26-59 (derived on the host
 * by the Python layer so the constants are bit-identical to the reference's) */
typedef struct mlv_params {
    int32_t nx, nz;        /* power of two along every transformed axis: 16..16384 (fully spectral), nx 16..8192 (FDM-z) */
    int32_t fdm_z;         /* 0: ["spectral","spectral"]   1: ["spectral","fdm"] */
    int32_t fd_order;      /* spatial_derivative_order: 2 or 4 */
    double lx, lz;
    double kx0, kz0;       /* |i 2 pi / lx|, |i 2 pi / lz| */
    double d2x, d2z;       /* -(2 pi)^2/lx^2, -(2 pi)^2/lz^2 */
} mlv_params;

typedef struct mlv_info {
    int32_t nn, nm;                 /* truncation (nm = -1 in FDM-z mode) */
    int32_t spec_rows, spec_cols;   /* spectral_shape */
    int32_t ipitch;                 /* row pitch (complex elements) of I buffers */
    int32_t nm_local;               /* retained columns owned by this rank (= nm when unsharded) */
    int64_t ibytes;                 /* bytes of one field of an I / exchange buffer */
    int64_t red_doubles;            /* doubles of a caller-owned buffer of reduction partials */
} mlv_info;

typedef struct mlv_view {           /* 2-D strided view, strides in elements */
    void* ptr;
    int64_t row_stride, col_stride;
} mlv_view;

typedef struct mlv_lin_terms {      /* sum_i (cre_i + i cim_i) * op_i(src_i) */
    int32_t n;
    int32_t op[MLV_MAXLIN];
    const void* src[MLV_MAXLIN];
    double cre[MLV_MAXLIN], cim[MLV_MAXLIN];
} mlv_lin_terms;

typedef struct mlv_integ {          /* melvin/Integrator.py:5-18,53-63; TimeDerivative.py:9-45 */
    int32_t ab_order;               /* 2 | 4 */
    int32_t scheme;                 /* MLV_SCHEME_* */
    double dt, alpha, lcoef;
    const double* larr;             /* scheme ARR: real spectral-shaped linear operator */
    const void* q_in;               /* state (may equal q_out) */
    void* q_out;
    void* f0;                       /* history level curr_idx (read; rewritten if lin.n > 0) */
    int32_t f0_set;                 /* mlv_integrate: 1 = f0 := the linear terms (whole right-hand side, not read) */
    int32_t reserved_;
    const void* fm1;                /* curr_idx-1 .. curr_idx-3 (ring order) */
    const void* fm2;
    const void* fm3;
} mlv_integ;

typedef struct mlv_xfwd {           /* forward x pass + epilogue */
    int32_t nf;                     /* 1..4 input I buffers */
    int32_t mode;                   /* 0: dst = value   1: f0 = value + lin, then integrate */
    const void* src[4];
    int32_t sym[4];                 /* MLV_SYM_* */
    double coef[4];
    void* dst;                      /* mode 0 */
    mlv_lin_terms lin;              /* mode 1 */
    mlv_integ integ;                /* mode 1 */
} mlv_xfwd;

typedef struct mlv_ew {             /* out = a (op) b on (rows, cols) views */
    int32_t op, rows, cols;
    int32_t out_kind, a_kind, b_kind;
    mlv_view out, a, b;
    double a_re, a_im, b_re, b_im;
} mlv_ew;

typedef struct mlv_trig {           /* pruned DFT of period `period` along ONE axis by direct summation */
    int32_t inverse;                /* 0: samples -> modes (e^{-2 pi i jk/M}); 1: modes -> samples (e^{+...}) */
    int32_t ext;                    /* how the n_samp samples fill a period: MLV_EXT_* */
    int32_t period;                 /* M: n_samp (periodic) or 2(n_samp-1) (mirrored) */
    int32_t n_samp, n_modes;        /* entries stored along the axis on either side */
    int32_t two_sided;              /* modes k = 0..nn,-nn..-1 (n_modes = 2nn+1); else k = 0..n_modes-1 */
    int32_t hermitian;              /* inverse of a one-sided spectrum to REAL samples (numpy irfft); real scale only */
    int32_t samp_complex;           /* samples are complex128 (else float64) */
    int32_t batch_fastest;          /* adjacent threads take adjacent lines (unit batch stride) */
    int32_t nbatch;                 /* number of lines */
    int64_t samp_stride, samp_batch_stride;   /* elements along the axis / between lines */
    int64_t mode_stride, mode_batch_stride;
    const void* in;
    void* out;
    double scale_re, scale_im;      /* complex factor of the result */
    double w0;                      /* weight of mode 0 (cosine: forward 1/2, inverse 2; else 1) */
} mlv_trig;
#define MLV_EXT_PERIODIC 0
#define MLV_EXT_EVEN 1              /* x0..x_{n-1}, x_{n-2}..x_1     SpectralTransformer.py:169-172 */
#define MLV_EXT_ODD 2               /* x0..x_{n-2}, -x_{n-1}..-x_1   SpectralTransformer.py:174-178 */

/* ---- context ------------------------------------------------------------- */
int mlv_create(const mlv_params* p, mlv_ctx** ctx);
int mlv_destroy(mlv_ctx* ctx);
int mlv_set_stream(mlv_ctx* ctx, void* cuda_stream);
int mlv_get_info(const mlv_ctx* ctx, mlv_info* out);
/* bit 0: x passes run as two half-length transforms, bit 1: z stage runs on single real
 * rows (lines of 16384 points, which one SM cannot hold in registers; no reference
 * counterpart -- the reference calls numpy.fft, SpectralTransformer.py:132,191) */
int mlv_long_lines(const mlv_ctx* ctx);
/* Slab decomposition over `nranks` processes (one per GPU; fully spectral mode).  The
 * reference is single-device, so this has no counterpart there.  Afterwards
 * every call works on local slabs: spectral arrays are (2nn+1, nml) column slabs
 * (kz-slabs), the z stage sees nx/nranks rows (x-slabs).  `mlv_x_inverse` writes and
 * `mlv_advect_z` reads the inverse exchange buffer [peer][field][rows][nml];
 * `mlv_advect_z` writes and `mlv_x_forward` reads the forward exchange buffer
 * [peer][field][tiles][rows][CT]; the caller moves the peer blocks with one all-to-all
 * per direction (NCCL) between those calls.  inv_fields / fwd_fields = fields batched
 * per exchange (block strides).  mlv_get_info then reports the local shapes. */
int mlv_set_sharding(mlv_ctx* ctx, int rank, int nranks, int inv_fields, int fwd_fields);
/* Cuts the forward exchange buffers into row blocks of `rows_per_block` local rows (a power
 * of two dividing nx/nranks; default: one block per rank): receive side
 * [global row block][field][tile][rows][CT], send side [peer][local row block][...].  A caller
 * that runs mlv_advect_z_rows block by block can ship every finished block while the next
 * one is computed.  Reset by mlv_set_sharding. */
int mlv_set_forward_blocks(mlv_ctx* ctx, int rows_per_block);
/* Compute + collective in one kernel: receive buffers allocated with mlv_p2p_alloc are
 * exported to the peer processes (64-byte CUDA IPC handle), opened there with
 * mlv_p2p_open and registered with mlv_set_peer_buffers (which = 0: inverse exchange,
 * 1: forward exchange; bufs[h] = rank h's receive buffer, own entry = own buffer).
 * From then on mlv_x_inverse / mlv_advect_z store every peer's block straight into
 * that peer's receive buffer over NVLink (destination pointers passed to those calls
 * must point into the caller's own receive buffer); the caller only has to order the
 * producer and consumer kernels across ranks (a barrier), no all-to-all is needed. */
int mlv_p2p_alloc(mlv_ctx* ctx, int64_t bytes, void** ptr, void* handle64);
int mlv_p2p_open(mlv_ctx* ctx, const void* handle64, void** ptr);
int mlv_p2p_close(mlv_ctx* ctx, void* ptr, int opened);
int mlv_set_peer_buffers(mlv_ctx* ctx, int which, void* const* bufs);
/* Device-side ordering of the peer-store exchange: counters[h] = rank h's pair of 64-bit arrival
 * counters (zero-initialised peer-mapped memory, [0] inverse, [1] forward).  Producer CTAs bump
 * the counter of every rank after their stores (release.sys), consumer CTAs spin on their own
 * (acquire.sys) -- no collective and no host synchronisation between the kernels of a step.
 * All ranks must issue the same sequence of calls.  NULL switches it off. */
int mlv_set_peer_flags(mlv_ctx* ctx, void* const* counters);
/* Alternative for large grids: the kernels write their blocks locally and the caller moves
 * each contiguous peer block with an asynchronous device-to-device copy into the peer's
 * (IPC-mapped) receive buffer on `cuda_stream` -- copy engines drive NVLink at full width and
 * leave the SMs to the next x pass. */
int mlv_p2p_copy(mlv_ctx* ctx, void* dst, const void* src, int64_t bytes, void* cuda_stream);
const char* mlv_last_error(void);
int mlv_abi_version(void);
long long mlv_launch_count(void);   /* kernels launched by the library so far (process-wide) */

/* ---- transforms: SpectralTransformer.to_physical / to_spectral ------------
 * (melvin/SpectralTransformer.py:33-88 1-D, :90-199 2-D, COMPLEX_EXP bases) */
int mlv_to_physical(mlv_ctx* ctx, const void* spec, void* iscratch, double* phys);
int mlv_to_spectral(mlv_ctx* ctx, const double* phys, void* iscratch, void* spec);

/* the two passes of the 2-D transforms, exposed so that callers can fuse around
 * the private intermediate (fully spectral mode only) */
int mlv_x_inverse(mlv_ctx* ctx, int nf, const void* const* spec, const int32_t* op,
                  void* const* idst);                       /* S -> I with prologue op  */
int mlv_z_inverse(mlv_ctx* ctx, const void* isrc, double* phys);   /* I -> P             */
int mlv_z_forward(mlv_ctx* ctx, const double* phys, void* idst);   /* P -> I             */
int mlv_x_forward(mlv_ctx* ctx, const mlv_xfwd* d);                /* I -> S + epilogue  */

/* ---- nonlinear term: Variable.vec_dot_nabla (melvin/Variable.py:119-128) ---
 * fused physical-space stage on x-transformed operands:
 *   ia = Fz[ux*q], ib = Fz[uz*q]; red4 (device, 4 doubles) = max ux, max uz,
 *   sum ux^2, sum uz^2 (Integrator.py:35-44, utility.py:42-59). */
int mlv_advect_z(mlv_ctx* ctx, const void* iux, const void* iuz, const void* iq,
                 void* ia, void* ib, double* red4);
/* same on the local rows [row0, row0 + nrows) only (both even); red4, if given, reduces the
 * partials of all rows below row0 + nrows (pass it with the last range) */
int mlv_advect_z_rows(mlv_ctx* ctx, const void* iux, const void* iuz, const void* iq,
                      void* ia, void* ib, int row0, int nrows, double* red4);
/* The four reductions are needed at ticker cadence only (Simulation.end_loop: CFL every
 * cfl_cadence loops, trackers every tracker_cadence).  With red4 == NULL the fused stage leaves
 * its per-CTA partials behind -- in the context's scratch, or in a caller-owned buffer of
 * mlv_info.red_doubles doubles registered with mlv_set_reduction_partials (NULL = scratch again)
 * -- and mlv_reduce_partials combines them into red4 whenever somebody asks. */
int mlv_set_reduction_partials(mlv_ctx* ctx, double* partials);
/* on = 0: the fused z stage skips its CFL / kinetic-energy partials (Integrator.py:35-44, utility.py:42-59
 * are ticker-cadence work: melvin/simulation.py asks for them only for steps a ticker will read);
 * mlv_reduce_partials then fails until a launch with reductions on has run.  Default: on. */
int mlv_set_reductions(mlv_ctx* ctx, int on);
int mlv_reduce_partials(mlv_ctx* ctx, const double* partials, double* red4);
/* materialised physical operands: out = pddx(ux*q) + pddz(uz*q) */
int mlv_advect_phys(mlv_ctx* ctx, const double* ux, const double* uz, const double* q,
                    double* out);

/* ---- COSINE / SINE bases (melvin/SpectralTransformer.py:108-125,134-146,169-196): the mirrored
 * rfft2 / irfft2 of the reference, one axis per call, only the retained modes evaluated. */
int mlv_trig_axis(mlv_ctx* ctx, const mlv_trig* d);

/* ---- SpatialDifferentiator ------------------------------------------------
 * spectral: sddx/sddz/sd2dx2/sd2dz2/calc_lap (:50-74) through mlv_spec_lincomb;
 * physical: pddx/pddz/pd2dz2 (:76-185) through mlv_stencil. */
int mlv_spec_lincomb(mlv_ctx* ctx, const mlv_lin_terms* t, void* out);
int mlv_lap_array(mlv_ctx* ctx, double coef, double* out);
int mlv_stencil(mlv_ctx* ctx, const void* in, void* out, int rows, int cols, int ncomp,
                int axis, int order, int periodic, int second, double h);

/* ---- LaplacianSolver.solve (melvin/LaplacianSolver.py:58-79) --------------
 * fully spectral: mlv_spec_lincomb with MLV_OP_INVLAP; FDM-z: batched tridiagonal */
int mlv_solve_fdm(mlv_ctx* ctx, const void* rhs, void* out);
/* 4th-order variant (EXTENSION, no reference code path: BASELINE names the pentadiagonal case, the
 * reference's FDM Laplacian is always the tridiagonal one): the 4th-order central second difference of
 * SpatialDifferentiator.py:121-130 on rows 2..nz-3, the 2nd-order one on rows 1 and nz-2, identity
 * boundary rows as in LaplacianSolver.py:46-51.  Parity unpinned; own accuracy / convergence tests. */
int mlv_solve_fdm_o4(mlv_ctx* ctx, const void* rhs, void* out);

/* ---- fused Fourier-x / FDM-z step (examples/rayleigh_benard_convection.py:95-145) ---------
 * mlv_fdm_velocity: utility.calc_velocity_from_vorticity, FDM branch (utility.py:62-79) in one
 *   row-wise pass: psi = solve(-w); uxh = -pddz(psi) and uzh = (i kx n) psi are the x spectra of
 *   the two velocity components (the z stencil commutes with the x transform).  (nn, nz) each.
 * mlv_fdm_advect: the physical-space stage of Variable.vec_dot_nabla (Variable.py:119-128) on
 *   x spectra: a = Fx[ux q]/nx, b = Fx[uz q]/nx; the caller forms
 *   d/dx(ux q) + d/dz(uz q) = MLV_OP_FDX_SYM(a) + MLV_OP_FDM_DDZ(b) as linear terms of
 *   mlv_integrate.  Reductions as mlv_advect_z (red4 / mlv_set_reduction_partials). */
int mlv_fdm_velocity(mlv_ctx* ctx, const void* w, void* psi, void* uxh, void* uzh);
int mlv_fdm_advect(mlv_ctx* ctx, const void* uxh, const void* uzh, const void* q, void* a, void* b,
                   double* red4);

/* ---- Integrator.integrate (melvin/Integrator.py:53-63) -------------------- */
int mlv_integrate(mlv_ctx* ctx, const mlv_lin_terms* extra, const mlv_integ* g);

/* ---- array-namespace surface (xp.max / xp.sum / xp.mean, arithmetic, slicing) */
int mlv_elementwise(mlv_ctx* ctx, const mlv_ew* d);
int mlv_reduce(mlv_ctx* ctx, int op, int rows, int cols, const mlv_view* a,
               const mlv_view* b, double* out_dev);

#ifdef __cplusplus
}
#endif
#endif /* MELVIN_B200_H */
