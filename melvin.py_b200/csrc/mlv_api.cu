// melvin-b200: C ABI (include/melvin_b200.h) -- context, plan tables, launch logic.
#include "../../include/melvin_b200.h"

#include "mlv_kernels_pw.cuh"
#include "mlv_kernels_split.cuh"
#include "mlv_kernels_trig.cuh"
#include "mlv_rt.h"

#include <stdarg.h>
#include <new>
#include <utility>
#include <vector>

namespace mlv {

static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;     // kernels launched by this library (process-wide)

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------ context
struct Plan {            // twiddle tables of one line length
    int log2n = 0;
    cplx* dev = nullptr; // one allocation, tables back to back
    FftTw tw{};
};

}  // namespace mlv

struct mlv_ctx {
    mlv_params p;
    int nn, nm, spec_rows, spec_cols, ipitch, ct;
    int log2nx, log2nz;
    // lines too long for the register transform (mlv_kernels_split.cuh): x passes in two
    // half-length transforms, z stage on single real rows; plan lengths are log2 - 1 then
    bool xsplit = false, zreal = false;
    int planlx = 0, planlz = 0;
    mlv::cplx* tws_x = nullptr;   // e^{-2 pi i k/nx}, k < nx/2
    mlv::cplx* tws_z = nullptr;   // e^{-2 pi i k/nz}, k < nz/2
    double dx, dz;
    mlv::SpecConsts k;
    mlv::Plan planx, planz;
    double* symz = nullptr;     // [nm]
    double* tri_inv = nullptr;  // FDM-z: (nn, nz) pivots of the Thomas factorisation
    double* penta = nullptr;    // FDM-z, 4th-order solve: 5 x (nn, nz) LU coefficients (built on first use)
    double* symx = nullptr;     // FDM-z: [nn] Fourier symbol of the periodic central x stencil
    double* red = nullptr;      // reduction partials
    size_t red_cap = 0;
    double* red_user = nullptr; // caller-owned partials of the fused z stage (mlv_set_reduction_partials)
    int red_count = 0;          // per-CTA partials written by the latest fused z stage over all local rows
    bool red_on = true;         // mlv_set_reductions: the fused z stage computes its CFL / energy partials
    // slab decomposition (mlv_set_sharding); defaults describe the unsharded case
    int rank = 0, nranks = 1;
    int nml = 0;                // local column pitch of spectral arrays / inverse buffers
    int nm_loc = 0;             // valid local columns
    int nxl = 0;                // local rows
    int inv_fields = 1, fwd_fields = 1;
    int fwd_rows = 0;           // rows per block of the forward exchange buffers (divides nxl)
    mlv::Shard sh{};
    // peer-mapped receive buffers (mlv_set_peer_buffers); null = exchange by all-to-all
    mlv::cplx* peer_inv[MLV_MAXPEER] = {};
    mlv::cplx* peer_fwd[MLV_MAXPEER] = {};
    bool p2p_inv = false, p2p_fwd = false;
    // device-side ordering of the peer-store exchange (mlv_set_peer_flags): per rank two arrival
    // counters [0] inverse blocks, [1] forward blocks; expect[] = arrivals due at this rank so far
    unsigned long long* peer_flags[MLV_MAXPEER] = {};
    bool flags_on = false;
    unsigned long long* expect = nullptr;       // device: arrivals due at this rank so far, [0] inverse, [1] forward
    // COSINE / SINE bases: roots of unity e^{-2 pi i j/M} per period M (mlv_trig_axis)
    std::vector<std::pair<int, mlv::cplx*>> trig_tables;
    mlv::stream_t stream = 0;
};

namespace mlv {

static int ilog2_exact(int n) {
    int l = 0;
    while ((1 << l) < n) ++l;
    return ((1 << l) == n) ? l : -1;
}

// table of one pass (radix r, remaining length n3): plane b < log2(r) holds exp(-2 pi i n 2^b / (r n3)), n < n3
static size_t append_table(std::vector<cplx>& host, int r, int n3) {
    const size_t off = host.size();
    const long double ncur = (long double)r * n3;
    int lr = 0;
    while ((1 << lr) < r) ++lr;
    for (int b = 0; b < lr; ++b)
        for (int n = 0; n < n3; ++n) {
            // argument reduced exactly in integers
            const long long e = ((long long)n << b) % (long long)ncur;
            const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)e / ncur;
            host.push_back(mk((double)cosl(ang), (double)sinl(ang)));
        }
    return off;
}

template <int LOG2N>
static void build_tables(std::vector<cplx>& host, size_t (&offs)[MLV_MAX_PASS + 2]) {
    typedef FftCfg<LOG2N> F;
    for (int p = 0; p < MLV_MAX_PASS + 2; ++p) offs[p] = (size_t)-1;
    for (int p = 0; p < F::NPASS; ++p) {
        const int n3 = F::n3(p), r = F::radix(p);
        if (n3 <= 1) continue;
        offs[p] = append_table(host, r, n3);
    }
    if (F::NPASS == 3) {        // mirrored pass order 16 -> 16 -> R0 (fft_grp2nat)
        offs[MLV_MAX_PASS] = append_table(host, 16, F::T);
        offs[MLV_MAX_PASS + 1] = append_table(host, 16, F::R0);
    }
}

static int make_plan(Plan& plan, int log2n, stream_t s) {
    std::vector<cplx> host;
    size_t offs[MLV_MAX_PASS + 2];
    switch (log2n) {
#define MLV_CASE(L) case L: build_tables<L>(host, offs); break;
        MLV_CASE(4) MLV_CASE(5) MLV_CASE(6) MLV_CASE(7) MLV_CASE(8) MLV_CASE(9)
        MLV_CASE(10) MLV_CASE(11) MLV_CASE(12) MLV_CASE(13)
#undef MLV_CASE
        default:
            set_error("unsupported transform length 2^%d (supported: 16..8192)", log2n);
            return MLV_ERR_UNSUPPORTED;
    }
    plan.log2n = log2n;
    int rc = rt_malloc((void**)&plan.dev, (host.size() + 1) * sizeof(cplx));
    if (rc) return rc;
    if (!host.empty()) {
        rc = rt_h2d(plan.dev, host.data(), host.size() * sizeof(cplx), s);
        if (rc) return rc;
    }
    for (int p = 0; p < MLV_MAX_PASS; ++p)
        plan.tw.p[p] = (offs[p] == (size_t)-1) ? nullptr : plan.dev + offs[p];
    for (int p = 0; p < 2; ++p)
        plan.tw.g[p] = (offs[MLV_MAX_PASS + p] == (size_t)-1) ? nullptr : plan.dev + offs[MLV_MAX_PASS + p];
    return 0;
}

// e^{-2 pi i k/nfull}, k < nfull/2: the twiddles of the radix-2 step around the register transform
static int make_split_table(cplx** dev, int nfull, stream_t s) {
    std::vector<cplx> host((size_t)nfull / 2);
    for (int k = 0; k < nfull / 2; ++k) {
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)nfull;
        host[k] = mk((double)cosl(ang), (double)sinl(ang));
    }
    int rc = rt_malloc((void**)dev, host.size() * sizeof(cplx));
    if (!rc) rc = rt_h2d(*dev, host.data(), host.size() * sizeof(cplx), s);
    return rc;
}

// imaginary part of the Fourier symbol of the reference's central stencils
// (SpatialDifferentiator.py:76-104 order 2, :130-185 order 4); SURVEY F2.
static double stencil_symbol(int order, long long mode, long long npts, double h) {
    const long double th = 2.0L * 3.14159265358979323846264338327950288L * (long double)mode / (long double)npts;
    if (order == 2) return (double)(sinl(th) / (long double)h);
    return (double)((8.0L * sinl(th) - sinl(2.0L * th)) / (6.0L * (long double)h));
}

// per-size launch geometry
static constexpr int xcols(int log2n) {          // adjacent columns per x-pass CTA
    // measured at 4096: C = 1 (two 256-thread CTAs per SM) is 1.2-1.7x slower than C = 2
    return log2n >= 13 ? 1 : ((512 >> (log2n - 4)) > 32 ? 32 : (512 >> (log2n - 4)));
}
static constexpr int zlines(int log2n) {         // row pairs per z-pass CTA
    return log2n >= 12 ? 1 : ((256 >> (log2n - 4)) > 16 ? 16 : (256 >> (log2n - 4)));
}

#define MLV_SWITCH_LOG2(L, MACRO)                                                         \
    switch (L) {                                                                          \
        case 4: MACRO(4); break;                                                          \
        case 5: MACRO(5); break;                                                          \
        case 6: MACRO(6); break;                                                          \
        case 7: MACRO(7); break;                                                          \
        case 8: MACRO(8); break;                                                          \
        case 9: MACRO(9); break;                                                          \
        case 10: MACRO(10); break;                                                        \
        case 11: MACRO(11); break;                                                        \
        case 12: MACRO(12); break;                                                        \
        case 13: MACRO(13); break;                                                        \
        default: set_error("unsupported transform length 2^%d", (L)); return MLV_ERR_UNSUPPORTED; \
    }

// the long-line kernels are instantiated for the production lengths (8192 = 2 x 4096,
// 16384 = 2 x 8192) and two small lengths that the tests force through them (MLV_FORCE_SPLIT)
#define MLV_SWITCH_SPLIT(L, MACRO)                                                        \
    switch (L) {                                                                          \
        case 6: MACRO(6); break;                                                          \
        case 7: MACRO(7); break;                                                          \
        case 12: MACRO(12); break;                                                        \
        case 13: MACRO(13); break;                                                        \
        default: set_error("split transforms: unsupported half length 2^%d", (L)); return MLV_ERR_UNSUPPORTED; \
    }

static unsigned grid1d(size_t total, unsigned block = 256) {
    size_t g = (total + block - 1) / block;
    const size_t cap = 148 * 16;                 // grid-stride: a few waves of 148 SMs
    if (g > cap) g = cap;
    if (g == 0) g = 1;
    return (unsigned)g;
}

static int ensure_red(mlv_ctx* c, size_t n) {
    if (n <= c->red_cap) return 0;
    if (c->red) rt_free(c->red);
    c->red = nullptr;
    c->red_cap = 0;
    int rc = rt_malloc((void**)&c->red, n * sizeof(double));
    if (rc) return rc;
    c->red_cap = n;
    return 0;
}

// Partition: column tiles of width CT are dealt out in equal contiguous runs of tpr tiles,
// rows in equal runs of nx/G.
static void configure_shard(mlv_ctx* c, int rank, int nranks, int inv_fields, int fwd_fields) {
    const int ct = c->ct;
    const int ntile = (c->nm + ct - 1) / ct;
    const int tpr = (ntile + nranks - 1) / nranks;
    c->rank = rank; c->nranks = nranks;
    c->inv_fields = inv_fields; c->fwd_fields = fwd_fields;
    c->nxl = c->p.nx / nranks;
    if (nranks == 1) {
        c->nml = c->ipitch;
        c->nm_loc = c->nm;
    } else {
        c->nml = tpr * ct;
        int left = c->nm - rank * c->nml;
        c->nm_loc = left < 0 ? 0 : (left > c->nml ? c->nml : left);
    }
    int shift = 0;
    while ((1 << shift) < c->nxl) ++shift;
    c->sh.m_off = rank * (nranks == 1 ? 0 : c->nml);
    c->sh.nm_glob = c->nm;
    c->sh.nml = c->nml;
    c->sh.rpc_shift = shift;
    c->sh.tpr = tpr;
    c->sh.inv_chunk = nranks == 1 ? 0 : (long long)inv_fields * c->nxl * c->nml;
    if (c->fwd_rows <= 0 || c->fwd_rows > c->nxl || nranks == 1) c->fwd_rows = c->nxl;
    int fshift = 0;
    while ((1 << fshift) < c->fwd_rows) ++fshift;
    c->sh.fwd_rshift = fshift;
    c->sh.fwd_chunk = nranks == 1 ? 0 : (long long)fwd_fields * tpr * c->fwd_rows * ct;
    c->sh.fwd_peer = c->sh.fwd_chunk * (c->nxl / c->fwd_rows);
    c->spec_cols = c->p.fdm_z ? c->p.nz : (nranks == 1 ? c->nm : c->nml);
}

static int need_2d(const mlv_ctx* c, const char* fn) {
    if (c->p.fdm_z) {
        set_error("%s: only available in fully spectral mode", fn);
        return MLV_ERR_INVALID;
    }
    return 0;
}

static void fill_lin(LinTerms& o, const mlv_lin_terms* t) {
    o.n = 0;
    if (!t) return;
    o.n = t->n;
    for (int i = 0; i < t->n && i < MLV_MAXLIN; ++i) {
        o.src[i] = (const cplx*)t->src[i];
        o.op[i] = t->op[i];
        o.cre[i] = t->cre[i];
        o.cim[i] = t->cim[i];
    }
}

static int check_lin(const mlv_ctx* c, const mlv_lin_terms* t, const char* fn) {
    if (!t) return 0;
    if (t->n < 0 || t->n > MLV_MAXLIN) {
        set_error("%s: at most %d linear terms", fn, MLV_MAXLIN);
        return MLV_ERR_INVALID;
    }
    for (int i = 0; i < t->n; ++i) {
        const int op = t->op[i];
        if (op < MLV_OP_IDENT || op > MLV_OP_FDX_SYM) {
            set_error("%s: bad operator code %d", fn, op);
            return MLV_ERR_INVALID;
        }
        if (c->p.fdm_z && !(op == MLV_OP_IDENT || op == MLV_OP_DDX || op == MLV_OP_D2DX2 || op >= MLV_OP_FDM_D2DZ2)) {
            set_error("%s: operator %d needs a spectral z axis", fn, op);
            return MLV_ERR_INVALID;
        }
        if (!c->p.fdm_z && op >= MLV_OP_FDM_D2DZ2) {
            set_error("%s: operator %d needs a finite-difference z axis", fn, op);
            return MLV_ERR_INVALID;
        }
        if (!t->src[i]) { set_error("%s: null source", fn); return MLV_ERR_INVALID; }
    }
    return 0;
}

static FdmConsts fdm_consts(const mlv_ctx* c) {
    FdmConsts f;
    f.nz = c->p.nz; f.order = c->p.fd_order; f.dz = c->dz; f.symx = c->symx;
    return f;
}

static void fill_integ(IntegArgs& o, const mlv_integ* g) {
    o.ab_order = g->ab_order;
    o.scheme = g->scheme;
    o.dt = g->dt;
    o.alpha = g->alpha;
    o.lcoef = g->lcoef;
    o.larr = g->larr;
    o.q_in = (const cplx*)g->q_in;
    o.q_out = (cplx*)g->q_out;
    o.f0 = (cplx*)g->f0;
    o.fm1 = (const cplx*)g->fm1;
    o.fm2 = (const cplx*)g->fm2;
    o.fm3 = (const cplx*)g->fm3;
}

static int check_integ(const mlv_ctx* c, const mlv_integ* g, const char* fn) {
    if (!g || !g->q_in || !g->q_out || !g->f0) { set_error("%s: null pointer", fn); return MLV_ERR_INVALID; }
    if (g->ab_order != 2 && g->ab_order != 4) { set_error("%s: integrator_order must be 2 or 4", fn); return MLV_ERR_INVALID; }
    if (!g->fm1 || (g->ab_order == 4 && (!g->fm2 || !g->fm3))) { set_error("%s: missing history level", fn); return MLV_ERR_INVALID; }
    if (g->scheme < 0 || g->scheme > 2) { set_error("%s: bad scheme", fn); return MLV_ERR_INVALID; }
    if (g->scheme == MLV_SCHEME_SEMI_IMPLICIT_ARR && !g->larr) { set_error("%s: null linear operator", fn); return MLV_ERR_INVALID; }
    if (g->scheme == MLV_SCHEME_SEMI_IMPLICIT_LAP && c->p.fdm_z) { set_error("%s: symbolic Laplacian needs fully spectral mode", fn); return MLV_ERR_INVALID; }
    return 0;
}

// ----------------------------------------------------------- launch helpers
// 4096-point lines, unsharded, tensor maps available: column-serial persistent kernel, two CTAs per SM
template <int L>
static int launch_xinv_cols(mlv_ctx* c, XInvArgs& a, bool& done) {
    typedef FftCfg<L> F;
    done = false;
    if constexpr (L == 12) {
        // measured slower than the two-column kernel (0.215 vs 0.168 ms at 4096^2): 16-byte-wide tensor
        // boxes move one row per ~3 cycles whatever their width, so a column-at-a-time kernel is
        // bound by the copy engine (profiles/r02_experiments.md); kept as a selectable variant
        if (c->nranks != 1 || !rt_tma_enabled() || !rt_env_flag("MLV_XINV_COLS")) return 0;
        const int rows = 2 * a.nn + 1;
        const int h = rows < 256 ? rows : 256, nb = (rows + h - 1) / h;
        const size_t smem = ((size_t)F::N + (size_t)h * nb) * sizeof(cplx) + 16;
        if (smem > 113 * 1024) return 0;
        a.load_tma = 1; a.ld_rows = h; a.ld_boxes = nb;
        a.use_tma = 1; a.tma_rows = 256;
        for (int f = 0; f < a.nf; ++f) {
            if (!rt_make_tmap(&a.smap[f], const_cast<cplx*>(a.src[f]), 2ull * a.spitch, (unsigned long long)rows,
                              16ull * a.spitch, 2u, (unsigned)h)) return 0;
            if (!rt_make_tmap(&a.tmap[f], a.dst[f], 2ull * a.ipitch, (unsigned long long)F::N,
                              16ull * a.ipitch, 2u, 256u)) return 0;
        }
        unsigned grid = (unsigned)a.nm < 296u ? (unsigned)a.nm : 296u;
        if (const char* g = getenv("MLV_XINV_GRID")) if (atoi(g) > 0 && (unsigned)atoi(g) < grid) grid = (unsigned)atoi(g);
        auto kfn = k_xinv_cols<L>;
        MLV_LAUNCH(kfn, grid, (unsigned)F::T, smem, c->stream, a);
        done = true;
    }
    return 0;
}

template <int L>
static int launch_xinv(mlv_ctx* c, XInvArgs& a) {
    {
        bool done = false;
        if (int rc = launch_xinv_cols<L>(c, a, done)) return rc;
        if (done) return 0;
    }
    constexpr int C = xcols(L);
    typedef FftCfg<L> F;
    auto kfn = k_xinv<L, C>;
    const int rows = 2 * a.nn + 1;
    size_t smem = ((size_t)F::XSLOTS + (size_t)rows) * C * sizeof(cplx) + 16;     // + mbarrier
    const unsigned grid = (unsigned)((a.nm + C - 1) / C);
    a.use_tma = 0;
    a.load_tma = 0; a.ld_rows = rows; a.ld_boxes = 1;
    if (rt_tma_enabled()) {
        // source tiles through tensor loads: boxes of ld_rows rows; the stash is rounded up to whole
        // boxes, so pick the largest box height whose padding still fits the shared memory
        for (int lr = 256; lr >= 16 && !a.load_tma; lr >>= 1) {
            const int h = rows < lr ? rows : lr, nb = (rows + h - 1) / h;
            const size_t need = ((size_t)F::XSLOTS + (size_t)h * nb) * C * sizeof(cplx) + 16;
            if (need > rt_max_smem()) continue;
            if (smem <= 113 * 1024 && need > 113 * 1024) continue;     // keep two CTAs per SM
            a.load_tma = 1; a.ld_rows = h; a.ld_boxes = nb;
            for (int f = 0; f < a.nf && a.load_tma; ++f)
                if (!rt_make_tmap(&a.smap[f], const_cast<cplx*>(a.src[f]), 2ull * a.spitch, (unsigned long long)rows,
                                  16ull * a.spitch, 2u * C, (unsigned)h))
                    a.load_tma = 0;
            if (a.load_tma) smem = need;
            else { a.ld_rows = rows; a.ld_boxes = 1; break; }
        }
    }
    a.wave = 148 * (smem > 113 * 1024 ? 1 : 2);
    unsigned launch_grid = grid;
    // tensor loads on: resident CTAs only, each walks over tiles and requests its next source tile early
    if (a.load_tma && !rt_env_flag("MLV_XINV_ONESHOT") && grid > (unsigned)a.wave) launch_grid = (unsigned)a.wave;
    if (a.load_tma) if (const char* g = getenv("MLV_XINV_GRID")) if (atoi(g) > 0 && (unsigned)atoi(g) < launch_grid) launch_grid = (unsigned)atoi(g);
    if (c->nranks == 1 && rt_tma_enabled()) {
        // column tiles leave through tensor stores: one (rows x 2C doubles) box per 256 rows
        a.tma_rows = F::N < 256 ? F::N : 256;
        a.use_tma = 1;
        for (int f = 0; f < a.nf && a.use_tma; ++f)
            if (!rt_make_tmap(&a.tmap[f], a.dst[f], 2ull * a.ipitch, (unsigned long long)F::N,
                              16ull * a.ipitch, 2u * C, (unsigned)a.tma_rows))
                a.use_tma = 0;
    }
    MLV_LAUNCH(kfn, launch_grid, (unsigned)(C * F::T), smem, c->stream, a);
    return 0;
}

template <int L>
static int launch_xinv_split(mlv_ctx* c, XInvArgs& a) {
    constexpr int C = xcols(L);
    typedef FftCfg<L> F;
    auto kfn = k_xinv_split<L, C>;
    const size_t smem = (size_t)F::XSLOTS * C * sizeof(cplx);
    const unsigned grid = (unsigned)((a.nm + C - 1) / C);
    a.wave = 148;
    MLV_LAUNCH(kfn, grid, (unsigned)(C * F::T), smem, c->stream, a);
    return 0;
}

// tensor maps of the epilogue's state / history column tiles (L2 prefetch by box instead of by row)
template <int C>
static void xfwd_prefetch_maps(XFwdArgs& a, int nthreads) {
    a.pf_tma = 0;
    if (a.mode != 1 || !rt_tma_enabled() || rt_env_flag("MLV_XFWD_LINEPF")) return;
    const int rows = 2 * a.nn + 1;
    a.pf_rows = rows < 256 ? rows : 256;
    a.pf_boxes = (rows + a.pf_rows - 1) / a.pf_rows;
    if (2 * a.pf_boxes > nthreads) return;                     // one thread per box
    a.pf_tma = rt_make_tmap(&a.qmap, const_cast<cplx*>(a.integ.q_in), 2ull * a.spitch, (unsigned long long)rows,
                            16ull * a.spitch, 2u * C, (unsigned)a.pf_rows) &&
               rt_make_tmap(&a.fmap, const_cast<cplx*>(a.integ.fm1), 2ull * a.spitch, (unsigned long long)rows,
                            16ull * a.spitch, 2u * C, (unsigned)a.pf_rows) ? 1 : 0;
}

template <int L>
static int launch_xfwd_split(mlv_ctx* c, XFwdArgs& a) {
    constexpr int C = xcols(L);
    typedef FftCfg<L> F;
    auto kfn = k_xfwd<L, C, 2>;
    a.stage = 0;
    a.pf_tma = 0;            // (per-row hints: tensor prefetch measured 2-3 % slower on the long-line kernels)
    const size_t smem = (size_t)F::XSLOTS * C * sizeof(cplx) + 16;
    const unsigned grid = (unsigned)((a.nm + C - 1) / C);
    a.wave = 148 * (smem > 113 * 1024 ? 1 : 2);
    MLV_LAUNCH(kfn, grid, (unsigned)(C * F::T), smem, c->stream, a);
    return 0;
}

template <int L>
static int launch_zreal(mlv_ctx* c, ZRealArgs& a, bool inverse) {
    typedef FftCfg<L> F;
    const size_t smem = (size_t)F::XSLOTS * sizeof(double);
    if (inverse) {
        auto kfn = k_zr_c2r<L>;
        MLV_LAUNCH(kfn, (unsigned)a.nx, (unsigned)F::T, smem, c->stream, a);
    } else {
        auto kfn = k_zr_r2c<L>;
        MLV_LAUNCH(kfn, (unsigned)a.nx, (unsigned)F::T, smem, c->stream, a);
    }
    return 0;
}

template <int L>
static int launch_zadv_real(mlv_ctx* c, ZAdvArgs& a, unsigned& grid_out) {
    typedef FftCfg<L> F;
    const size_t smem = F::XSLOTS * sizeof(double) + (size_t)F::N * sizeof(cplx) + (size_t)4 * F::T * sizeof(double);
    const unsigned grid = (unsigned)a.nrows;
    grid_out = grid;
    int rc = ensure_red(c, (size_t)(a.nx / a.nrows) * grid * 4);
    if (rc) return rc;
    a.red = c->red_on ? (c->red_user ? c->red_user : c->red) + (size_t)(a.row0 / a.nrows) * grid * 4 : nullptr;
    a.wave = 148;
    const bool single = c->nranks == 1 && a.sh.fwd_chunk == 0 && a.sh.inv_chunk == 0 && (1 << a.sh.fwd_rshift) == a.nx &&
                        !rt_env_flag("MLV_ZADV_GENERIC");
#define MLV_ZR_GO(RED_, SH_)                                                             \
    do {                                                                                  \
        auto kfn = k_zr_advect<L, RED_, SH_>;                                             \
        MLV_LAUNCH(kfn, grid, (unsigned)F::T, smem, c->stream, a);                        \
    } while (0)
    if (a.red) { if (single) MLV_ZR_GO(true, false); else MLV_ZR_GO(true, true); }
    else { if (single) MLV_ZR_GO(false, false); else MLV_ZR_GO(false, true); }
#undef MLV_ZR_GO
    return 0;
}

template <int L>
static int launch_xfwd(mlv_ctx* c, XFwdArgs& a) {
    constexpr int C = xcols(L);
    typedef FftCfg<L> F;
    auto kfn = k_xfwd<L, C, 1>;
    const size_t smem = (size_t)F::XSLOTS * C * sizeof(cplx) + 16;       // + mbarrier
    const unsigned grid = (unsigned)((a.nm + C - 1) / C);
    a.wave = 148 * (smem > 113 * 1024 ? 1 : 2);
    a.stage = 0;
    // operand pair of an advected scalar (d/dx by the order-2 stencil, d/dz by its symbol):
    // stage the stencil operand in shared memory with bulk copies
    if (rt_tma_enabled() && a.nf >= 2 && a.sym[0] == XSYM_FDX && a.sym[1] == XSYM_FDZ && a.order == 2 &&
        ((1u << a.sh.fwd_rshift) * C * sizeof(cplx)) % 16 == 0)
        a.stage = 1;
    // the epilogue's state / history column tiles: 22 tensor prefetches per CTA instead of 5462 per-row hints
    // at 4096^2 (which filled the load/store queue: lg_throttle 4.5 warps per issue), 0.1286 -> 0.1161 ms
    a.pf_tma = 0;
    // the single-scalar step on one GPU (KH / TG loops): specialised kernel
    if (a.stage && c->nranks == 1 && (1 << a.sh.fwd_rshift) == F::N && a.sh.fwd_chunk == 0 && a.nf == 2 && a.mode == 1 &&
        a.integ.ab_order == 2 && a.integ.scheme == 0 && a.lin.n == 0 && !rt_env_flag("MLV_XFWD_GENERIC")) {
        xfwd_prefetch_maps<C>(a, C * F::T);
#define MLV_XFWD_GO(UN_)                                                                  \
        do {                                                                              \
            auto kh = k_xfwd_scalar<L, C, UN_>;                                           \
            unsigned g_ = grid;                                                           \
            if (rt_env_flag("MLV_XFWD_PERSISTENT") && g_ > (unsigned)a.wave) g_ = (unsigned)a.wave; \
            MLV_LAUNCH(kh, g_, (unsigned)(C * F::T), smem, c->stream, a);                 \
        } while (0)
        MLV_XFWD_GO(4);        // 2, 3, 4, 6 outputs per trip measure the same (0.1254 .. 0.1259 ms)
#undef MLV_XFWD_GO
        return 0;
    }
    MLV_LAUNCH(kfn, grid, (unsigned)(C * F::T), smem, c->stream, a);
    return 0;
}

template <int L>
static int launch_zc2r(mlv_ctx* c, ZArgs& a) {
    constexpr int LPC = zlines(L);
    typedef FftCfg<L> F;
    auto kfn = k_z_c2r<L, LPC>;
    const size_t smem = (size_t)F::XSLOTS * LPC * sizeof(double);
    const unsigned grid = (unsigned)((a.nx / 2 + LPC - 1) / LPC);
    MLV_LAUNCH(kfn, grid, (unsigned)(LPC * F::T), smem, c->stream, a);
    return 0;
}

template <int L>
static int launch_zr2c(mlv_ctx* c, ZArgs& a) {
    constexpr int LPC = zlines(L);
    typedef FftCfg<L> F;
    auto kfn = k_z_r2c<L, LPC>;
    const size_t smem = (size_t)F::XSLOTS * LPC * sizeof(double);
    const unsigned grid = (unsigned)((a.nx / 2 + LPC - 1) / LPC);
    MLV_LAUNCH(kfn, grid, (unsigned)(LPC * F::T), smem, c->stream, a);
    return 0;
}

template <int L>
static int launch_zadv(mlv_ctx* c, ZAdvArgs& a, unsigned& grid_out) {
    constexpr int LPC = zlines(L);
    typedef FftCfg<L> F;
    const size_t smem = (size_t)LPC * (F::XSLOTS * sizeof(double) + (size_t)F::N * sizeof(cplx)) +
                        (size_t)4 * LPC * F::T * sizeof(double);
    unsigned grid = (unsigned)((a.nrows / 2 + LPC - 1) / LPC);
    a.wave = 148 * (smem > 113 * 1024 ? 1 : 2);
    // three-pass lengths: persistent CTAs (one per resident slot) with grouped transforms -- measured
    // 2 % slower than one row pair per CTA with natural-order transforms (profiles/r02_experiments.md),
    // kept as a selectable, parity-tested variant
    const bool grouped = F::NPASS == 3 && rt_env_flag("MLV_ZADV_GROUPED");
    if (grouped && grid > (unsigned)a.wave) grid = (unsigned)a.wave;
    if (grouped) {                           // test switch: force several row pairs per CTA on small grids
        const char* g = getenv("MLV_ZADV_GRID");
        if (g && atoi(g) > 0 && (unsigned)atoi(g) < grid) grid = (unsigned)atoi(g);
    }
    grid_out = grid;
    // per-CTA partials: the launch over rows [row0, row0 + nrows) owns slots [k grid, (k+1) grid), k = row0 / nrows
    int rc = ensure_red(c, (size_t)(a.nx / a.nrows) * grid * 4);
    if (rc) return rc;
    a.red = (c->red_user ? c->red_user : c->red) + (size_t)(a.row0 / a.nrows) * grid * 4;
    if (!c->red_on && !grouped) a.red = nullptr;          // no ticker reads the reductions of this step
    if constexpr (F::NPASS == 3) {
        if (grouped) {
            auto kfn = k_z_advect_grouped<L, LPC>;
            MLV_LAUNCH(kfn, grid, (unsigned)(LPC * F::T), smem, c->stream, a);
            return 0;
        }
    }
    const bool single = c->nranks == 1 && a.sh.fwd_chunk == 0 && a.sh.inv_chunk == 0 && (1 << a.sh.fwd_rshift) == a.nx &&
                        !rt_env_flag("MLV_ZADV_GENERIC");
#define MLV_ZADV_GO(RED_, SH_)                                                           \
    do {                                                                                  \
        auto kfn = k_z_advect<L, LPC, RED_, SH_>;                                         \
        MLV_LAUNCH(kfn, grid, (unsigned)(LPC * F::T), smem, c->stream, a);                \
    } while (0)
    if (a.red) { if (single) MLV_ZADV_GO(true, false); else MLV_ZADV_GO(true, true); }
    else { if (single) MLV_ZADV_GO(false, false); else MLV_ZADV_GO(false, true); }
#undef MLV_ZADV_GO
    return 0;
}

template <int L>
static int launch_x1d(mlv_ctx* c, X1dArgs& a, bool inverse) {
    constexpr int C = xcols(L);
    typedef FftCfg<L> F;
    const size_t smem = (size_t)F::XSLOTS * C * sizeof(cplx);
    const unsigned grid = (unsigned)(((a.nz + 1) / 2 + C - 1) / C);
    if (inverse) {
        auto kfn = k_x1d_c2r<L, C>;
        MLV_LAUNCH(kfn, grid, (unsigned)(C * F::T), smem, c->stream, a);
    } else {
        auto kfn = k_x1d_r2c<L, C>;
        MLV_LAUNCH(kfn, grid, (unsigned)(C * F::T), smem, c->stream, a);
    }
    return 0;
}

static int launch_fdm_solve(mlv_ctx* c, const void* rhs, double sign, void* out, void* uxh, void* uzh) {
    FdmSolveArgs a;
    a.rhs = (const cplx*)rhs; a.sign = sign; a.out = (cplx*)out; a.uxh = (cplx*)uxh; a.uzh = (cplx*)uzh;
    a.inv = c->tri_inv; a.nn = c->nn; a.nz = c->p.nz; a.off = 1.0 / (c->dz * c->dz); a.kx0 = c->k.kx0;
    a.f = fdm_consts(c);
    // MLV_FDM_PER consecutive unknowns per thread
    const unsigned nt = a.nz <= 256 * MLV_FDM_PER ? 256u : 512u;
    if ((long long)nt * MLV_FDM_PER < a.nz) {
        set_error("FDM-z solve: nz=%d unsupported (at most %d)", a.nz, 512 * MLV_FDM_PER);
        return MLV_ERR_UNSUPPORTED;
    }
    size_t smem = (size_t)(a.nz + (a.nz >> 3) + 1) * sizeof(cplx) + 96 * sizeof(double);
#ifdef MLV_EMU
    smem += 3 * (size_t)nt * sizeof(double);
#endif
    auto kfn = k_fdm_solve;
    MLV_LAUNCH(kfn, (unsigned)a.nn, nt, smem, c->stream, a);
    return MLV_OK;
}

template <int L>
static int launch_x1d_advect(mlv_ctx* c, X1dAdvArgs& a, unsigned& grid_out) {
    constexpr int C = xcols(L);
    typedef FftCfg<L> F;
    const size_t smem = (size_t)F::XSLOTS * C * sizeof(double) + (size_t)F::N * C * sizeof(cplx) +
                        (size_t)4 * C * F::T * sizeof(double);
    const unsigned grid = (unsigned)(((a.nz + 1) / 2 + C - 1) / C);
    grid_out = grid;
    int rc = ensure_red(c, (size_t)grid * 4);
    if (rc) return rc;
    a.red = c->red_on ? (c->red_user ? c->red_user : c->red) : nullptr;
    if (a.red) {
        auto kfn = k_x1d_advect<L, C, true>;
        MLV_LAUNCH(kfn, grid, (unsigned)(C * F::T), smem, c->stream, a);
    } else {                                        // no ticker reads the reductions of this step
        auto kfn = k_x1d_advect<L, C, false>;
        MLV_LAUNCH(kfn, grid, (unsigned)(C * F::T), smem, c->stream, a);
    }
    return 0;
}

// one-warp kernel behind a producer launch: arrival counter `which` of every rank += 1 (if this
// rank produced anything), this rank's count of arrivals due += `due`
static int signal_arrival(mlv_ctx* c, int which, bool produced, int due) {
    PeerSignals s;
    s.n = produced ? c->nranks : 0;
    for (int r = 0; r < c->nranks; ++r) s.counter[r] = c->peer_flags[r] + which;
    auto kfn = k_signal_peers;
    MLV_LAUNCH(kfn, 1u, 32u, 0, c->stream, s, c->expect + which, (unsigned long long)due);
    return 0;
}

static int reduce_final(mlv_ctx* c, const double* partial, int n, int stride, int off, int op,
                        double* out) {
    auto kfn = k_reduce_final;
    MLV_LAUNCH(kfn, 1u, 256u, 256 * sizeof(double), c->stream, partial, n, stride, off, op, out);
    return 0;
}

}  // namespace mlv

using namespace mlv;

// =========================================================================== ABI
extern "C" {

int mlv_abi_version(void) { return 2; }

long long mlv_launch_count(void) { return (long long)g_launches; }

const char* mlv_last_error(void) { return g_err; }

int mlv_create(const mlv_params* p, mlv_ctx** out) {
    if (!p || !out) { set_error("mlv_create: null argument"); return MLV_ERR_INVALID; }
    *out = nullptr;
    if (p->fd_order != 2 && p->fd_order != 4) {
        set_error("mlv_create: spatial_derivative_order must be 2 or 4");
        return MLV_ERR_INVALID;
    }
    const int lx = ilog2_exact(p->nx);
    const int lz = ilog2_exact(p->nz);
    const int lxmax = p->fdm_z ? 13 : 14;
    if (lx < 4 || lx > lxmax) {
        set_error("mlv_create: nx=%d unsupported (power of two, 16..%d)", p->nx, 1 << lxmax);
        return MLV_ERR_UNSUPPORTED;
    }
    if (!p->fdm_z && (lz < 4 || lz > 14)) {
        set_error("mlv_create: nz=%d unsupported (power of two, 16..16384)", p->nz);
        return MLV_ERR_UNSUPPORTED;
    }
    // test hook: MLV_FORCE_SPLIT (bit 0: x passes, bit 1: z stage) sends small grids through
    // the long-line kernels (instantiated for half lengths 64, 128 and 8192)
    int force = 0;
    if (const char* e = getenv("MLV_FORCE_SPLIT")) force = atoi(e);
    // 8192-point lines fit the register transform, but only as one 512-thread line per SM;
    // measured at 8192^2 the long-line forms are faster (x inverse 1.37 -> 1.01 ms, fused z
    // stage 1.60 -> 1.20 ms, whole step 3.71 -> 3.06 ms), so they start at 8192
    const bool xsplit = !p->fdm_z && (lx >= 13 || ((force & 1) && (lx == 7 || lx == 8)));
    const bool zreal = !p->fdm_z && (lz >= 13 || ((force & 2) && (lz == 7 || lz == 8)));
    if (p->fdm_z && p->nz < 5) { set_error("mlv_create: nz too small"); return MLV_ERR_INVALID; }
    mlv_ctx* c = new (std::nothrow) mlv_ctx();
    if (!c) { set_error("out of host memory"); return MLV_ERR_NOMEM; }
    c->p = *p;
    c->log2nx = lx;
    c->log2nz = lz;
    c->xsplit = xsplit; c->zreal = zreal;
    c->planlx = lx - (xsplit ? 1 : 0);
    c->planlz = lz - (zreal ? 1 : 0);
    c->nn = (p->nx - 1) / 3;                       // Parameters.py:67-70
    c->nm = p->fdm_z ? -1 : (p->nz - 1) / 3;
    c->spec_rows = p->fdm_z ? c->nn : 2 * c->nn + 1;
    c->spec_cols = p->fdm_z ? p->nz : c->nm;
    c->ct = xcols(c->planlx);                        // columns per x-pass CTA = tile width
    {   // pitch: multiple of the tile width (and even: 32-byte aligned column pairs)
        const int q = c->ct > 2 ? c->ct : 2;
        c->ipitch = p->fdm_z ? 0 : ((c->nm + q - 1) / q) * q;
    }
    c->dx = p->lx / p->nx;
    c->dz = p->lz / p->nz;
    c->k.kx0 = p->kx0; c->k.kz0 = p->kz0; c->k.d2x = p->d2x; c->k.d2z = p->d2z;
    int rc = make_plan(c->planx, c->planlx, c->stream);
    if (!rc && !p->fdm_z) rc = make_plan(c->planz, c->planlz, c->stream);
    if (!rc && xsplit) rc = make_split_table(&c->tws_x, p->nx, c->stream);
    if (!rc && zreal) rc = make_split_table(&c->tws_z, p->nz, c->stream);
    if (!rc && !p->fdm_z) {
        std::vector<double> sz(c->nm > 0 ? c->nm : 1);
        for (int m = 0; m < c->nm; ++m) sz[m] = stencil_symbol(p->fd_order, m, p->nz, c->dz);
        rc = rt_malloc((void**)&c->symz, sz.size() * sizeof(double));
        if (!rc) rc = rt_h2d(c->symz, sz.data(), sz.size() * sizeof(double), c->stream);
    }
    if (!rc && p->fdm_z) {
        // Thomas factors of the nn tridiagonal systems (LaplacianSolver.py:22-49)
        // inv[i] = 1/(b - a c'[i-1]), c'[i] = a inv[i]; identity first and last rows
        const size_t tot = (size_t)c->nn * p->nz;
        std::vector<double> inv(tot);
        const double off = 1.0 / (c->dz * c->dz);
        for (int n = 0; n < c->nn; ++n) {
            const double kx = n * p->kx0;
            const double b = -(kx * kx + 2.0 / (c->dz * c->dz));
            double* ivr = &inv[(size_t)n * p->nz];
            double cprev = 0.0;
            ivr[0] = 1.0;
            for (int i = 1; i < p->nz - 1; ++i) {
                const double den = b - off * cprev;
                ivr[i] = 1.0 / den;
                cprev = off * ivr[i];
            }
            ivr[p->nz - 1] = 1.0;
        }
        rc = rt_malloc((void**)&c->tri_inv, tot * sizeof(double));
        if (!rc) rc = rt_h2d(c->tri_inv, inv.data(), tot * sizeof(double), c->stream);
        std::vector<double> sx(c->nn > 0 ? c->nn : 1);
        for (int n = 0; n < c->nn; ++n) sx[n] = stencil_symbol(p->fd_order, n, p->nx, c->dx);
        if (!rc) rc = rt_malloc((void**)&c->symx, sx.size() * sizeof(double));
        if (!rc) rc = rt_h2d(c->symx, sx.data(), sx.size() * sizeof(double), c->stream);
    }
    if (!rc) rc = ensure_red(c, 148 * 16 * 4);
    if (!rc && !p->fdm_z) configure_shard(c, 0, 1, 1, 1);

    if (rc) { mlv_destroy(c); return rc; }
    *out = c;
    return MLV_OK;
}

int mlv_destroy(mlv_ctx* c) {
    if (!c) return MLV_OK;
    if (c->planx.dev) rt_free(c->planx.dev);
    if (c->planz.dev) rt_free(c->planz.dev);
    if (c->tws_x) rt_free(c->tws_x);
    if (c->tws_z) rt_free(c->tws_z);
    if (c->symz) rt_free(c->symz);
    if (c->tri_inv) rt_free(c->tri_inv);
    if (c->penta) rt_free(c->penta);
    if (c->symx) rt_free(c->symx);
    if (c->red) rt_free(c->red);
    if (c->expect) rt_free(c->expect);
    for (auto& t : c->trig_tables) rt_free(t.second);
    delete c;
    return MLV_OK;
}

int mlv_set_stream(mlv_ctx* c, void* s) {
    if (!c) { set_error("null context"); return MLV_ERR_INVALID; }
    c->stream = (stream_t)s;
    return MLV_OK;
}

int mlv_set_sharding(mlv_ctx* c, int rank, int nranks, int inv_fields, int fwd_fields) {
    if (!c) { set_error("null context"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_set_sharding")) return rc;
    if (nranks < 1 || nranks > MLV_MAXPEER || rank < 0 || rank >= nranks || (nranks & (nranks - 1)) || c->p.nx / nranks < 2 ||
        inv_fields < 1 || fwd_fields < 1) {
        set_error("mlv_set_sharding: need a power-of-two rank count dividing nx/2 and field counts >= 1");
        return MLV_ERR_INVALID;
    }
    c->fwd_rows = 0;
    configure_shard(c, rank, nranks, inv_fields, fwd_fields);
    return MLV_OK;
}

int mlv_set_forward_blocks(mlv_ctx* c, int rows_per_block) {
    if (!c) { set_error("null context"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_set_forward_blocks")) return rc;
    if (rows_per_block < 2 || rows_per_block > c->nxl || (rows_per_block & (rows_per_block - 1)) ||
        c->nxl % rows_per_block) {
        set_error("mlv_set_forward_blocks: need a power of two >= 2 dividing the %d local rows", c->nxl);
        return MLV_ERR_INVALID;
    }
    if (c->nranks == 1) return MLV_OK;                 // unsharded: one block
    c->fwd_rows = rows_per_block;
    configure_shard(c, c->rank, c->nranks, c->inv_fields, c->fwd_fields);
    return MLV_OK;
}

// ---- peer memory (CUDA IPC): receive buffers that other ranks' kernels store into
int mlv_p2p_alloc(mlv_ctx* c, int64_t bytes, void** ptr, void* handle64) {
    if (!c || !ptr || !handle64 || bytes <= 0) { set_error("mlv_p2p_alloc: bad argument"); return MLV_ERR_INVALID; }
#ifdef MLV_EMU
    set_error("peer memory is not available in the emulation build");
    return MLV_ERR_UNSUPPORTED;
#else
    if (int rc = rt_check(cudaMalloc(ptr, (size_t)bytes), "cudaMalloc")) return rc;
    if (int rc = rt_check(cudaMemset(*ptr, 0, (size_t)bytes), "cudaMemset")) return rc;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    return rt_check(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, *ptr), "cudaIpcGetMemHandle");
#endif
}

int mlv_p2p_open(mlv_ctx* c, const void* handle64, void** ptr) {
    if (!c || !ptr || !handle64) { set_error("mlv_p2p_open: bad argument"); return MLV_ERR_INVALID; }
#ifdef MLV_EMU
    set_error("peer memory is not available in the emulation build");
    return MLV_ERR_UNSUPPORTED;
#else
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    return rt_check(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
#endif
}

int mlv_p2p_close(mlv_ctx* c, void* ptr, int opened) {
    if (!c || !ptr) return MLV_OK;
#ifdef MLV_EMU
    return MLV_OK;
#else
    return rt_check(opened ? cudaIpcCloseMemHandle(ptr) : cudaFree(ptr), "mlv_p2p_close");
#endif
}

int mlv_p2p_copy(mlv_ctx* c, void* dst, const void* src, int64_t bytes, void* stream) {
    if (!c || !dst || !src || bytes < 0) { set_error("mlv_p2p_copy: bad argument"); return MLV_ERR_INVALID; }
    if (bytes == 0) return MLV_OK;
#ifdef MLV_EMU
    (void)stream;
    memcpy(dst, src, (size_t)bytes);
    return MLV_OK;
#else
    // MLV_COPY_CTAS=n (n > 0): a few CTAs push the block with coalesced 16-byte stores instead.  The
    // copy engines of one GPU move ~350 GB/s in total when 7 peers are served at once; SM stores
    // reach the NVLink rate with a handful of CTAs (they share the SMs with the transform kernels)
    static int ctas = -1;
    if (ctas < 0) { const char* e = getenv("MLV_COPY_CTAS"); ctas = e ? atoi(e) : 0; if (ctas < 0) ctas = 0; }
    if (ctas > 0 && (bytes & 15) == 0 && !((uintptr_t)dst & 15) && !((uintptr_t)src & 15)) {
        auto kfn = k_peer_copy;
        kfn<<<(unsigned)ctas, 512, 0, (cudaStream_t)stream>>>((cplx*)dst, (const cplx*)src, (size_t)bytes / 16);
        ++mlv::g_launches;
        return rt_check(cudaGetLastError(), "k_peer_copy");
    }
    // device-to-device copy (peer-mapped destination): runs on a copy engine, not on the SMs
    return rt_check(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream),
                    "cudaMemcpyAsync (peer)");
#endif
}

int mlv_set_peer_buffers(mlv_ctx* c, int which, void* const* bufs) {
    if (!c || (which != 0 && which != 1)) { set_error("mlv_set_peer_buffers: bad argument"); return MLV_ERR_INVALID; }
    if (c->nranks > MLV_MAXPEER) { set_error("at most %d peers", MLV_MAXPEER); return MLV_ERR_UNSUPPORTED; }
    bool& flag = which == 0 ? c->p2p_inv : c->p2p_fwd;
    cplx** tab = which == 0 ? c->peer_inv : c->peer_fwd;
    flag = bufs != nullptr;
    for (int h = 0; h < c->nranks; ++h) tab[h] = bufs ? (cplx*)bufs[h] : nullptr;
    return MLV_OK;
}

int mlv_set_peer_flags(mlv_ctx* c, void* const* counters) {
    if (!c) { set_error("mlv_set_peer_flags: null context"); return MLV_ERR_INVALID; }
    c->flags_on = counters != nullptr;
    for (int h = 0; h < c->nranks && h < MLV_MAXPEER; ++h)
        c->peer_flags[h] = counters ? (unsigned long long*)counters[h] : nullptr;
    if (c->flags_on && !c->expect) {
        const unsigned long long zero[2] = {0, 0};
        if (int rc = rt_malloc((void**)&c->expect, sizeof(zero))) return rc;
        if (int rc = rt_h2d(c->expect, zero, sizeof(zero), c->stream)) return rc;
    }
    return MLV_OK;
}

int mlv_long_lines(const mlv_ctx* c) {
    return c ? ((c->xsplit ? 1 : 0) | (c->zreal ? 2 : 0)) : 0;
}

int mlv_get_info(const mlv_ctx* c, mlv_info* o) {
    if (!c || !o) { set_error("null argument"); return MLV_ERR_INVALID; }
    o->nn = c->nn; o->nm = c->nm;
    o->spec_rows = c->spec_rows; o->spec_cols = c->spec_cols;
    o->ipitch = c->nranks == 1 ? c->ipitch : c->nml; o->nm_local = c->nm_loc;
    // one field of an exchange buffer: inverse nx*nml, forward nranks*tpr*nxl*ct (<= the former + pad)
    {
        const int64_t inv = (int64_t)c->p.nx * (c->nranks == 1 ? c->ipitch : c->nml);
        const int64_t fwd = (int64_t)c->nranks * c->sh.tpr * c->nxl * c->ct;
        o->ibytes = (inv > fwd ? inv : fwd) * (int64_t)sizeof(cplx);
    }
    // >= one partial per row (+ row-block slack); FDM-z: per pair of z columns
    o->red_doubles = 4 * ((int64_t)(c->p.fdm_z ? c->p.nz : (c->nxl > 0 ? c->nxl : c->p.nx)) + 64);
    return MLV_OK;
}

// ---------------------------------------------------------------- transforms
int mlv_x_inverse(mlv_ctx* c, int nf, const void* const* spec, const int32_t* op, void* const* idst) {
    if (!c || !spec || !op || !idst) { set_error("mlv_x_inverse: null argument"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_x_inverse")) return rc;
    if (nf < 1 || nf > MLV_XMAXF) { set_error("mlv_x_inverse: 1..%d fields", MLV_XMAXF); return MLV_ERR_INVALID; }
    XInvArgs a;
    a.nn = c->nn; a.nm = c->nm_loc; a.spitch = c->spec_cols; a.ipitch = c->nml; a.nf = nf;
    a.sh = c->sh;
    const bool signal = c->p2p_inv && c->flags_on;
    int owners = 0;                                 // ranks that own columns: each signals once per launch
    for (int r = 0; r < c->nranks; ++r) owners += (c->nm - r * c->nml > 0);
    if (c->nm_loc <= 0)                             // this rank owns no retained column
        return signal ? signal_arrival(c, 0, false, owners) : MLV_OK;
    for (int f = 0; f < nf; ++f) {
        if (!spec[f] || !idst[f] || op[f] < MLV_OP_IDENT || op[f] > MLV_OP_INVLAP) {
            set_error("mlv_x_inverse: bad field %d", f);
            return MLV_ERR_INVALID;
        }
        a.src[f] = (const cplx*)spec[f]; a.op[f] = op[f]; a.dst[f] = (cplx*)idst[f];
    }
    for (int f = 0; f < nf; ++f) a.dstoff[f] = (cplx*)idst[f] - (cplx*)idst[0];
    for (int h = 0; h < c->nranks; ++h) {
        if (c->p2p_inv) {   // idst[] point into this rank's own receive buffer (block 0 position)
            a.out.blk[h] = c->peer_inv[h] + (size_t)c->rank * c->sh.inv_chunk +
                           ((cplx*)idst[0] - c->peer_inv[c->rank]);
        } else {
            a.out.blk[h] = (cplx*)idst[0] + (size_t)h * c->sh.inv_chunk;
        }
    }
    a.k = c->k; a.tw = c->planx.tw; a.tws = c->tws_x;
    int rc = MLV_OK;
    if (c->xsplit) {
#define MLV_GO(L) rc = launch_xinv_split<L>(c, a)
        MLV_SWITCH_SPLIT(c->planlx, MLV_GO)
#undef MLV_GO
    } else {
#define MLV_GO(L) rc = launch_xinv<L>(c, a)
        MLV_SWITCH_LOG2(c->log2nx, MLV_GO)
#undef MLV_GO
    }
    if (!rc && signal) rc = signal_arrival(c, 0, true, owners);
    return rc;
}

int mlv_z_inverse(mlv_ctx* c, const void* isrc, double* phys) {
    if (!c || !isrc || !phys) { set_error("mlv_z_inverse: null argument"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_z_inverse")) return rc;
    if (c->zreal) {
        ZRealArgs r{};
        r.nx = c->nxl; r.nm = c->nm; r.ipitch = c->nml; r.ct = c->ct; r.sh = c->sh;
        r.I = (const cplx*)isrc; r.P = phys; r.tw = c->planz.tw; r.tws = c->tws_z;
#define MLV_GO(L) return launch_zreal<L>(c, r, true)
        MLV_SWITCH_SPLIT(c->planlz, MLV_GO)
#undef MLV_GO
    }
    ZArgs a{};
    a.nx = c->nxl; a.nm = c->nm; a.ipitch = c->nml; a.ct = c->ct; a.sh = c->sh;
    a.I = (const cplx*)isrc; a.P = phys; a.tw = c->planz.tw;
#define MLV_GO(L) return launch_zc2r<L>(c, a)
    MLV_SWITCH_LOG2(c->log2nz, MLV_GO)
#undef MLV_GO
    return MLV_OK;
}

int mlv_z_forward(mlv_ctx* c, const double* phys, void* idst) {
    if (!c || !idst || !phys) { set_error("mlv_z_forward: null argument"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_z_forward")) return rc;
    if (c->zreal) {
        ZRealArgs r{};
        r.nx = c->nxl; r.nm = c->nm; r.ipitch = c->nml; r.ct = c->ct; r.sh = c->sh;
        r.Pin = phys; r.Iout = (cplx*)idst; r.tw = c->planz.tw; r.tws = c->tws_z;
#define MLV_GO(L) return launch_zreal<L>(c, r, false)
        MLV_SWITCH_SPLIT(c->planlz, MLV_GO)
#undef MLV_GO
    }
    ZArgs a{};
    a.nx = c->nxl; a.nm = c->nm; a.ipitch = c->nml; a.ct = c->ct; a.sh = c->sh;
    a.Pin = phys; a.Iout = (cplx*)idst; a.tw = c->planz.tw;
#define MLV_GO(L) return launch_zr2c<L>(c, a)
    MLV_SWITCH_LOG2(c->log2nz, MLV_GO)
#undef MLV_GO
    return MLV_OK;
}

int mlv_x_forward(mlv_ctx* c, const mlv_xfwd* d) {
    if (!c || !d) { set_error("mlv_x_forward: null argument"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_x_forward")) return rc;
    if (d->nf < 1 || d->nf > MLV_XMAXF) { set_error("mlv_x_forward: 1..%d fields", MLV_XMAXF); return MLV_ERR_INVALID; }
    XFwdArgs a;
    a.nn = c->nn; a.nm = c->nm_loc; a.spitch = c->spec_cols; a.ipitch = c->nml; a.nf = d->nf;
    a.sh = c->sh;
    a.wait_counter = nullptr; a.wait_expect = nullptr;
    if (c->p2p_fwd && c->flags_on) { a.wait_counter = c->peer_flags[c->rank] + 1; a.wait_expect = c->expect + 1; }
    if (c->nm_loc <= 0) return MLV_OK;
    for (int f = 0; f < d->nf; ++f) {
        if (!d->src[f] || d->sym[f] < MLV_SYM_ONE || d->sym[f] > MLV_SYM_FDZ) {
            set_error("mlv_x_forward: bad field %d", f);
            return MLV_ERR_INVALID;
        }
        a.src[f] = (const cplx*)d->src[f]; a.sym[f] = d->sym[f]; a.coef[f] = d->coef[f];
    }
    a.symz = c->symz; a.order = c->p.fd_order; a.rdx = 1.0 / c->dx;
    a.scale = 1.0 / ((double)c->p.nx * (double)c->p.nz);      // SpectralTransformer.py:191
    a.mode = d->mode;
    a.dst = (cplx*)d->dst;
    a.lin.n = 0;
    if (d->mode == 0) {
        if (!d->dst) { set_error("mlv_x_forward: null destination"); return MLV_ERR_INVALID; }
        memset(&a.integ, 0, sizeof(a.integ));
    } else if (d->mode == 1) {
        if (int rc = check_lin(c, &d->lin, "mlv_x_forward")) return rc;
        if (int rc = check_integ(c, &d->integ, "mlv_x_forward")) return rc;
        fill_lin(a.lin, &d->lin);
        fill_integ(a.integ, &d->integ);
    } else {
        set_error("mlv_x_forward: bad mode %d", d->mode);
        return MLV_ERR_INVALID;
    }
    a.k = c->k; a.tw = c->planx.tw; a.tws = c->tws_x;
    if (c->xsplit) {
#define MLV_GO(L) return launch_xfwd_split<L>(c, a)
        MLV_SWITCH_SPLIT(c->planlx, MLV_GO)
#undef MLV_GO
    }
#define MLV_GO(L) return launch_xfwd<L>(c, a)
    MLV_SWITCH_LOG2(c->log2nx, MLV_GO)
#undef MLV_GO
    return MLV_OK;
}

int mlv_to_physical(mlv_ctx* c, const void* spec, void* iscratch, double* phys) {
    if (!c || !spec || !phys) { set_error("mlv_to_physical: null argument"); return MLV_ERR_INVALID; }
    if (c->p.fdm_z) {
        X1dArgs a{};
        a.nn = c->nn; a.nz = c->p.nz; a.S = (const cplx*)spec; a.P = phys; a.tw = c->planx.tw;
#define MLV_GO(L) return launch_x1d<L>(c, a, true)
        MLV_SWITCH_LOG2(c->log2nx, MLV_GO)
#undef MLV_GO
        return MLV_OK;
    }
    if (!iscratch) { set_error("mlv_to_physical: null scratch"); return MLV_ERR_INVALID; }
    const void* s[1] = {spec};
    const int32_t op[1] = {MLV_OP_IDENT};
    void* d[1] = {iscratch};
    if (int rc = mlv_x_inverse(c, 1, s, op, d)) return rc;
    return mlv_z_inverse(c, iscratch, phys);
}

int mlv_to_spectral(mlv_ctx* c, const double* phys, void* iscratch, void* spec) {
    if (!c || !spec || !phys) { set_error("mlv_to_spectral: null argument"); return MLV_ERR_INVALID; }
    if (c->p.fdm_z) {
        X1dArgs a{};
        a.nn = c->nn; a.nz = c->p.nz; a.Pin = phys; a.Sout = (cplx*)spec;
        a.scale = 1.0 / (double)c->p.nx;                        // SpectralTransformer.py:85
        a.tw = c->planx.tw;
#define MLV_GO(L) return launch_x1d<L>(c, a, false)
        MLV_SWITCH_LOG2(c->log2nx, MLV_GO)
#undef MLV_GO
        return MLV_OK;
    }
    if (!iscratch) { set_error("mlv_to_spectral: null scratch"); return MLV_ERR_INVALID; }
    if (int rc = mlv_z_forward(c, phys, iscratch)) return rc;
    mlv_xfwd d;
    memset(&d, 0, sizeof(d));
    d.nf = 1; d.mode = 0; d.src[0] = iscratch; d.sym[0] = MLV_SYM_ONE; d.coef[0] = 1.0; d.dst = spec;
    return mlv_x_forward(c, &d);
}

// ------------------------------------------------------------ nonlinear term
int mlv_advect_z(mlv_ctx* c, const void* iux, const void* iuz, const void* iq, void* ia, void* ib,
                 double* red4) {
    if (!c) { set_error("mlv_advect_z: null argument"); return MLV_ERR_INVALID; }
    return mlv_advect_z_rows(c, iux, iuz, iq, ia, ib, 0, c->nxl, red4);
}

int mlv_advect_z_rows(mlv_ctx* c, const void* iux, const void* iuz, const void* iq, void* ia, void* ib,
                      int row0, int nrows, double* red4) {
    if (!c || !iux || !iuz || !iq || !ia || !ib) { set_error("mlv_advect_z: null argument"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_advect_z")) return rc;
    if (row0 < 0 || nrows < 2 || (nrows & 1) || row0 % nrows || row0 + nrows > c->nxl || c->nxl % nrows) {
        set_error("mlv_advect_z_rows: need equal even row ranges tiling the %d local rows", c->nxl);
        return MLV_ERR_INVALID;
    }
    ZAdvArgs a{};
    a.row0 = row0; a.nrows = nrows;
    a.nx = c->nxl; a.nm = c->nm; a.ipitch = c->nml; a.ct = c->ct; a.sh = c->sh;
    a.Iux = (const cplx*)iux; a.Iuz = (const cplx*)iuz; a.Iq = (const cplx*)iq;
    a.IA = (cplx*)ia; a.IB = (cplx*)ib; a.tw = c->planz.tw;
    a.outoff[0] = 0; a.outoff[1] = (cplx*)ib - (cplx*)ia;
    for (int h = 0; h < c->nranks; ++h) {
        if (c->p2p_fwd) {
            a.out.blk[h] = c->peer_fwd[h] + (size_t)c->rank * c->sh.fwd_peer +
                           ((cplx*)ia - c->peer_fwd[c->rank]);
        } else {
            a.out.blk[h] = (cplx*)ia + (size_t)h * c->sh.fwd_peer;
        }
    }
    if (c->p2p_inv && c->flags_on) { a.wait_counter = c->peer_flags[c->rank]; a.wait_expect = c->expect; }
    unsigned grid = 0;
    int rc = 0;
    a.tws = c->tws_z;
    if (c->zreal) {
#define MLV_GO(L) rc = launch_zadv_real<L>(c, a, grid)
        MLV_SWITCH_SPLIT(c->planlz, MLV_GO)
#undef MLV_GO
    } else {
#define MLV_GO(L) rc = launch_zadv<L>(c, a, grid)
        MLV_SWITCH_LOG2(c->log2nz, MLV_GO)
#undef MLV_GO
    }
    if (rc) return rc;
    if (c->p2p_fwd && c->flags_on)         // every rank bumps the forward counter of every rank once per launch
        if (int rs = signal_arrival(c, 1, true, c->nranks)) return rs;
    // per-CTA partials of the rows [0, row0 + nrows) launched so far
    c->red_count = a.red ? (int)((long long)grid * (row0 + nrows) / nrows) : 0;
    if (red4) return mlv_reduce_partials(c, nullptr, red4);
    return MLV_OK;
}

int mlv_set_reductions(mlv_ctx* c, int on) {
    if (!c) { set_error("mlv_set_reductions: null context"); return MLV_ERR_INVALID; }
    c->red_on = on != 0;
    return MLV_OK;
}

int mlv_set_reduction_partials(mlv_ctx* c, double* partials) {
    if (!c) { set_error("mlv_set_reduction_partials: null context"); return MLV_ERR_INVALID; }
    c->red_user = partials;
    return MLV_OK;
}

int mlv_reduce_partials(mlv_ctx* c, const double* partials, double* red4) {
    if (!c || !red4) { set_error("mlv_reduce_partials: null argument"); return MLV_ERR_INVALID; }
    if (c->red_count <= 0) { set_error("mlv_reduce_partials: no fused z stage has run"); return MLV_ERR_INVALID; }
    const double* src = partials ? partials : (c->red_user ? c->red_user : c->red);
    auto kfn = k_reduce_final4;
    MLV_LAUNCH(kfn, 4u, 256u, 256 * sizeof(double), c->stream, src, c->red_count, red4);
    return MLV_OK;
}

int mlv_advect_phys(mlv_ctx* c, const double* ux, const double* uz, const double* q, double* out) {
    if (!c || !ux || !uz || !q || !out) { set_error("mlv_advect_phys: null argument"); return MLV_ERR_INVALID; }
    AdvectArgs a;
    a.ux = ux; a.uz = uz; a.q = q; a.out = out;
    a.nx = c->p.nx; a.nz = c->p.nz; a.order = c->p.fd_order; a.z_periodic = !c->p.fdm_z;
    a.dx = c->dx; a.dz = c->dz;
    auto kfn = k_advect_phys;
    MLV_LAUNCH(kfn, grid1d((size_t)a.nx * a.nz), 256u, 0, c->stream, a);
    return MLV_OK;
}

// ------------------------------------------------------ spectral pointwise
int mlv_spec_lincomb(mlv_ctx* c, const mlv_lin_terms* t, void* out) {
    if (!c || !t || !out) { set_error("mlv_spec_lincomb: null argument"); return MLV_ERR_INVALID; }
    if (int rc = check_lin(c, t, "mlv_spec_lincomb")) return rc;
    SpecLinArgs a;
    a.rows = c->spec_rows; a.cols = c->spec_cols; a.nn = c->nn; a.fdm = c->p.fdm_z;
    a.m_off = c->sh.m_off; a.nm_glob = c->p.fdm_z ? 0 : c->nm;
    fill_lin(a.lin, t);
    a.out = (cplx*)out; a.k = c->k; a.f = fdm_consts(c);
    auto kfn = k_spec_lincomb;
    MLV_LAUNCH(kfn, grid1d((size_t)a.rows * a.cols), 256u, 0, c->stream, a);
    return MLV_OK;
}

int mlv_lap_array(mlv_ctx* c, double coef, double* out) {
    if (!c || !out) { set_error("mlv_lap_array: null argument"); return MLV_ERR_INVALID; }
    if (int rc = need_2d(c, "mlv_lap_array")) return rc;
    auto kfn = k_lap_array;
    MLV_LAUNCH(kfn, grid1d((size_t)c->spec_rows * c->spec_cols), 256u, 0, c->stream, out,
               c->spec_rows, c->spec_cols, c->nn, c->k, coef, c->sh.m_off);
    return MLV_OK;
}

int mlv_stencil(mlv_ctx* c, const void* in, void* out, int rows, int cols, int ncomp, int axis,
                int order, int periodic, int second, double h) {
    if (!c || !in || !out) { set_error("mlv_stencil: null argument"); return MLV_ERR_INVALID; }
    if ((order != 2 && order != 4) || (ncomp != 1 && ncomp != 2) || (axis != 0 && axis != 1) ||
        rows < 1 || cols < 1) {
        set_error("mlv_stencil: bad argument");
        return MLV_ERR_INVALID;
    }
    if (second && axis != 1) { set_error("mlv_stencil: second derivative only along z"); return MLV_ERR_INVALID; }
    StencilArgs a;
    a.in = (const double*)in; a.out = (double*)out; a.rows = rows; a.cols = cols; a.ncomp = ncomp;
    a.axis = axis; a.order = order; a.periodic = periodic; a.second = second; a.h = h;
    auto kfn = k_stencil;
    MLV_LAUNCH(kfn, grid1d((size_t)rows * cols * ncomp), 256u, 0, c->stream, a);
    return MLV_OK;
}

int mlv_solve_fdm(mlv_ctx* c, const void* rhs, void* out) {
    if (!c || !rhs || !out) { set_error("mlv_solve_fdm: null argument"); return MLV_ERR_INVALID; }
    if (!c->p.fdm_z) { set_error("mlv_solve_fdm: context is fully spectral"); return MLV_ERR_INVALID; }
    return launch_fdm_solve(c, rhs, 1.0, out, nullptr, nullptr);
}

// LU factors (no pivoting) of the nn pentadiagonal systems of the 4th-order solve, in extended
// precision on the host; stored as the coefficients of the two substitution recurrences.
static int build_penta(mlv_ctx* c) {
    const int nz = c->p.nz, nn = c->nn;
    const size_t tot = (size_t)nn * nz;
    std::vector<double> co(5 * tot);
    double *fa = co.data(), *fb = fa + tot, *dinv = fb + tot, *ba = dinv + tot, *bb = ba + tot;
    const long double h2 = (long double)c->dz * (long double)c->dz;
    std::vector<long double> e(nz), cc(nz), d(nz), f(nz), g(nz), du(nz), fu(nz);
    for (int n = 0; n < nn; ++n) {
        const long double k2 = ((long double)n * (long double)c->p.kx0) * ((long double)n * (long double)c->p.kx0);
        for (int i = 0; i < nz; ++i) {
            e[i] = cc[i] = f[i] = g[i] = 0.0L;
            if (i == 0 || i == nz - 1) { d[i] = 1.0L; continue; }             // solution = right-hand side
            if (i == 1 || i == nz - 2) {                                        // 2nd-order closure
                cc[i] = f[i] = 1.0L / h2; d[i] = -2.0L / h2 - k2;
                continue;
            }
            e[i] = g[i] = -1.0L / (12.0L * h2);                                 // SpatialDifferentiator.py:121-130
            cc[i] = f[i] = 4.0L / (3.0L * h2);
            d[i] = -5.0L / (2.0L * h2) - k2;
        }
        double* FA = fa + (size_t)n * nz; double* FB = fb + (size_t)n * nz; double* DI = dinv + (size_t)n * nz;
        double* BA = ba + (size_t)n * nz; double* BB = bb + (size_t)n * nz;
        for (int i = 0; i < nz; ++i) {
            long double l2 = 0.0L, l1 = 0.0L, ci = cc[i], di = d[i], fi = f[i];
            if (i >= 2) { l2 = e[i] / du[i - 2]; ci -= l2 * fu[i - 2]; di -= l2 * g[i - 2]; }
            if (i >= 1) { l1 = ci / du[i - 1]; di -= l1 * fu[i - 1]; fi -= l1 * g[i - 1]; }
            du[i] = di; fu[i] = fi;
            FA[i] = (double)(-l1); FB[i] = (double)(-l2);
            DI[i] = (double)(1.0L / di);
            BA[i] = (double)(-fi / di); BB[i] = (double)(-g[i] / di);
        }
    }
    int rc = rt_malloc((void**)&c->penta, co.size() * sizeof(double));
    if (!rc) rc = rt_h2d(c->penta, co.data(), co.size() * sizeof(double), c->stream);
    return rc;
}

int mlv_solve_fdm_o4(mlv_ctx* c, const void* rhs, void* out) {
    if (!c || !rhs || !out) { set_error("mlv_solve_fdm_o4: null argument"); return MLV_ERR_INVALID; }
    if (!c->p.fdm_z) { set_error("mlv_solve_fdm_o4: context is fully spectral"); return MLV_ERR_INVALID; }
    if (c->p.nz < 6) { set_error("mlv_solve_fdm_o4: nz >= 6"); return MLV_ERR_UNSUPPORTED; }
    if (!c->penta) if (int rc = build_penta(c)) return rc;
    FdmSolve5Args a;
    const size_t tot = (size_t)c->nn * c->p.nz;
    a.rhs = (const cplx*)rhs; a.out = (cplx*)out; a.nn = c->nn; a.nz = c->p.nz;
    a.fa = c->penta; a.fb = c->penta + tot; a.dinv = c->penta + 2 * tot; a.ba = c->penta + 3 * tot; a.bb = c->penta + 4 * tot;
    const unsigned nt = a.nz <= 256 * MLV_FDM_PER ? 256u : 512u;
    if ((long long)nt * MLV_FDM_PER < a.nz) {
        set_error("FDM-z solve: nz=%d unsupported (at most %d)", a.nz, 512 * MLV_FDM_PER);
        return MLV_ERR_UNSUPPORTED;
    }
    size_t smem = (size_t)(a.nz + (a.nz >> 3) + 1) * sizeof(cplx) + 256 * sizeof(double);
#ifdef MLV_EMU
    smem += 8 * (size_t)nt * sizeof(double);
#endif
    auto kfn = k_fdm_solve5;
    MLV_LAUNCH(kfn, (unsigned)a.nn, nt, smem, c->stream, a);
    return MLV_OK;
}

int mlv_fdm_velocity(mlv_ctx* c, const void* w, void* psi, void* uxh, void* uzh) {
    if (!c || !w || !psi || !uxh || !uzh) { set_error("mlv_fdm_velocity: null argument"); return MLV_ERR_INVALID; }
    if (!c->p.fdm_z) { set_error("mlv_fdm_velocity: context is fully spectral"); return MLV_ERR_INVALID; }
    return launch_fdm_solve(c, w, -1.0, psi, uxh, uzh);          // psi = solve(-w), utility.py:65
}

int mlv_fdm_advect(mlv_ctx* c, const void* uxh, const void* uzh, const void* q, void* ia, void* ib,
                   double* red4) {
    if (!c || !uxh || !uzh || !q || !ia || !ib) { set_error("mlv_fdm_advect: null argument"); return MLV_ERR_INVALID; }
    if (!c->p.fdm_z) { set_error("mlv_fdm_advect: context is fully spectral"); return MLV_ERR_INVALID; }
    X1dAdvArgs a{};
    a.nn = c->nn; a.nz = c->p.nz;
    a.uxh = (const cplx*)uxh; a.uzh = (const cplx*)uzh; a.q = (const cplx*)q;
    a.A = (cplx*)ia; a.B = (cplx*)ib;
    a.scale = 1.0 / (double)c->p.nx;                          // SpectralTransformer.py:85
    a.tw = c->planx.tw;
    unsigned grid = 0;
    int rc = 0;
#define MLV_GO(L) rc = launch_x1d_advect<L>(c, a, grid)
    MLV_SWITCH_LOG2(c->log2nx, MLV_GO)
#undef MLV_GO
    if (rc) return rc;
    c->red_count = a.red ? (int)grid : 0;
    if (red4) return mlv_reduce_partials(c, nullptr, red4);
    return MLV_OK;
}

int mlv_integrate(mlv_ctx* c, const mlv_lin_terms* extra, const mlv_integ* g) {
    if (!c) { set_error("mlv_integrate: null context"); return MLV_ERR_INVALID; }
    if (int rc = check_lin(c, extra, "mlv_integrate")) return rc;
    if (int rc = check_integ(c, g, "mlv_integrate")) return rc;
    IntegKArgs a;
    a.rows = c->spec_rows; a.cols = c->spec_cols; a.nn = c->nn; a.fdm = c->p.fdm_z;
    a.m_off = c->sh.m_off;
    a.f0_set = g->f0_set ? 1 : 0;
    fill_lin(a.lin, extra);
    fill_integ(a.integ, g);
    a.k = c->k; a.f = fdm_consts(c);
    auto kfn = k_integrate;
    MLV_LAUNCH(kfn, grid1d((size_t)a.rows * a.cols), 256u, 0, c->stream, a);
    return MLV_OK;
}

// -------------------------------------------------------- array namespace
int mlv_elementwise(mlv_ctx* c, const mlv_ew* d) {
    if (!c || !d || !d->out.ptr) { set_error("mlv_elementwise: null argument"); return MLV_ERR_INVALID; }
    if (d->rows < 0 || d->cols < 0 || d->op < MLV_EW_ADD || d->op > MLV_EW_POW) {
        set_error("mlv_elementwise: bad argument");
        return MLV_ERR_INVALID;
    }
    if (d->rows == 0 || d->cols == 0) return MLV_OK;
    EwArgs a;
    a.op = d->op; a.rows = d->rows; a.cols = d->cols;
    a.out.p = d->out.ptr; a.out.rs = d->out.row_stride; a.out.cs = d->out.col_stride; a.out_kind = d->out_kind;
    a.a.p = d->a.ptr; a.a.rs = d->a.row_stride; a.a.cs = d->a.col_stride; a.a_kind = d->a_kind;
    a.b.p = d->b.ptr; a.b.rs = d->b.row_stride; a.b.cs = d->b.col_stride; a.b_kind = d->b_kind;
    a.a_re = d->a_re; a.a_im = d->a_im; a.b_re = d->b_re; a.b_im = d->b_im;
    auto kfn = k_elementwise;
    MLV_LAUNCH(kfn, grid1d((size_t)a.rows * a.cols), 256u, 0, c->stream, a);
    return MLV_OK;
}

int mlv_reduce(mlv_ctx* c, int op, int rows, int cols, const mlv_view* av, const mlv_view* bv,
               double* out_dev) {
    if (!c || !av || !av->ptr || !out_dev) { set_error("mlv_reduce: null argument"); return MLV_ERR_INVALID; }
    if (op < MLV_RED_SUM || op > MLV_RED_SUMPROD || rows < 1 || cols < 1) {
        set_error("mlv_reduce: bad argument");
        return MLV_ERR_INVALID;
    }
    if (op == MLV_RED_SUMPROD && (!bv || !bv->ptr)) { set_error("mlv_reduce: second operand missing"); return MLV_ERR_INVALID; }
    RedArgs a;
    a.op = op; a.rows = rows; a.cols = cols;
    a.a.p = av->ptr; a.a.rs = av->row_stride; a.a.cs = av->col_stride;
    a.b.p = bv ? bv->ptr : nullptr; a.b.rs = bv ? bv->row_stride : 0; a.b.cs = bv ? bv->col_stride : 0;
    const unsigned grid = grid1d((size_t)rows * cols);
    if (int rc = ensure_red(c, grid)) return rc;
    a.partial = c->red;
    auto kfn = k_reduce;
    MLV_LAUNCH(kfn, grid, 256u, 256 * sizeof(double), c->stream, a);
    const int fop = (op == MLV_RED_MAX) ? RED_MAX : (op == MLV_RED_MIN ? RED_MIN : RED_SUM);
    return reduce_final(c, c->red, (int)grid, 1, 0, fop, out_dev);
}

// ------------------------------------------------ COSINE / SINE bases
static int trig_table(mlv_ctx* c, int M, const cplx** out) {
    for (auto& t : c->trig_tables)
        if (t.first == M) { *out = t.second; return MLV_OK; }
    std::vector<cplx> host((size_t)M);
    for (int j = 0; j < M; ++j) {
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)M;
        host[j] = mk((double)cosl(ang), (double)sinl(ang));
    }
    cplx* dev = nullptr;
    int rc = rt_malloc((void**)&dev, host.size() * sizeof(cplx));
    if (!rc) rc = rt_h2d(dev, host.data(), host.size() * sizeof(cplx), c->stream);
    if (rc) return rc;
    c->trig_tables.push_back(std::make_pair(M, dev));
    *out = dev;
    return MLV_OK;
}

int mlv_trig_axis(mlv_ctx* c, const mlv_trig* d) {
    if (!c || !d || !d->in || !d->out) { set_error("mlv_trig_axis: null argument"); return MLV_ERR_INVALID; }
    const int mirrored = d->ext == MLV_EXT_EVEN || d->ext == MLV_EXT_ODD;
    if (d->ext < MLV_EXT_PERIODIC || d->ext > MLV_EXT_ODD || d->n_samp < 2 || d->n_modes < 1 || d->nbatch < 1 ||
        d->period != (mirrored ? 2 * (d->n_samp - 1) : d->n_samp) ||
        (d->two_sided && !(d->n_modes & 1)) || (d->two_sided ? d->n_modes / 2 : d->n_modes - 1) > d->period / 2 ||
        (d->hermitian && (!d->inverse || d->two_sided || d->scale_im != 0.0))) {
        set_error("mlv_trig_axis: inconsistent descriptor");
        return MLV_ERR_INVALID;
    }
    TrigArgs a;
    a.inverse = d->inverse; a.ext = d->ext; a.M = d->period; a.n_samp = d->n_samp; a.n_modes = d->n_modes;
    a.two_sided = d->two_sided; a.hermitian = d->hermitian; a.samp_complex = d->samp_complex;
    a.batch_fastest = d->batch_fastest; a.w0 = d->w0; a.nbatch = d->nbatch;
    a.samp_stride = d->samp_stride; a.samp_batch = d->samp_batch_stride;
    a.mode_stride = d->mode_stride; a.mode_batch = d->mode_batch_stride;
    a.in = d->in; a.out = d->out; a.sre = d->scale_re; a.sim = d->scale_im;
    if (int rc = trig_table(c, a.M, &a.E)) return rc;
    auto kfn = k_trig_axis;
    MLV_LAUNCH(kfn, grid1d((size_t)(a.inverse ? a.n_samp : a.n_modes) * a.nbatch), 256u, 0, c->stream, a);
    return MLV_OK;
}

}  // extern "C"
