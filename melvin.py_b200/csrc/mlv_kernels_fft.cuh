// melvin-b200: FFT-based kernels of the pseudo-spectral step.
//
// Data layouts (all row-major, last index contiguous):
//   S  spectral   (2nn+1, nm) complex128, rows n = 0..nn,-nn..-1   [API layout,
//                  reference melvin/ArrayFactory.py:8-45]
//   P  physical   (nx, nz) float64                                  [API layout]
//   I  x-transformed intermediate, private.  Two layouts, chosen so that every
//      *read* is contiguous and only *writes* scatter (the L2 merges scattered
//      32-byte writes; scattered reads would cost a DRAM row activation each):
//        row layout   (nx, ipitch)          written by the inverse x pass (column
//                      tiles), read row-wise by the z stage;
//        tile layout  [m/CT][nx][CT]        written by the forward z stage (CT = column
//                      count of an x-pass CTA), read as one contiguous block per CTA by
//                      the forward x pass.
//   FDM-z mode:  S_f (nn, nz) complex128 <-> P, one 1-D transform along x.
//
// A 2-D transform is an x pass (complex, strided columns, C adjacent columns per
// CTA so that every global access is a C*16-byte segment) and a z pass (two real
// rows packed into one complex line; contiguous).  Pruning (2/3 rule) and all
// scaling happen in the load/store of the passes; no padded array ever exists.
#pragma once

#include "mlv_fft.cuh"

namespace mlv {

// ---------------------------------------------------------------- op codes
// Diagonal spectral operators: value = g(n, m) * src[n][m].
enum {
    XOP_IDENT = 0,
    XOP_PSI = 1,   // psi = (-w)/lap, lap(0,0) := 1   (LaplacianSolver.py:58-68, utility.py:65)
    XOP_UX = 2,    // ux  = -(i kz m) psi             (utility.py:71, SpatialDifferentiator.py:55)
    XOP_UZ = 3,    // uz  =  (i kx n) psi             (utility.py:78, SpatialDifferentiator.py:50)
    XOP_DDX = 4,   // (i kx n) src                    (SpatialDifferentiator.py:50)
    XOP_DDZ = 5,   // (i kz m) src                    (SpatialDifferentiator.py:55)
    XOP_D2DX2 = 6, // d2x n^2 src                     (SpatialDifferentiator.py:60)
    XOP_D2DZ2 = 7, // d2z m^2 src                     (SpatialDifferentiator.py:65)
    XOP_LAP = 8,   // (d2x n^2 + d2z m^2) src         (Variable.py:111-113)
    XOP_INVLAP = 9 // src / lap, lap(0,0) := 1        (LaplacianSolver.py:58-68)
};
// Multiplier symbols of the forward x pass epilogue.
enum {
    XSYM_ONE = 0,
    XSYM_FDX = 1,  // central x stencil (order 2/4), applied along the line before the transform
    XSYM_FDZ = 2,  // Fourier symbol of the central z stencil, table symz[m]
};

struct SpecConsts {
    double kx0, kz0;   // 2 pi / lx, 2 pi / lz
    double d2x, d2z;   // -(2 pi)^2/lx^2, -(2 pi)^2/lz^2
};

MLV_DEV double lap_symbol(int n, int m, const SpecConsts& k) {
    return k.d2x * ((double)n * (double)n) + k.d2z * ((double)m * (double)m);
}

MLV_DEV cplx spectral_op(int op, cplx s, int n, int m, const SpecConsts& k) {
    switch (op) {
        case XOP_IDENT: return s;
        case XOP_DDX: { const double b = k.kx0 * n; return mk(-b * s.y, b * s.x); }
        case XOP_DDZ: { const double b = k.kz0 * m; return mk(-b * s.y, b * s.x); }
        case XOP_D2DX2: { const double b = k.d2x * ((double)n * (double)n); return mk(b * s.x, b * s.y); }
        case XOP_D2DZ2: { const double b = k.d2z * ((double)m * (double)m); return mk(b * s.x, b * s.y); }
        case XOP_LAP: { const double b = lap_symbol(n, m, k); return mk(b * s.x, b * s.y); }
        default: break;
    }
    double lap = lap_symbol(n, m, k);
    if (n == 0 && m == 0) lap = 1.0;
    const double r = fast_rcp(lap);
    if (op == XOP_INVLAP) return mk(s.x * r, s.y * r);
    if (op == XOP_PSI) return mk(-s.x * r, -s.y * r);
    if (op == XOP_UX) { const double c = (k.kz0 * m) * r; return mk(-c * s.y, c * s.x); }
    /* XOP_UZ */ { const double c = (k.kx0 * n) * r; return mk(c * s.y, -c * s.x); }
}

// 2/3-rule invariant used for static pruning: a transform of length N keeps at most
// (N-1)/3 modes per side (Parameters.py:67-70), and thread tau owns the FFT indices
// tau + (N/16) j.  Indices with j = 6..9 lie in [6N/16, 10N/16) and are therefore ALWAYS
// truncated; retained low modes have j <= 5, retained high (negative) modes j >= 10.  Loops
// skip those j statically, so the zeros also fold through the first butterflies of the
// inverse transforms and the unused outputs of the forward transforms are never computed.
#define MLV_MID(j) ((j) >= 6 && (j) <= 9)

// FFT index k (0..N-1) -> spectral row r and signed mode n; false if truncated.
MLV_DEV bool xrow_of(int k, int N, int nn, int& r, int& n) {
    if (k <= nn) { r = k; n = k; return true; }
    if (k >= N - nn) { n = k - N; r = n + 2 * nn + 1; return true; }
    return false;
}

// Same for the point j (compile-time after unrolling) of thread tau, k = tau + (N/16) j:
// by the invariant above j <= 5 can only be a low mode and j >= 10 only a high one, so the
// mapping is branch-free.
template <int N>
MLV_DEV bool xrow_static(int j, int k, int nn, int& r, int& n) {
    if (j <= 5) { r = k; n = k; return k <= nn; }
    if (j >= 10) { n = k - N; r = n + 2 * nn + 1; return k >= N - nn; }
    r = 0; n = 0;
    return false;
}

// Slab decomposition over G ranks (G = 1: one chunk, everything local).
//   spectral state : kz-slabs, rank g owns columns [g*nml, (g+1)*nml) (nml = tpr*CT)
//   physical side  : x-slabs,  rank g owns rows    [g*nxl, (g+1)*nxl)
// Exchange buffers are laid out so that the block for every peer is contiguous:
//   inverse (x pass -> z stage): [peer h][field][nxl rows of h][nml cols]  (row layout)
//   forward (z stage -> x pass): [peer h][row block][field][tpr tiles of h][RB rows][CT]  (tile layout)
struct Shard {
    int m_off;            // global index of local column 0
    int nm_glob;          // global number of retained columns (nm)
    int nml;              // columns per rank = pitch of local spectral arrays / inverse buffers
    int rpc_shift;        // log2(rows per rank)
    int tpr;              // column tiles per rank
    long long inv_chunk;  // elements between the blocks of consecutive peers, inverse buffers
    // forward buffers: row blocks of RB = 2^fwd_rshift rows (RB divides the rows of a rank; one
    // block per rank by default, several when the caller pipelines the exchange by row ranges).
    // Receive side: [global row block][tile][RB][CT];  send side: [dest peer][local row block][tile][RB][CT]
    long long fwd_chunk;  // elements between consecutive row blocks (0 when unsharded)
    long long fwd_peer;   // elements between the send regions of consecutive destination peers
    int fwd_rshift;       // log2(RB)
};

// Where the producer kernels store the block destined for peer h.  Without peer
// memory: blk[h] = local send buffer + h*chunk (an all-to-all follows).  With peer
// memory (mlv_set_peer_buffers): blk[h] = peer h's receive buffer + rank*chunk, i.e. the
// exchange is fused into the stores of the kernel and crosses NVLink directly.
#define MLV_MAXPEER 8
struct PeerBlocks {
    cplx* blk[MLV_MAXPEER];
};
// Arrival counters of one exchange direction, one per rank (null: no device-side ordering).
struct PeerSignals {
    int n;                                   // ranks to signal (0: none)
    unsigned long long* counter[MLV_MAXPEER];
};
// Runs right after a producer kernel on the same stream: the kernel boundary has completed all
// of the producer's peer stores, one thread per rank publishes the arrival.  (Signalling from
// every producer CTA instead costs a system-scope fence per CTA, which waits for thousands of
// outstanding NVLink stores with one CTA per SM: measured slower than this extra launch.)
// It also advances this rank's own count of arrivals due (`expect`, device memory, += inc): the
// consumer kernels read it from there, so no launch argument changes from step to step and a whole
// step can be replayed as a CUDA graph.
__global__ void __launch_bounds__(32) k_signal_peers(const PeerSignals s, unsigned long long* expect,
                                                     unsigned long long inc) {
    if ((int)threadIdx.x < s.n) flag_signal(s.counter[threadIdx.x]);
    if (threadIdx.x == 31) *expect += inc;
}
// first statement of a consumer CTA
MLV_DEV void wait_arrivals(const unsigned long long* counter, const unsigned long long* expect) {
    if (counter == nullptr) return;
    if (threadIdx.x == 0) flag_wait(counter, *reinterpret_cast<const volatile unsigned long long*>(expect));
    __syncthreads();
}

// element (local row xl, tile tl of its owner, column cc of the tile) inside the send region of
// the tile owner
MLV_DEV size_t fwd_store_off(int xl, int tl, int cc, int ct, const Shard& sh) {
    const int rbm = (1 << sh.fwd_rshift) - 1;
    return (size_t)(xl >> sh.fwd_rshift) * sh.fwd_chunk +
           ((((size_t)tl) << sh.fwd_rshift) + (size_t)(xl & rbm)) * ct + cc;
}
// element (local row xl, global column m) of a forward intermediate, tile layout
MLV_DEV size_t fwd_off(int xl, int m, int ct, const Shard& sh) {
    const int t = m / ct;
    const int h = t / sh.tpr, tl = t - h * sh.tpr;
    return (size_t)h * sh.fwd_peer + fwd_store_off(xl, tl, m % ct, ct, sh);
}
// element m (global column) of a row of an inverse intermediate whose chunk-0 part starts at `row`
MLV_DEV size_t inv_col_off(int m, const Shard& sh) {
    const int h = m / sh.nml;
    return (size_t)h * sh.inv_chunk + (m - h * sh.nml);
}

// ------------------------------------------------------ time integration
// One spectral coefficient of Integrator.py:5-18 (AB2/AB4 predictor) and
// :53-63 (explicit / theta-scheme update).  History levels are passed oldest
// last; unused levels are ignored.
struct IntegArgs {
    int ab_order;      // 2 or 4
    int scheme;        // 0: semi-implicit, L = lcoef*lap symbol; 1: semi-implicit, L from array
                       // 2: explicit (q += AB)
    double dt, alpha, lcoef;
    const double* larr;        // scheme 1: real (spectral-shaped) linear operator
    const cplx* q_in;
    cplx* q_out;
    cplx* f0;                  // current history level (written by the caller, read here)
    const cplx* fm1;
    const cplx* fm2;
    const cplx* fm3;
};

MLV_DEV cplx ab_predict(const IntegArgs& g, cplx f0, size_t idx) {
    if (g.ab_order == 2) {
        const cplx f1 = g.fm1[idx];
        const double h = g.dt / 2;
        return mk(h * (3 * f0.x - f1.x), h * (3 * f0.y - f1.y));
    }
    const cplx f1 = g.fm1[idx], f2 = g.fm2[idx], f3 = g.fm3[idx];
    const double h = g.dt / 24;
    return mk(h * (55 * f0.x - 59 * f1.x + 37 * f2.x - 9 * f3.x),
              h * (55 * f0.y - 59 * f1.y + 37 * f2.y - 9 * f3.y));
}

MLV_DEV void integrate_point(const IntegArgs& g, cplx f0, size_t idx, int n, int m,
                             const SpecConsts& k) {
    const cplx inc = ab_predict(g, f0, idx);
    const cplx q = g.q_in[idx];
    if (g.scheme == 2) {
        g.q_out[idx] = cadd(q, inc);
        return;
    }
    const double L = (g.scheme == 0) ? g.lcoef * lap_symbol(n, m, k) : g.larr[idx];
    const double a = 1 + ((1 - g.alpha) * g.dt) * L;
    const double rb = fast_rcp(1 - (g.alpha * g.dt) * L);
    g.q_out[idx] = mk((a * q.x + inc.x) * rb, (a * q.y + inc.y) * rb);
}

// same update with the state and history values already in registers
MLV_DEV cplx integrate_value(const IntegArgs& g, cplx f0, cplx q, cplx f1, cplx f2, cplx f3,
                             size_t idx, int n, int m, const SpecConsts& k) {
    cplx inc;
    if (g.ab_order == 2) {
        const double h = g.dt / 2;
        inc = mk(h * (3 * f0.x - f1.x), h * (3 * f0.y - f1.y));
    } else {
        const double h = g.dt / 24;
        inc = mk(h * (55 * f0.x - 59 * f1.x + 37 * f2.x - 9 * f3.x),
                 h * (55 * f0.y - 59 * f1.y + 37 * f2.y - 9 * f3.y));
    }
    if (g.scheme == 2) return cadd(q, inc);
    const double L = (g.scheme == 0) ? g.lcoef * lap_symbol(n, m, k) : g.larr[idx];
    const double a = 1 + ((1 - g.alpha) * g.dt) * L;
    const double rb = fast_rcp(1 - (g.alpha * g.dt) * L);
    return mk((a * q.x + inc.x) * rb, (a * q.y + inc.y) * rb);
}

// Extra linear right-hand-side terms  sum_i coef_i * op_i(src_i)
#define MLV_MAXLIN 6
struct LinTerms {
    int n;
    const cplx* src[MLV_MAXLIN];
    int op[MLV_MAXLIN];
    double cre[MLV_MAXLIN], cim[MLV_MAXLIN];   // complex coefficient
};

MLV_DEV cplx lin_terms_at(const LinTerms& lt, size_t idx, int n, int m, const SpecConsts& k) {
    cplx acc = mk(0.0, 0.0);
    for (int i = 0; i < lt.n; ++i) {
        const cplx t = spectral_op(lt.op[i], lt.src[i][idx], n, m, k);
        acc = cadd(acc, cmul(mk(lt.cre[i], lt.cim[i]), t));
    }
    return acc;
}

// ===================================================================== x inverse
#define MLV_XMAXF 4
struct XInvArgs {
    int nn, nm, spitch, ipitch, nf;
    int wave;                    // CTAs resident at once (prefetch distance)
    const cplx* src[MLV_XMAXF];
    int op[MLV_XMAXF];
    cplx* dst[MLV_XMAXF];
    Shard sh;                    // nm = local valid columns; spectral ops use m + sh.m_off
    PeerBlocks out;              // destination block per row owner; dst[f] are offsets into it
    long long dstoff[MLV_XMAXF];
    SpecConsts k;
    FftTw tw;
    const cplx* tws;             // split lines: e^{-2 pi i k/nx}, k < nx/2
    int use_tma;                 // 1: column tiles leave through TMA tensor stores (unsharded, GPU build)
    int tma_rows;                // rows per tensor-store box (<= 256)
    CUtensorMap tmap[MLV_XMAXF]; // per destination: (nx, 2*ipitch) float64, box (tma_rows, 2*C)
    int load_tma;                // 1: source column tiles arrive in the stash through TMA tensor loads
    int ld_rows, ld_boxes;       // rows per load box, boxes per column tile (stash holds ld_rows*ld_boxes rows)
    CUtensorMap smap[MLV_XMAXF]; // per source: (2nn+1, 2*spitch) float64, box (ld_rows, 2*C)
};

// spectral (2nn+1, nm) -> I (nx, ipitch); nf fields, each with its own prologue.
// Shared memory: [ exchange XSLOTS*C cplx | raw-column stash (2nn+1)*C cplx ].
// Consecutive fields that read the same spectral array (ux, uz, w all come from
// the vorticity) load it from global memory once and re-read it from the
// thread-private stash.
template <int LOG2N, int C>
__global__ void __launch_bounds__(C * FftCfg<LOG2N>::T, (C * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_xinv(const __grid_constant__ XInvArgs a) {
    typedef FftCfg<LOG2N> F;
    const int c = threadIdx.x % C, tau = threadIdx.x / C;
    XchgFull<C> xc;
    cplx* const xbase = reinterpret_cast<cplx*>(MLV_SMEM_BASE());
    xc.buf = xbase;
    xc.c = c;
    cplx* stash = xbase + (size_t)F::XSLOTS * C;
    // source column tiles (C*16-byte pieces of 2nn+1 rows) are gathered by the copy engine into
    // the stash: one tensor load per ld_rows rows instead of a scattered load per row and warp
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(stash + (size_t)a.ld_rows * a.ld_boxes * C);
    unsigned ld_phase = 0;
    auto stash_fetch = [&](int f, int tile) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (unsigned)((size_t)a.ld_rows * a.ld_boxes * C * sizeof(cplx)));
            for (int b = 0; b < a.ld_boxes; ++b)
                tma_load_2d(stash + (size_t)b * a.ld_rows * C, &a.smap[f], 2 * C * tile, b * a.ld_rows, bar);
        }
    };
    const int ntiles = (a.nm + C - 1) / C;
    if (a.load_tma) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        stash_fetch(0, (int)blockIdx.x);             // flies during the prefetch announcements below
    }
    const int rmask = (1 << a.sh.rpc_shift) - 1;
    bool tma_pending = false;
    bool prefetched = a.load_tma != 0;          // the stash (being) filled belongs to field 0 of this tile
    // the CTA walks over column tiles (grid = ntiles unless the launch asks for resident CTAs only):
    // the source tile of the next trip is requested as soon as the last field of this one has left
    // the stash, so that its ~2700 tensor-load rows travel during that field's transform
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int mreal = tile * C + c;
    const bool valid = mreal < a.nm;
    const int m = valid ? mreal : a.nm - 1;          // clamp: loads stay unpredicated
    // announce the columns of the CTA that follows this one on the SM
    if (tile + a.wave < ntiles && gridDim.x == (unsigned)ntiles) {
        const int rows = 2 * a.nn + 1;
        for (int r = threadIdx.x; r < rows; r += C * F::T)
            l2_prefetch_line(a.src[0] + (size_t)r * a.spitch + (size_t)(tile + a.wave) * C);
    }
    const int mg = m + a.sh.m_off;                   // global column (spectral symbols)
    bool stash_psi = false;     // stash holds psi = -src/lap instead of the raw column
    for (int f = 0; f < a.nf; ++f) {
        cplx v[16];
        const cplx* __restrict__ src = a.src[f];
        const int op = a.op[f];
        const bool wants_psi = (op == XOP_PSI || op == XOP_UX || op == XOP_UZ);
        const bool same_prev = f > 0 && a.src[f] == a.src[f - 1];
        const bool reuse = same_prev && (!stash_psi || wants_psi);
        if (!reuse) stash_psi = false;
        const bool keep = f + 1 < a.nf && a.src[f + 1] == a.src[f];
        const int opn = keep ? a.op[f + 1] : -1;
        const bool next_wants_psi = (opn == XOP_PSI || opn == XOP_UX || opn == XOP_UZ);
        // branch-free loads (all issued before the first is consumed); truncated rows read 0
        const bool via_tma = !reuse && a.load_tma;
        if (via_tma) {
            if (!(f == 0 && prefetched)) {           // the stash still held the previous source
                __syncthreads();
                stash_fetch(f, tile);
            }
            mbar_wait(bar, ld_phase);
            ld_phase ^= 1;
        }
        prefetched = false;
        if (reuse || via_tma) {
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                if (MLV_MID(j)) { v[j] = mk(0.0, 0.0); continue; }
                int r = 0, n = 0;
                const bool ok = xrow_static<F::N>(j, tau + F::T * j, a.nn, r, n);
                const cplx t = stash[(size_t)(ok ? r : 0) * C + c];
                v[j] = ok ? t : mk(0.0, 0.0);
            }
        } else {
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                if (MLV_MID(j)) { v[j] = mk(0.0, 0.0); continue; }
                int r = 0, n = 0;
                const bool ok = xrow_static<F::N>(j, tau + F::T * j, a.nn, r, n);
                v[j] = ldg_pred(src + (size_t)(ok ? r : 0) * a.spitch + m, ok);
            }
        }
        if (keep && !reuse && !via_tma && !(wants_psi && next_wants_psi)) {      // park the raw column
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                if (MLV_MID(j)) continue;
                const int kk = tau + F::T * j;
                int r, n;
                if (xrow_static<F::N>(j, kk, a.nn, r, n)) stash[(size_t)r * C + c] = v[j];
            }
        }
        if (wants_psi) {
            const bool have_psi = reuse && stash_psi;
            const bool park_psi = keep && next_wants_psi && !have_psi;
            // ux = -(i kz m) psi, uz = (i kx n) psi   (utility.py:71,78); branch-free over the
            // points: truncated rows hold 0 and stay 0
            const double bz = a.k.kz0 * mg;
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                if (MLV_MID(j)) continue;
                const int kk = tau + F::T * j;
                int r, n;
                const bool ok = xrow_static<F::N>(j, kk, a.nn, r, n);
                cplx psi = v[j];
                if (!have_psi) {
                    double lap = lap_symbol(n, mg, a.k);
                    if (n == 0 && mg == 0) lap = 1.0;
                    const double rl = -fast_rcp(lap);             // psi = (-w)/lap
                    psi = mk(v[j].x * rl, v[j].y * rl);
                    if (park_psi && ok) stash[(size_t)r * C + c] = psi;
                }
                if (op == XOP_UX) v[j] = mk(bz * psi.y, -bz * psi.x);
                else if (op == XOP_UZ) { const double b = a.k.kx0 * n; v[j] = mk(-b * psi.y, b * psi.x); }
                else v[j] = psi;
            }
            if (park_psi) stash_psi = true;
        } else if (op != XOP_IDENT) {
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                if (MLV_MID(j)) continue;
                const int kk = tau + F::T * j;
                int r, n;
                xrow_static<F::N>(j, kk, a.nn, r, n);
                v[j] = spectral_op(op, v[j], n, mg, a.k);      // truncated rows: 0 stays 0
            }
        }
        if (a.load_tma && f == a.nf - 1 && tile + (int)gridDim.x < ntiles) {
            // last field of the tile: nobody reads the stash any more -> request the next tile now
            __syncthreads();
            stash_fetch(0, tile + (int)gridDim.x);
            prefetched = true;
        }
        if (tma_pending && threadIdx.x == 0) tma_wait_read();   // previous tile has left the buffer
        fft_line<LOG2N, true>(v, tau, a.tw, xc);
        if (a.use_tma) {
            // park the tile in the exchange buffer (dense [row][C]) and hand it to the copy
            // engine; the stores overlap the next field's loads, prologue and first butterflies
            __syncthreads();
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) xbase[(size_t)(tau + F::T * j) * C + c] = v[j];
            tma_fence_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int r0 = 0; r0 < F::N; r0 += a.tma_rows)
                    tma_store_2d(&a.tmap[f], xbase + (size_t)r0 * C, 2 * C * tile, r0);
                tma_commit();
            }
            tma_pending = true;
        } else if (valid) {
            const size_t off = (size_t)a.dstoff[f] + m;
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                const int x = tau + F::T * j;          // global row -> block of its owner
                a.out.blk[x >> a.sh.rpc_shift][off + (size_t)(x & rmask) * a.ipitch] = v[j];
            }
        }
    }
    }
    if (tma_pending && threadIdx.x == 0) tma_wait_read();       // shared memory must outlive the reads
}

// ---- column-serial form of the inverse x pass (4096-point lines, unsharded, TMA both ways).
// k_xinv runs ONE 512-thread CTA per SM (two columns side by side): its 16 warps meet at every
// barrier of every exchange, so the shared-memory phases and the fp64 phases of the SM never
// overlap.  Here a CTA is one 256-thread line (like the z stage) and two of them share an SM
// and drift apart: one is in an exchange while the other is in its butterflies.  What makes two
// CTAs fit: one column at a time (stash of 2nn+1 cplx = 44 KB instead of 87 KB) and an exchange
// buffer of exactly N slots that is also the parking area of the result (XchgLineSwz, 64 KB).
// A CTA walks over columns (persistent); the column enters through 16-byte-wide tensor loads and
// each result leaves through 16-byte-wide tensor stores (the neighbouring column's halves of the
// sectors follow within microseconds and meet them in the L2).
template <int LOG2N>
__global__ void __launch_bounds__(FftCfg<LOG2N>::T, 2)
k_xinv_cols(const __grid_constant__ XInvArgs a) {
    typedef FftCfg<LOG2N> F;
    const int tau = threadIdx.x;
    XchgLineSwz xc;
    cplx* const xbase = reinterpret_cast<cplx*>(MLV_SMEM_BASE());
    xc.buf = xbase;
    cplx* const stash = xbase + F::N;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(stash + (size_t)a.ld_rows * a.ld_boxes);
    unsigned ld_phase = 0;
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    bool tma_pending = false;
    for (int m = blockIdx.x; m < a.nm; m += gridDim.x) {
        const int mg = m + a.sh.m_off;                   // global column (spectral symbols)
        bool stash_psi = false;     // stash holds psi = -src/lap instead of the raw column
        for (int f = 0; f < a.nf; ++f) {
            cplx v[16];
            const int op = a.op[f];
            const bool wants_psi = (op == XOP_PSI || op == XOP_UX || op == XOP_UZ);
            const bool same_prev = f > 0 && a.src[f] == a.src[f - 1];
            const bool reuse = same_prev && (!stash_psi || wants_psi);
            if (!reuse) stash_psi = false;
            const bool keep = f + 1 < a.nf && a.src[f + 1] == a.src[f];
            const int opn = keep ? a.op[f + 1] : -1;
            const bool next_wants_psi = (opn == XOP_PSI || opn == XOP_UX || opn == XOP_UZ);
            if (!reuse) {
                __syncthreads();                         // every thread is done with the old stash
                if (threadIdx.x == 0) {
                    mbar_expect_tx(bar, (unsigned)((size_t)a.ld_rows * a.ld_boxes * sizeof(cplx)));
                    for (int b = 0; b < a.ld_boxes; ++b)
                        tma_load_2d(stash + (size_t)b * a.ld_rows, &a.smap[f], 2 * m, b * a.ld_rows, bar);
                }
                mbar_wait(bar, ld_phase);
                ld_phase ^= 1;
            }
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                if (MLV_MID(j)) { v[j] = mk(0.0, 0.0); continue; }
                int r = 0, n = 0;
                const bool ok = xrow_static<F::N>(j, tau + F::T * j, a.nn, r, n);
                const cplx t = stash[ok ? r : 0];
                v[j] = ok ? t : mk(0.0, 0.0);
            }
            if (wants_psi) {
                const bool have_psi = reuse && stash_psi;
                const bool park_psi = keep && next_wants_psi && !have_psi;
                const double bz = a.k.kz0 * mg;
                MLV_UNROLL
                for (int j = 0; j < 16; ++j) {
                    if (MLV_MID(j)) continue;
                    int r, n;
                    const bool ok = xrow_static<F::N>(j, tau + F::T * j, a.nn, r, n);
                    cplx psi = v[j];
                    if (!have_psi) {
                        double lap = lap_symbol(n, mg, a.k);
                        if (n == 0 && mg == 0) lap = 1.0;
                        const double rl = -fast_rcp(lap);             // psi = (-w)/lap
                        psi = mk(v[j].x * rl, v[j].y * rl);
                        if (park_psi && ok) stash[r] = psi;           // thread-private slot
                    }
                    if (op == XOP_UX) v[j] = mk(bz * psi.y, -bz * psi.x);
                    else if (op == XOP_UZ) { const double b = a.k.kx0 * n; v[j] = mk(-b * psi.y, b * psi.x); }
                    else v[j] = psi;
                }
                if (park_psi) stash_psi = true;
            } else if (op != XOP_IDENT) {
                MLV_UNROLL
                for (int j = 0; j < 16; ++j) {
                    if (MLV_MID(j)) continue;
                    int r, n;
                    xrow_static<F::N>(j, tau + F::T * j, a.nn, r, n);
                    v[j] = spectral_op(op, v[j], n, mg, a.k);      // truncated rows: 0 stays 0
                }
            }
            if (tma_pending && threadIdx.x == 0) tma_wait_read();   // previous column has left the buffer
            fft_line<LOG2N, true>(v, tau, a.tw, xc);
            // park the column in the exchange buffer (dense) and hand it to the copy engine; the
            // stores overlap the next field's prologue and first butterflies
            __syncthreads();
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) xbase[tau + F::T * j] = v[j];
            tma_fence_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int r0 = 0; r0 < F::N; r0 += a.tma_rows)
                    tma_store_2d(&a.tmap[f], xbase + r0, 2 * m, r0);
                tma_commit();
            }
            tma_pending = true;
        }
    }
    if (tma_pending && threadIdx.x == 0) tma_wait_read();       // shared memory must outlive the reads
}

// ===================================================================== x forward
struct XFwdArgs {
    int nn, nm, spitch, ipitch, nf;
    const cplx* src[MLV_XMAXF];
    int sym[MLV_XMAXF];
    double coef[MLV_XMAXF];      // real coefficient per field
    const double* symz;          // [nm] imaginary part of the z stencil symbol
    int order;                   // central x stencil: 2 or 4 (SpatialDifferentiator.py:76-104,130-185)
    double rdx;                  // 1/dx
    double scale;                // 1/(nx nz)
    int wave;                    // CTAs resident at once (prefetch distance)
    int mode;                    // 0: dst = value; 1: f0 = value + lin terms, then integrate
    cplx* dst;                   // mode 0: spectral (2nn+1, nm)
    LinTerms lin;                // mode 1
    IntegArgs integ;             // mode 1
    Shard sh;                    // nm = local valid columns; symbols use m + sh.m_off
    SpecConsts k;
    FftTw tw;
    const cplx* tws;             // SPLIT = 2: e^{-2 pi i p/nx}, p < nx/2
    int stage;                   // 1: the block of operand 0 is staged in shared memory by bulk copies
    const unsigned long long* wait_counter;   // peer stores: forward blocks of all ranks have arrived
    const unsigned long long* wait_expect;    // when *wait_counter >= *wait_expect
    int pf_tma;                  // k_xfwd_scalar: 1 = the epilogue's state / history column tiles are announced to
    int pf_rows, pf_boxes;       //    the L2 by tensor prefetches (boxes of pf_rows rows), not one hint per row
    CUtensorMap qmap, fmap;      // integ.q_in, integ.fm1: (2nn+1, 2*spitch) float64, box (pf_rows, 2*C)
};

// I (tile layout) x nf -> spectral: value = scale * FFT_x( sum_f coef_f * D_f[src_f] ), rows
// truncated to |n| <= nn, where D_f is the identity, the multiplication by the z stencil
// symbol i*symz[m] (a constant of the line) or the reference's periodic central x stencil
// (SpatialDifferentiator.py:76-104 order 2, :130-185 order 4) applied along the line
// *before* the transform -- the x stencil commutes with the z transform that produced the
// intermediate, so all fields of a right-hand side share ONE transform per column.
// The epilogue (right-hand-side assembly + time integration, Integrator.py:5-63) runs
// from the registers that hold the transform output.
// SPLIT = 2: lines of 2N points through two N-point transforms (mlv_kernels_split.cuh):
//   X[2q+s] = DFT_N( (x[p] + (-1)^s x[p+N]) e^{-2 pi i s p/2N} )[q],  s = 0, 1.
template <int LOG2N, int C, int SPLIT>
__global__ void __launch_bounds__(C * FftCfg<LOG2N>::T, (C * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_xfwd(const __grid_constant__ XFwdArgs a) {
    typedef FftCfg<LOG2N> F;
    constexpr int NF = F::N * SPLIT;                  // line length
    const int c = threadIdx.x % C, tau = threadIdx.x / C;
    const int mreal = blockIdx.x * C + c;
    const bool valid = mreal < a.nm;
    const int m = valid ? mreal : a.nm - 1;
    XchgFull<C> xc;
    cplx* const xbase = reinterpret_cast<cplx*>(MLV_SMEM_BASE());
    xc.buf = xbase;
    xc.c = c;
    const double sz = a.symz[m + a.sh.m_off];
    const int rpc = 1 << a.sh.fwd_rshift;             // rows per block of the forward buffers
    wait_arrivals(a.wait_counter, a.wait_expect);
    {   // L2 prefetch: the other fields' blocks, the next CTA's first block, and the
        // state / history columns the epilogue will read
        constexpr unsigned CHUNK = 16384;
        constexpr unsigned BLOCK = (unsigned)NF * C * (unsigned)sizeof(cplx);
        constexpr int NCH = (int)(BLOCK / CHUNK) > 0 ? (int)(BLOCK / CHUNK) : 1;
        const int t = threadIdx.x;
        if (rpc == NF && t < NCH * a.nf) {
            const int f = t / NCH, ch = t % NCH;
            const unsigned bytes = BLOCK < CHUNK ? BLOCK : CHUNK;
            if (f > 0) {
                l2_prefetch_bulk(reinterpret_cast<const char*>(a.src[f] + (size_t)blockIdx.x * NF * C) + (size_t)ch * CHUNK, bytes);
            } else if ((int)(blockIdx.x + a.wave) < (int)gridDim.x) {
                l2_prefetch_bulk(reinterpret_cast<const char*>(a.src[0] + (size_t)(blockIdx.x + a.wave) * NF * C) + (size_t)ch * CHUNK, bytes);
            }
        }
        if (a.mode == 1) {
            if (a.pf_tma) {                      // one tensor prefetch per 256 rows and array
                if (t < 2 * a.pf_boxes)
                    tma_prefetch_2d(t & 1 ? &a.fmap : &a.qmap, 2 * C * (int)blockIdx.x, (t >> 1) * a.pf_rows);
            } else {
                const int rows = 2 * a.nn + 1;
                for (int r = t; r < rows; r += C * F::T) {
                    const size_t idx = (size_t)r * a.spitch + blockIdx.x * C;
                    l2_prefetch_line(a.integ.q_in + idx);
                    l2_prefetch_line(a.integ.fm1 + idx);
                }
            }
        }
    }
    const size_t blk0 = (size_t)blockIdx.x * rpc * C + c;
    const int mg = m + a.sh.m_off;
    // the x stencil reads every element of operand 0 twice (rows x-1, x+1): its block of this
    // column tile is brought into the (still idle) exchange buffer by the copy engine -- one
    // request for the whole block instead of four dependent groups of loads per thread
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(xbase + (size_t)F::XSLOTS * C);
    if (SPLIT == 1 && a.stage) {
        if (threadIdx.x == 0) mbar_init(bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (unsigned)(NF * C * sizeof(cplx)));
            for (int x0 = 0; x0 < NF; x0 += rpc)
                bulk_load(xbase + (size_t)x0 * C,
                          a.src[0] + (size_t)(x0 >> a.sh.fwd_rshift) * a.sh.fwd_chunk + (size_t)blockIdx.x * rpc * C,
                          (unsigned)(rpc * C * sizeof(cplx)), bar);
        }
    }
    for (int s = 0; s < SPLIT; ++s) {
        cplx v[16];
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) v[j] = mk(0.0, 0.0);
        const double sg = s ? -1.0 : 1.0;
        for (int f = 0; f < a.nf; ++f) {
            const cplx* __restrict__ src = a.src[f];
            const int sym = a.sym[f];
            const double cf = a.coef[f];
            // element x (global row, periodic) of this thread's column: block of the row owner
            const int rshift = a.sh.fwd_rshift;
            const size_t chunk = (size_t)a.sh.fwd_chunk;
            auto at = [=](int x) -> cplx {
                x &= NF - 1;
                return src[(size_t)(x >> rshift) * chunk + blk0 + (size_t)(x & (rpc - 1)) * C];
            };
            const double w1 = a.order == 2 ? 0.5 * a.rdx : 2.0 / 3.0 * a.rdx;
            const double w2 = a.order == 2 ? 0.0 : -0.25 / 3.0 * a.rdx;
            // D_f at row x, without the coefficient
            auto d2 = [=](int x) -> cplx {                    // central stencil, order 2
                const cplx p = at(x + 1), q = at(x - 1);
                return mk(w1 * (p.x - q.x), w1 * (p.y - q.y));
            };
            auto d4 = [=](int x) -> cplx {                    // central stencil, order 4
                const cplx p1 = at(x + 1), q1 = at(x - 1), p2 = at(x + 2), q2 = at(x - 2);
                return mk(fma(w1, p1.x - q1.x, w2 * (p2.x - q2.x)), fma(w1, p1.y - q1.y, w2 * (p2.y - q2.y)));
            };
            // the pair every advected scalar produces (Variable.py:119-128): d/dx of src[f] by the
            // order-2 stencil + d/dz of src[f+1] by its symbol -- one fused sweep, so that the
            // loads of both operands are in flight together
            if (SPLIT == 1 && sym == XSYM_FDX && a.order == 2 && f + 1 < a.nf && a.sym[f + 1] == XSYM_FDZ) {
                const cplx* __restrict__ srcb = a.src[f + 1];
                const double ci = sz * a.coef[f + 1], cw = cf * w1;
                if (SPLIT == 1 && a.stage && f == 0) {
                    // second operand straight from global memory (16 loads in flight per thread)
                    // while the staged block arrives
                    MLV_UNROLL
                    for (int j = 0; j < 16; ++j) {
                        const int x = tau + F::T * j;
                        const cplx b = srcb[(size_t)(x >> rshift) * chunk + blk0 + (size_t)(x & (rpc - 1)) * C];
                        v[j] = mk(fma(-ci, b.y, v[j].x), fma(ci, b.x, v[j].y));
                    }
                    mbar_wait(bar, 0);
                    MLV_UNROLL
                    for (int j = 0; j < 16; ++j) {
                        const int x = tau + F::T * j;
                        const cplx p = xbase[(size_t)((x + 1) & (NF - 1)) * C + c];
                        const cplx q = xbase[(size_t)((x - 1) & (NF - 1)) * C + c];
                        v[j] = mk(fma(cw, p.x - q.x, v[j].x), fma(cw, p.y - q.y, v[j].y));
                    }
                    ++f;
                    continue;
                }
                MLV_UNROLL
                for (int j0 = 0; j0 < 16; j0 += 4) {
                    MLV_SCHED_FENCE();
                    const int tq = opaque_int(tau);
                    cplx p[4], q[4], b[4];
                    MLV_UNROLL
                    for (int u = 0; u < 4; ++u) {
                        const int x = tq + F::T * (j0 + u);
                        p[u] = at(x + 1);
                        q[u] = at(x - 1);
                        b[u] = srcb[(size_t)(x >> rshift) * chunk + blk0 + (size_t)(x & (rpc - 1)) * C];
                    }
                    MLV_UNROLL
                    for (int u = 0; u < 4; ++u) {
                        const int j = j0 + u;
                        v[j] = mk(fma(cw, p[u].x - q[u].x, fma(-ci, b[u].y, v[j].x)),
                                  fma(cw, p[u].y - q[u].y, fma(ci, b[u].x, v[j].y)));
                    }
                }
                ++f;
                continue;
            }
            // groups of 4 points: bounds the number of loads in flight (registers)
            MLV_UNROLL
            for (int j0 = 0; j0 < 16; j0 += 4) {
                MLV_SCHED_FENCE();
                const int tq = opaque_int(tau);      // address arithmetic stays inside the group
                if (sym == XSYM_FDX) {
                    if (a.order == 2) {
                        MLV_UNROLL
                        for (int j = j0; j < j0 + 4; ++j) {
                            const int x = tq + F::T * j;
                            cplx t = d2(x);
                            if constexpr (SPLIT == 2) { const cplx u = d2(x + F::N); t = mk(t.x + sg * u.x, t.y + sg * u.y); }
                            v[j] = mk(fma(cf, t.x, v[j].x), fma(cf, t.y, v[j].y));
                        }
                    } else {
                        MLV_UNROLL
                        for (int j = j0; j < j0 + 4; ++j) {
                            const int x = tq + F::T * j;
                            cplx t = d4(x);
                            if constexpr (SPLIT == 2) { const cplx u = d4(x + F::N); t = mk(t.x + sg * u.x, t.y + sg * u.y); }
                            v[j] = mk(fma(cf, t.x, v[j].x), fma(cf, t.y, v[j].y));
                        }
                    }
                } else {
                    // identity, or * (i sz) for the z stencil symbol
                    const double cr = sym == XSYM_FDZ ? 0.0 : cf, ci = sym == XSYM_FDZ ? sz * cf : 0.0;
                    MLV_UNROLL
                    for (int j = j0; j < j0 + 4; ++j) {
                        const int x = tq + F::T * j;
                        cplx t = at(x);
                        if constexpr (SPLIT == 2) { const cplx u = at(x + F::N); t = mk(t.x + sg * u.x, t.y + sg * u.y); }
                        v[j] = mk(fma(cr, t.x, fma(-ci, t.y, v[j].x)), fma(cr, t.y, fma(ci, t.x, v[j].y)));
                    }
                }
            }
        }
        MLV_SCHED_FENCE();
        if constexpr (SPLIT == 2) {
            if (s) {
                const int tq = opaque_int(tau);
                MLV_UNROLL
                for (int j = 0; j < 16; ++j) v[j] = cmul(v[j], a.tws[tq + F::T * j]);   // e^{-2 pi i p/NF}
            }
        }
        fft_line<LOG2N, false>(v, tau, a.tw, xc);
        if (!valid) continue;                                 // no barrier below in this trip
        // ---- epilogue from registers: UN outputs per trip, all global loads of a trip issued
        //      before the arithmetic (memory-level parallelism)
        constexpr int UN = 2;
        MLV_UNROLL
        for (int j0 = 0; j0 < 16; j0 += UN) {
            if (SPLIT == 1 && MLV_MID(j0) && MLV_MID(j0 + UN - 1)) continue;       // always truncated
            cplx q[UN], f1[UN];
            size_t idx[UN];
            int nmode[UN];
            bool ok[UN];
            MLV_SCHED_FENCE();
            const int tq = opaque_int(tau);
            MLV_UNROLL
            for (int u = 0; u < UN; ++u) {
                int r = 0, n = 0;
                if constexpr (SPLIT == 1) ok[u] = xrow_static<NF>(j0 + u, tq + F::T * (j0 + u), a.nn, r, n);
                else ok[u] = xrow_of(SPLIT * (tq + F::T * (j0 + u)) + s, NF, a.nn, r, n);
                nmode[u] = n;
                idx[u] = ok[u] ? (size_t)r * a.spitch + m : (size_t)m;
            }
            if (a.mode == 1) {
                MLV_UNROLL
                for (int u = 0; u < UN; ++u) {
                    q[u] = a.integ.q_in[idx[u]];
                    f1[u] = a.integ.fm1[idx[u]];
                }
            }
            MLV_UNROLL
            for (int u = 0; u < UN; ++u) {
                if (!ok[u]) continue;
                const cplx t = cscale(v[j0 + u], a.scale);
                cplx f2 = mk(0.0, 0.0), f3 = f2;
                if (a.mode == 1 && a.integ.ab_order == 4) { f2 = a.integ.fm2[idx[u]]; f3 = a.integ.fm3[idx[u]]; }
                if (a.mode == 0) {
                    a.dst[idx[u]] = t;
                    continue;
                }
                const cplx f0 = cadd(t, lin_terms_at(a.lin, idx[u], nmode[u], mg, a.k));
                a.integ.f0[idx[u]] = f0;
                a.integ.q_out[idx[u]] = integrate_value(a.integ, f0, q[u], f1[u], f2, f3, idx[u], nmode[u], mg, a.k);
            }
        }
    }
}

// ---- specialised forward x pass of the single-scalar step (Kelvin-Helmholtz / Taylor-Green loop,
// examples/kelvin_helmholtz_instability.py:115-131): one GPU, two operands (d/dx by the order-2
// stencil, d/dz by its symbol), staged stencil operand, no extra linear terms, AB2 + theta-scheme with
// the symbolic Laplacian.  Same arithmetic, in the same order, as the generic k_xfwd<.., 1> on that
// path; everything the generic kernel decides at run time is fixed here, which frees enough
// registers for UN epilogue outputs per trip (fewer exposed load round trips).
template <int LOG2N, int C, int UN>
__global__ void __launch_bounds__(C * FftCfg<LOG2N>::T, (C * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_xfwd_scalar(const __grid_constant__ XFwdArgs a) {
    typedef FftCfg<LOG2N> F;
    constexpr int NF = F::N;
    const int c = threadIdx.x % C, tau = threadIdx.x / C;
    XchgFull<C> xc;
    cplx* const xbase = reinterpret_cast<cplx*>(MLV_SMEM_BASE());
    xc.buf = xbase;
    xc.c = c;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(xbase + (size_t)F::XSLOTS * C);
    const int ntiles = (a.nm + C - 1) / C;
    const bool persistent = gridDim.x != (unsigned)ntiles;
    auto stage_fetch = [&](int tile) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, (unsigned)(NF * C * sizeof(cplx)));
            bulk_load(xbase, a.src[0] + (size_t)tile * NF * C, (unsigned)(NF * C * sizeof(cplx)), bar);
        }
    };
    // L2 prefetch of what tile `tl` needs beyond its staged operand: the second operand (unless
    // `first_only`) and the state / history columns of its epilogue; `first_only`: the staged operand
    auto announce = [&](int tl, bool first_only) {
        constexpr unsigned CHUNK = 16384;
        constexpr unsigned BLOCK = (unsigned)NF * C * (unsigned)sizeof(cplx);
        constexpr int NCH = (int)(BLOCK / CHUNK) > 0 ? (int)(BLOCK / CHUNK) : 1;
        const int t = threadIdx.x;
        const unsigned bytes = BLOCK < CHUNK ? BLOCK : CHUNK;
        if (t < NCH)
            l2_prefetch_bulk(reinterpret_cast<const char*>(a.src[first_only ? 0 : 1] + (size_t)tl * NF * C) + (size_t)t * CHUNK, bytes);
        if (first_only) return;
        if (a.pf_tma) {
            if (t < 2 * a.pf_boxes) tma_prefetch_2d(t & 1 ? &a.fmap : &a.qmap, 2 * C * tl, (t >> 1) * a.pf_rows);
            return;
        }
        const int rows = 2 * a.nn + 1;
        for (int r = t; r < rows; r += C * F::T) {
            const size_t idx = (size_t)r * a.spitch + (size_t)tl * C;
            l2_prefetch_line(a.integ.q_in + idx);
            l2_prefetch_line(a.integ.fm1 + idx);
        }
    };
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    stage_fetch((int)blockIdx.x);
    announce((int)blockIdx.x, false);
    unsigned phase = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int mreal = tile * C + c;
    const bool valid = mreal < a.nm;
    const int m = valid ? mreal : a.nm - 1;
    {
        const int nxt = persistent ? tile + (int)gridDim.x : tile + a.wave;
        if (nxt < ntiles) announce(nxt, !persistent);      // resident CTAs: everything of their next tile
    }
    const double w1 = 0.5 * a.rdx;
    const double ci = a.symz[m] * a.coef[1], cw = a.coef[0] * w1;
    const cplx* __restrict__ srcb = a.src[1] + (size_t)tile * NF * C + c;
    cplx v[16];
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) {
        const cplx b = srcb[(size_t)(tau + F::T * j) * C];
        v[j] = mk(fma(-ci, b.y, 0.0), fma(ci, b.x, 0.0));
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) {
        const int x = tau + F::T * j;
        const cplx p = xbase[(size_t)((x + 1) & (NF - 1)) * C + c];
        const cplx q = xbase[(size_t)((x - 1) & (NF - 1)) * C + c];
        v[j] = mk(fma(cw, p.x - q.x, v[j].x), fma(cw, p.y - q.y, v[j].y));
    }
    MLV_SCHED_FENCE();
    fft_line<LOG2N, false>(v, tau, a.tw, xc);
    if (tile + (int)gridDim.x < ntiles) {
        __syncthreads();                                  // the transform has left the buffer
        stage_fetch(tile + (int)gridDim.x);               // travels during the epilogue
    }
    if (!valid) continue;
    // ---- epilogue from registers (Integrator.py:5-18 AB2, :58-63 theta-scheme with L = lcoef * lap)
    const double h = a.integ.dt / 2;
    const double c1 = (1 - a.integ.alpha) * a.integ.dt, c2 = a.integ.alpha * a.integ.dt;
    // the 12 registers that can hold retained modes: i = 0..11 -> j = 0..5, 10..15
    MLV_UNROLL
    for (int i0 = 0; i0 < 12; i0 += UN) {
        cplx q[UN], f1[UN];
        size_t idx[UN];
        int nmode[UN];
        bool ok[UN];
        MLV_SCHED_FENCE();
        const int tq = opaque_int(tau);
        MLV_UNROLL
        for (int u = 0; u < UN; ++u) {
            const int j = (i0 + u) < 6 ? (i0 + u) : (i0 + u) + 4;
            int r = 0, n = 0;
            ok[u] = xrow_static<NF>(j, tq + F::T * j, a.nn, r, n);
            nmode[u] = n;
            idx[u] = ok[u] ? (size_t)r * a.spitch + m : (size_t)m;
        }
        MLV_UNROLL
        for (int u = 0; u < UN; ++u) {
            q[u] = a.integ.q_in[idx[u]];
            f1[u] = a.integ.fm1[idx[u]];
        }
        MLV_UNROLL
        for (int u = 0; u < UN; ++u) {
            if (!ok[u]) continue;
            const int j = (i0 + u) < 6 ? (i0 + u) : (i0 + u) + 4;
            const cplx f0 = cadd(cscale(v[j], a.scale), mk(0.0, 0.0));      // (+ the empty sum of linear terms)
            a.integ.f0[idx[u]] = f0;
            const cplx inc = mk(h * (3 * f0.x - f1[u].x), h * (3 * f0.y - f1[u].y));
            const double L = a.integ.lcoef * lap_symbol(nmode[u], m, a.k);
            const double aa = 1 + c1 * L;
            const double rb = fast_rcp(1 - c2 * L);
            a.integ.q_out[idx[u]] = mk((aa * q[u].x + inc.x) * rb, (aa * q[u].y + inc.y) * rb);
        }
    }
    }
}

// NaN-propagating maximum (numpy.max semantics)
MLV_DEV double nan_max(double x, double y) { return (x != x || y != y) ? NAN : fmax(x, y); }

// ===================================================================== z passes
// Packed-pair element idx of the Hermitian-extended spectrum of two real rows
// whose one-sided spectra are rowA, rowB (F4: Im of the m=0 bin is dropped):
//   idx < nm        : A[idx] + i B[idx]            (idx = 0: Re A + i Re B)
//   idx > N - nm    : conj(A[N-idx]) + i conj(B[N-idx])
//   otherwise       : 0   (2/3-rule truncation)
template <int LOG2N, bool SHARDED>
MLV_DEV void zpair_load_line_(cplx (&v)[16], const cplx* __restrict__ rowA,
                              const cplx* __restrict__ rowB, int tau_, int nm, const Shard& sh) {
    typedef FftCfg<LOG2N> F;
    const int tau = opaque_int(tau_);
    // branch-free: the loads are all issued before the first one is consumed.  By the 2/3-rule
    // invariant (MLV_MID) registers j <= 5 can only hold low modes (A + iB) and j >= 10 only the
    // mirrored high ones (conj A + i conj B): which form applies is known at compile time.
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) {
        if (MLV_MID(j)) { v[j] = mk(0.0, 0.0); continue; }
        const int idx = tau + F::T * j;
        if (j <= 5) {
            const bool ok = idx < nm;
            const int mm = ok ? idx : 0;
            const size_t off = SHARDED ? inv_col_off(mm, sh) : (size_t)mm;
            const cplx A = ldg_pred(rowA + off, ok), B = ldg_pred(rowB + off, ok);
            if (j == 0) {                                  // F4: Im of bin 0 dropped
                const bool z = idx == 0;
                v[j] = mk(A.x - (z ? 0.0 : B.y), (z ? 0.0 : A.y) + B.x);
            } else {
                v[j] = mk(A.x - B.y, A.y + B.x);
            }
        } else {
            const bool ok = idx > F::N - nm;
            const int mm = ok ? F::N - idx : 0;
            const size_t off = SHARDED ? inv_col_off(mm, sh) : (size_t)mm;
            const cplx A = ldg_pred(rowA + off, ok), B = ldg_pred(rowB + off, ok);
            v[j] = mk(A.x + B.y, B.x - A.y);
        }
    }
}
template <int LOG2N>
MLV_DEV void zpair_load_line(cplx (&v)[16], const cplx* __restrict__ rowA,
                             const cplx* __restrict__ rowB, int tau, int nm, const Shard& sh) {
    if (sh.inv_chunk == 0) zpair_load_line_<LOG2N, false>(v, rowA, rowB, tau, nm, sh);   // one block per row
    else zpair_load_line_<LOG2N, true>(v, rowA, rowB, tau, nm, sh);
}

// After a forward transform of z = a + ib every thread holds Zf[tau + T j].
// Publishes the upper part (index > N-nm) to `pbuf` so that the owner of
// output k can read Zf[N-k] from pbuf[k].  Caller must __syncthreads() between
// publish and unpack (and before pbuf is reused).
template <int LOG2N>
MLV_DEV void zpair_publish(const cplx (&v)[16], int tau, int nm, cplx* pbuf) {
    typedef FftCfg<LOG2N> F;
    MLV_UNROLL
    for (int j = 10; j < 16; ++j) {                      // retained high modes: j >= 10
        const int kk = tau + F::T * j;
        if (kk > F::N - nm) pbuf[F::N - kk] = v[j];
    }
}
// One-sided spectra of the two rows at index k (k < nm), given Zk and P = Zf[N-k]
MLV_DEV void zpair_unpack(cplx Zk, cplx P, cplx& A, cplx& B) {
    A = mk(0.5 * (Zk.x + P.x), 0.5 * (Zk.y - P.y));
    B = mk(0.5 * (Zk.y + P.y), -0.5 * (Zk.x - P.x));
}

struct ZArgs {
    int nx, nm, ipitch, ct;      // nx = local rows, nm = global retained columns
    Shard sh;
    const cplx* I;     // c2r input / unused
    cplx* Iout;        // r2c output
    const double* Pin; // r2c input
    double* P;         // c2r output (nx, nz)
    FftTw tw;
};

// I (nx, ipitch) -> P (nx, nz): inverse z pass, two rows per line
template <int LOG2N, int LPC>
__global__ void __launch_bounds__(LPC * FftCfg<LOG2N>::T, (LPC * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_z_c2r(const ZArgs a) {
    typedef FftCfg<LOG2N> F;
    const int l = threadIdx.x / F::T, tau = threadIdx.x % F::T;
    const int rpreal = blockIdx.x * LPC + l;
    const bool valid = 2 * rpreal < a.nx;
    const int rp = valid ? rpreal : 0;               // clamp: loads stay unpredicated
    XchgSplit xc;
    xc.buf = reinterpret_cast<double*>(MLV_SMEM_BASE()) + (size_t)l * F::XSLOTS;
    const cplx* rowA = a.I + (size_t)(2 * rp) * a.ipitch;
    const cplx* rowB = rowA + a.ipitch;
    cplx v[16];
    zpair_load_line<LOG2N>(v, rowA, rowB, tau, a.nm, a.sh);
    fft_line<LOG2N, true>(v, tau, a.tw, xc);
    if (valid) {
        double* pa = a.P + (size_t)(2 * rp) * F::N;
        double* pb = pa + F::N;
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) {
            pa[tau + F::T * j] = v[j].x;
            pb[tau + F::T * j] = v[j].y;
        }
    }
}

// P (nx, nz) -> I (tile layout): forward z pass (unnormalised), truncated to m < nm
template <int LOG2N, int LPC>
__global__ void __launch_bounds__(LPC * FftCfg<LOG2N>::T, (LPC * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_z_r2c(const ZArgs a) {
    typedef FftCfg<LOG2N> F;
    const int l = threadIdx.x / F::T, tau = threadIdx.x % F::T;
    const int rp = blockIdx.x * LPC + l;
    const bool valid = 2 * rp < a.nx;
    XchgSplit xc;
    xc.buf = reinterpret_cast<double*>(MLV_SMEM_BASE()) + (size_t)l * F::XSLOTS;
    cplx v[16];
    {
        const double* pa = a.Pin + (size_t)(2 * rp) * F::N;
        const double* pb = pa + F::N;
        MLV_UNROLL
        for (int j = 0; j < 16; ++j)
            v[j] = valid ? mk(pa[tau + F::T * j], pb[tau + F::T * j]) : mk(0.0, 0.0);
    }
    fft_line<LOG2N, false>(v, tau, a.tw, xc);
    cplx* pbuf = reinterpret_cast<cplx*>(xc.buf);
    __syncthreads();
    zpair_publish<LOG2N>(v, tau, a.nm, pbuf);
    __syncthreads();
    if (valid) {
        MLV_UNROLL
        for (int j = 0; j < 6; ++j) {                        // retained low modes: j <= 5
            const int kk = tau + F::T * j;
            if (kk < a.nm) {
                const cplx P = (kk == 0) ? v[j] : pbuf[kk];
                cplx A, B;
                zpair_unpack(v[j], P, A, B);
                cplx* o = a.Iout + fwd_off(2 * rp, kk, a.ct, a.sh);
                o[0] = A;
                o[a.ct] = B;                                   // row 2rp+1
            }
        }
    }
}

// ============================================================ 1-D x transforms
// Fourier-x / finite-difference-z mode (SpectralTransformer.py:33-88): spectral
// arrays are (nn, nz) with modes n = 0..nn-1, the transform runs along x only and
// is real-to-complex.  Two adjacent z columns are packed into one complex line.
struct X1dArgs {
    int nn, nz;
    const cplx* S;       // c2r input  (nn, nz)
    double* P;           // c2r output (nx, nz)
    const double* Pin;   // r2c input
    cplx* Sout;          // r2c output
    double scale;        // r2c: 1/nx
    FftTw tw;
};

template <int LOG2N, int C>
__global__ void __launch_bounds__(C * FftCfg<LOG2N>::T, (C * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_x1d_c2r(const X1dArgs a) {
    typedef FftCfg<LOG2N> F;
    const int c = threadIdx.x % C, tau = threadIdx.x / C;
    const int z0 = 2 * (blockIdx.x * C + c);
    const bool valid = z0 < a.nz, has2 = z0 + 1 < a.nz;
    XchgFull<C> xc;
    xc.buf = reinterpret_cast<cplx*>(MLV_SMEM_BASE());
    xc.c = c;
    cplx v[16];
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) {
        const int kk = tau + F::T * j;
        v[j] = mk(0.0, 0.0);
        if (!valid || MLV_MID(j)) continue;
        if (kk < a.nn) {
            const cplx A = a.S[(size_t)kk * a.nz + z0];
            const cplx B = has2 ? a.S[(size_t)kk * a.nz + z0 + 1] : mk(0.0, 0.0);
            v[j] = (kk == 0) ? mk(A.x, B.x) : mk(A.x - B.y, A.y + B.x);
        } else if (kk > F::N - a.nn) {
            const int mm = F::N - kk;
            const cplx A = a.S[(size_t)mm * a.nz + z0];
            const cplx B = has2 ? a.S[(size_t)mm * a.nz + z0 + 1] : mk(0.0, 0.0);
            v[j] = mk(A.x + B.y, B.x - A.y);
        }
    }
    fft_line<LOG2N, true>(v, tau, a.tw, xc);
    if (valid) {
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) {
            double* o = a.P + (size_t)(tau + F::T * j) * a.nz + z0;
            o[0] = v[j].x;
            if (has2) o[1] = v[j].y;
        }
    }
}

template <int LOG2N, int C>
__global__ void __launch_bounds__(C * FftCfg<LOG2N>::T, (C * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_x1d_r2c(const X1dArgs a) {
    typedef FftCfg<LOG2N> F;
    const int c = threadIdx.x % C, tau = threadIdx.x / C;
    const int z0 = 2 * (blockIdx.x * C + c);
    const bool valid = z0 < a.nz, has2 = z0 + 1 < a.nz;
    XchgFull<C> xc;
    xc.buf = reinterpret_cast<cplx*>(MLV_SMEM_BASE());
    xc.c = c;
    cplx v[16];
    MLV_UNROLL
    for (int j = 0; j < 16; ++j) {
        v[j] = mk(0.0, 0.0);
        if (valid) {
            const double* in = a.Pin + (size_t)(tau + F::T * j) * a.nz + z0;
            v[j] = mk(in[0], has2 ? in[1] : 0.0);
        }
    }
    fft_line<LOG2N, false>(v, tau, a.tw, xc);
    __syncthreads();
    MLV_UNROLL
    for (int j = 10; j < 16; ++j) {
        const int kk = tau + F::T * j;
        if (kk > F::N - a.nn) xc.buf[(size_t)(F::N - kk) * C + c] = v[j];
    }
    __syncthreads();
    if (valid) {
        MLV_UNROLL
        for (int j = 0; j < 6; ++j) {
            const int kk = tau + F::T * j;
            if (kk < a.nn) {
                const cplx P = (kk == 0) ? v[j] : xc.buf[(size_t)kk * C + c];
                cplx A, B;
                zpair_unpack(v[j], P, A, B);
                cplx* o = a.Sout + (size_t)kk * a.nz + z0;
                o[0] = cscale(A, a.scale);
                if (has2) o[1] = cscale(B, a.scale);
            }
        }
    }
}

// ============================================================ fused z stage
// The physical-space stage of Variable.vec_dot_nabla (Variable.py:119-128):
//   inverse z pass of ux, uz, q  ->  A = ux*q, B = uz*q  ->  forward z pass of
//   A and B (truncated to m < nm).
// The two derivatives d/dx(A) + d/dz(B) of the conservative form are applied
// later as the exact Fourier symbols of the reference's central stencils
// (SpatialDifferentiator.py:76-185) in the forward x pass.
// Also produces the reductions the tickers need (Integrator.py:35-44 signed max
// of ux, uz; utility.py:42-59 sum ux^2, uz^2) as per-CTA partials.
struct ZAdvArgs {
    int nx, nm, ipitch, ct;        // nx = local rows, nm = global retained columns
    int row0, nrows;               // this launch covers local rows [row0, row0 + nrows) (both even)
    Shard sh;
    int wave;                      // CTAs resident at once (prefetch distance)
    const cplx* Iux;
    const cplx* Iuz;
    const cplx* Iq;
    PeerBlocks out;                // destination block per tile owner; IA/IB given as offsets
    const unsigned long long* wait_counter;   // peer stores: inverse blocks of all ranks have arrived
    const unsigned long long* wait_expect;
    long long outoff[2];
    cplx* IA;                      // out: z-spectrum of ux q   (tile layout)
    cplx* IB;                      // out: z-spectrum of uz q
    double* red;                   // [gridDim.x][4] partials: max ux, max uz, sum ux^2, sum uz^2
    FftTw tw;
    const cplx* tws;               // real-row lines: e^{-2 pi i k/nz}, k < nz/2
};

// RED = false: no ticker reads the reductions of this step (mlv_set_reductions): the maxima / sums
// and their shared-memory tree are compiled out (a run-time test instead costs registers this
// kernel does not have: spills 88 -> 324 bytes, 0.255 -> 0.270 ms)
// SHARDED = false: one rank, one row block (the single-GPU step): no owner look-ups, contiguous rows
template <int LOG2N, int LPC, bool RED, bool SHARDED>
__global__ void __launch_bounds__(LPC * FftCfg<LOG2N>::T, (LPC * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_z_advect(const ZAdvArgs a) {
    typedef FftCfg<LOG2N> F;
    constexpr int NT = LPC * F::T;
    const int l = threadIdx.x / F::T, tau = threadIdx.x % F::T;
    const int rplocal = blockIdx.x * LPC + l;
    const bool valid = 2 * rplocal < a.nrows;
    const int rp = a.row0 / 2 + (valid ? rplocal : 0);   // clamp: loads stay unpredicated
    // shared memory: [ LPC*XSLOTS doubles exchange | LPC*N cplx thread-private stash | 4*NT doubles ]
    unsigned char* base = MLV_SMEM_BASE();
    XchgSplit xc;
    xc.buf = reinterpret_cast<double*>(base) + (size_t)l * F::XSLOTS;
    cplx* stash = reinterpret_cast<cplx*>(base + (size_t)LPC * F::XSLOTS * sizeof(double)) +
                  (size_t)l * F::N + tau;
    double* rbuf = reinterpret_cast<double*>(base + (size_t)LPC * F::XSLOTS * sizeof(double) +
                                             (size_t)LPC * F::N * sizeof(cplx));
    const size_t rowoff = (size_t)(2 * rp) * a.ipitch;
    const int cts = log2_pow2(a.ct);                 // tile width is a power of two
    constexpr bool sharded = SHARDED;
    if constexpr (SHARDED) wait_arrivals(a.wait_counter, a.wait_expect);

    // announce the rows of the two velocity components (needed one and two transforms
    // from now) and the scalar rows of the CTA that will follow this one on the SM
    if (tau < 6 && (!SHARDED || a.sh.nml >= a.nm)) {           // rows are contiguous only when unsharded
        const unsigned rowbytes = (unsigned)a.nm * (unsigned)sizeof(cplx);
        if (tau < 4) {
            l2_prefetch_bulk((tau < 2 ? a.Iux : a.Iuz) + rowoff + (size_t)(tau & 1) * a.ipitch, rowbytes);
        } else {
            const int rpn = rp + a.wave * LPC;
            if (2 * rpn < a.nx)
                l2_prefetch_bulk(a.Iq + (size_t)(2 * rpn + (tau & 1)) * a.ipitch, rowbytes);
        }
    }

    cplx v[16];
    {   // q -> physical, parked in the thread-private stash
        const cplx* rowA = a.Iq + rowoff;
        zpair_load_line_<LOG2N, SHARDED>(v, rowA, rowA + a.ipitch, tau, a.nm, a.sh);
        fft_line<LOG2N, true>(v, tau, a.tw, xc);
        MLV_UNROLL
        for (int j = 0; j < 16; ++j) stash[j * F::T] = v[j];
    }
    for (int pass = 0; pass < 2; ++pass) {        // pass 0: A = ux q, pass 1: B = uz q
        MLV_SCHED_FENCE();
        {
            const cplx* src = (pass == 0 ? a.Iux : a.Iuz) + rowoff;
            zpair_load_line_<LOG2N, SHARDED>(v, src, src + a.ipitch, tau, a.nm, a.sh);
        }
        MLV_SCHED_FENCE();
        fft_line<LOG2N, true>(v, tau, a.tw, xc);
        if constexpr (RED) {
            double mx = -INFINITY, ss = 0.0;
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                mx = fmax(mx, fmax(v[j].x, v[j].y));
                ss += v[j].x * v[j].x + v[j].y * v[j].y;
                const cplx q = stash[j * F::T];
                v[j] = mk(v[j].x * q.x, v[j].y * q.y);
            }
            // fmax drops NaNs; the sum of squares does not: a NaN anywhere in the row makes the
            // maximum NaN too, as numpy.max does (Integrator.py:41 tests np.isnan(cfl_dt))
            rbuf[pass * NT + threadIdx.x] = valid ? (ss != ss ? NAN : mx) : -INFINITY;
            rbuf[(2 + pass) * NT + threadIdx.x] = valid ? ss : 0.0;
        } else {                                 // no ticker reads the reductions of this step
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) {
                const cplx q = stash[j * F::T];
                v[j] = mk(v[j].x * q.x, v[j].y * q.y);
            }
        }
        fft_line<LOG2N, false>(v, tau, a.tw, xc);
        cplx* pbuf = reinterpret_cast<cplx*>(xc.buf);
        __syncthreads();
        zpair_publish<LOG2N>(v, tau, a.nm, pbuf);
        __syncthreads();
        if (valid) {
            const size_t foff = (size_t)a.outoff[pass];
            MLV_UNROLL
            for (int j0 = 0; j0 < 6; j0 += 3) {            // retained low modes: j <= 5
                cplx P[3];
                MLV_UNROLL
                for (int u = 0; u < 3; ++u) {              // partner values: unconditional reads
                    const int kk = tau + F::T * (j0 + u);
                    P[u] = pbuf[kk < a.nm ? kk : 0];
                }
                MLV_UNROLL
                for (int u = 0; u < 3; ++u) {
                    const int j = j0 + u;
                    const int kk = tau + F::T * j;
                    if (kk < a.nm) {
                        cplx A, B;
                        zpair_unpack(v[j], (kk == 0) ? v[j] : P[u], A, B);
                        const int t = kk >> cts;
                        cplx* o;
                        if constexpr (sharded) {
                            const int h = t / a.sh.tpr, tl = t - h * a.sh.tpr;       // tile owner
                            o = a.out.blk[h] + foff + fwd_store_off(2 * rp, tl, kk & (a.ct - 1), a.ct, a.sh);
                        } else {                                                 // [tile][nx][ct]
                            o = a.IA + foff + ((((size_t)t) << a.sh.fwd_rshift) + (size_t)(2 * rp)) * a.ct + (kk & (a.ct - 1));
                        }
                        o[0] = A;
                        o[a.ct] = B;                               // row 2rp+1
                    }
                }
            }
        }
    }
    // ---- reductions: per-CTA partials (deterministic two-stage reduction)
    if constexpr (RED) {
    __syncthreads();
    // tree over the 4 x NT table: thread t reduces column block of quantity w = t / (NT/4)
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
                const double x = rbuf[w * NT + g], y = rbuf[w * NT + g + s2];
                rbuf[w * NT + g] = (w < 2) ? nan_max(x, y) : x + y;
            }
        }
        if (w < 4 && g == 0) a.red[(size_t)blockIdx.x * 4 + w] = rbuf[w * NT];
    }
    }
}

// ---- persistent form with grouped transforms (line lengths 512 .. 4096: three passes).
// Same arithmetic as k_z_advect; differences:
//  * the three inverse transforms leave their result in grouped order and the two forward
//    transforms start from it (mlv_fft.cuh: fft_nat2grp / fft_grp2nat): one CTA-wide exchange per
//    transform instead of two, the other one is a 16x16 transpose inside half a warp;
//  * a CTA walks over row pairs (grid = resident CTAs), keeps the CFL / energy partials in
//    registers and reduces them once, and announces the rows of its next pair to the L2 while it
//    works on the current one.
template <int LOG2N, int LPC>
__global__ void __launch_bounds__(LPC * FftCfg<LOG2N>::T, (LPC * FftCfg<LOG2N>::T <= 256) ? 2 : 1)
k_z_advect_grouped(const ZAdvArgs a) {
    typedef FftCfg<LOG2N> F;
    constexpr int NT = LPC * F::T;
    const int l = threadIdx.x / F::T, tau = threadIdx.x % F::T;
    unsigned char* base = MLV_SMEM_BASE();
    double* const xbuf = reinterpret_cast<double*>(base) + (size_t)l * F::XSLOTS;
    cplx* stash = reinterpret_cast<cplx*>(base + (size_t)LPC * F::XSLOTS * sizeof(double)) +
                  (size_t)l * F::N + tau;
    double* rbuf = reinterpret_cast<double*>(base + (size_t)LPC * F::XSLOTS * sizeof(double) +
                                             (size_t)LPC * F::N * sizeof(cplx));
    const int cts = log2_pow2(a.ct);                 // tile width is a power of two
    const bool sharded = a.sh.fwd_chunk != 0;
    const int npairs = a.nrows / 2;
    const int stride = (int)gridDim.x * LPC;
    wait_arrivals(a.wait_counter, a.wait_expect);

    double acc_mx[2] = {-INFINITY, -INFINITY}, acc_ss[2] = {0.0, 0.0};
    for (int rp0 = (int)blockIdx.x * LPC; rp0 < npairs; rp0 += stride) {     // uniform trip count per CTA
        const int rplocal = rp0 + l;
        const bool valid = rplocal < npairs;
        const int rp = a.row0 / 2 + (valid ? rplocal : 0);   // clamp: loads stay unpredicated
        const size_t rowoff = (size_t)(2 * rp) * a.ipitch;
        // announce the velocity rows of this pair (needed one and two transforms from now) and the
        // scalar rows of the pair this line works on next
        if (tau < 6 && a.sh.nml >= a.nm) {           // rows are contiguous only when unsharded
            const unsigned rowbytes = (unsigned)a.nm * (unsigned)sizeof(cplx);
            if (tau < 4) {
                l2_prefetch_bulk((tau < 2 ? a.Iux : a.Iuz) + rowoff + (size_t)(tau & 1) * a.ipitch, rowbytes);
            } else if (rplocal + stride < npairs) {
                l2_prefetch_bulk(a.Iq + (size_t)(2 * (rp + stride) + (tau & 1)) * a.ipitch, rowbytes);
            }
        }
        cplx v[16];
        {   // q -> physical (grouped order), parked in the thread-private stash
            const cplx* rowA = a.Iq + rowoff;
            zpair_load_line<LOG2N>(v, rowA, rowA + a.ipitch, tau, a.nm, a.sh);
            fft_nat2grp<LOG2N, true>(v, tau, a.tw, xbuf);
            MLV_UNROLL
            for (int j = 0; j < 16; ++j) stash[j * F::T] = v[j];
        }
        for (int pass = 0; pass < 2; ++pass) {        // pass 0: A = ux q, pass 1: B = uz q
            MLV_SCHED_FENCE();
            {
                const cplx* src = (pass == 0 ? a.Iux : a.Iuz) + rowoff;
                zpair_load_line<LOG2N>(v, src, src + a.ipitch, tau, a.nm, a.sh);
            }
            MLV_SCHED_FENCE();
            fft_nat2grp<LOG2N, true>(v, tau, a.tw, xbuf);
            {
                double mx = -INFINITY, ss = 0.0;
                MLV_UNROLL
                for (int j = 0; j < 16; ++j) {
                    mx = fmax(mx, fmax(v[j].x, v[j].y));
                    ss += v[j].x * v[j].x + v[j].y * v[j].y;
                    const cplx q = stash[j * F::T];
                    v[j] = mk(v[j].x * q.x, v[j].y * q.y);
                }
                if (valid) {
                    acc_mx[pass] = fmax(acc_mx[pass], mx);
                    acc_ss[pass] += ss;
                }
            }
            fft_grp2nat<LOG2N, false>(v, tau, a.tw, xbuf);
            cplx* pbuf = reinterpret_cast<cplx*>(xbuf);
            __syncthreads();
            zpair_publish<LOG2N>(v, tau, a.nm, pbuf);
            __syncthreads();
            if (valid) {
                const size_t foff = (size_t)a.outoff[pass];
                MLV_UNROLL
                for (int j0 = 0; j0 < 6; j0 += 3) {            // retained low modes: j <= 5
                    cplx P[3];
                    MLV_UNROLL
                    for (int u = 0; u < 3; ++u) {              // partner values: unconditional reads
                        const int kk = tau + F::T * (j0 + u);
                        P[u] = pbuf[kk < a.nm ? kk : 0];
                    }
                    MLV_UNROLL
                    for (int u = 0; u < 3; ++u) {
                        const int j = j0 + u;
                        const int kk = tau + F::T * j;
                        if (kk < a.nm) {
                            cplx A, B;
                            zpair_unpack(v[j], (kk == 0) ? v[j] : P[u], A, B);
                            const int t = kk >> cts;
                            int h = 0, tl = t;                                       // tile owner
                            if (sharded) { h = t / a.sh.tpr; tl = t - h * a.sh.tpr; }
                            cplx* o = a.out.blk[h] + foff + fwd_store_off(2 * rp, tl, kk & (a.ct - 1), a.ct, a.sh);
                            o[0] = A;
                            o[a.ct] = B;                               // row 2rp+1
                        }
                    }
                }
            }
        }
    }
    // ---- reductions: per-CTA partials (deterministic two-stage reduction).  fmax drops NaNs,
    // the sum of squares does not: a NaN anywhere makes the maximum NaN too, as numpy.max does
    // (Integrator.py:41 tests np.isnan(cfl_dt))
    __syncthreads();
    MLV_UNROLL
    for (int pass = 0; pass < 2; ++pass) {
        rbuf[pass * NT + threadIdx.x] = acc_ss[pass] != acc_ss[pass] ? NAN : acc_mx[pass];
        rbuf[(2 + pass) * NT + threadIdx.x] = acc_ss[pass];
    }
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
                const double x = rbuf[w * NT + g], y = rbuf[w * NT + g + s2];
                rbuf[w * NT + g] = (w < 2) ? nan_max(x, y) : x + y;
            }
        }
        if (w < 4 && g == 0) a.red[(size_t)blockIdx.x * 4 + w] = rbuf[w * NT];
    }
}

}  // namespace mlv
