// melvin-b200: COSINE / SINE bases (reference melvin/SpectralTransformer.py:108-125,134-146,169-196).
//
// The reference builds the even / odd mirror image of the physical field (period M = 2(n-1)) and
// calls numpy's rfft2 / irfft2 on it, then keeps the 2/3-rule modes.  For the grid sizes the package
// supports (n a power of two) M = 2(n-1) has large odd factors (n = 256: M = 510 = 2*3*5*17), so the
// power-of-two register transform of mlv_fft.cuh does not apply; and only (2nn+1) or nm of the M
// modes survive the truncation.  The kernel below evaluates exactly those modes (or, on the way back,
// the n physical samples) of the length-M transform by direct summation along ONE axis, with the
// roots of unity from an exact table: O(n * retained modes) per line, one pass per axis, no mirrored
// or padded array ever exists.  A 2-D transform is two launches (z axis then x axis forward, x then z
// inverse: the order of rfft2 / irfft2).  No example script uses these bases; this path is about
// parity (rounding-level agreement with pocketfft), not about the roofline.
#pragma once

#include "mlv_common.cuh"

namespace mlv {

struct TrigArgs {
    int inverse;        // 0: samples -> modes (e^{-2 pi i jk/M});  1: modes -> samples (e^{+...})
    int ext;            // how n_samp samples fill a period: 0 periodic (M = n_samp), 1 even mirror, 2 odd mirror
    int M;              // period
    int n_samp, n_modes;
    int two_sided;      // stored modes k = 0..nn,-nn..-1 (n_modes = 2nn+1); else k = 0..n_modes-1
    int hermitian;      // inverse of a one-sided spectrum to REAL samples: Re U0 + 2 Re sum_{k>=1} (irfft)
    int samp_complex;   // samples complex128 (else float64)
    int batch_fastest;  // adjacent threads take adjacent batch entries (else adjacent outputs)
    double w0;          // weight of mode 0 (cosine: forward 1/2, inverse 2)
    int nbatch;
    long long samp_stride, samp_batch, mode_stride, mode_batch;   // elements
    const void* in;
    void* out;
    double sre, sim;    // complex scale of the result
    const cplx* E;      // e^{-2 pi i j/M}, j < M
};

// signed mode number of stored mode index i
MLV_DEV int trig_mode(const TrigArgs& a, int i) {
    if (!a.two_sided) return i;
    const int nn = (a.n_modes - 1) / 2;
    return i <= nn ? i : i - a.n_modes;
}

__global__ void __launch_bounds__(256) k_trig_axis(const TrigArgs a) {
    const int nout = a.inverse ? a.n_samp : a.n_modes;
    const long long total = (long long)nout * a.nbatch;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(a.batch_fastest ? t / a.nbatch : t % nout);
        const int b = (int)(a.batch_fastest ? t % a.nbatch : t / nout);
        double ar = 0.0, ai = 0.0;
        if (!a.inverse) {
            // mode k = sum over the period of the (mirrored) samples
            int kk = trig_mode(a, o) % a.M;
            if (kk < 0) kk += a.M;
            const double* src = (const double*)a.in;
            int idx = 0;                                     // (j * kk) mod M
            for (int j = 0; j < a.M; ++j) {
                // even: x0 .. x_{n-1}, x_{n-2} .. x_1;  odd: x0 .. x_{n-2}, -x_{n-1}, -x_{n-2} .. -x_1
                const bool direct = a.ext == 0 || j <= a.n_samp - 2;
                const int js = direct ? j : a.M - j;
                const double sg = (direct || a.ext == 1) ? 1.0 : -1.0;
                const long long off = (long long)js * a.samp_stride + (long long)b * a.samp_batch;
                double xr, xi = 0.0;
                if (a.samp_complex) { xr = src[2 * off]; xi = src[2 * off + 1]; }
                else xr = src[off];
                const cplx e = a.E[idx];
                ar += sg * (xr * e.x - xi * e.y);
                ai += sg * (xr * e.y + xi * e.x);
                idx += kk;
                if (idx >= a.M) idx -= a.M;
            }
            if (trig_mode(a, o) == 0) { ar *= a.w0; ai *= a.w0; }
            cplx* dst = (cplx*)a.out + (long long)o * a.mode_stride + (long long)b * a.mode_batch;
            *dst = mk(ar * a.sre - ai * a.sim, ar * a.sim + ai * a.sre);
        } else {
            // sample o = sum over the stored modes
            const cplx* src = (const cplx*)a.in;
            for (int i = 0; i < a.n_modes; ++i) {
                int kk = trig_mode(a, i) % a.M;
                if (kk < 0) kk += a.M;
                const int idx = (int)(((long long)o * kk) % a.M);
                cplx u = src[(long long)i * a.mode_stride + (long long)b * a.mode_batch];
                double w = 1.0;
                if (trig_mode(a, i) == 0) {
                    w = a.w0;
                    if (a.hermitian) u.y = 0.0;              // irfft drops Im of the mean mode
                } else if (a.hermitian) {
                    w = 2.0;
                }
                const cplx e = a.E[idx];                     // conj(e): e^{+2 pi i o k/M}
                ar += w * (u.x * e.x + u.y * e.y);
                ai += w * (u.y * e.x - u.x * e.y);
            }
            const double rr = ar * a.sre - ai * a.sim, ri = ar * a.sim + ai * a.sre;
            const long long off = (long long)o * a.samp_stride + (long long)b * a.samp_batch;
            if (a.samp_complex) ((cplx*)a.out)[off] = mk(rr, ri);
            else ((double*)a.out)[off] = rr;
        }
    }
}

}  // namespace mlv
