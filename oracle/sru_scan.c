/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never linked into or called by the product path.
 *
 * Plain-C restatement of the element-wise recurrence of the third-party `sru` package
 * (v2.x; pinned in /root/reference/setup/requirements.yaml:18,33, source NOT in /root/reference),
 * as used by DualPathRNN (/root/reference/src/models/layers/rnn_layers.py:100-105,150).
 * PARITY UNPINNED: restated from the published algorithm; no upstream
 * source or golden vector is available to check against.
 *
 * U      : (L, B, D, k)   D = ndir*d; column index (dir*d + j); k = 3 or 4
 * x      : (L, B, D)      highway input, only read when k == 3
 * wc     : (2*D)          [v_f | v_r]
 * bias   : (2*D)          [b_f | b_r]
 * h      : (L, B, D)      output
 * c_last : (B, D)         final cell state
 *
 *   f_t = sigmoid(U1_t + v_f * c_{t-1} + b_f)
 *   r_t = sigmoid(U2_t + v_r * c_{t-1} + b_r)
 *   c_t = f_t * c_{t-1} + (1 - f_t) * U0_t
 *   h_t = r_t * c_t + (1 - r_t) * x'_t          x' = U3 (k == 4) or x (k == 3)
 * forward half scans t = 0..L-1, backward half (dir == 1) t = L-1..0, c_{-1} = 0.
 */
#include <math.h>
#include <stddef.h>

static inline float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

void sru_scan_f32(const float *U, const float *x, const float *wc, const float *bias, float *h, float *c_last,
                  int L, int B, int d, int ndir, int k) {
    const int D = d * ndir;
#pragma omp parallel for schedule(static)
    for (int col = 0; col < B * D; ++col) {
        const int b = col / D, j = col % D;
        const int reverse = (j >= d);
        const float vf = wc[j], vr = wc[D + j], bf = bias[j], br = bias[D + j];
        float c = 0.0f;
        for (int s = 0; s < L; ++s) {
            const int t = reverse ? (L - 1 - s) : s;
            const size_t row = ((size_t)t * B + b) * D + j;
            const float *u = U + row * k;
            const float f = sigmoidf_(u[1] + vf * c + bf);
            const float r = sigmoidf_(u[2] + vr * c + br);
            const float xp = (k == 4) ? u[3] : x[row];
            c = f * c + (1.0f - f) * u[0];
            h[row] = r * c + (1.0f - r) * xp;
        }
        c_last[col] = c;
    }
}

void sru_scan_f64(const double *U, const double *x, const double *wc, const double *bias, double *h, double *c_last,
                  int L, int B, int d, int ndir, int k) {
    const int D = d * ndir;
#pragma omp parallel for schedule(static)
    for (int col = 0; col < B * D; ++col) {
        const int b = col / D, j = col % D;
        const int reverse = (j >= d);
        const double vf = wc[j], vr = wc[D + j], bf = bias[j], br = bias[D + j];
        double c = 0.0;
        for (int s = 0; s < L; ++s) {
            const int t = reverse ? (L - 1 - s) : s;
            const size_t row = ((size_t)t * B + b) * D + j;
            const double *u = U + row * k;
            const double f = 1.0 / (1.0 + exp(-(u[1] + vf * c + bf)));
            const double r = 1.0 / (1.0 + exp(-(u[2] + vr * c + br)));
            const double xp = (k == 4) ? u[3] : x[row];
            c = f * c + (1.0 - f) * u[0];
            h[row] = r * c + (1.0 - r) * xp;
        }
        c_last[col] = c;
    }
}
