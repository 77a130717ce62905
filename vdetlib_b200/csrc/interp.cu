// interp.cu -- densify strided tubelets.
//
// Replaces the per-tubelet scipy loop of score_proto_interpolation (vdet/tubelet_cls.py:430-490):
// six fields (x1, y1, x2, y2, det_score, anchor) are interpolated linearly over the dense frame
// range of each tubelet, with linear EXTRApolation one frame beyond either end (extrap1d,
// tubelet_cls.py:416-428).  Arithmetic follows what the reference executes, in float64:
//   inside the knots   scipy.interpolate.interp1d(kind='linear') delegates to numpy.interp (SciPy >= 0.17; the
//                      golden vectors were generated with SciPy 1.18.1 / NumPy 2.3.5 -- older SciPy used the
//                      searchsorted form, which rounds differently AT a knot):
//                      j = last knot <= x;  x == xs[j] -> ys[j];  else
//                      slope = (ys[j+1]-ys[j]) / (xs[j+1]-xs[j]);  slope*(x-xs[j]) + ys[j]
//   left of the knots  ys[0]  + (x-xs[0])  * (ys[1]-ys[0])   / (xs[1]-xs[0])
//   right of them      ys[-1] + (x-xs[-1]) * (ys[-1]-ys[-2]) / (xs[-1]-xs[-2])
// One thread per dense frame of a tubelet: one binary search shared by the six fields.
#include "common.cuh"

namespace vdet {

__global__ void __launch_bounds__(256) tubelet_interp_kernel(const double* __restrict__ knot_x,
                                                             const double* __restrict__ knot_y, int64_t knot_ld,
                                                             const int32_t* __restrict__ knot_off,
                                                             const int32_t* __restrict__ dense_off,
                                                             const int32_t* __restrict__ dense_first,
                                                             const int32_t* __restrict__ dense_tub, int n_fields,
                                                             int64_t n_dense, double* __restrict__ out) {
    const int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dense) return;
    const int k = dense_tub[d];
    const int ko = knot_off[k];
    const int n = knot_off[k + 1] - ko;
    const double* xs = knot_x + ko;
    const double x = (double)(dense_first[k] + (int)(d - dense_off[k]));
    int mode, j;            // 0: knot value, 1: interior, 2: left extrapolation, 3: right extrapolation
    if (x < xs[0]) { mode = 2; j = 0; }
    else if (x > xs[n - 1]) { mode = 3; j = n - 1; }
    else {
        int lo = 0, hi = n;                     // largest j with xs[j] <= x
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (xs[mid] <= x) lo = mid; else hi = mid;
        }
        j = lo;
        mode = (j == n - 1 || xs[j] == x) ? 0 : 1;
    }
    for (int f = 0; f < n_fields; ++f) {
        const double* ys = knot_y + (int64_t)f * knot_ld + ko;
        double r;
        if (mode == 0) {
            r = ys[j];
        } else if (mode == 1) {
            const double slope = __ddiv_rn(__dsub_rn(ys[j + 1], ys[j]), __dsub_rn(xs[j + 1], xs[j]));
            r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xs[j])), ys[j]);
        } else if (mode == 2) {
            r = __dadd_rn(ys[0], __ddiv_rn(__dmul_rn(__dsub_rn(x, xs[0]), __dsub_rn(ys[1], ys[0])),
                                         __dsub_rn(xs[1], xs[0])));
        } else {
            r = __dadd_rn(ys[n - 1], __ddiv_rn(__dmul_rn(__dsub_rn(x, xs[n - 1]), __dsub_rn(ys[n - 1], ys[n - 2])),
                                             __dsub_rn(xs[n - 1], xs[n - 2])));
        }
        out[(int64_t)f * n_dense + d] = r;
    }
}

}  // namespace vdet

using namespace vdet;

extern "C" int vdet_tubelet_interpolate_f64(const double* knot_x, const double* knot_y, int64_t n_knots,
                                            const int32_t* knot_off, const int32_t* dense_off,
                                            const int32_t* dense_first, const int32_t* dense_tub,
                                            int n_tubelets, int n_fields, int64_t n_dense, double* out,
                                            void* stream) {
    VDET_REQUIRE(n_knots >= 0 && n_tubelets >= 0 && n_fields >= 1 && n_dense >= 0, "tubelet_interpolate: bad size");
    if (n_dense == 0 || n_tubelets == 0) return VDET_OK;
    tubelet_interp_kernel<<<(unsigned)((n_dense + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        knot_x, knot_y, n_knots, knot_off, dense_off, dense_first, dense_tub, n_fields, n_dense, out);
    VDET_LAUNCH_CHECK();
    return VDET_OK;
}
