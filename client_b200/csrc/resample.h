// resample.h -- host-side coefficient tables of the resize kernel (plain C++, also compiled
// by tests/host_emul on the CPU).
//
// Coefficients as Pillow computes them (libImaging/Resample.c: precompute_coeffs +
// normalize_coeffs_8bpc; third-party to the reference, which calls it through Image.resize at
// src/python/examples/image_client.py:166): for output index xx, center = (xx + 0.5) * scale,
// triangle filter of half-width support = max(scale, 1), taps [xmin, xmax) rounded to the
// nearest source index, weights normalised to sum 1 in double precision and quantised to
// 22-bit fixed point, round half up.
#ifndef TB200_CSRC_RESAMPLE_H_
#define TB200_CSRC_RESAMPLE_H_

#include <cmath>
#include <cstdint>
#include <vector>

namespace tb200 {

struct ResampleBound {
  int first;  // first source index
  int count;  // taps
};

inline void resample_coefficients(int in_size, int out_size, std::vector<ResampleBound>* bounds,
                                  std::vector<int32_t>* coeffs, int* ksize_out) {
  const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;  // bilinear: filter support 1.0
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  bounds->assign(static_cast<size_t>(out_size), ResampleBound{0, 0});
  coeffs->assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(static_cast<size_t>(ksize));
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double t = (x + xmin - center + 0.5) * ss;
      if (t < 0.0) t = -t;
      const double w = t < 1.0 ? 1.0 - t : 0.0;
      k[static_cast<size_t>(x)] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[static_cast<size_t>(x)] /= ww;
    }
    (*bounds)[static_cast<size_t>(xx)] = ResampleBound{xmin, xmax};
    for (int x = 0; x < xmax; ++x) {
      const double v = k[static_cast<size_t>(x)];
      (*coeffs)[static_cast<size_t>(xx) * ksize + x] =
          v < 0 ? static_cast<int32_t>(-0.5 + v * (1 << 22)) : static_cast<int32_t>(0.5 + v * (1 << 22));
    }
  }
  *ksize_out = ksize;
}

}  // namespace tb200

#endif  // TB200_CSRC_RESAMPLE_H_
