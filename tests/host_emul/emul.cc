// Host build of the device math header (client_b200/csrc/philox.cuh) for
// CPU-only unit tests: the same fill_group / scale_pixel code the kernels
// inline, driven by a plain loop.  TEST AID ONLY -- never linked into
// libtb200.so, never used as a fallback.
#include <cstdint>
#include <cstring>

#include "../../client_b200/csrc/philox.cuh"
#include "../../client_b200/csrc/resample.h"

using namespace tb200;

extern "C" {

void emul_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  const U32x4 r = philox4x32<10>(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

// mirrors fill_kernel's per-job setup + per-group body
void emul_fill(uint8_t* dst, uint64_t nbytes, uint32_t dtype, uint64_t seed, uint64_t stream,
               double lo, double span, int64_t ilo, uint64_t irange) {
  FillParams p;
  p.lo_f = static_cast<float>(lo);
  p.span_f = static_cast<float>(span);
  p.lo_d = lo;
  p.span_d = span;
  p.ilo = ilo;
  p.irange = irange;
  p.unit = (span == 0.0) ? 1u : 0u;
  uint32_t dt = dtype;
  if (dt == kU64) dt = kI64;
  if (dt == kU32) dt = kI32;
  if (dt == kU16) dt = kI16;
  if (dt == kU8) dt = kI8;
  for (uint64_t g = 0; g * 16 < nbytes; ++g) {
    const U32x4 r = philox4x32<10>(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32),
                                   static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32),
                                   static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const U32x4 o = fill_group(dt, r, p);
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
    const uint64_t left = nbytes - g * 16;
    memcpy(dst + g * 16, w, left < 16 ? left : 16);
  }
}

uint32_t emul_scale_f32_bits(uint32_t px, uint32_t scaling, int c, int ch) {
  return f32_bits(scale_pixel_f32(px, scaling, c, ch));
}
uint32_t emul_scale_f16_bits(uint32_t px, uint32_t scaling, int c, int ch) {
  return scale_pixel_f16(px, scaling, c, ch);
}
uint32_t emul_f32_to_f16(uint32_t bits) { return f32_to_f16_bits(bits_f32(bits)); }
uint32_t emul_f16_to_f32(uint32_t h) { return f32_bits(f16_bits_to_f32(static_cast<uint16_t>(h))); }

// BYTES fill: group g of a fixed-length string tensor (what fill_segment_random<kBytes> stores)
void emul_fill_bytes(uint8_t* dst, uint64_t nbytes, uint64_t seed, uint64_t stream, uint32_t len) {
  for (uint64_t g = 0; g * 16 < nbytes; ++g) {
    const U32x4 r = philox4x32<10>(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32),
                                   static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32),
                                   static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const U32x4 o = fill_group_bytes(r, g, len);
    const uint32_t w[4] = {o.x, o.y, o.z, o.w};
    const uint64_t left = nbytes - g * 16;
    memcpy(dst + g * 16, w, left < 16 ? left : 16);
  }
}

// the resize kernel's host tables; returns ksize, fills bounds[2*out] and coeffs[out*cap]
int emul_resample_tables(int in_size, int out_size, int* bounds, int* coeffs, int cap) {
  std::vector<ResampleBound> b;
  std::vector<int32_t> c;
  int ks = 0;
  resample_coefficients(in_size, out_size, &b, &c, &ks);
  if (ks > cap) return -ks;
  for (int i = 0; i < out_size; ++i) {
    bounds[2 * i] = b[i].first;
    bounds[2 * i + 1] = b[i].count;
    for (int t = 0; t < ks; ++t) coeffs[i * cap + t] = c[static_cast<size_t>(i) * ks + t];
  }
  return ks;
}
}
