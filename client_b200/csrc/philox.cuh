// philox.cuh -- Philox4x32-10 counter RNG and the "fill contract": how the four
// 32-bit words of one Philox call become the 16 bytes of one group of a tensor.
//
// Everything here is __host__ __device__ so the very same arithmetic can be
// unit-tested on a CPU-only box (tests/host_emul) before GPU time is spent; the
// product library only ever calls it from kernels.
//
// Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy
// as 1, 2, 3" (SC'11); constants and known-answer vectors are the published
// Random123 ones (oracle/ pins them, and checks cuRAND's host generator).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define TB200_HD __host__ __device__ __forceinline__
#else
#define TB200_HD inline
#endif

namespace tb200 {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct U32x4 {
  uint32_t x, y, z, w;
};

// One Philox4x32 block with R rounds. ptxas turns each 32x32->64 product into
// a single IMAD.WIDE.U32 and each xor3 into one LOP3.
template <int R = 10>
TB200_HD U32x4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return U32x4{c0, c1, c2, c3};
}

// Same block function with the key schedule precomputed (rk[2r], rk[2r+1] are the
// round-r keys).  The kernel passes rk in the launch parameters so every xor3 takes
// its key from the constant bank and the 18 per-call key additions disappear.
struct RoundKeys {
  uint32_t k[20];
};
inline void make_round_keys(uint64_t seed, RoundKeys* out) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    out->k[2 * r] = k0;
    out->k[2 * r + 1] = k1;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
}
template <int R = 10>
TB200_HD U32x4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                const RoundKeys& rk) {
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;
    const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ rk.k[2 * r];
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ rk.k[2 * r + 1];
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
  }
  return U32x4{c0, c1, c2, c3};
}

// Stream-hoisted form.  For a counter (g, 0, s_lo, s_hi) two of the four multiplies of rounds
// 0 and 1 see only the stream id: M1 * s_lo in round 0 and, through it, M0 * c0 in round 1.
// philox_stream_const folds them (and the key words they meet) into four words per stream;
// philox4x32_10_hoisted then needs 18 wide multiplies per call instead of 20 and takes
// p0 = M0 * g from the caller (a multiply, or a running 64-bit sum when g advances by a
// constant).  Bit-identical to philox4x32_10_rk(g, 0, s_lo, s_hi, rk) -- checked on the CPU
// by tests/host_emul and on the device against the C oracle.
struct PhiloxStreamConst {
  uint32_t x, y, z, e;
};
TB200_HD PhiloxStreamConst philox_stream_const(uint32_t s_lo, uint32_t s_hi, const RoundKeys& rk) {
  const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * s_lo;               // round 0, second lane
  const uint32_t c0 = static_cast<uint32_t>(p1 >> 32) ^ rk.k[0];            // ^ g_hi (= 0)
  const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c0;                // round 1, first lane
  PhiloxStreamConst sc;
  sc.x = s_hi ^ rk.k[1];
  sc.y = static_cast<uint32_t>(p1) ^ rk.k[2];
  sc.z = static_cast<uint32_t>(p0 >> 32) ^ rk.k[3];
  sc.e = static_cast<uint32_t>(p0);
  return sc;
}
template <int R = 10>
TB200_HD U32x4 philox4x32_10_hoisted(uint64_t m0_times_g, const PhiloxStreamConst& sc, const RoundKeys& rk) {
  static_assert(R >= 2, "the hoisted form covers rounds 0 and 1");
  uint32_t c2 = static_cast<uint32_t>(m0_times_g >> 32) ^ sc.x;             // after round 0
  const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c2;                // round 1, second lane
  uint32_t c0 = static_cast<uint32_t>(p1 >> 32) ^ sc.y;
  uint32_t c1 = static_cast<uint32_t>(p1);
  c2 = static_cast<uint32_t>(m0_times_g) ^ sc.z;
  uint32_t c3 = sc.e;
#pragma unroll
  for (int r = 2; r < R; ++r) {
    const uint64_t q0 = static_cast<uint64_t>(kPhiloxM0) * c0;
    const uint64_t q1 = static_cast<uint64_t>(kPhiloxM1) * c2;
    const uint32_t n0 = static_cast<uint32_t>(q1 >> 32) ^ c1 ^ rk.k[2 * r];
    const uint32_t n2 = static_cast<uint32_t>(q0 >> 32) ^ c3 ^ rk.k[2 * r + 1];
    c1 = static_cast<uint32_t>(q1);
    c3 = static_cast<uint32_t>(q0);
    c0 = n0;
    c2 = n2;
  }
  return U32x4{c0, c1, c2, c3};
}

// ---- float <-> bit helpers that behave identically on host and device ------
TB200_HD uint32_t f32_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union {
    float f;
    uint32_t u;
  } v;
  v.f = f;
  return v.u;
#endif
}
TB200_HD float bits_f32(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union {
    float f;
    uint32_t u;
  } v;
  v.u = u;
  return v.f;
#endif
}
TB200_HD uint64_t f64_bits(double d) {
#if defined(__CUDA_ARCH__)
  return static_cast<uint64_t>(__double_as_longlong(d));
#else
  union {
    double d;
    uint64_t u;
  } v;
  v.d = d;
  return v.u;
#endif
}

TB200_HD float fma_f32(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}
TB200_HD double fma_f64(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}

// fp32 -> fp16 bits, round to nearest even (numpy astype(float16) semantics:
// overflow -> inf, subnormals kept, NaN stays NaN).
TB200_HD uint16_t f32_to_f16_bits(float f) {
#if defined(__CUDA_ARCH__)
  return __half_as_ushort(__float2half_rn(f));
#else
  const uint32_t x = f32_bits(f);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const uint32_t absx = x & 0x7FFFFFFFu;
  if (absx >= 0x7F800000u) {  // inf / nan
    if (absx > 0x7F800000u) return static_cast<uint16_t>(sign | 0x7E00u | ((absx >> 13) & 0x3FFu));
    return static_cast<uint16_t>(sign | 0x7C00u);
  }
  if (absx >= 0x477FF000u) {  // rounds to >= 65520 -> inf
    return static_cast<uint16_t>(sign | 0x7C00u);
  }
  if (absx < 0x38800000u) {  // below the smallest normal half: subnormal/zero
    if (absx < 0x33000000u) return static_cast<uint16_t>(sign);  // < 2^-25 -> 0
    const int e = static_cast<int>(absx >> 23);                  // biased fp32 exp
    const uint32_t mant = (absx & 0x7FFFFFu) | 0x800000u;        // 24-bit
    const int shift = 126 - e;  // bits to drop: result = mant >> shift, 14..24
    const uint32_t q = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1u);
    const uint32_t half = 1u << (shift - 1);
    uint32_t r = q;
    if (rem > half || (rem == half && (q & 1u))) r += 1;
    return static_cast<uint16_t>(sign | r);
  }
  // normal range
  uint32_t r = absx - 0x38000000u;  // rebias exponent (127-15)<<23
  const uint32_t rem = r & 0x1FFFu;
  r >>= 13;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r += 1;
  return static_cast<uint16_t>(sign | r);
#endif
}

TB200_HD float f16_bits_to_f32(uint16_t h) {
#if defined(__CUDA_ARCH__)
  return __half2float(__ushort_as_half(h));
#else
  const uint32_t sign = (static_cast<uint32_t>(h) & 0x8000u) << 16;
  const uint32_t e = (h >> 10) & 0x1Fu;
  uint32_t m = h & 0x3FFu;
  if (e == 0) {
    if (m == 0) return bits_f32(sign);
    int k = 0;  // normalise the subnormal
    while (!(m & 0x400u)) {
      m <<= 1;
      ++k;
    }
    m &= 0x3FFu;
    return bits_f32(sign | static_cast<uint32_t>(113 - k) << 23 | (m << 13));
  }
  if (e == 31) return bits_f32(sign | 0x7F800000u | (m << 13));
  return bits_f32(sign | ((e + 112u) << 23) | (m << 13));
#endif
}

// fp32 -> bf16 the way tritonclient serialises BF16: keep the upper 16 bits
// (PY/utils/__init__.py:327-331 "struct.pack('<f', x)[2:4]"), no rounding.
TB200_HD uint16_t f32_to_bf16_trunc(float f) {
  return static_cast<uint16_t>(f32_bits(f) >> 16);
}

TB200_HD uint64_t mulhi_u64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return static_cast<uint64_t>((static_cast<unsigned __int128>(a) * b) >> 64);
#endif
}

// ---- the fill contract -------------------------------------------------------
// dtype codes are tb200_dtype (include/tb200.h); kept as plain ints here so the
// header has no dependency on the ABI header.
enum : uint32_t {
  kBool = 1, kU8 = 2, kU16 = 3, kU32 = 4, kU64 = 5, kI8 = 6, kI16 = 7,
  kI32 = 8, kI64 = 9, kF16 = 10, kF32 = 11, kF64 = 12, kBytes = 13, kBF16 = 14
};

struct FillParams {
  float lo_f;       // float dtypes (fp16/bf16/fp32 compute in fp32)
  float span_f;
  double lo_d;      // fp64
  double span_d;
  int64_t ilo;
  uint64_t irange;  // 0 -> raw bits
  uint32_t unit;    // 1 -> floats are the unit interval, skip the fma
};

// Uniform values in [0,1) by mantissa insertion: the top mantissa-width bits of the
// random field become the fraction of a float in [1,2), then 1 is subtracted (exact).
//   fp32: (w   >> 9 ) * 2^-23      fp16: (x16 >> 6) * 2^-10
//   bf16: (x16 >> 9 ) * 2^-7       fp64: (x64 >> 12) * 2^-52
// One shift/or + one add per value and no I2F conversions (those run on a slow pipe).
TB200_HD float unit_f32(uint32_t w) {
#if defined(__CUDA_ARCH__)
  // one funnel shift: ({0x7F : w} >> 9) = (w >> 9) | 0x3F800000
  return __uint_as_float(__funnelshift_r(w, 0x7Fu, 9)) - 1.0f;
#else
  return bits_f32(0x3F800000u | (w >> 9)) - 1.0f;
#endif
}
TB200_HD float unit_f16(uint32_t x16) {  // 10 random bits, exactly representable in fp16
  return bits_f32(0x3F800000u | ((x16 >> 6) << 13)) - 1.0f;
}
TB200_HD float unit_bf16(uint32_t x16) {  // 7 random bits, exactly representable in bf16
  return bits_f32(0x3F800000u | ((x16 >> 9) << 16)) - 1.0f;
}
TB200_HD double unit_f64(uint32_t lo, uint32_t hi) {  // 52 random bits
  const uint64_t x = (static_cast<uint64_t>(hi) << 32) | lo;
  const uint64_t bits = 0x3FF0000000000000ull | (x >> 12);
#if defined(__CUDA_ARCH__)
  // two funnel shifts build 0x3FF0000000000000 | (x >> 12)
  const uint32_t blo = __funnelshift_r(lo, hi, 12);
  const uint32_t bhi = __funnelshift_r(hi, 0x3FFu, 12);
  (void)bits;
  return __hiloint2double(static_cast<int>(bhi), static_cast<int>(blo)) - 1.0;
#else
  union {
    uint64_t u;
    double d;
  } v;
  v.u = bits;
  return v.d - 1.0;
#endif
}

TB200_HD uint32_t fill_word_f32(uint32_t w, const FillParams& p) {
  float u = unit_f32(w);
  if (!p.unit) u = fma_f32(u, p.span_f, p.lo_f);
  return f32_bits(u);
}
TB200_HD uint32_t fill_word_f16(uint32_t w, const FillParams& p) {
  if (p.unit) {
    // both halves at once: 0x3C00 | (x16 >> 6) is a half in [1,2); minus 1 is exact
    const uint32_t one_plus = 0x3C003C00u | ((w >> 6) & 0x03FF03FFu);
#if defined(__CUDA_ARCH__)
    const __half2 v = __hsub2(*reinterpret_cast<const __half2*>(&one_plus), __float2half2_rn(1.0f));
    return *reinterpret_cast<const uint32_t*>(&v);
#else
    const float a = f16_bits_to_f32(static_cast<uint16_t>(one_plus & 0xFFFFu)) - 1.0f;
    const float b = f16_bits_to_f32(static_cast<uint16_t>(one_plus >> 16)) - 1.0f;
    return static_cast<uint32_t>(f32_to_f16_bits(a)) | (static_cast<uint32_t>(f32_to_f16_bits(b)) << 16);
#endif
  }
  const float a = fma_f32(unit_f16(w & 0xFFFFu), p.span_f, p.lo_f);
  const float b = fma_f32(unit_f16(w >> 16), p.span_f, p.lo_f);
  return static_cast<uint32_t>(f32_to_f16_bits(a)) |
         (static_cast<uint32_t>(f32_to_f16_bits(b)) << 16);
}
TB200_HD uint32_t fill_word_bf16(uint32_t w, const FillParams& p) {
  float a = unit_bf16(w & 0xFFFFu);
  float b = unit_bf16(w >> 16);
  if (!p.unit) {
    a = fma_f32(a, p.span_f, p.lo_f);
    b = fma_f32(b, p.span_f, p.lo_f);
  }
  return static_cast<uint32_t>(f32_to_bf16_trunc(a)) |
         (static_cast<uint32_t>(f32_to_bf16_trunc(b)) << 16);
}
TB200_HD uint32_t fill_word_i32(uint32_t w, const FillParams& p) {
  if (p.irange == 0) return w;
  const uint32_t r = static_cast<uint32_t>((static_cast<uint64_t>(w) * p.irange) >> 32);
  return static_cast<uint32_t>(p.ilo) + r;
}
TB200_HD uint32_t fill_word_i16(uint32_t w, const FillParams& p) {
  if (p.irange == 0) return w;
  const uint32_t rg = static_cast<uint32_t>(p.irange);
  const uint32_t a = (((w & 0xFFFFu) * rg) >> 16) + static_cast<uint32_t>(p.ilo);
  const uint32_t b = (((w >> 16) * rg) >> 16) + static_cast<uint32_t>(p.ilo);
  return (a & 0xFFFFu) | (b << 16);
}
TB200_HD uint32_t fill_word_i8(uint32_t w, const FillParams& p) {
  if (p.irange == 0) return w;
  const uint32_t rg = static_cast<uint32_t>(p.irange);
  const uint32_t lo = static_cast<uint32_t>(p.ilo);
  uint32_t out = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t b = (w >> (8 * i)) & 0xFFu;
    out |= ((((b * rg) >> 8) + lo) & 0xFFu) << (8 * i);
  }
  return out;
}
TB200_HD uint64_t fill_pair_i64(uint32_t lo, uint32_t hi, const FillParams& p) {
  const uint64_t x = (static_cast<uint64_t>(hi) << 32) | lo;
  if (p.irange == 0) return x;
  return static_cast<uint64_t>(p.ilo) + mulhi_u64(x, p.irange);
}
TB200_HD uint64_t fill_pair_f64(uint32_t lo, uint32_t hi, const FillParams& p) {
  double u = unit_f64(lo, hi);
  if (!p.unit) u = fma_f64(u, p.span_d, p.lo_d);
  return f64_bits(u);
}

// 4 Philox words -> the 16 bytes of one group (as 4 little-endian words).
TB200_HD U32x4 fill_group(uint32_t dtype, U32x4 r, const FillParams& p) {
  U32x4 o;
  switch (dtype) {
    case kF32:
      o.x = fill_word_f32(r.x, p); o.y = fill_word_f32(r.y, p);
      o.z = fill_word_f32(r.z, p); o.w = fill_word_f32(r.w, p);
      break;
    case kF16:
      o.x = fill_word_f16(r.x, p); o.y = fill_word_f16(r.y, p);
      o.z = fill_word_f16(r.z, p); o.w = fill_word_f16(r.w, p);
      break;
    case kBF16:
      o.x = fill_word_bf16(r.x, p); o.y = fill_word_bf16(r.y, p);
      o.z = fill_word_bf16(r.z, p); o.w = fill_word_bf16(r.w, p);
      break;
    case kF64: {
      const uint64_t a = fill_pair_f64(r.x, r.y, p);
      const uint64_t b = fill_pair_f64(r.z, r.w, p);
      o.x = static_cast<uint32_t>(a); o.y = static_cast<uint32_t>(a >> 32);
      o.z = static_cast<uint32_t>(b); o.w = static_cast<uint32_t>(b >> 32);
      break;
    }
    case kI64:
    case kU64: {
      const uint64_t a = fill_pair_i64(r.x, r.y, p);
      const uint64_t b = fill_pair_i64(r.z, r.w, p);
      o.x = static_cast<uint32_t>(a); o.y = static_cast<uint32_t>(a >> 32);
      o.z = static_cast<uint32_t>(b); o.w = static_cast<uint32_t>(b >> 32);
      break;
    }
    case kI32:
    case kU32:
      o.x = fill_word_i32(r.x, p); o.y = fill_word_i32(r.y, p);
      o.z = fill_word_i32(r.z, p); o.w = fill_word_i32(r.w, p);
      break;
    case kI16:
    case kU16:
      o.x = fill_word_i16(r.x, p); o.y = fill_word_i16(r.y, p);
      o.z = fill_word_i16(r.z, p); o.w = fill_word_i16(r.w, p);
      break;
    case kI8:
    case kU8:
      o.x = fill_word_i8(r.x, p); o.y = fill_word_i8(r.y, p);
      o.z = fill_word_i8(r.z, p); o.w = fill_word_i8(r.w, p);
      break;
    case kBool:
    default:
      o.x = r.x & 0x01010101u; o.y = r.y & 0x01010101u;
      o.z = r.z & 0x01010101u; o.w = r.w & 0x01010101u;
      break;
  }
  return o;
}

// BYTES tensors of fixed-length strings (perf_analyzer's --string-length inputs): the
// serialised form of PY/utils/__init__.py:208-261 generated in place -- element e occupies
// bytes [e*(4+L), (e+1)*(4+L)): a little-endian u32 length L, then L characters.  Byte i
// of group g takes Philox byte i of that group; a character is "0-9A-Za-z"[(b * 62) >> 8].
TB200_HD uint32_t alnum_char(uint32_t b) {
  const uint32_t k = (b * 62u) >> 8;
  return k < 10u ? 48u + k : (k < 36u ? 55u + k : 61u + k);
}
TB200_HD U32x4 fill_group_bytes(U32x4 r, uint64_t g, uint32_t len) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
  uint32_t out[4] = {0u, 0u, 0u, 0u};
  const uint32_t per = len + 4u;
  uint32_t o = static_cast<uint32_t>((g * 16u) % per);
  for (uint32_t i = 0; i < 16u; ++i) {
    const uint32_t rb = (w[i >> 2] >> (8u * (i & 3u))) & 0xFFu;
    const uint32_t byte = o < 4u ? ((len >> (8u * o)) & 0xFFu) : alnum_char(rb);
    out[i >> 2] |= byte << (8u * (i & 3u));
    o = (o + 1u == per) ? 0u : o + 1u;
  }
  U32x4 v;
  v.x = out[0]; v.y = out[1]; v.z = out[2]; v.w = out[3];
  return v;
}

// ---- image scaling arithmetic (image_client.py:171-181) ---------------------
// The source is a uint8 pixel, so every formula below has only 256 inputs per
// channel; tests enumerate all of them against numpy.
//
// numpy float32:  q = fl32(x / 127.5);         y = fl32(q - 1)
// numpy float16:  q = fl16(fl32(x) / 127.5f);  y = fl16(fl32(q) - 1)   (numpy
//                 evaluates half ops in float and rounds once per op)
// x / 127.5 is evaluated as fl32(x * r + e) with r = fl32(1/127.5) and one
// Newton correction; that equals IEEE division for all 256 inputs (checked
// exhaustively in tests/test_host_emul.py).
TB200_HD float div_127_5(float x) {
  constexpr float r = 1.0f / 127.5f;  // fl32(1/127.5), folded at compile time
  const float q0 = x * r;
  const float rem = fma_f32(-q0, 127.5f, x);
  return fma_f32(rem, r, q0);
}
TB200_HD float vgg_mean(int c, int channel) {
  if (c == 1) return 128.0f;
  return channel == 0 ? 123.0f : (channel == 1 ? 117.0f : 104.0f);
}
// scaled value in the *destination precision's* arithmetic, returned as fp32
// (for fp16 destinations the value is exactly representable in fp16).
TB200_HD float scale_pixel_f32(uint32_t px, uint32_t scaling, int c, int channel) {
  const float x = static_cast<float>(px);
  if (scaling == 1) return div_127_5(x) - 1.0f;
  if (scaling == 2) return x - vgg_mean(c, channel);
  return x;
}
TB200_HD uint16_t scale_pixel_f16(uint32_t px, uint32_t scaling, int c, int channel) {
  const float x = static_cast<float>(px);
  if (scaling == 1) {
    const float q = f16_bits_to_f32(f32_to_f16_bits(div_127_5(x)));
    return f32_to_f16_bits(q - 1.0f);
  }
  if (scaling == 2) return f32_to_f16_bits(x - vgg_mean(c, channel));
  return f32_to_f16_bits(x);
}

}  // namespace tb200
