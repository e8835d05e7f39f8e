// Host build of the device math header (client_b200/csrc/philox.cuh) for
// CPU-only unit tests: the same fill_group / scale_pixel code the kernels
// inline, driven by a plain loop.  TEST AID ONLY -- never linked into
// libtb200.so, never used as a fallback.
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../../client_b200/csrc/philox.cuh"
#include "../../client_b200/csrc/resample.h"
#include "../../client_b200/csrc/deflate.cuh"

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

// mirrors fill_uniform_kernel: round keys + philox_stream_const once per tensor, then the
// hoisted block function on M0 * g per 16-byte group
void emul_fill_hoisted(uint8_t* dst, uint64_t nbytes, uint32_t dtype, uint64_t seed, uint64_t stream,
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
  RoundKeys rk;
  make_round_keys(seed, &rk);
  const PhiloxStreamConst sc = philox_stream_const(static_cast<uint32_t>(stream), static_cast<uint32_t>(stream >> 32), rk);
  for (uint64_t g = 0; g * 16 < nbytes; ++g) {
    const U32x4 r = philox4x32_10_hoisted<10>(static_cast<uint64_t>(kPhiloxM0) * static_cast<uint32_t>(g), sc, rk);
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

// The device deflate encoder's pipeline, sequentially: chunk -> candidates in rounds of
// kDeflateThreads -> two-pass parse per sub-block -> stored fallback -> checksums -> container.
// Orchestration mirrors deflate.cu; every bit-level routine is the shared deflate.cuh code.
uint64_t emul_deflate(const uint8_t* src, uint64_t nbytes, uint32_t gzip, uint8_t* dst) {
  uint64_t out = 0;
  if (gzip) {
    const uint8_t h[10] = {0x1F, 0x8B, 0x08, 0x00, 0, 0, 0, 0, 0x00, 0xFF};
    memcpy(dst, h, 10);
    out = 10;
  } else {
    dst[0] = 0x78;
    dst[1] = 0x01;
    out = 2;
  }
  uint32_t A = 0, B = 0, crc = 0;
  std::vector<uint8_t> in(kDeflateInBytes);
  std::vector<uint16_t> cand(kDeflateChunk), table(1 << kDeflateHashBits);
  std::vector<uint32_t> words(kDeflateOutWords);
  for (uint64_t base = 0; base < nbytes; base += kDeflateChunk) {
    const uint32_t n = static_cast<uint32_t>(nbytes - base < static_cast<uint64_t>(kDeflateChunk) ? nbytes - base : kDeflateChunk);
    for (uint32_t i = 0; i < n; ++i) in[deflate_at(i)] = src[base + i];
    std::fill(table.begin(), table.end(), static_cast<uint16_t>(kDeflateNoCand));
    std::fill(words.begin(), words.end(), 0u);
    for (uint32_t r0 = 0; r0 < n; r0 += kDeflateThreads) {
      for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads) && r0 + t < n; ++t) {
        const uint32_t p = r0 + t;
        cand[p] = p + 3 < n ? table[deflate_hash(in.data(), p)] : static_cast<uint16_t>(kDeflateNoCand);
      }
      for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads) && r0 + t + 3 < n; ++t) table[deflate_hash(in.data(), r0 + t)] = static_cast<uint16_t>(r0 + t);
    }
    uint32_t off[kDeflateThreads], total = 3;
    uint32_t chunk_crc = 0;
    for (int t = 0; t < kDeflateThreads; ++t) {
      const uint32_t b0 = t * kDeflateSub;
      const uint32_t e0 = b0 < n ? (b0 + kDeflateSub < n ? b0 + kDeflateSub : n) : b0;
      off[t] = total;
      if (e0 > b0) total += deflate_parse(in.data(), b0, e0, cand.data());
      uint32_t a, b;
      adler_piece(in.data() + deflate_at(b0), e0 - b0, &a, &b);
      adler_append(&A, &B, a, b, e0 - b0);
      chunk_crc = crc32_concat_raw(chunk_crc, crc32_raw(0u, in.data() + deflate_at(b0), e0 - b0), e0 - b0);
    }
    crc = crc32_concat_raw(crc, chunk_crc, n);
    const uint32_t flush_at = (total + 7 + 3 + 7) >> 3;
    const uint32_t comp = flush_at + 4;
    if (comp >= n + 5) {
      dst[out] = 0;
      dst[out + 1] = static_cast<uint8_t>(n & 0xFF);
      dst[out + 2] = static_cast<uint8_t>(n >> 8);
      dst[out + 3] = static_cast<uint8_t>(~n & 0xFF);
      dst[out + 4] = static_cast<uint8_t>((~n >> 8) & 0xFF);
      for (uint32_t i = 0; i < n; ++i) dst[out + 5 + i] = in[deflate_at(i)];
      out += n + 5;
    } else {
      deflate_put(words.data(), 0, 2u, 3);
      for (int t = 0; t < kDeflateThreads; ++t) {
        const uint32_t b0 = t * kDeflateSub;
        const uint32_t e0 = b0 < n ? (b0 + kDeflateSub < n ? b0 + kDeflateSub : n) : b0;
        if (e0 > b0) deflate_emit(in.data(), b0, e0, cand.data(), words.data(), off[t]);
      }
      deflate_put(words.data(), (flush_at + 2) * 8, 0xFFFFu, 16);
      memcpy(dst + out, words.data(), comp);
      out += comp;
    }
  }
  const uint8_t fin[5] = {0x01, 0x00, 0x00, 0xFF, 0xFF};
  memcpy(dst + out, fin, 5);
  out += 5;
  if (gzip) {
    const uint32_t v = (crc32_mulmod(0xFFFFFFFFu, crc32_xpow8n(nbytes)) ^ crc) ^ 0xFFFFFFFFu;
    const uint32_t isize = static_cast<uint32_t>(nbytes);
    for (int i = 0; i < 4; ++i) dst[out + i] = static_cast<uint8_t>(v >> (8 * i));
    for (int i = 0; i < 4; ++i) dst[out + 4 + i] = static_cast<uint8_t>(isize >> (8 * i));
    out += 8;
  } else {
    uint32_t A1 = 1, B1 = 0;
    adler_append(&A1, &B1, A, B, nbytes);
    const uint32_t v = (B1 << 16) | A1;
    for (int i = 0; i < 4; ++i) dst[out + i] = static_cast<uint8_t>(v >> (8 * (3 - i)));
    out += 4;
  }
  return out;
}

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
