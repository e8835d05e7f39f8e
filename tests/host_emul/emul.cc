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

// The device deflate encoder's pipeline, sequentially: the phases of deflate_chunk_kernel with a
// loop over the 128 threads where the kernel has a barrier.  Orchestration mirrors deflate.cu; every
// bit-level routine, the parse, the code construction and the token walks are the shared deflate.cuh
// code.  modes[0..2] (optional) count the chunks that ended up stored / fixed / dynamic.
uint64_t emul_deflate_modes(const uint8_t* src, uint64_t nbytes, uint32_t gzip, uint8_t* dst, uint32_t* modes) {
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
  constexpr int NS = kDeflateLitSyms + kDeflateDistSyms;
  std::vector<uint8_t> in(kDeflateInBytes);
  std::vector<uint16_t> tok(kDeflateChunk);
  std::vector<uint32_t> table(1 << kDeflateHashBits), words(kDeflateOutWords);
  std::vector<DeflateThread> th(kDeflateThreads);
  for (uint64_t base = 0; base < nbytes; base += kDeflateChunk) {
    const uint32_t n = static_cast<uint32_t>(nbytes - base < static_cast<uint64_t>(kDeflateChunk) ? nbytes - base : kDeflateChunk);
    uint32_t* inw = reinterpret_cast<uint32_t*>(in.data());
    std::fill(in.begin(), in.end(), 0);
    // load
    for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads); ++t) {
      DeflateThread& T = th[t];
      T.begin = t * kDeflateSub;
      T.end = T.begin < n ? (T.begin + kDeflateSub < n ? T.begin + kDeflateSub : n) : T.begin;
      for (int k = 0; k < 16; ++k) {
        uint32_t v = 0;
        for (uint32_t b = 0; b < 4u; ++b) {
          const uint32_t p = T.begin + 4u * k + b;
          if (p < n) v |= static_cast<uint32_t>(src[base + p]) << (8u * b);
        }
        T.w[2 + k] = v;
        inw[17u * t + k] = v;
      }
    }
    for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads); ++t) inw[17u * t + 16u] = t + 1 < static_cast<uint32_t>(kDeflateThreads) ? th[t + 1].w[2] : 0u;
    for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads); ++t) {
      th[t].w[0] = t > 0 ? inw[17u * (t - 1u) + 14u] : 0u;
      th[t].w[1] = t > 0 ? inw[17u * (t - 1u) + 15u] : 0u;
      th[t].w[18] = inw[17u * t + 16u];
    }
    // checksums
    uint32_t chunk_crc = 0;
    for (int t = 0; t < kDeflateThreads; ++t) {
      uint32_t a, b;
      adler_piece(in.data() + deflate_at(th[t].begin), th[t].end - th[t].begin, &a, &b);
      adler_append(&A, &B, a, b, th[t].end - th[t].begin);
      chunk_crc = crc32_concat_raw(chunk_crc, crc32_raw(0u, in.data() + deflate_at(th[t].begin), th[t].end - th[t].begin), th[t].end - th[t].begin);
    }
    crc = crc32_concat_raw(crc, chunk_crc, n);
    // masks + first-occurrence table
    std::fill(table.begin(), table.end(), kDeflateNoCand);
    for (int t = 0; t < kDeflateThreads; ++t) {
      deflate_masks(th[t]);
      for (uint32_t i = 0; i < 64u; ++i) {
        const uint32_t p = th[t].begin + i;
        if (p + 3u < n) {
          const uint32_t v = deflate_word_at(th[t], i);
          const uint32_t h = deflate_hash4(v);
          if (deflate_hashable(v) && p < table[h]) table[h] = p;
        }
      }
    }
    // parse + histograms
    uint32_t hist[NS] = {0}, ntok = 0, nmatch = 0;
    for (int t = 0; t < kDeflateThreads; ++t) {
      if (th[t].end > th[t].begin) {
        deflate_parse(in.data(), th[t], table.data(), tok.data(), n, [&](uint32_t a, uint32_t b, uint32_t times) {
          hist[a] += times;
          if (b != 0xFFFFFFFFu) hist[b] += times;
        });
        ntok += static_cast<uint32_t>(__builtin_popcountll(th[t].is_start));
        nmatch += static_cast<uint32_t>(__builtin_popcountll(th[t].is_match));
      } else {
        th[t].is_start = th[t].is_match = 0;
      }
    }
    hist[256] = 1;
    // codes: Shannon classes, completed histogram, lengths by (class, index), canonical codes
    uint8_t len_tab[NS];
    uint16_t code_tab[NS];
    uint32_t cls[NS], bl[2][16] = {{0}}, bl0[2][16], next_code[2][16];
    for (int s2 = 0; s2 < NS; ++s2) {
      const int a = s2 < kDeflateLitSyms ? 0 : 1;
      cls[s2] = hist[s2] != 0u ? deflate_shannon_len(hist[s2], a == 0 ? ntok + 1u : nmatch, 15u) : 0u;
      if (cls[s2]) bl[a][cls[s2]] += 1u;
    }
    memcpy(bl0, bl, sizeof(bl));
    deflate_complete_code(bl[0], 15u);
    {
      uint32_t used = 0;
      for (int k = 1; k <= 15; ++k) used += bl[1][k];
      if (used >= 2u) deflate_complete_code(bl[1], 15u);
    }
    for (int s2 = 0; s2 < NS; ++s2) {
      len_tab[s2] = 0;
      if (cls[s2] == 0u) continue;
      const int a = s2 < kDeflateLitSyms ? 0 : 1;
      uint32_t pos = 0;
      for (int j = a == 0 ? 0 : kDeflateLitSyms; j < s2; ++j) pos += cls[j] == cls[s2] ? 1u : 0u;
      for (uint32_t k = 1; k < cls[s2]; ++k) pos += bl0[a][k];
      uint32_t k = 1, cum = bl[a][1];
      while (pos >= cum && k < 15u) {
        ++k;
        cum += bl[a][k];
      }
      len_tab[s2] = static_cast<uint8_t>(k);
    }
    deflate_next_codes(bl[0], 15u, next_code[0]);
    deflate_next_codes(bl[1], 15u, next_code[1]);
    for (int s2 = 0; s2 < NS; ++s2) {
      code_tab[s2] = 0;
      if (len_tab[s2] == 0) continue;
      const int a = s2 < kDeflateLitSyms ? 0 : 1;
      uint32_t idx = 0;
      for (int j = a == 0 ? 0 : kDeflateLitSyms; j < s2; ++j) idx += len_tab[j] == len_tab[s2] ? 1u : 0u;
      code_tab[s2] = static_cast<uint16_t>(deflate_reverse(next_code[a][len_tab[s2]] + idx, len_tab[s2]));
    }
    // header (thread 0 of the kernel)
    uint8_t hdr_sym[kDeflateHdrMax + 8], hdr_extra[kDeflateHdrMax + 8];
    uint32_t cl_len[19], cl_code[19];
    uint32_t nlit = 286, ndist = 30;
    while (nlit > 257u && len_tab[nlit - 1u] == 0u) --nlit;
    while (ndist > 1u && len_tab[kDeflateLitSyms + ndist - 1u] == 0u) --ndist;
    uint32_t cnt[19] = {0}, lbl[8] = {0}, lnext[8], ord[19];
    uint32_t ne = 0;
    for (uint32_t seg = 0; seg < 32u; ++seg) {  // one lane per segment on the device
      const uint32_t entries = deflate_segment_entries(len_tab, nlit, nlit + ndist, seg);
      deflate_segment_write(len_tab, nlit, nlit + ndist, seg, entries, ne, hdr_sym, hdr_extra, [&](uint32_t v) { cnt[v] += 1u; });
      ne += entries;
    }
    uint32_t used = 0;
    for (int i = 0; i < 19; ++i) {
      cl_len[i] = cnt[i] != 0u ? deflate_shannon_len(cnt[i], ne, 7u) : 0u;
      if (cnt[i] != 0u) {
        lbl[cl_len[i]] += 1u;
        ++used;
      }
    }
    if (used == 1u) {
      for (int i = 0; i < 19; ++i) {
        if (cnt[i] == 0u) {
          cl_len[i] = 1;
          lbl[1] += 1u;
          break;
        }
      }
    }
    uint32_t no = 0;
    for (uint32_t k = 1; k <= 7u; ++k) {
      for (uint32_t i = 0; i < 19u; ++i) {
        if (cl_len[i] == k) ord[no++] = i;
      }
    }
    deflate_complete_code(lbl, 7u);
    {
      uint32_t k = 1, left = lbl[1];
      for (uint32_t o = 0; o < no; ++o) {
        while (left == 0u && k < 7u) {
          ++k;
          left = lbl[k];
        }
        cl_len[ord[o]] = k;
        --left;
      }
    }
    deflate_next_codes(lbl, 7u, lnext);
    for (uint32_t i = 0; i < 19u; ++i) cl_code[i] = cl_len[i] != 0u ? deflate_reverse(lnext[cl_len[i]]++, cl_len[i]) : 0u;
    uint32_t ncl = 19;
    while (ncl > 4u && cl_len[deflate_cl_order(ncl - 1u)] == 0u) --ncl;
    uint32_t hdr_bits = 3u + 5u + 5u + 4u + 3u * ncl;
    for (uint32_t e = 0; e < ne; ++e) hdr_bits += cl_len[hdr_sym[e]] + deflate_cl_extra_bits(hdr_sym[e]);
    // counts, mode, offsets
    uint32_t sub_dyn[kDeflateThreads], sub_fix[kDeflateThreads], sub_off[kDeflateThreads];
    uint32_t dyn = hdr_bits, fix = 3u;
    for (int t = 0; t < kDeflateThreads; ++t) {
      sub_dyn[t] = sub_fix[t] = 0;
      if (th[t].end > th[t].begin) deflate_count(in.data(), th[t], tok.data(), len_tab, len_tab + kDeflateLitSyms, &sub_dyn[t], &sub_fix[t]);
      dyn += sub_dyn[t];
      fix += sub_fix[t];
    }
    dyn += len_tab[256];
    fix += 7u;
    const uint32_t best = dyn < fix ? dyn : fix;
    const uint32_t comp_bytes = ((best + 3u + 7u) >> 3) + 4u;
    const uint32_t mode = comp_bytes >= n + 5u ? 0u : (dyn < fix ? 2u : 1u);
    if (modes != nullptr) {
      modes[mode] += 1u;
      if (mode == 2u) {  // diagnostics: header bits, body bits, tokens, matches of the dynamic chunks
        modes[3] += hdr_bits;
        modes[4] += dyn - hdr_bits;
        modes[5] += ntok;
        modes[6] += nmatch;
        modes[7] += ne;
      }
    }
    if (mode == 0u) {
      dst[out] = 0;
      dst[out + 1] = static_cast<uint8_t>(n & 0xFF);
      dst[out + 2] = static_cast<uint8_t>(n >> 8);
      dst[out + 3] = static_cast<uint8_t>(~n & 0xFF);
      dst[out + 4] = static_cast<uint8_t>((~n >> 8) & 0xFF);
      for (uint32_t i = 0; i < n; ++i) dst[out + 5 + i] = in[deflate_at(i)];
      out += n + 5;
      continue;
    }
    std::fill(words.begin(), words.end(), 0u);
    uint32_t off = mode == 2u ? hdr_bits : 3u;
    for (int t = 0; t < kDeflateThreads; ++t) {
      sub_off[t] = off;
      off += mode == 2u ? sub_dyn[t] : sub_fix[t];
    }
    const uint32_t total_bits = off;
    if (mode == 2u) {
      deflate_put(words.data(), 0, 4u, 3);
      deflate_put(words.data(), 3, nlit - 257u, 5);
      deflate_put(words.data(), 8, ndist - 1u, 5);
      deflate_put(words.data(), 13, ncl - 4u, 4);
      for (uint32_t i = 0; i < ncl; ++i) deflate_put(words.data(), 17u + 3u * i, cl_len[deflate_cl_order(i)], 3);
      uint32_t pos = 17u + 3u * ncl;
      for (uint32_t e = 0; e < ne; ++e) {
        const uint32_t sy = hdr_sym[e], l = cl_len[sy], xb = deflate_cl_extra_bits(sy);
        deflate_put(words.data(), pos, cl_code[sy] | (static_cast<uint32_t>(hdr_extra[e]) << l), l + xb);
        pos += l + xb;
      }
      for (int t = 0; t < kDeflateThreads; ++t) {
        if (th[t].end > th[t].begin) {
          deflate_emit(in.data(), th[t], tok.data(), code_tab, len_tab, code_tab + kDeflateLitSyms, len_tab + kDeflateLitSyms, words.data(), sub_off[t]);
        }
      }
      deflate_put(words.data(), total_bits, code_tab[256], len_tab[256]);
    } else {
      deflate_put(words.data(), 0, 2u, 3);
      for (int t = 0; t < kDeflateThreads; ++t) {
        if (th[t].end > th[t].begin) deflate_emit(in.data(), th[t], tok.data(), nullptr, nullptr, nullptr, nullptr, words.data(), sub_off[t]);
      }
    }
    const uint32_t eob_bits = mode == 2u ? len_tab[256] : 7u;
    const uint32_t flush_at = (total_bits + eob_bits + 3u + 7u) >> 3;
    deflate_put(words.data(), (flush_at + 2u) * 8u, 0xFFFFu, 16);
    memcpy(dst + out, words.data(), flush_at + 4u);
    out += flush_at + 4u;
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
uint64_t emul_deflate(const uint8_t* src, uint64_t nbytes, uint32_t gzip, uint8_t* dst) { return emul_deflate_modes(src, nbytes, gzip, dst, nullptr); }

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

// diagnostics: the tokens of the first chunk (is_start / is_match / tok), out[3*k] = position, [3*k+1] = len (0 literal), [3*k+2] = dist
extern "C" uint32_t emul_deflate_tokens(const uint8_t* src, uint32_t n, uint32_t* out, uint32_t cap) {
  using namespace tb200;
  std::vector<uint8_t> in(kDeflateInBytes, 0);
  std::vector<uint16_t> tok(kDeflateChunk);
  std::vector<uint32_t> table(1 << kDeflateHashBits, kDeflateNoCand);
  std::vector<DeflateThread> th(kDeflateThreads);
  uint32_t* inw = reinterpret_cast<uint32_t*>(in.data());
  for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads); ++t) {
    th[t].begin = t * kDeflateSub;
    th[t].end = th[t].begin < n ? (th[t].begin + kDeflateSub < n ? th[t].begin + kDeflateSub : n) : th[t].begin;
    for (int k = 0; k < 16; ++k) {
      uint32_t v = 0;
      for (uint32_t b = 0; b < 4u; ++b) {
        const uint32_t p = th[t].begin + 4u * k + b;
        if (p < n) v |= static_cast<uint32_t>(src[p]) << (8u * b);
      }
      th[t].w[2 + k] = v;
      inw[17u * t + k] = v;
    }
  }
  for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads); ++t) inw[17u * t + 16u] = t + 1 < static_cast<uint32_t>(kDeflateThreads) ? th[t + 1].w[2] : 0u;
  for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads); ++t) {
    th[t].w[0] = t > 0 ? inw[17u * (t - 1u) + 14u] : 0u;
    th[t].w[1] = t > 0 ? inw[17u * (t - 1u) + 15u] : 0u;
    th[t].w[18] = inw[17u * t + 16u];
    deflate_masks(th[t]);
    for (uint32_t i = 0; i < 64u; ++i) {
      const uint32_t p = th[t].begin + i;
      if (p + 3u < n) {
        const uint32_t v = deflate_word_at(th[t], i);
        const uint32_t h = deflate_hash4(v);
        if (deflate_hashable(v) && p < table[h]) table[h] = p;
      }
    }
  }
  uint32_t k = 0;
  for (uint32_t t = 0; t < static_cast<uint32_t>(kDeflateThreads); ++t) {
    if (th[t].end <= th[t].begin) continue;
    deflate_parse(in.data(), th[t], table.data(), tok.data(), n, [&](uint32_t, uint32_t, uint32_t) {});
    for (uint32_t i = 0; i < 64u; ++i) {
      if (!((th[t].is_start >> i) & 1ull) || k >= cap) continue;
      const uint32_t p = th[t].begin + i;
      out[3 * k] = p;
      out[3 * k + 1] = ((th[t].is_match >> i) & 1ull) ? tok[p + 1] : 0u;
      out[3 * k + 2] = ((th[t].is_match >> i) & 1ull) ? tok[p] : 0u;
      ++k;
    }
  }
  return k;
}
