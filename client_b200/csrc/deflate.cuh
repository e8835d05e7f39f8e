// deflate.cuh -- the per-chunk logic of the device deflate encoder (RFC 1951 fixed-Huffman
// blocks with LZ77 matches), written as host+device functions so that tests/host_emul can
// run the exact code on the CPU and hand its output to zlib.
//
// Why: the reference compresses a request body with zlib / gzip on one host core
// (PY/http/_client.py:1440-1460 -> gzip.compress / zlib.compress; CC/http_client.cc:146-221).
// When the body is generated on the device (wire mode) the encoder belongs there too.  Any
// valid deflate stream is acceptable to the peer, so parity here means: the reference's own
// decompressor (zlib) returns the original bytes -- not byte equality with zlib's encoder.
//
// Layout of the work: the input is cut into chunks of kDeflateChunk bytes, one CTA each; a
// chunk becomes ONE fixed-Huffman block followed by an empty stored block (a "sync flush":
// 00 00 FF FF), which ends on a byte boundary, so chunks concatenate bytewise (pigz does the
// same between threads).  Inside a chunk every thread parses its own kDeflateSub-byte
// sub-block: matches may point anywhere earlier in the chunk but do not run past the
// sub-block's end, so sub-blocks parse independently; a first pass counts bits, a prefix
// sum gives every sub-block its bit offset, a second pass emits.  A chunk that does not
// shrink is stored (BTYPE=00).
#ifndef TB200_CSRC_DEFLATE_CUH_
#define TB200_CSRC_DEFLATE_CUH_

#include <cstdint>

#include "philox.cuh"  // TB200_HD

namespace tb200 {

constexpr int kDeflateChunk = 8192;   // bytes per CTA
constexpr int kDeflateSub = 64;       // bytes per thread
constexpr int kDeflateSubShift = 6;
constexpr int kDeflateThreads = kDeflateChunk / kDeflateSub;  // 128
constexpr int kDeflateHashBits = 11;   // 2048 heads of 32 bits: atomicMax keeps the latest position
constexpr uint32_t kDeflateNoCand = 0xFFFFu;
// worst case of a fixed-Huffman chunk: 9 bits per byte + header/EOB/flush
constexpr int kDeflateOutWords = (kDeflateChunk * 9 / 8 + 64) / 4;
constexpr int kDeflateMaxChunkOut = kDeflateChunk + 16;  // stored fallback bounds the output

// ---- bit stream (LSB-first within bytes, RFC 1951 section 3.1.1) ---------------------------
TB200_HD void deflate_or(uint32_t* word, uint32_t bits) {
#ifdef __CUDA_ARCH__
  atomicOr(word, bits);  // neighbouring sub-blocks share boundary words
#else
  *word |= bits;
#endif
}
// value's low `nbits` (<= 25) bits at absolute bit position `pos`
TB200_HD void deflate_put(uint32_t* words, uint32_t pos, uint32_t value, uint32_t nbits) {
  if (nbits == 0) return;
  const uint32_t w = pos >> 5, sh = pos & 31u;
  deflate_or(words + w, value << sh);
  if (sh + nbits > 32u) deflate_or(words + w + 1, value >> (32u - sh));
}
TB200_HD uint32_t deflate_reverse(uint32_t code, uint32_t nbits) {
  uint32_t r = 0;
  for (uint32_t i = 0; i < nbits; ++i) r |= ((code >> i) & 1u) << (nbits - 1u - i);
  return r;
}

// ---- fixed Huffman tables (RFC 1951 section 3.2.6), codes are sent MSB first ---------------
TB200_HD void deflate_litlen_code(uint32_t sym, uint32_t* code, uint32_t* nbits) {
  if (sym < 144u) { *code = 0x30u + sym; *nbits = 8; }
  else if (sym < 256u) { *code = 0x190u + (sym - 144u); *nbits = 9; }
  else if (sym < 280u) { *code = sym - 256u; *nbits = 7; }
  else { *code = 0xC0u + (sym - 280u); *nbits = 8; }
}
// length 3..258 -> (symbol 257..285, extra bit count, extra value)
TB200_HD void deflate_len_symbol(uint32_t len, uint32_t* sym, uint32_t* ebits, uint32_t* eval) {
  if (len == 258u) { *sym = 285; *ebits = 0; *eval = 0; return; }
  const uint32_t l = len - 3u;  // 0..254
  if (l < 8u) { *sym = 257u + l; *ebits = 0; *eval = 0; return; }
  // groups of 4 symbols share an extra-bit count e = 1..5: base offset 8, 16, 32, 64, 128
  uint32_t e = 1;
  while ((8u << e) <= l) ++e;  // l in [8<<(e-1), 8<<e)
  const uint32_t base = 8u << (e - 1u);
  *sym = 261u + 4u * e + ((l - base) >> e);
  *ebits = e;
  *eval = (l - base) & ((1u << e) - 1u);
}
// distance 1..32768 -> (code 0..29, extra bit count, extra value)
TB200_HD void deflate_dist_symbol(uint32_t dist, uint32_t* code, uint32_t* ebits, uint32_t* eval) {
  const uint32_t d = dist - 1u;
  if (d < 4u) { *code = d; *ebits = 0; *eval = 0; return; }
  uint32_t e = 1;
  while ((4u << e) <= d) ++e;  // d in [4<<(e-1), 4<<e)
  const uint32_t base = 4u << (e - 1u);
  *code = 2u + 2u * e + ((d - base) >> e);
  *ebits = e;
  *eval = (d - base) & ((1u << e) - 1u);
}

TB200_HD uint32_t deflate_literal_bits(uint32_t byte) { return byte < 144u ? 8u : 9u; }
TB200_HD uint32_t deflate_match_bits(uint32_t len, uint32_t dist) {
  uint32_t sym, eb, ev, code, nb, dc, deb, dev;
  deflate_len_symbol(len, &sym, &eb, &ev);
  deflate_litlen_code(sym, &code, &nb);
  deflate_dist_symbol(dist, &dc, &deb, &dev);
  return nb + eb + 5u + deb;
}
TB200_HD uint32_t deflate_emit_literal(uint32_t* words, uint32_t pos, uint32_t byte) {
  uint32_t code, nb;
  deflate_litlen_code(byte, &code, &nb);
  deflate_put(words, pos, deflate_reverse(code, nb), nb);
  return nb;
}
TB200_HD uint32_t deflate_emit_match(uint32_t* words, uint32_t pos, uint32_t len, uint32_t dist) {
  uint32_t sym, eb, ev, code, nb, dc, deb, dev;
  deflate_len_symbol(len, &sym, &eb, &ev);
  deflate_litlen_code(sym, &code, &nb);
  deflate_dist_symbol(dist, &dc, &deb, &dev);
  uint32_t p = pos;
  deflate_put(words, p, deflate_reverse(code, nb), nb);
  p += nb;
  deflate_put(words, p, ev, eb);
  p += eb;
  deflate_put(words, p, deflate_reverse(dc, 5), 5);
  p += 5;
  deflate_put(words, p, dev, deb);
  p += deb;
  return p - pos;
}

// ---- match finding --------------------------------------------------------------------------
// The chunk sits in shared memory with every sub-block shifted by one more word: thread t
// walks sub-block t, and without the skew the threads of a warp would hit the same banks on
// every byte they compare.  A sub-block itself stays contiguous.
TB200_HD uint32_t deflate_at(uint32_t p) { return p + ((p >> kDeflateSubShift) << 2); }
constexpr int kDeflateInBytes = kDeflateChunk + (kDeflateChunk / kDeflateSub) * 4 + 16;

TB200_HD uint32_t deflate_hash(const uint8_t* in, uint32_t p) {
  const uint32_t v = static_cast<uint32_t>(in[deflate_at(p)]) | (static_cast<uint32_t>(in[deflate_at(p + 1)]) << 8) |
                     (static_cast<uint32_t>(in[deflate_at(p + 2)]) << 16) | (static_cast<uint32_t>(in[deflate_at(p + 3)]) << 24);
  return (v * 2654435761u) >> (32 - kDeflateHashBits);
}
TB200_HD uint32_t deflate_match_len(const uint8_t* in, uint32_t a, uint32_t b, uint32_t maxlen) {
  uint32_t n = 0;
  while (n < maxlen && in[deflate_at(a + n)] == in[deflate_at(b + n)]) ++n;
  return n;
}
// best (len, dist) at position p, match not longer than `room`; candidates: the hash chain
// head recorded for p and the short periodic distances typed tensors are full of
TB200_HD void deflate_best(const uint8_t* in, uint32_t p, uint32_t room, uint32_t cand, uint32_t* len, uint32_t* dist) {
  uint32_t best = 0, bd = 0;
  const uint32_t maxlen = room < 258u ? room : 258u;
  if (maxlen >= 3u) {
    const uint32_t fixed[4] = {1u, 2u, 4u, 8u};
    for (int i = 0; i < 5; ++i) {
      uint32_t d;
      if (i < 4) d = fixed[i];
      else if (cand != kDeflateNoCand && cand < p) d = p - cand;
      else continue;
      if (d > p) continue;
      const uint32_t l = deflate_match_len(in, p - d, p, maxlen);
      if (l > best) { best = l; bd = d; }
    }
  }
  if (best < 3u) { best = 0; bd = 0; }
  *len = best;
  *dist = bd;
}

// Greedy parse of sub-block [begin, end) of the chunk `in`; cand[p] = an earlier position with
// the same 4-byte hash (or kDeflateNoCand).  Counts the bits and leaves the decisions in the
// cand array itself for the emitting pass: cand[p] = distance (0 = literal) and, for a match,
// cand[p + 1] = length (a match covers >= 3 positions of its own sub-block, and the positions
// it covers are never parsed, so their candidates are dead).
TB200_HD uint32_t deflate_parse(const uint8_t* in, uint32_t begin, uint32_t end, uint16_t* cand) {
  uint32_t bits = 0;
  uint32_t p = begin;
  while (p < end) {
    uint32_t len, dist;
    deflate_best(in, p, end - p, cand[p], &len, &dist);
    if (len >= 3u) {
      bits += deflate_match_bits(len, dist);
      cand[p] = static_cast<uint16_t>(dist);
      cand[p + 1] = static_cast<uint16_t>(len);
      p += len;
    } else {
      bits += deflate_literal_bits(in[deflate_at(p)]);
      cand[p] = 0;
      p += 1;
    }
  }
  return bits;
}
// Emit the tokens deflate_parse left behind, starting at bit `bitpos`.
TB200_HD void deflate_emit(const uint8_t* in, uint32_t begin, uint32_t end, const uint16_t* tokens, uint32_t* words, uint32_t bitpos) {
  uint32_t pos = bitpos;
  uint32_t p = begin;
  while (p < end) {
    const uint32_t dist = tokens[p];
    if (dist != 0u) {
      const uint32_t len = tokens[p + 1];
      pos += deflate_emit_match(words, pos, len, dist);
      p += len;
    } else {
      pos += deflate_emit_literal(words, pos, in[deflate_at(p)]);
      p += 1;
    }
  }
}

// ---- checksums ---------------------------------------------------------------------------------
// Adler-32 of a piece as (a, b) with a starting from 0 (not 1): pieces combine linearly
constexpr uint32_t kAdlerMod = 65521u;
TB200_HD void adler_piece(const uint8_t* d, uint32_t n, uint32_t* a_out, uint32_t* b_out) {
  uint32_t a = 0, b = 0;
  for (uint32_t i = 0; i < n; ++i) {  // n <= kDeflateSub: no overflow
    a += d[i];
    b += a;
  }
  *a_out = a % kAdlerMod;
  *b_out = b % kAdlerMod;
}
// running (A, B) over the stream so far, then a piece (a, b) of n bytes:
//   A' = A + a ;  B' = B + n*A + b      (all mod 65521)
TB200_HD void adler_append(uint32_t* A, uint32_t* B, uint32_t a, uint32_t b, uint64_t n) {
  const uint64_t nb = n % kAdlerMod;
  *B = static_cast<uint32_t>((*B + nb * *A + b) % kAdlerMod);
  *A = (*A + a) % kAdlerMod;
}

// CRC-32 (IEEE, reflected, poly 0xEDB88320).  crc_raw = register value without the final xor,
// starting from `init`.
TB200_HD uint32_t crc32_raw(uint32_t init, const uint8_t* d, uint32_t n) {
  uint32_t c = init;
  for (uint32_t i = 0; i < n; ++i) {
    c ^= d[i];
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ (0xEDB88320u & (0u - (c & 1u)));
  }
  return c;
}
// the same through a 256-entry table (tbl[i] = crc32_raw(0, {i}, 1))
TB200_HD uint32_t crc32_raw_tbl(const uint32_t* tbl, uint32_t init, const uint8_t* d, uint32_t n) {
  uint32_t c = init;
  for (uint32_t i = 0; i < n; ++i) c = tbl[(c ^ d[i]) & 0xFFu] ^ (c >> 8);
  return c;
}
// multiply two polynomials mod P in the reflected representation (zlib's multmodp)
TB200_HD uint32_t crc32_mulmod(uint32_t a, uint32_t b) {
  if (a == 0u || b == 0u) return 0u;
  uint32_t m = 1u << 31, p = 0;
  for (;;) {
    if (a & m) {
      p ^= b;
      if ((a & (m - 1u)) == 0) break;
    }
    m >>= 1;
    b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
  }
  return p;
}
// x^(8n) mod P
TB200_HD uint32_t crc32_xpow8n(uint64_t n) {
  uint32_t result = 1u << 31;          // x^0
  uint32_t sq = 1u << 23;              // x^8 in the reflected representation
  while (n) {
    if (n & 1u) result = crc32_mulmod(result, sq);
    sq = crc32_mulmod(sq, sq);
    n >>= 1;
  }
  return result;
}
// raw register after stream X || Y, given raw(X) (any init), and raw0(Y) computed from init 0
TB200_HD uint32_t crc32_concat_raw(uint32_t raw_x, uint32_t raw0_y, uint64_t len_y) {
  return crc32_mulmod(raw_x, crc32_xpow8n(len_y)) ^ raw0_y;
}

}  // namespace tb200

#endif  // TB200_CSRC_DEFLATE_CUH_
