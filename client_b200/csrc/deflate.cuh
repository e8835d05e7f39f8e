// deflate.cuh -- the per-chunk logic of the device deflate encoder (RFC 1951: LZ77 matches +
// a DYNAMIC Huffman code per chunk, fixed-Huffman and stored fallbacks), written as host+device
// functions so that tests/host_emul can run the exact code on the CPU and hand its output to zlib.
//
// Why: the reference compresses a request body with zlib / gzip on one host core
// (PY/http/_client.py:1440-1460 -> gzip.compress / zlib.compress; CC/http_client.cc:146-221).
// When the body is generated on the device (wire mode) the encoder belongs there too.  Any
// valid deflate stream is acceptable to the peer, so parity here means: the reference's own
// decompressor (zlib) returns the original bytes -- not byte equality with zlib's encoder.
//
// Layout of the work (round 2).  The input is cut into chunks of kDeflateChunk bytes, one CTA of
// kDeflateThreads threads each; a chunk becomes ONE deflate block followed by an empty stored block
// (a "sync flush": 00 00 FF FF), which ends on a byte boundary, so chunks concatenate bytewise.
// Thread t owns sub-block t: kDeflateSub bytes held in registers.
//   1. match finding.  Typed tensors repeat at distances 1, 2, 4, 8: for each of them a 64-bit
//      equality mask of the sub-block (byte i == byte i-d, SIMD byte compares on the register
//      words) turns "how long is the match at p" into a count of trailing ones -- no byte loop.
//      Longer distances: a hash table of the FIRST position of every 4-byte hash in the chunk
//      (one atomicMin pass), verified against shared memory only when it hits.
//   2. greedy parse per sub-block (matches do not cross a sub-block's end, so sub-blocks parse
//      independently), tokens left in shared memory, symbol histograms by shared atomics.
//   3. a code per chunk: code lengths = ceil(log2(total / count)) (a Shannon code, Kraft sum <= 1),
//      made complete on the 15-bin length histogram (promote the longest steps that still fit),
//      lengths handed out by frequency rank, canonical codes -- every step is parallel over the
//      symbols except a 15-iteration fix of the histogram; within ~1 % of a true Huffman code on
//      token ids.  The code-length header (RLE symbols 16 / 17 / 18, code-length code of <= 7 bits)
//      is formed by one thread and emitted by all.
//   4. bit counts per sub-block under the dynamic and the fixed code, the cheaper of dynamic /
//      fixed / stored wins per chunk; prefix sum; every thread emits its sub-block.
#ifndef TB200_CSRC_DEFLATE_CUH_
#define TB200_CSRC_DEFLATE_CUH_

#include <cstdint>

#include "philox.cuh"  // TB200_HD

namespace tb200 {

constexpr int kDeflateChunk = 8192;   // bytes per CTA
constexpr int kDeflateSub = 64;       // bytes per thread
constexpr int kDeflateSubShift = 6;
constexpr int kDeflateThreads = kDeflateChunk / kDeflateSub;  // 128
constexpr int kDeflateHashBits = 11;   // 2048 entries: first position of a 4-byte hash in the chunk
constexpr uint32_t kDeflateNoCand = 0xFFFFFFFFu;
constexpr int kDeflateLitSyms = 288, kDeflateDistSyms = 32;  // 286 / 30 used
constexpr int kDeflateHdrMax = 320;                          // code-length entries of a header (<= 286 + 30)
// worst case of a chunk: 9 bits per byte (fixed code) or 15 bits per byte in a hypothetical dynamic
// code never chosen (the cheapest of dynamic / fixed / stored is taken), + header + flush
constexpr int kDeflateOutWords = (kDeflateChunk * 9 / 8 + 512) / 4;
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
#ifdef __CUDA_ARCH__
  return nbits == 0 ? 0u : (__brev(code) >> (32u - nbits));
#else
  uint32_t r = 0;
  for (uint32_t i = 0; i < nbits; ++i) r |= ((code >> i) & 1u) << (nbits - 1u - i);
  return r;
#endif
}

// ---- fixed Huffman tables (RFC 1951 section 3.2.6), codes are sent MSB first ---------------
TB200_HD void deflate_litlen_code(uint32_t sym, uint32_t* code, uint32_t* nbits) {
  if (sym < 144u) { *code = 0x30u + sym; *nbits = 8; }
  else if (sym < 256u) { *code = 0x190u + (sym - 144u); *nbits = 9; }
  else if (sym < 280u) { *code = sym - 256u; *nbits = 7; }
  else { *code = 0xC0u + (sym - 280u); *nbits = 8; }
}
TB200_HD uint32_t deflate_fixed_litlen_bits(uint32_t sym) { return sym < 144u ? 8u : (sym < 256u ? 9u : (sym < 280u ? 7u : 8u)); }
// length 3..258 -> (symbol 257..285, extra bit count, extra value)
TB200_HD void deflate_len_symbol(uint32_t len, uint32_t* sym, uint32_t* ebits, uint32_t* eval) {
  if (len == 258u) { *sym = 285; *ebits = 0; *eval = 0; return; }
  const uint32_t l = len - 3u;  // 0..254
  if (l < 8u) { *sym = 257u + l; *ebits = 0; *eval = 0; return; }
  // groups of 4 symbols share an extra-bit count e = 1..5: base offset 8, 16, 32, 64, 128
#ifdef __CUDA_ARCH__
  const uint32_t e = 29u - static_cast<uint32_t>(__clz(l));  // l in [8<<(e-1), 8<<e)
#else
  uint32_t e = 1;
  while ((8u << e) <= l) ++e;
#endif
  const uint32_t base = 8u << (e - 1u);
  *sym = 261u + 4u * e + ((l - base) >> e);
  *ebits = e;
  *eval = (l - base) & ((1u << e) - 1u);
}
// distance 1..32768 -> (code 0..29, extra bit count, extra value)
TB200_HD void deflate_dist_symbol(uint32_t dist, uint32_t* code, uint32_t* ebits, uint32_t* eval) {
  const uint32_t d = dist - 1u;
  if (d < 4u) { *code = d; *ebits = 0; *eval = 0; return; }
#ifdef __CUDA_ARCH__
  const uint32_t e = 30u - static_cast<uint32_t>(__clz(d));  // d in [4<<(e-1), 4<<e)
#else
  uint32_t e = 1;
  while ((4u << e) <= d) ++e;
#endif
  const uint32_t base = 4u << (e - 1u);
  *code = 2u + 2u * e + ((d - base) >> e);
  *ebits = e;
  *eval = (d - base) & ((1u << e) - 1u);
}

// ---- the chunk in shared memory ----------------------------------------------------------------
// Every sub-block is shifted by one more word: thread t reads sub-block t, and without the skew the
// threads of a warp would hit the same two banks on every access.  A sub-block itself stays contiguous.
TB200_HD uint32_t deflate_at(uint32_t p) { return p + ((p >> kDeflateSubShift) << 2); }
constexpr int kDeflateInBytes = kDeflateChunk + (kDeflateChunk / kDeflateSub) * 4 + 16;

// what a thread keeps in registers between the phases
struct DeflateThread {
  uint32_t w[19];        // w[0..1]: the 8 bytes before the sub-block, w[2..17]: the sub-block, w[18]: the 4 bytes after it
  uint64_t eq[4];        // bit i: byte i == byte i - d, d = 1, 2, 4, 8
  uint64_t is_match;     // bit i: the token that starts at byte i is a match
  uint64_t is_start;     // bit i: a token starts at byte i
  uint32_t begin, end;   // the sub-block's range in the chunk
};

// 4-bit mask: byte k of a equals byte k of b
TB200_HD uint32_t deflate_eq4(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  const uint32_t m = __vcmpeq4(a, b) & 0x01010101u;  // 0x01 per equal byte
  return (m * 0x01020408u) >> 24;                    // gather bits 0, 8, 16, 24 -> 0..3
#else
  uint32_t r = 0;
  for (int k = 0; k < 4; ++k) r |= (((a >> (8 * k)) & 0xFFu) == ((b >> (8 * k)) & 0xFFu) ? 1u : 0u) << k;
  return r;
#endif
}

// equality masks of the sub-block for distances 1, 2, 4, 8 from the register words
TB200_HD void deflate_masks(DeflateThread& t) {
  uint64_t m1 = 0, m2 = 0, m4 = 0, m8 = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const uint32_t cur = t.w[2 + k], a = t.w[1 + k], b = t.w[k];
    m1 |= static_cast<uint64_t>(deflate_eq4(cur, (cur << 8) | (a >> 24))) << (4 * k);
    m2 |= static_cast<uint64_t>(deflate_eq4(cur, (cur << 16) | (a >> 16))) << (4 * k);
    m4 |= static_cast<uint64_t>(deflate_eq4(cur, a)) << (4 * k);
    m8 |= static_cast<uint64_t>(deflate_eq4(cur, b)) << (4 * k);
  }
  const uint32_t len = t.end - t.begin;
  const uint64_t valid = len >= 64u ? ~0ull : ((1ull << len) - 1ull);
  // the first sub-block has nothing before it: positions < d cannot match at distance d
  const uint64_t first = t.begin == 0u ? 1ull : 0ull;
  t.eq[0] = m1 & valid & ~(first * 0x01ull);
  t.eq[1] = m2 & valid & ~(first * 0x03ull);
  t.eq[2] = m4 & valid & ~(first * 0x0Full);
  t.eq[3] = m8 & valid & ~(first * 0xFFull);
}

// the 4 bytes at sub-block offset i (0..63) as a little-endian word, from the register words
TB200_HD uint32_t deflate_word_at(const DeflateThread& t, uint32_t i) {
  // dynamic register indexing would go through local memory: select by compare instead
  uint32_t lo = 0, hi = 0;
  const uint32_t k = i >> 2;
#pragma unroll
  for (uint32_t j = 0; j < 16; ++j) {
    if (j == k) {
      lo = t.w[2 + j];
      hi = t.w[3 + j];
    }
  }
  const uint32_t sh = (i & 3u) * 8u;
  return sh == 0u ? lo : ((lo >> sh) | (hi << (32u - sh)));
}
TB200_HD uint32_t deflate_hash4(uint32_t v) { return (v * 2654435761u) >> (32 - kDeflateHashBits); }
// Windows of four equal bytes (runs of zeros above all) are what the distance-1 mask finds anyway; kept out
// of the hash table they do not pile thousands of shared-memory atomics onto one bucket.
TB200_HD bool deflate_hashable(uint32_t v) { return v != ((v << 8) | (v >> 24)); }
// the 4 bytes at chunk position p from the skewed image: two aligned words of the sub-block (the gap word
// behind a sub-block repeats the first word of the next one, so this works up to the sub-block's last byte)
TB200_HD uint32_t deflate_load4(const uint8_t* in, uint32_t p) {
  const uint32_t a = deflate_at(p);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(in + (a & ~3u));
  const uint32_t sh = (a & 3u) * 8u;
  return sh == 0u ? w[0] : ((w[0] >> sh) | (w[1] << (32u - sh)));
}

TB200_HD uint32_t deflate_ctz64(uint64_t x) {  // x != 0
#ifdef __CUDA_ARCH__
  return static_cast<uint32_t>(__ffsll(static_cast<long long>(x)) - 1);
#else
  return static_cast<uint32_t>(__builtin_ctzll(x));
#endif
}
// number of consecutive one bits of m starting at bit p
TB200_HD uint32_t deflate_run(uint64_t m, uint32_t p) {
  const uint64_t inv = ~(m >> p);
  return inv == 0ull ? 64u - p : deflate_ctz64(inv);
}

// match length between chunk positions a < b in the skewed byte image, at most maxlen
TB200_HD uint32_t deflate_match_len(const uint8_t* in, uint32_t a, uint32_t b, uint32_t maxlen) {
  uint32_t n = 0;
  while (n < maxlen && in[deflate_at(a + n)] == in[deflate_at(b + n)]) ++n;
  return n;
}

// extra bits of a length symbol (257..285) / a distance code (0..29), RFC 1951 3.2.5
TB200_HD uint32_t deflate_len_extra_bits(uint32_t sym) { return (sym < 265u || sym == 285u) ? 0u : (sym - 261u) >> 2; }
TB200_HD uint32_t deflate_dist_extra_bits(uint32_t dc) { return dc < 4u ? 0u : (dc - 2u) >> 1; }
// bits of a match token under a code given as length tables
TB200_HD uint32_t deflate_match_cost(const uint8_t* lit_len, const uint8_t* dist_len, uint32_t len, uint32_t dist) {
  uint32_t sym, eb, ev, dc, deb, dev;
  deflate_len_symbol(len, &sym, &eb, &ev);
  deflate_dist_symbol(dist, &dc, &deb, &dev);
  return lit_len[sym] + eb + dist_len[dc] + deb;
}
TB200_HD uint32_t deflate_fixed_match_bits(uint32_t len, uint32_t dist) {
  uint32_t sym, eb, ev, dc, deb, dev;
  deflate_len_symbol(len, &sym, &eb, &ev);
  deflate_dist_symbol(dist, &dc, &deb, &dev);
  return deflate_fixed_litlen_bits(sym) + eb + 5u + deb;
}

// Greedy parse of the thread's sub-block.  table: first chunk position of each 4-byte hash.
// Leaves the decisions in t.is_start / t.is_match and, for a match that starts at chunk position
// p, tok[p] = distance, tok[p + 1] = length (a match covers >= 3 positions of its own sub-block);
// counts the symbols into the histograms (shared atomics on the device).
template <typename AddFn>
TB200_HD void deflate_parse(const uint8_t* in, DeflateThread& t, const uint32_t* table, uint16_t* tok, uint32_t chunk_bytes, AddFn add) {
  uint64_t starts = 0, matches = 0;
  const uint32_t n = t.end - t.begin;
  uint32_t i = 0;
  // typed tensors repeat the same (length, distance) pair element after element: count such repeats
  // locally and update the shared histograms once per run, not once per token
  uint32_t run_sym = 0xFFFFFFFFu, run_dsym = 0, run_n = 0;
  while (i < n) {
    const uint32_t room = n - i;
    uint32_t best = 0, bd = 0;
    if (room >= 3u) {
      // the periodic distances: longest run of equal bytes at distance d starting here; the
      // shortest distance wins ties (cheapest distance code).  Most positions of incompressible data
      // match at none of them: one test of the united masks skips the four run counts.
      if (((t.eq[0] | t.eq[1] | t.eq[2] | t.eq[3]) >> i) & 1ull) {
        const uint32_t r1 = deflate_run(t.eq[0], i), r2 = deflate_run(t.eq[1], i), r4 = deflate_run(t.eq[2], i), r8 = deflate_run(t.eq[3], i);
        best = r1;
        bd = 1;
        if (r2 > best) { best = r2; bd = 2; }
        if (r4 > best) { best = r4; bd = 4; }
        if (r8 > best) { best = r8; bd = 8; }
        if (best > room) best = room;
      }
      const uint32_t p = t.begin + i;
      const uint32_t v4 = best < 8u && p + 3u < chunk_bytes ? deflate_load4(in, p) : 0u;
      if (best < 8u && p + 3u < chunk_bytes && deflate_hashable(v4)) {  // a farther candidate is only worth a look when the near ones are short
        const uint32_t c = table[deflate_hash4(v4)];
        // most hits are hash collisions: one word compare rejects them before the byte loop
        if (c < p && deflate_load4(in, c) == v4) {
          const uint32_t cap = room < 258u ? room : 258u;
          const uint32_t l = cap <= 4u ? cap : 4u + deflate_match_len(in, c + 4u, p + 4u, cap - 4u);
          const uint32_t d = p - c;
          // A far distance costs up to 11 extra bits more than a near one: it has to beat the near
          // match here by two bytes, must not be a short match far away, and -- lazy evaluation --
          // must beat what a literal now and the near match one byte later would cover.
          uint32_t next = 0;
          if (i + 1u < n && (((t.eq[0] | t.eq[1] | t.eq[2] | t.eq[3]) >> (i + 1u)) & 1ull)) {
            const uint32_t q1 = deflate_run(t.eq[0], i + 1u), q2 = deflate_run(t.eq[1], i + 1u), q4 = deflate_run(t.eq[2], i + 1u), q8 = deflate_run(t.eq[3], i + 1u);
            next = q1 > q2 ? q1 : q2;
            next = q4 > next ? q4 : next;
            next = q8 > next ? q8 : next;
          }
          const uint32_t need = d > 4096u ? 6u : (d > 256u ? 5u : (d > 32u ? 4u : 3u));
          if (l > best + 1u && l >= need && l > next + 2u) { best = l; bd = d; }
        }
      }
      if (best > 258u) best = 258u;
    }
    starts |= 1ull << i;
    if (best >= 3u) {
      matches |= 1ull << i;
      const uint32_t p = t.begin + i;
      uint32_t sym, eb, ev, dc, deb, dev;
      deflate_len_symbol(best, &sym, &eb, &ev);
      deflate_dist_symbol(bd, &dc, &deb, &dev);
      // the three positions a match covers at least hold what the later walks need, so that the
      // symbols are computed once: [p] distance, [p + 1] length, [p + 2] length symbol - 257 | distance code << 5
      tok[p] = static_cast<uint16_t>(bd);
      tok[p + 1] = static_cast<uint16_t>(best);
      tok[p + 2] = static_cast<uint16_t>((sym - 257u) | (dc << 5));
      if (sym == run_sym && kDeflateLitSyms + dc == run_dsym) {
        ++run_n;
      } else {
        if (run_n != 0u) add(run_sym, run_dsym, run_n);
        run_sym = sym;
        run_dsym = kDeflateLitSyms + dc;
        run_n = 1;
      }
      i += best;
    } else {
      add(static_cast<uint32_t>(in[deflate_at(t.begin + i)]), 0xFFFFFFFFu, 1u);
      i += 1;
    }
  }
  if (run_n != 0u) add(run_sym, run_dsym, run_n);
  t.is_start = starts;
  t.is_match = matches;
}

// ---- code construction ---------------------------------------------------------------------------
// Shannon length of a symbol: smallest l with count << l >= total, clamped to [1, maxl]
TB200_HD uint32_t deflate_shannon_len(uint32_t count, uint32_t total, uint32_t maxl) {
  const uint32_t q = (total + count - 1u) / count;  // ceil(total / count) >= 1
  uint32_t l = 0;
#ifdef __CUDA_ARCH__
  l = q <= 1u ? 0u : 32u - static_cast<uint32_t>(__clz(q - 1u));
#else
  while ((1u << l) < q) ++l;
#endif
  return l < 1u ? 1u : (l > maxl ? maxl : l);
}
// Make the length histogram bl[1..MAXL] describe a COMPLETE prefix code (Kraft sum exactly 1):
// demote from the longest lengths while over-subscribed (only after clamping), then promote
// the largest steps that still fit.  Needs >= 2 symbols in use.  All loops have constant bounds
// and a local copy of the histogram, so on the device the whole thing runs in registers.
template <int MAXL>
TB200_HD void deflate_complete_code_t(uint32_t* bl_io) {
  uint32_t bl[MAXL + 2];
#pragma unroll
  for (int k = 0; k <= MAXL; ++k) bl[k] = bl_io[k];
  bl[MAXL + 1] = 0;
  uint32_t S = 0;
  const uint32_t T = 1u << MAXL;
#pragma unroll
  for (int k = 1; k <= MAXL; ++k) S += bl[k] << (MAXL - k);
#pragma unroll
  for (int k = MAXL - 1; k >= 1; --k) {
    while (S > T && bl[k] > 0u) {
      bl[k] -= 1u;
      bl[k + 1] += 1u;
      S -= 1u << (MAXL - k - 1);
    }
  }
  uint32_t gap = T - S;
  for (int pass = 0; pass < 16 && gap != 0u; ++pass) {
#pragma unroll
    for (int k = 2; k <= MAXL; ++k) {
      uint32_t take = gap >> (MAXL - k);
      if (take > bl[k]) take = bl[k];
      bl[k] -= take;
      bl[k - 1] += take;
      gap -= take << (MAXL - k);
    }
  }
#pragma unroll
  for (int k = 0; k <= MAXL; ++k) bl_io[k] = bl[k];
}
TB200_HD void deflate_complete_code(uint32_t* bl, uint32_t maxl) {
  if (maxl == 15u) deflate_complete_code_t<15>(bl);
  else deflate_complete_code_t<7>(bl);
}
// canonical first codes per length (RFC 1951 3.2.2)
TB200_HD void deflate_next_codes(const uint32_t* bl, uint32_t maxl, uint32_t* next) {
  uint32_t code = 0;
  next[0] = 0;
  for (uint32_t k = 1; k <= maxl; ++k) {
    code = (code + (k == 1u ? 0u : bl[k - 1u])) << 1;
    next[k] = code;
  }
}

// order in which the code-length code's own lengths are sent (RFC 1951 3.2.7)
TB200_HD uint32_t deflate_cl_order(uint32_t i) {
  const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  return order[i];
}

// The code-length sequence of a dynamic header: the literal/length lengths [0, nlit) followed directly
// by the distance lengths [0, ndist).
TB200_HD uint32_t deflate_hdr_len(const uint8_t* len_tab, uint32_t nlit, uint32_t i) {
  return i < nlit ? len_tab[i] : len_tab[kDeflateLitSyms + (i - nlit)];
}
// Run-length form of that sequence (RFC 1951 3.2.7), block-wise so that it is parallel: the sequence is
// cut into segments of 16 entries; a segment of >= 11 entries that is all zero becomes ONE entry
// (symbol 18, "repeat zero 11..138 times"), every other segment sends its lengths as they are.  Segment s
// -> number of entries it contributes.
constexpr uint32_t kDeflateHdrSeg = 16;
TB200_HD uint32_t deflate_segment_entries(const uint8_t* len_tab, uint32_t nlit, uint32_t n, uint32_t s) {
  const uint32_t b = s * kDeflateHdrSeg, e = b + kDeflateHdrSeg < n ? b + kDeflateHdrSeg : n;
  if (b >= n) return 0u;
  bool zero = e - b >= 11u;
  for (uint32_t i = b; zero && i < e; ++i) zero = deflate_hdr_len(len_tab, nlit, i) == 0u;
  return zero ? 1u : e - b;
}
// write segment s's entries at hdr_sym / hdr_extra[at ...]; count(sym) is called once per entry
template <typename CountFn>
TB200_HD void deflate_segment_write(const uint8_t* len_tab, uint32_t nlit, uint32_t n, uint32_t s, uint32_t entries, uint32_t at, uint8_t* hdr_sym,
                                    uint8_t* hdr_extra, CountFn count) {
  const uint32_t b = s * kDeflateHdrSeg, e = b + kDeflateHdrSeg < n ? b + kDeflateHdrSeg : n;
  if (entries == 0u) return;
  if (entries == 1u && e - b >= 11u) {
    hdr_sym[at] = 18;
    hdr_extra[at] = static_cast<uint8_t>(e - b - 11u);
    count(18u);
    return;
  }
  for (uint32_t i = b; i < e; ++i) {
    const uint32_t v = deflate_hdr_len(len_tab, nlit, i);
    hdr_sym[at + (i - b)] = static_cast<uint8_t>(v);
    hdr_extra[at + (i - b)] = 0;
    count(v);
  }
}
TB200_HD uint32_t deflate_cl_extra_bits(uint32_t sym) { return sym == 16u ? 2u : (sym == 17u ? 3u : (sym == 18u ? 7u : 0u)); }

// ---- token walks ---------------------------------------------------------------------------------
// bits of the thread's sub-block under the dynamic code (length tables) and under the fixed code
TB200_HD void deflate_count(const uint8_t* in, const DeflateThread& t, const uint16_t* tok, const uint8_t* lit_len, const uint8_t* dist_len,
                            uint32_t* dyn_bits, uint32_t* fix_bits) {
  uint32_t dyn = 0, fix = 0;
  uint64_t s = t.is_start;
  while (s != 0ull) {
    const uint32_t i = deflate_ctz64(s);
    s &= s - 1ull;
    const uint32_t p = t.begin + i;
    if ((t.is_match >> i) & 1ull) {
      const uint32_t sd = tok[p + 2], sym = 257u + (sd & 31u), dc = sd >> 5;
      const uint32_t extra = deflate_len_extra_bits(sym) + deflate_dist_extra_bits(dc);
      dyn += lit_len[sym] + dist_len[dc] + extra;
      fix += deflate_fixed_litlen_bits(sym) + 5u + extra;
    } else {
      const uint32_t b = in[deflate_at(p)];
      dyn += lit_len[b];
      fix += deflate_fixed_litlen_bits(b);
    }
  }
  *dyn_bits = dyn;
  *fix_bits = fix;
}
// emit the sub-block's tokens from bit `bitpos`; dynamic: reversed codes + lengths from the tables,
// fixed (lit_code == nullptr): RFC 1951 3.2.6
TB200_HD void deflate_emit(const uint8_t* in, const DeflateThread& t, const uint16_t* tok, const uint16_t* lit_code, const uint8_t* lit_len,
                           const uint16_t* dist_code, const uint8_t* dist_len, uint32_t* words, uint32_t bitpos) {
  uint32_t pos = bitpos;
  uint64_t s = t.is_start;
  while (s != 0ull) {
    const uint32_t i = deflate_ctz64(s);
    s &= s - 1ull;
    const uint32_t p = t.begin + i;
    if ((t.is_match >> i) & 1ull) {
      const uint32_t dist = tok[p], len = tok[p + 1], sd = tok[p + 2];
      const uint32_t sym = 257u + (sd & 31u), dc = sd >> 5;
      const uint32_t eb = deflate_len_extra_bits(sym), deb = deflate_dist_extra_bits(dc);
      // extra values: offset of the length / distance inside its symbol's range
      const uint32_t ev = eb == 0u ? 0u : ((len - 3u) & ((1u << eb) - 1u));
      const uint32_t dev = deb == 0u ? 0u : ((dist - 1u) & ((1u << deb) - 1u));
      if (lit_code != nullptr) {
        // code, extra bits, distance code, extra bits: at most 15 + 5 + 15 + 13 = 48 bits, two puts
        deflate_put(words, pos, static_cast<uint32_t>(lit_code[sym]) | (ev << lit_len[sym]), lit_len[sym] + eb);
        pos += lit_len[sym] + eb;
        deflate_put(words, pos, static_cast<uint32_t>(dist_code[dc]) | (dev << dist_len[dc]), dist_len[dc] + deb);
        pos += dist_len[dc] + deb;
      } else {
        uint32_t code, nb;
        deflate_litlen_code(sym, &code, &nb);
        deflate_put(words, pos, deflate_reverse(code, nb) | (ev << nb), nb + eb);
        pos += nb + eb;
        deflate_put(words, pos, deflate_reverse(dc, 5) | (dev << 5), 5u + deb);
        pos += 5u + deb;
      }
    } else {
      const uint32_t b = in[deflate_at(p)];
      if (lit_code != nullptr) {
        deflate_put(words, pos, lit_code[b], lit_len[b]);
        pos += lit_len[b];
      } else {
        uint32_t code, nb;
        deflate_litlen_code(b, &code, &nb);
        deflate_put(words, pos, deflate_reverse(code, nb), nb);
        pos += nb;
      }
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
