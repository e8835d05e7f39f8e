// deflate.cu -- device deflate encoder (zlib / gzip containers) built on deflate.cuh.
//   deflate_chunk_kernel     one CTA of 128 threads per 8 KiB chunk: equality masks + first-occurrence
//                            hash, greedy parse per 64-byte sub-block, a dynamic Huffman code per
//                            chunk (fixed / stored when cheaper), bit packing in shared memory,
//                            per-chunk Adler-32 / CRC-32 pieces
//   deflate_finalize_kernel  sizes -> offsets, checksum combination across chunks, header/trailer
//   deflate_gather_kernel    chunk bytes -> their final places
#include "deflate.cuh"
#include "kernels.cuh"

namespace tb200 {

namespace {

// rank of this lane's symbol among the symbols of the same key with a smaller index, for symbols laid
// out as s = tid + 128 * round (so a warp of a round holds 32 consecutive symbols = one "unit").
// unit_cnt: [12 units][16 keys] scratch.  Phase A (this function) fills the unit counts and returns the
// rank inside the unit; phase B (after a barrier) adds the counts of the earlier units of the alphabet.
__device__ __forceinline__ uint32_t rank_in_unit(uint32_t key, uint32_t unit, uint32_t (*unit_cnt)[16]) {
  const uint32_t mask = __match_any_sync(0xFFFFFFFFu, key);
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t before = __popc(mask & ((1u << lane) - 1u));
  if (before == 0u) unit_cnt[unit][key] = static_cast<uint32_t>(__popc(mask));  // the group's first lane
  return before;
}

}  // namespace

__global__ void __launch_bounds__(kDeflateThreads) deflate_chunk_kernel(const uint8_t* __restrict__ src, uint64_t nbytes,
                                                                         uint8_t* __restrict__ scratch, DeflateChunkMeta* __restrict__ meta) {
  __shared__ __align__(16) uint8_t in[kDeflateInBytes];  // skewed layout, deflate_at(); the gap word after a sub-block
                                                         // repeats the first word of the next one
  __shared__ uint32_t words[kDeflateOutWords];           // hash table during the parse, bit buffer afterwards
  __shared__ uint16_t tok[kDeflateChunk];
  __shared__ uint32_t hist[kDeflateLitSyms + kDeflateDistSyms];
  __shared__ uint8_t len_tab[kDeflateLitSyms + kDeflateDistSyms];
  __shared__ uint16_t code_tab[kDeflateLitSyms + kDeflateDistSyms];
  __shared__ uint32_t bl[2][16], bl0[2][16], next_code[2][16];
  __shared__ uint32_t unit_cnt[12][16];
  __shared__ uint8_t hdr_sym[kDeflateHdrMax + 8], hdr_extra[kDeflateHdrMax + 8];
  __shared__ uint32_t cl_len[19], cl_code[19];
  __shared__ uint32_t sub_dyn[kDeflateThreads], sub_fix[kDeflateThreads], sub_off[kDeflateThreads];
  __shared__ uint32_t crc_tbl[256];
  __shared__ uint32_t crc_s[kDeflateThreads], len_s[kDeflateThreads], adl_a[kDeflateThreads], adl_b[kDeflateThreads];
  __shared__ uint32_t pw[8];  // x^(8 * 64 * 2^k)
  __shared__ uint32_t ntok_s, nmatch_s, hdr_n_s, hdr_bits_s, ncl_s, nlit_s, ndist_s, mode_s, total_bits_s, scan_s[4];
  uint32_t* table = words;
  static_assert((1 << kDeflateHashBits) <= kDeflateOutWords, "table aliases words");

  const uint32_t tid = threadIdx.x;
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kDeflateChunk;
  const uint32_t n = static_cast<uint32_t>(nbytes - base < static_cast<uint64_t>(kDeflateChunk) ? nbytes - base : kDeflateChunk);
  const uint8_t* g = src + base;
  uint32_t* inw = reinterpret_cast<uint32_t*>(in);

  DeflateThread t;
  t.begin = tid * kDeflateSub;
  t.end = t.begin < n ? (t.begin + kDeflateSub < n ? t.begin + kDeflateSub : n) : t.begin;

  // ---- load: 64 bytes per thread into registers and into the skewed shared image
  if (t.begin + kDeflateSub <= n && (reinterpret_cast<uintptr_t>(g) & 15u) == 0u) {
    const uint4* g4 = reinterpret_cast<const uint4*>(g + t.begin);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 v = __ldg(g4 + q);
      t.w[2 + 4 * q] = v.x;
      t.w[3 + 4 * q] = v.y;
      t.w[4 + 4 * q] = v.z;
      t.w[5 + 4 * q] = v.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      uint32_t v = 0;
      for (uint32_t b = 0; b < 4u; ++b) {
        const uint32_t p = t.begin + 4u * k + b;
        if (p < n) v |= static_cast<uint32_t>(g[p]) << (8u * b);
      }
      t.w[2 + k] = v;
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) inw[17u * tid + k] = t.w[2 + k];
  if (tid > 0) inw[17u * (tid - 1u) + 16u] = t.w[2];  // look-ahead word of the previous sub-block
  if (tid == kDeflateThreads - 1) inw[17u * tid + 16u] = 0u;
  for (uint32_t i = tid; i < 256u; i += kDeflateThreads) {
    const uint8_t b = static_cast<uint8_t>(i);
    crc_tbl[i] = crc32_raw(0u, &b, 1);
  }
  for (uint32_t i = tid; i < (1u << kDeflateHashBits); i += kDeflateThreads) table[i] = kDeflateNoCand;
  for (uint32_t i = tid; i < static_cast<uint32_t>(kDeflateLitSyms + kDeflateDistSyms); i += kDeflateThreads) hist[i] = 0u;
  if (tid < 32) (&bl[0][0])[tid] = 0u;
  if (tid < 8) pw[tid] = crc32_xpow8n(static_cast<uint64_t>(kDeflateSub) << tid);
  if (tid == 0) {
    ntok_s = 0;
    nmatch_s = 0;
  }
  __syncthreads();
  t.w[0] = tid > 0 ? inw[17u * (tid - 1u) + 14u] : 0u;
  t.w[1] = tid > 0 ? inw[17u * (tid - 1u) + 15u] : 0u;
  t.w[18] = inw[17u * tid + 16u];

  // ---- checksum pieces of the raw bytes (a sub-block is contiguous in the image)
  {
    uint32_t a = 0, b = 0;
    adler_piece(in + deflate_at(t.begin), t.end - t.begin, &a, &b);
    adl_a[tid] = a;
    adl_b[tid] = b;
    crc_s[tid] = crc32_raw_tbl(crc_tbl, 0u, in + deflate_at(t.begin), t.end - t.begin);
    len_s[tid] = t.end - t.begin;
  }

  // ---- match finding: equality masks, first-occurrence hash table
  deflate_masks(t);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t p = t.begin + 4u * k + j;
      if (p + 3u < n) {
        const uint32_t v = j == 0 ? t.w[2 + k] : ((t.w[2 + k] >> (8 * j)) | (t.w[3 + k] << (32 - 8 * j)));
        atomicMin(table + deflate_hash4(v), p);
      }
    }
  }
  __syncthreads();

  // ---- greedy parse, histograms
  if (t.end > t.begin) {
    deflate_parse(in, t, table, tok, n, [&](uint32_t a, uint32_t b) {
      atomicAdd(hist + a, 1u);
      if (b != 0xFFFFFFFFu) atomicAdd(hist + b, 1u);
    });
    atomicAdd(&ntok_s, static_cast<uint32_t>(__popcll(t.is_start)));
    atomicAdd(&nmatch_s, static_cast<uint32_t>(__popcll(t.is_match)));
  } else {
    t.is_start = t.is_match = 0ull;
  }
  if (tid == 0) hist[256] = 1u;  // end of block
  __syncthreads();
  // the hash table is dead: the bit buffer takes its place
  for (uint32_t i = tid; i < static_cast<uint32_t>(kDeflateOutWords); i += kDeflateThreads) words[i] = 0;

  // ---- a code for this chunk.  Symbols s = tid + 128 * round; s < 288: literal / length alphabet,
  //      288 <= s < 320: distance alphabet (one warp-sized unit of its own).
  const uint32_t total_lit = ntok_s + 1u, total_dist = nmatch_s;
  uint32_t cls[3], unit_rank[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint32_t s = tid + 128u * r;
    uint32_t c = 0;
    if (s < static_cast<uint32_t>(kDeflateLitSyms + kDeflateDistSyms)) {
      const uint32_t cnt = hist[s];
      if (cnt != 0u) c = deflate_shannon_len(cnt, s < static_cast<uint32_t>(kDeflateLitSyms) ? total_lit : total_dist, 15u);
    }
    cls[r] = c;
    if (c != 0u) atomicAdd(&bl[s < static_cast<uint32_t>(kDeflateLitSyms) ? 0 : 1][c], 1u);
    unit_rank[r] = rank_in_unit(c, 4u * r + (tid >> 5), unit_cnt);
  }
  __syncthreads();
  if (tid < 32) (&bl0[0][0])[tid] = (&bl[0][0])[tid];  // the Shannon classes, before the fix
  __syncthreads();
  if (tid == 0) deflate_complete_code(bl[0], 15u);
  if (tid == 32) {
    uint32_t used = 0;
    for (int k = 1; k <= 15; ++k) used += bl[1][k];
    if (used >= 2u) deflate_complete_code(bl[1], 15u);
  }
  __syncthreads();
  // position of a symbol in the order (class, index) -> its final length; then the canonical code
  uint32_t flen[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint32_t s = tid + 128u * r;
    flen[r] = 0;
    if (cls[r] != 0u) {
      const int a = s < static_cast<uint32_t>(kDeflateLitSyms) ? 0 : 1;
      const uint32_t unit = 4u * r + (tid >> 5), first_unit = a == 0 ? 0u : 9u;
      uint32_t pos = unit_rank[r];
      for (uint32_t u = first_unit; u < unit; ++u) pos += unit_cnt[u][cls[r]];
      for (uint32_t k = 1; k < cls[r]; ++k) pos += bl0[a][k];
      uint32_t k = 1, cum = bl[a][1];
      while (pos >= cum && k < 15u) {
        ++k;
        cum += bl[a][k];
      }
      flen[r] = k;
    }
    if (s < static_cast<uint32_t>(kDeflateLitSyms + kDeflateDistSyms)) len_tab[s] = static_cast<uint8_t>(flen[r]);
  }
  __syncthreads();  // unit_cnt is reused below
  if (tid == 0) deflate_next_codes(bl[0], 15u, next_code[0]);
  if (tid == 32) deflate_next_codes(bl[1], 15u, next_code[1]);
#pragma unroll
  for (int r = 0; r < 3; ++r) unit_rank[r] = rank_in_unit(flen[r], 4u * r + (tid >> 5), unit_cnt);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint32_t s = tid + 128u * r;
    if (flen[r] != 0u) {
      const int a = s < static_cast<uint32_t>(kDeflateLitSyms) ? 0 : 1;
      const uint32_t unit = 4u * r + (tid >> 5), first_unit = a == 0 ? 0u : 9u;
      uint32_t idx = unit_rank[r];
      for (uint32_t u = first_unit; u < unit; ++u) idx += unit_cnt[u][flen[r]];
      code_tab[s] = static_cast<uint16_t>(deflate_reverse(next_code[a][flen[r]] + idx, flen[r]));
    } else if (s < static_cast<uint32_t>(kDeflateLitSyms + kDeflateDistSyms)) {
      code_tab[s] = 0;
    }
  }
  __syncthreads();

  // ---- the header of a dynamic block: thread 0 forms the run-length entries and the code-length code
  if (tid == 0) {
    uint32_t nlit = 286, ndist = 30;
    while (nlit > 257u && len_tab[nlit - 1u] == 0u) --nlit;
    while (ndist > 1u && len_tab[kDeflateLitSyms + ndist - 1u] == 0u) --ndist;
    // the two length sequences back to back (the distance lengths follow the literal ones directly)
    uint8_t* seq = reinterpret_cast<uint8_t*>(unit_cnt);  // 768 bytes of scratch, dead now
    for (uint32_t i = 0; i < nlit; ++i) seq[i] = len_tab[i];
    for (uint32_t i = 0; i < ndist; ++i) seq[nlit + i] = len_tab[kDeflateLitSyms + i];
    const uint32_t ne = deflate_rle_lengths(seq, nlit + ndist, hdr_sym, hdr_extra);
    uint32_t cnt[19], lbl[8], lnext[8], ord[19];
    for (int i = 0; i < 19; ++i) cnt[i] = 0;
    for (uint32_t e = 0; e < ne; ++e) cnt[hdr_sym[e]] += 1u;
    for (int k = 0; k < 8; ++k) lbl[k] = 0;
    uint32_t used = 0;
    for (int i = 0; i < 19; ++i) {
      cl_len[i] = cnt[i] != 0u ? deflate_shannon_len(cnt[i], ne, 7u) : 0u;
      if (cnt[i] != 0u) {
        lbl[cl_len[i]] += 1u;
        ++used;
      }
    }
    if (used == 1u) {  // a single symbol: give it one bit and a dummy partner (a complete code)
      for (int i = 0; i < 19; ++i) {
        if (cnt[i] == 0u) {
          cl_len[i] = 1;
          lbl[1] += 1u;
          break;
        }
      }
    }
    // order (class, index), lengths by position in the completed histogram
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
    for (uint32_t i = 0; i < 19u; ++i) {
      if (cl_len[i] != 0u) cl_code[i] = deflate_reverse(lnext[cl_len[i]]++, cl_len[i]);
      else cl_code[i] = 0;
    }
    uint32_t ncl = 19;
    while (ncl > 4u && cl_len[deflate_cl_order(ncl - 1u)] == 0u) --ncl;
    uint32_t bits = 3u + 5u + 5u + 4u + 3u * ncl;
    for (uint32_t e = 0; e < ne; ++e) bits += cl_len[hdr_sym[e]] + deflate_cl_extra_bits(hdr_sym[e]);
    hdr_n_s = ne;
    hdr_bits_s = bits;
    ncl_s = ncl;
    nlit_s = nlit;
    ndist_s = ndist;
  }

  // ---- bits per sub-block under both codes
  {
    uint32_t dyn = 0, fix = 0;
    if (t.end > t.begin) deflate_count(in, t, tok, len_tab, len_tab + kDeflateLitSyms, &dyn, &fix);
    sub_dyn[tid] = dyn;
    sub_fix[tid] = fix;
  }
  // CRC tree: raw0(X || Y) = raw0(X) * x^(8|Y|) + raw0(Y)
  for (uint32_t stride = 1, level = 0; stride < kDeflateThreads; stride <<= 1, ++level) {
    __syncthreads();
    if ((tid & (2 * stride - 1)) == 0) {
      const uint32_t ly = len_s[tid + stride];
      const uint32_t mult = ly == (static_cast<uint32_t>(kDeflateSub) << level) ? pw[level] : crc32_xpow8n(ly);
      crc_s[tid] = crc32_mulmod(crc_s[tid], mult) ^ crc_s[tid + stride];
      len_s[tid] += ly;
    }
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t dyn = hdr_bits_s, fix = 3u;
    for (int i = 0; i < kDeflateThreads; ++i) {
      dyn += sub_dyn[i];
      fix += sub_fix[i];
    }
    dyn += len_tab[256];
    fix += 7u;
    // EOB, then the empty stored block: 3 header bits, pad to a byte, 00 00 FF FF
    const uint32_t best = dyn < fix ? dyn : fix;
    const uint32_t comp_bytes = ((best + 3u + 7u) >> 3) + 4u;
    const uint32_t mode = comp_bytes >= n + 5u ? 0u : (dyn < fix ? 2u : 1u);
    uint32_t off = mode == 2u ? hdr_bits_s : 3u;
    for (int i = 0; i < kDeflateThreads; ++i) {
      sub_off[i] = off;
      off += mode == 2u ? sub_dyn[i] : sub_fix[i];
    }
    total_bits_s = off;  // position of the end-of-block code
    mode_s = mode;
  }
  __syncthreads();

  const uint32_t mode = mode_s;
  uint8_t* out = scratch + static_cast<size_t>(blockIdx.x) * kDeflateMaxChunkOut;
  uint32_t out_bytes = 0;
  if (mode != 0u) {
    if (mode == 2u) {
      // header: fixed part by thread 0, the entries by everyone (three consecutive entries per thread)
      const uint32_t ne = hdr_n_s, ncl = ncl_s;
      if (tid == 0) {
        deflate_put(words, 0, 4u, 3);  // BFINAL = 0, BTYPE = 10
        deflate_put(words, 3, nlit_s - 257u, 5);
        deflate_put(words, 8, ndist_s - 1u, 5);
        deflate_put(words, 13, ncl - 4u, 4);
        for (uint32_t i = 0; i < ncl; ++i) deflate_put(words, 17u + 3u * i, cl_len[deflate_cl_order(i)], 3);
      }
      uint32_t mine = 0;
      for (uint32_t e = 3u * tid; e < 3u * tid + 3u && e < ne; ++e) mine += cl_len[hdr_sym[e]] + deflate_cl_extra_bits(hdr_sym[e]);
      // exclusive scan over the 128 threads
      uint32_t incl = mine;
      const uint32_t lane = tid & 31u;
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= static_cast<uint32_t>(d)) incl += v;
      }
      if (lane == 31u) scan_s[tid >> 5] = incl;
      __syncthreads();
      uint32_t pos = 17u + 3u * ncl + incl - mine;
      for (uint32_t wq = 0; wq < (tid >> 5); ++wq) pos += scan_s[wq];
      for (uint32_t e = 3u * tid; e < 3u * tid + 3u && e < ne; ++e) {
        const uint32_t s = hdr_sym[e], l = cl_len[s], xb = deflate_cl_extra_bits(s);
        deflate_put(words, pos, cl_code[s] | (static_cast<uint32_t>(hdr_extra[e]) << l), l + xb);
        pos += l + xb;
      }
      if (t.end > t.begin) {
        deflate_emit(in, t, tok, code_tab, len_tab, code_tab + kDeflateLitSyms, len_tab + kDeflateLitSyms, words, sub_off[tid]);
      }
      if (tid == 0) deflate_put(words, total_bits_s, code_tab[256], len_tab[256]);
    } else {
      if (tid == 0) deflate_put(words, 0, 2u, 3);  // BFINAL = 0, BTYPE = 01; the fixed EOB is 7 zero bits
      if (t.end > t.begin) deflate_emit(in, t, tok, nullptr, nullptr, nullptr, nullptr, words, sub_off[tid]);
    }
    const uint32_t eob_bits = mode == 2u ? len_tab[256] : 7u;
    const uint32_t flush_at = (total_bits_s + eob_bits + 3u + 7u) >> 3;  // byte index of LEN
    out_bytes = flush_at + 4u;
    __syncthreads();
    if (tid == 0) deflate_put(words, (flush_at + 2u) * 8u, 0xFFFFu, 16);  // LEN = 0 is already there
    __syncthreads();
    const uint8_t* wb = reinterpret_cast<const uint8_t*>(words);
    for (uint32_t i = tid; i < out_bytes; i += kDeflateThreads) out[i] = wb[i];
  } else {
    if (tid == 0) {
      out[0] = 0x00;  // BFINAL=0, BTYPE=00
      out[1] = static_cast<uint8_t>(n & 0xFF);
      out[2] = static_cast<uint8_t>(n >> 8);
      out[3] = static_cast<uint8_t>(~n & 0xFF);
      out[4] = static_cast<uint8_t>((~n >> 8) & 0xFF);
    }
    for (uint32_t i = tid; i < n; i += kDeflateThreads) out[5 + i] = in[deflate_at(i)];
    out_bytes = n + 5u;
  }
  if (tid == 0) {
    uint32_t A = 0, B = 0;
    for (uint32_t i = 0; i < static_cast<uint32_t>(kDeflateThreads); ++i) {
      const uint32_t b0 = i * kDeflateSub;
      const uint32_t li = b0 < n ? (n - b0 < static_cast<uint32_t>(kDeflateSub) ? n - b0 : kDeflateSub) : 0u;
      adler_append(&A, &B, adl_a[i], adl_b[i], li);
    }
    DeflateChunkMeta m;
    m.out_bytes = out_bytes;
    m.in_bytes = n;
    m.adler_a = A;
    m.adler_b = B;
    m.crc_raw0 = crc_s[0];
    meta[blockIdx.x] = m;
  }
}

// One CTA: offsets of the chunks in the final stream, checksums of the whole input, container
// header and trailer.  Threads own contiguous ranges of chunks.  Both checksums are linear in
// the pieces, so each partial is moved to the end of the stream on its own
// (crc * x^(8*bytes_after), b + bytes_after * a) and the results are summed -- no serial chain.
__global__ void __launch_bounds__(1024) deflate_finalize_kernel(DeflateChunkMeta* __restrict__ meta, uint32_t nchunks, uint64_t nbytes,
                                                                uint32_t format, uint8_t* __restrict__ dst, uint64_t* __restrict__ out_size) {
  __shared__ unsigned long long sz[1024];   // inclusive scan of output sizes
  __shared__ unsigned long long ln[1024];   // inclusive scan of input lengths
  __shared__ uint32_t red[32 * 3];
  const uint32_t tid = threadIdx.x;
  const uint32_t per = (nchunks + 1023u) / 1024u;
  const uint32_t c0 = min(tid * per, nchunks), c1 = min(c0 + per, nchunks);
  unsigned long long bytes = 0, len = 0;
  uint32_t c = 0, A = 0, B = 0;
  const uint32_t pw_chunk = crc32_xpow8n(kDeflateChunk);
  for (uint32_t i = c0; i < c1; ++i) {
    const DeflateChunkMeta m = meta[i];
    bytes += m.out_bytes;
    c = crc32_mulmod(c, m.in_bytes == static_cast<uint32_t>(kDeflateChunk) ? pw_chunk : crc32_xpow8n(m.in_bytes)) ^ m.crc_raw0;
    adler_append(&A, &B, m.adler_a, m.adler_b, m.in_bytes);
    len += m.in_bytes;
  }
  sz[tid] = bytes;
  ln[tid] = len;
  __syncthreads();
  for (uint32_t d = 1; d < 1024u; d <<= 1) {  // Hillis-Steele inclusive scans
    const unsigned long long s0 = tid >= d ? sz[tid - d] : 0ull;
    const unsigned long long l0 = tid >= d ? ln[tid - d] : 0ull;
    __syncthreads();
    sz[tid] += s0;
    ln[tid] += l0;
    __syncthreads();
  }
  {
    unsigned long long off = sz[tid] - bytes;  // exclusive
    for (uint32_t i = c0; i < c1; ++i) {
      const uint32_t ob = meta[i].out_bytes;
      meta[i].offset = off;
      off += ob;
    }
  }
  // move this thread's partial to the end of the stream
  const unsigned long long after = nbytes - ln[tid];
  uint32_t crc_part = len != 0 ? crc32_mulmod(c, crc32_xpow8n(after)) : 0u;
  uint32_t a_part = A;
  uint32_t b_part = static_cast<uint32_t>((B + (after % kAdlerMod) * A) % kAdlerMod);
  // block reduction: xor for the CRC, modular sums for Adler
  for (int off = 16; off > 0; off >>= 1) {
    crc_part ^= __shfl_xor_sync(0xFFFFFFFFu, crc_part, off);
    a_part += __shfl_xor_sync(0xFFFFFFFFu, a_part, off);
    b_part += __shfl_xor_sync(0xFFFFFFFFu, b_part, off);  // 32 * 65520 < 2^32
  }
  if ((tid & 31u) == 0) {
    red[(tid >> 5) * 3 + 0] = crc_part;
    red[(tid >> 5) * 3 + 1] = a_part % kAdlerMod;
    red[(tid >> 5) * 3 + 2] = b_part % kAdlerMod;
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t crc_all = 0, a_all = 0, b_all = 0;
    for (int w = 0; w < 32; ++w) {
      crc_all ^= red[w * 3];
      a_all = (a_all + red[w * 3 + 1]) % kAdlerMod;
      b_all = (b_all + red[w * 3 + 2]) % kAdlerMod;
    }
    const unsigned long long body = sz[1023];
    const uint32_t hdr = format == TB200_DEFLATE_GZIP ? 10u : 2u;
    if (format == TB200_DEFLATE_GZIP) {
      const uint8_t h[10] = {0x1F, 0x8B, 0x08, 0x00, 0, 0, 0, 0, 0x00, 0xFF};
      for (int i = 0; i < 10; ++i) dst[i] = h[i];
    } else {
      dst[0] = 0x78;
      dst[1] = 0x01;
    }
    uint8_t* t = dst + hdr + body;
    t[0] = 0x01;  // final empty stored block
    t[1] = 0x00;
    t[2] = 0x00;
    t[3] = 0xFF;
    t[4] = 0xFF;
    t += 5;
    unsigned long long total = hdr + body + 5;
    if (format == TB200_DEFLATE_GZIP) {
      // register from init 0xFFFFFFFF over the whole stream, final xor
      const uint32_t raw = crc32_mulmod(0xFFFFFFFFu, crc32_xpow8n(nbytes)) ^ crc_all;
      const uint32_t v = raw ^ 0xFFFFFFFFu;
      const uint32_t isize = static_cast<uint32_t>(nbytes);
      for (int i = 0; i < 4; ++i) t[i] = static_cast<uint8_t>(v >> (8 * i));
      for (int i = 0; i < 4; ++i) t[4 + i] = static_cast<uint8_t>(isize >> (8 * i));
      total += 8;
    } else {
      // Adler-32 starts from a = 1: A = 1 + sum, B = n*1 + b
      uint32_t A1 = 1, B1 = 0;
      adler_append(&A1, &B1, a_all, b_all, nbytes);
      const uint32_t v = (B1 << 16) | A1;
      for (int i = 0; i < 4; ++i) t[i] = static_cast<uint8_t>(v >> (8 * (3 - i)));  // big endian
      total += 4;
    }
    *out_size = total;
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(256) deflate_gather_kernel(const uint8_t* __restrict__ scratch, const DeflateChunkMeta* __restrict__ meta,
                                                             uint8_t* __restrict__ dst, uint32_t hdr) {
  const DeflateChunkMeta m = meta[blockIdx.x];
  const uint8_t* s = scratch + static_cast<size_t>(blockIdx.x) * kDeflateMaxChunkOut;
  uint8_t* d = dst + hdr + m.offset;
  for (uint32_t i = threadIdx.x; i < m.out_bytes; i += 256) d[i] = s[i];
}

cudaError_t launch_deflate(const uint8_t* src, uint64_t nbytes, uint32_t format, uint8_t* scratch, DeflateChunkMeta* meta, uint8_t* dst,
                           uint64_t* out_size, cudaStream_t s) {
  const uint64_t nchunks64 = (nbytes + kDeflateChunk - 1) / kDeflateChunk;
  const uint32_t nchunks = static_cast<uint32_t>(nchunks64);
  if (nchunks > 0) deflate_chunk_kernel<<<nchunks, kDeflateThreads, 0, s>>>(src, nbytes, scratch, meta);
  deflate_finalize_kernel<<<1, 1024, 0, s>>>(meta, nchunks, nbytes, format, dst, out_size);
  if (nchunks > 0) deflate_gather_kernel<<<nchunks, 256, 0, s>>>(scratch, meta, dst, format == TB200_DEFLATE_GZIP ? 10u : 2u);
  return cudaGetLastError();
}

}  // namespace tb200
