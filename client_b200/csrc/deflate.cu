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

// CRC tables, computed by the host once per device (crc32_raw / crc32_xpow8n are host+device code):
// the byte-at-a-time table and x^(8 * 64 * k) for k = 0..128 -- what a sub-block's CRC has to be
// multiplied with to stand where it belongs in a full chunk (k sub-blocks follow it).
struct DeflateTables {
  uint32_t crc4[4][256];  // slicing-by-4: crc4[k][i] = CRC register after byte i and k zero bytes
  uint32_t pw_after[kDeflateThreads + 1];
  uint32_t pw_chunks[32];  // x^(8 * 8192 * 2^k): what moves a chunk's CRC past 2^k full chunks
};
__device__ DeflateTables g_deflate_tables;

cudaError_t ensure_deflate_tables() {
  static bool ready[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 0 && dev < 64 && ready[dev]) return cudaSuccess;
  DeflateTables t;
  for (uint32_t i = 0; i < 256u; ++i) {
    const uint8_t b = static_cast<uint8_t>(i);
    t.crc4[0][i] = crc32_raw(0u, &b, 1);
  }
  for (int k = 1; k < 4; ++k) {
    for (uint32_t i = 0; i < 256u; ++i) t.crc4[k][i] = (t.crc4[k - 1][i] >> 8) ^ t.crc4[0][t.crc4[k - 1][i] & 0xFFu];
  }
  for (uint32_t k = 0; k <= static_cast<uint32_t>(kDeflateThreads); ++k) t.pw_after[k] = crc32_xpow8n(static_cast<uint64_t>(kDeflateSub) * k);
  for (uint32_t k = 0; k < 32u; ++k) t.pw_chunks[k] = crc32_xpow8n(static_cast<uint64_t>(kDeflateChunk) << k);
  e = cudaMemcpyToSymbol(g_deflate_tables, &t, sizeof(t));
  if (e == cudaSuccess && dev >= 0 && dev < 64) ready[dev] = true;
  return e;
}

// sum over the block (128 threads = 4 warps); every thread gets the result.  red: 4 words of scratch.
__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t* red) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  if ((threadIdx.x & 31u) == 0u) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint32_t total = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return total;
}
__device__ __forceinline__ uint32_t block_xor(uint32_t v, uint32_t* red) {
  for (int d = 16; d > 0; d >>= 1) v ^= __shfl_xor_sync(0xFFFFFFFFu, v, d);
  if ((threadIdx.x & 31u) == 0u) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint32_t total = red[0] ^ red[1] ^ red[2] ^ red[3];
  __syncthreads();
  return total;
}
// exclusive prefix sum over the block; *total = the block's sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* red, uint32_t* total) {
  uint32_t incl = v;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= static_cast<uint32_t>(d)) incl += u;
  }
  if (lane == 31u) red[warp] = incl;
  __syncthreads();
  uint32_t before = 0;
  for (uint32_t w = 0; w < warp; ++w) before += red[w];
  *total = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return before + incl - v;
}

__device__ __forceinline__ uint64_t block_sum64(uint64_t v, unsigned long long* red) {
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
  if ((threadIdx.x & 31u) == 0u) red[threadIdx.x >> 5] = v;
  __syncthreads();
  const uint64_t total = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return total;
}
__device__ __forceinline__ uint64_t block_exclusive_scan64(uint64_t v, unsigned long long* red, uint64_t* total) {
  uint64_t incl = v;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= static_cast<uint32_t>(d)) incl += u;
  }
  if (lane == 31u) red[warp] = incl;
  __syncthreads();
  uint64_t before = 0;
  for (uint32_t w = 0; w < warp; ++w) before += red[w];
  *total = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return before + incl - v;
}

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

// xr = x^(8 * bytes of the last chunk): with pw_chunks it moves a chunk's CRC to the end of the stream
__global__ void __launch_bounds__(kDeflateThreads, 5) deflate_chunk_kernel(const uint8_t* __restrict__ src, uint64_t nbytes,
                                                                         uint8_t* __restrict__ scratch, DeflateChunkMeta* __restrict__ meta,
                                                                         uint32_t xr, uint32_t format) {
  __shared__ __align__(16) uint8_t in[kDeflateInBytes];  // skewed layout, deflate_at(); the gap word after a sub-block
                                                         // repeats the first word of the next one
  __shared__ uint32_t words[kDeflateOutWords];           // hash table during the parse, bit buffer afterwards
  __shared__ uint16_t tok[kDeflateChunk];
  __shared__ uint32_t hist[kDeflateLitSyms + kDeflateDistSyms];
  __shared__ uint8_t len_tab[kDeflateLitSyms + kDeflateDistSyms];
  __shared__ uint16_t code_tab[kDeflateLitSyms + kDeflateDistSyms];
  __shared__ uint32_t bl[2][16], bl0[2][16], next_code[2][16];
  __shared__ uint32_t unit_cnt[12][16], unit_cnt2[12][16];  // symbols per (unit, class) and per (unit, final length); zeroed: a unit
                                                            // only writes the keys it holds
  __shared__ uint8_t hdr_sym[kDeflateHdrMax + 8], hdr_extra[kDeflateHdrMax + 8];
  __shared__ uint32_t cl_len[19], cl_code[19];
  __shared__ unsigned long long red64_s[4];
  __shared__ uint32_t ntok_s, nmatch_s, hdr_n_s, hdr_fixed_bits_s, ncl_s, nlit_s, ndist_s, red_s[4], cl_cnt[19], mult_s;
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
  for (uint32_t i = tid; i < (1u << kDeflateHashBits); i += kDeflateThreads) table[i] = kDeflateNoCand;
  for (uint32_t i = tid; i < static_cast<uint32_t>(kDeflateLitSyms + kDeflateDistSyms); i += kDeflateThreads) hist[i] = 0u;
  if (tid < 32) (&bl[0][0])[tid] = 0u;
  for (uint32_t i = tid; i < 12u * 16u; i += kDeflateThreads) (&unit_cnt[0][0])[i] = (&unit_cnt2[0][0])[i] = 0u;
  if (tid == 0) {
    ntok_s = 0;
    nmatch_s = 0;
    nlit_s = 257u;
    ndist_s = 1u;
  }
  if (tid < 19) cl_cnt[tid] = 0u;
  __syncthreads();
  t.w[0] = tid > 0 ? inw[17u * (tid - 1u) + 14u] : 0u;
  t.w[1] = tid > 0 ? inw[17u * (tid - 1u) + 15u] : 0u;
  t.w[18] = inw[17u * tid + 16u];

  // ---- checksum of the raw bytes: CRC-32 for gzip, Adler-32 for zlib, straight from the register words
  //      (slicing-by-4 tables through the read-only cache; no byte loop through shared memory).  Both are
  //      linear in the pieces: every thread moves its piece to the end of the chunk (crc * x^(8 * bytes
  //      after it); b + bytes_after * a) and the block adds them up.
  uint32_t chunk_crc = 0, chunk_a = 0, chunk_b = 0;
  {
    const uint32_t len = t.end - t.begin;
    const uint32_t after = n - t.end;  // bytes of the chunk behind this sub-block
    if (format == TB200_DEFLATE_GZIP) {
      uint32_t c = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (4u * k + 4u <= len) {
          const uint32_t x = c ^ t.w[2 + k];
          c = __ldg(&g_deflate_tables.crc4[3][x & 0xFFu]) ^ __ldg(&g_deflate_tables.crc4[2][(x >> 8) & 0xFFu]) ^
              __ldg(&g_deflate_tables.crc4[1][(x >> 16) & 0xFFu]) ^ __ldg(&g_deflate_tables.crc4[0][x >> 24]);
        } else {
          for (uint32_t j = 0; j < 4u; ++j) {
            if (4u * k + j < len) c = __ldg(&g_deflate_tables.crc4[0][(c ^ (t.w[2 + k] >> (8u * j))) & 0xFFu]) ^ (c >> 8);
          }
        }
      }
      if (len != 0u && after != 0u) {
        const uint32_t mult = (after & (kDeflateSub - 1u)) == 0u ? g_deflate_tables.pw_after[after >> kDeflateSubShift] : crc32_xpow8n(after);
        c = crc32_mulmod(c, mult);
      }
      chunk_crc = block_xor(c, red_s);
    } else {
      uint32_t a = 0, b = 0;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (4u * k + j < len) {
            a += (t.w[2 + k] >> (8 * j)) & 0xFFu;  // <= 64 * 255: no overflow before the reduction
            b += a;
          }
        }
      }
      b = (b % kAdlerMod + (after % kAdlerMod) * a) % kAdlerMod;  // after < 8192, a < 2^14: fits 32 bits
      const uint64_t both = block_sum64((static_cast<uint64_t>(b) << 32) | a, red64_s);  // 128 * 65520 < 2^32 in each half
      chunk_a = static_cast<uint32_t>(both) % kAdlerMod;
      chunk_b = static_cast<uint32_t>(both >> 32) % kAdlerMod;
    }
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
        if (deflate_hashable(v)) atomicMin(table + deflate_hash4(v), p);
      }
    }
  }
  __syncthreads();

  // ---- greedy parse, histograms
  if (t.end > t.begin) {
    deflate_parse(in, t, table, tok, n, [&](uint32_t a, uint32_t b, uint32_t times) {
      atomicAdd(hist + a, times);
      if (b != 0xFFFFFFFFu) atomicAdd(hist + b, times);
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
    if (flen[r] != 0u) {  // HLIT / HDIST: one past the last symbol in use
      if (s < static_cast<uint32_t>(kDeflateLitSyms)) atomicMax(&nlit_s, s + 1u);
      else atomicMax(&ndist_s, s - static_cast<uint32_t>(kDeflateLitSyms) + 1u);
    }
  }
  __syncthreads();
  if (tid == 0) deflate_next_codes(bl[0], 15u, next_code[0]);
  if (tid == 32) deflate_next_codes(bl[1], 15u, next_code[1]);
#pragma unroll
  for (int r = 0; r < 3; ++r) unit_rank[r] = rank_in_unit(flen[r], 4u * r + (tid >> 5), unit_cnt2);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const uint32_t s = tid + 128u * r;
    if (flen[r] != 0u) {
      const int a = s < static_cast<uint32_t>(kDeflateLitSyms) ? 0 : 1;
      const uint32_t unit = 4u * r + (tid >> 5), first_unit = a == 0 ? 0u : 9u;
      uint32_t idx = unit_rank[r];
      for (uint32_t u = first_unit; u < unit; ++u) idx += unit_cnt2[u][flen[r]];
      code_tab[s] = static_cast<uint16_t>(deflate_reverse(next_code[a][flen[r]] + idx, flen[r]));
    } else if (s < static_cast<uint32_t>(kDeflateLitSyms + kDeflateDistSyms)) {
      code_tab[s] = 0;
    }
  }
  __syncthreads();

  // ---- the header of a dynamic block, by warp 0: block-wise run-length entries (one lane per segment of
  //      16 code lengths), then the code-length code (one lane per symbol, ballots instead of loops)
  if (tid < 32) {
    const uint32_t lane = tid, lt = (1u << lane) - 1u;
    const uint32_t nlit = nlit_s, ndist = ndist_s, N = nlit + ndist;
    const uint32_t mine = deflate_segment_entries(len_tab, nlit, N, lane);
    uint32_t incl = mine;
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= static_cast<uint32_t>(d)) incl += v;
    }
    const uint32_t ne = __shfl_sync(0xFFFFFFFFu, incl, 31);
    deflate_segment_write(len_tab, nlit, N, lane, mine, incl - mine, hdr_sym, hdr_extra, [&](uint32_t v) { atomicAdd(cl_cnt + v, 1u); });
    __syncwarp();
    const uint32_t cnt = lane < 19u ? cl_cnt[lane] : 0u;
    const uint32_t used = __ballot_sync(0xFFFFFFFFu, cnt != 0u);
    uint32_t c = cnt != 0u ? deflate_shannon_len(cnt, ne, 7u) : 0u;
    if (__popc(used) == 1 && lane == static_cast<uint32_t>(__ffs(static_cast<int>(~used & 0x7FFFFu)) - 1)) c = 1u;  // a dummy partner
    uint32_t lbl[8], lbl0[8], lnext[8];
    lbl[0] = 0;
#pragma unroll
    for (int k = 1; k <= 7; ++k) lbl[k] = static_cast<uint32_t>(__popc(__ballot_sync(0xFFFFFFFFu, c == static_cast<uint32_t>(k))));
#pragma unroll
    for (int k = 0; k <= 7; ++k) lbl0[k] = lbl[k];
    deflate_complete_code_t<7>(lbl);
    uint32_t pos = static_cast<uint32_t>(__popc(__match_any_sync(0xFFFFFFFFu, c) & lt));
#pragma unroll
    for (int k = 1; k <= 7; ++k) pos += static_cast<uint32_t>(k) < c ? lbl0[k] : 0u;
    uint32_t fl = 0, cum = 0;
#pragma unroll
    for (int k = 1; k <= 7; ++k) {
      if (fl == 0u && pos < cum + lbl[k]) fl = static_cast<uint32_t>(k);
      cum += lbl[k];
    }
    if (c == 0u) fl = 0u;
    deflate_next_codes(lbl, 7u, lnext);
    const uint32_t idx = static_cast<uint32_t>(__popc(__match_any_sync(0xFFFFFFFFu, fl) & lt));
    uint32_t first = 0;
#pragma unroll
    for (int k = 1; k <= 7; ++k) first = fl == static_cast<uint32_t>(k) ? lnext[k] : first;
    if (lane < 19u) {
      cl_len[lane] = fl;
      cl_code[lane] = fl != 0u ? deflate_reverse(first + idx, fl) : 0u;
    }
    __syncwarp();
    const uint32_t sent = __ballot_sync(0xFFFFFFFFu, lane < 19u && cl_len[deflate_cl_order(lane < 19u ? lane : 0u)] != 0u);
    uint32_t ncl = 32u - static_cast<uint32_t>(__clz(sent | 0xFu));  // at least 4
    if (lane == 0) {
      hdr_n_s = ne;
      hdr_fixed_bits_s = 17u + 3u * ncl;
      ncl_s = ncl;
    }
  } else if (tid < 64) {
    // warp 1: x^(8 * bytes behind this chunk in the stream) = xr * prod over the bits of (full chunks behind it)
    const uint32_t lane = tid - 32u;
    const uint64_t nchunks = (nbytes + kDeflateChunk - 1) / kDeflateChunk;
    uint32_t f = 1u << 31;  // x^0
    if (static_cast<uint64_t>(blockIdx.x) + 1u < nchunks) {
      const uint64_t q = nchunks - 2u - blockIdx.x;
      if ((q >> lane) & 1ull) f = g_deflate_tables.pw_chunks[lane];
      if (lane == 0) f = crc32_mulmod(f, xr);
    }
    for (int d = 16; d > 0; d >>= 1) f = crc32_mulmod(f, __shfl_down_sync(0xFFFFFFFFu, f, d));
    if (lane == 0) mult_s = f;
  }

  // ---- bits per sub-block under both codes, bits of the header entries, mode, offsets
  uint32_t my_dyn = 0, my_fix = 0;
  if (t.end > t.begin) deflate_count(in, t, tok, len_tab, len_tab + kDeflateLitSyms, &my_dyn, &my_fix);
  __syncthreads();  // thread 0's header is complete
  const uint32_t ne = hdr_n_s, ncl = ncl_s;
  uint32_t my_hdr = 0;
  for (uint32_t e = 3u * tid; e < 3u * tid + 3u && e < ne; ++e) my_hdr += cl_len[hdr_sym[e]] + deflate_cl_extra_bits(hdr_sym[e]);
  // one scan for the three running sums: dynamic bits (< 2^20) | fixed bits << 20 | header bits << 40
  uint64_t totals;
  const uint64_t before = block_exclusive_scan64(static_cast<uint64_t>(my_dyn) | (static_cast<uint64_t>(my_fix) << 20) | (static_cast<uint64_t>(my_hdr) << 40),
                                                 red64_s, &totals);
  const uint32_t dyn_before = static_cast<uint32_t>(before) & 0xFFFFFu, fix_before = static_cast<uint32_t>(before >> 20) & 0xFFFFFu;
  const uint32_t hdr_before = static_cast<uint32_t>(before >> 40);
  const uint32_t dyn_body = static_cast<uint32_t>(totals) & 0xFFFFFu, fix_body = static_cast<uint32_t>(totals >> 20) & 0xFFFFFu;
  const uint32_t hdr_entries_bits = static_cast<uint32_t>(totals >> 40);
  const uint32_t hdr_bits = hdr_fixed_bits_s + hdr_entries_bits;
  const uint32_t dyn_total = hdr_bits + dyn_body + len_tab[256], fix_total = 3u + fix_body + 7u;
  // EOB, then the empty stored block: 3 header bits, pad to a byte, 00 00 FF FF
  const uint32_t best_bits = dyn_total < fix_total ? dyn_total : fix_total;
  const uint32_t mode = (((best_bits + 3u + 7u) >> 3) + 4u) >= n + 5u ? 0u : (dyn_total < fix_total ? 2u : 1u);
  const uint32_t my_off = mode == 2u ? hdr_bits + dyn_before : 3u + fix_before;
  const uint32_t eob_at = mode == 2u ? hdr_bits + dyn_body : 3u + fix_body;  // position of the end-of-block code

  uint8_t* out = scratch + static_cast<size_t>(blockIdx.x) * kDeflateMaxChunkOut;
  uint32_t out_bytes = 0;
  if (mode != 0u) {
    if (mode == 2u) {
      // header: fixed part by thread 0, the entries by everyone (three consecutive entries per thread)
      if (tid == 0) {
        deflate_put(words, 0, 4u, 3);  // BFINAL = 0, BTYPE = 10
        deflate_put(words, 3, nlit_s - 257u, 5);
        deflate_put(words, 8, ndist_s - 1u, 5);
        deflate_put(words, 13, ncl - 4u, 4);
        for (uint32_t i = 0; i < ncl; ++i) deflate_put(words, 17u + 3u * i, cl_len[deflate_cl_order(i)], 3);
      }
      uint32_t pos = hdr_fixed_bits_s + hdr_before;
      for (uint32_t e = 3u * tid; e < 3u * tid + 3u && e < ne; ++e) {
        const uint32_t sy = hdr_sym[e], l = cl_len[sy], xb = deflate_cl_extra_bits(sy);
        deflate_put(words, pos, cl_code[sy] | (static_cast<uint32_t>(hdr_extra[e]) << l), l + xb);
        pos += l + xb;
      }
      if (t.end > t.begin) {
        deflate_emit(in, t, tok, code_tab, len_tab, code_tab + kDeflateLitSyms, len_tab + kDeflateLitSyms, words, my_off);
      }
      if (tid == 0) deflate_put(words, eob_at, code_tab[256], len_tab[256]);
    } else {
      if (tid == 0) deflate_put(words, 0, 2u, 3);  // BFINAL = 0, BTYPE = 01; the fixed EOB is 7 zero bits
      if (t.end > t.begin) deflate_emit(in, t, tok, nullptr, nullptr, nullptr, nullptr, words, my_off);
    }
    const uint32_t eob_bits = mode == 2u ? len_tab[256] : 7u;
    const uint32_t flush_at = (eob_at + eob_bits + 3u + 7u) >> 3;  // byte index of LEN
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
    DeflateChunkMeta m;
    m.out_bytes = out_bytes;
    m.in_bytes = n;
    // both checksums as they count at the END OF THE STREAM: the finalize kernel only adds them up
    const uint64_t after_s = nbytes - (base + n);
    m.adler_a = chunk_a;
    m.adler_b = static_cast<uint32_t>((chunk_b + (after_s % kAdlerMod) * chunk_a) % kAdlerMod);
    m.crc_raw0 = after_s != 0ull ? crc32_mulmod(chunk_crc, mult_s) : chunk_crc;
    meta[blockIdx.x] = m;
  }
}

// One CTA: offsets of the chunks in the final stream (a scan of their sizes), the checksums of the whole
// input -- the chunk kernel already moved every chunk's piece to the end of the stream, so they only
// have to be added up (xor for the CRC, modular sums for Adler) --, container header and trailer.
// crc_init_term = 0xFFFFFFFF * x^(8 * nbytes): what the CRC register's initial value contributes (host).
__global__ void __launch_bounds__(1024) deflate_finalize_kernel(DeflateChunkMeta* __restrict__ meta, uint32_t nchunks, uint64_t nbytes,
                                                                uint32_t format, uint8_t* __restrict__ dst, uint64_t* __restrict__ out_size,
                                                                uint32_t crc_init_term) {
  __shared__ unsigned long long sz[1024];   // inclusive scan of output sizes
  __shared__ uint32_t red[32 * 3];
  const uint32_t tid = threadIdx.x;
  const uint32_t per = (nchunks + 1023u) / 1024u;
  const uint32_t c0 = min(tid * per, nchunks), c1 = min(c0 + per, nchunks);
  unsigned long long bytes = 0;
  uint32_t crc_part = 0, a_part = 0, b_part = 0;
  for (uint32_t i = c0; i < c1; ++i) {
    const DeflateChunkMeta m = meta[i];
    bytes += m.out_bytes;
    crc_part ^= m.crc_raw0;
    a_part = (a_part + m.adler_a) % kAdlerMod;
    b_part = (b_part + m.adler_b) % kAdlerMod;
  }
  sz[tid] = bytes;
  __syncthreads();
  for (uint32_t d = 1; d < 1024u; d <<= 1) {  // Hillis-Steele inclusive scan
    const unsigned long long s0 = tid >= d ? sz[tid - d] : 0ull;
    __syncthreads();
    sz[tid] += s0;
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
      const uint32_t v = (crc_init_term ^ crc_all) ^ 0xFFFFFFFFu;
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
  const cudaError_t te = ensure_deflate_tables();
  if (te != cudaSuccess) return te;
  // host side of the checksum algebra: x^(8 * bytes of the last chunk) and the CRC's initial-value term
  const uint64_t last = nbytes == 0 ? 0 : nbytes - static_cast<uint64_t>(nchunks - 1u) * kDeflateChunk;
  const uint32_t xr = crc32_xpow8n(last);
  const uint32_t crc_init_term = crc32_mulmod(0xFFFFFFFFu, crc32_xpow8n(nbytes));
  if (nchunks > 0) deflate_chunk_kernel<<<nchunks, kDeflateThreads, 0, s>>>(src, nbytes, scratch, meta, xr, format);
  deflate_finalize_kernel<<<1, 1024, 0, s>>>(meta, nchunks, nbytes, format, dst, out_size, crc_init_term);
  if (nchunks > 0) deflate_gather_kernel<<<nchunks, 256, 0, s>>>(scratch, meta, dst, format == TB200_DEFLATE_GZIP ? 10u : 2u);
  return cudaGetLastError();
}

}  // namespace tb200
