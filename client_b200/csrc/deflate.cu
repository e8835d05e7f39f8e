// deflate.cu -- device deflate encoder (zlib / gzip containers) built on deflate.cuh.
//   deflate_chunk_kernel     one CTA of 128 threads per 8 KiB chunk: hash candidates, greedy
//                            parse per 64-byte sub-block, emit pass, fixed-Huffman bit packing in
//                            shared memory, stored fallback, per-chunk Adler-32 / CRC-32 pieces
//   deflate_finalize_kernel  sizes -> offsets, checksum combination across chunks, header/trailer
//   deflate_gather_kernel    chunk bytes -> their final places
#include "deflate.cuh"
#include "kernels.cuh"

namespace tb200 {

__global__ void __launch_bounds__(kDeflateThreads) deflate_chunk_kernel(const uint8_t* __restrict__ src, uint64_t nbytes,
                                                                         uint8_t* __restrict__ scratch, DeflateChunkMeta* __restrict__ meta) {
  __shared__ __align__(16) uint8_t in[kDeflateInBytes];  // skewed layout, deflate_at()
  __shared__ uint32_t crc_tbl[256];
  __shared__ uint16_t cand[kDeflateChunk];
  __shared__ uint32_t words[kDeflateOutWords];
  uint32_t* table = words;  // hash heads (position + 1, 0 = none); dead before the bit buffer is used
  static_assert((1 << kDeflateHashBits) <= kDeflateOutWords, "table aliases words");
  __shared__ uint32_t sub_bits[kDeflateThreads];
  __shared__ uint32_t sub_off[kDeflateThreads];
  __shared__ uint32_t crc_s[kDeflateThreads];
  __shared__ uint32_t len_s[kDeflateThreads];
  __shared__ uint32_t adl_a[kDeflateThreads], adl_b[kDeflateThreads];
  __shared__ uint32_t pw[8];  // x^(8 * 128 * 2^k)
  __shared__ uint32_t total_bits_s;

  const uint32_t tid = threadIdx.x;
  const uint64_t base = static_cast<uint64_t>(blockIdx.x) * kDeflateChunk;
  const uint32_t n = static_cast<uint32_t>(nbytes - base < static_cast<uint64_t>(kDeflateChunk) ? nbytes - base : kDeflateChunk);
  const uint8_t* g = src + base;

  for (uint32_t i = tid; i < n; i += kDeflateThreads) in[deflate_at(i)] = g[i];
  for (uint32_t i = tid; i < 256u; i += kDeflateThreads) {
    const uint8_t b = static_cast<uint8_t>(i);
    crc_tbl[i] = crc32_raw(0u, &b, 1);
  }
  for (uint32_t i = tid; i < (1u << kDeflateHashBits); i += kDeflateThreads) table[i] = 0u;
  if (tid < 8) pw[tid] = crc32_xpow8n(static_cast<uint64_t>(kDeflateSub) << tid);
  __syncthreads();

  // candidates: round r looks up what earlier rounds inserted, then inserts its own positions
  for (uint32_t r0 = 0; r0 < n; r0 += kDeflateThreads) {
    const uint32_t p = r0 + tid;
    uint32_t h = 0;
    const bool ok = p + 3 < n;
    if (ok) {
      h = deflate_hash(in, p);
      const uint32_t head = table[h];
      cand[p] = head != 0u ? static_cast<uint16_t>(head - 1u) : static_cast<uint16_t>(kDeflateNoCand);
    } else if (p < n) {
      cand[p] = static_cast<uint16_t>(kDeflateNoCand);
    }
    __syncthreads();
    if (ok) atomicMax(table + h, p + 1u);  // several positions of a round may share a hash: the latest stays
    __syncthreads();
  }

  for (uint32_t i = tid; i < static_cast<uint32_t>(kDeflateOutWords); i += kDeflateThreads) words[i] = 0;  // table -> bit buffer
  // checksum pieces of the raw bytes
  const uint32_t begin = tid * kDeflateSub;
  const uint32_t end = begin < n ? (begin + kDeflateSub < n ? begin + kDeflateSub : n) : begin;
  {
    uint32_t a = 0, b = 0;
    adler_piece(in + deflate_at(begin), end - begin, &a, &b);  // a sub-block is contiguous
    adl_a[tid] = a;
    adl_b[tid] = b;
    crc_s[tid] = crc32_raw_tbl(crc_tbl, 0u, in + deflate_at(begin), end - begin);
    len_s[tid] = end - begin;
  }

  // pass 1: bits per sub-block
  sub_bits[tid] = end > begin ? deflate_parse(in, begin, end, cand) : 0u;
  __syncthreads();
  if (tid == 0) {
    uint32_t off = 3;  // block header: BFINAL=0, BTYPE=01
    for (int i = 0; i < kDeflateThreads; ++i) {
      sub_off[i] = off;
      off += sub_bits[i];
    }
    total_bits_s = off;
    deflate_put(words, 0, 2u, 3);
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

  const uint32_t body_bits = total_bits_s;
  // EOB (7 zero bits) + empty stored block header (3 zero bits), padded to a byte, + 00 00 FF FF
  const uint32_t flush_at = (body_bits + 7 + 3 + 7) >> 3;  // byte index of LEN
  const uint32_t comp_bytes = flush_at + 4;
  const bool stored = comp_bytes >= n + 5;
  uint8_t* out = scratch + static_cast<size_t>(blockIdx.x) * kDeflateMaxChunkOut;
  if (!stored) {
    if (end > begin) deflate_emit(in, begin, end, cand, words, sub_off[tid]);
    __syncthreads();
    if (tid == 0) {
      deflate_put(words, (flush_at + 2) * 8, 0xFFFFu, 16);  // LEN = 0 is already there
    }
    __syncthreads();
    const uint8_t* wb = reinterpret_cast<const uint8_t*>(words);
    for (uint32_t i = tid; i < comp_bytes; i += kDeflateThreads) out[i] = wb[i];
  } else {
    if (tid == 0) {
      out[0] = 0x00;  // BFINAL=0, BTYPE=00
      out[1] = static_cast<uint8_t>(n & 0xFF);
      out[2] = static_cast<uint8_t>(n >> 8);
      out[3] = static_cast<uint8_t>(~n & 0xFF);
      out[4] = static_cast<uint8_t>((~n >> 8) & 0xFF);
    }
    for (uint32_t i = tid; i < n; i += kDeflateThreads) out[5 + i] = in[deflate_at(i)];
  }
  if (tid == 0) {
    uint32_t A = 0, B = 0;
    for (uint32_t i = 0; i < static_cast<uint32_t>(kDeflateThreads); ++i) {
      const uint32_t b0 = i * kDeflateSub;
      const uint32_t li = b0 < n ? (n - b0 < static_cast<uint32_t>(kDeflateSub) ? n - b0 : kDeflateSub) : 0u;
      adler_append(&A, &B, adl_a[i], adl_b[i], li);
    }
    DeflateChunkMeta m;
    m.out_bytes = stored ? n + 5 : comp_bytes;
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
