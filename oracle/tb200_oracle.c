/*
 * tb200_oracle.c -- CPU restatement (plain C) of the device-side hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call this file.  The
 * product (client_b200/, libtb200.so) never links or imports it.
 *
 * What it restates, and from where:
 *   - Philox4x32-10: Salmon et al., SC'11 (Random123).  perf_analyzer's input
 *     generator is NOT in the reference, so the fill contract is
 *     this project's own; parity with perf_analyzer values is
 *     UNPINNED.  The block function itself is pinned by the published Random123
 *     known-answer vectors and by cuRAND's host generator (tests/test_oracle.py).
 *   - image pack: the numpy arithmetic of
 *     src/python/examples/image_client.py:154-193 (preprocess): astype, INCEPTION
 *     (x / 127.5) - 1, VGG x - (123,117,104), HWC -> CHW transpose.
 *   - bf16: truncation as in src/python/library/tritonclient/utils/__init__.py:327-331.
 *   - checks: the value comparisons the examples do after get_contents_as_numpy
 *     (src/python/examples/simple_http_cudashm_client.py:164-195).
 *
 * Written independently of client_b200/csrc/philox.cuh on purpose: fp16 rounding
 * uses the compiler's _Float16, division is a real IEEE division.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum {
  DT_BOOL = 1, DT_U8 = 2, DT_U16 = 3, DT_U32 = 4, DT_U64 = 5, DT_I8 = 6, DT_I16 = 7,
  DT_I32 = 8, DT_I64 = 9, DT_F16 = 10, DT_F32 = 11, DT_F64 = 12, DT_BYTES = 13, DT_BF16 = 14
};

static uint16_t half_bits(float f) {
  _Float16 h = (_Float16)f; /* round to nearest even, like numpy astype(float16) */
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
static float half_to_float(uint16_t u) {
  _Float16 h;
  memcpy(&h, &u, 2);
  return (float)h;
}
static uint32_t float_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; ++round) {
    uint64_t prod0 = (uint64_t)0xD2511F53u * c0;
    uint64_t prod1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t t0 = (uint32_t)(prod1 >> 32) ^ c1 ^ k0;
    uint32_t t1 = (uint32_t)prod1;
    uint32_t t2 = (uint32_t)(prod0 >> 32) ^ c3 ^ k1;
    uint32_t t3 = (uint32_t)prod0;
    c0 = t0; c1 = t1; c2 = t2; c3 = t3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* one 16-byte group of a tensor from the 4 words of one Philox call */
static void group_bytes(uint32_t dtype, const uint32_t w[4], double lo, double span, int64_t ilo,
                        uint64_t irange, uint8_t out[16]) {
  const int unit = (span == 0.0);
  const float lo_f = (float)lo, span_f = (float)span;
  switch (dtype) {
    case DT_F32:
      for (int i = 0; i < 4; ++i) {
        float u = (float)(w[i] >> 9) * 0x1p-23f; /* 23 random bits */
        if (!unit) u = fmaf(u, span_f, lo_f);
        memcpy(out + 4 * i, &u, 4);
      }
      break;
    case DT_F16:
    case DT_BF16:
      for (int i = 0; i < 8; ++i) {
        uint32_t x16 = (w[i / 2] >> (16 * (i & 1))) & 0xFFFFu;
        float u = (dtype == DT_F16) ? (float)(x16 >> 6) * 0x1p-10f : (float)(x16 >> 9) * 0x1p-7f;
        if (!unit) u = fmaf(u, span_f, lo_f);
        uint16_t b = (dtype == DT_F16) ? half_bits(u) : (uint16_t)(float_bits(u) >> 16);
        memcpy(out + 2 * i, &b, 2);
      }
      break;
    case DT_F64:
      for (int i = 0; i < 2; ++i) {
        uint64_t x = ((uint64_t)w[2 * i + 1] << 32) | w[2 * i];
        double u = (double)(x >> 12) * 0x1p-52; /* 52 random bits */
        if (!unit) u = fma(u, span, lo);
        memcpy(out + 8 * i, &u, 8);
      }
      break;
    case DT_I64:
    case DT_U64:
      for (int i = 0; i < 2; ++i) {
        uint64_t x = ((uint64_t)w[2 * i + 1] << 32) | w[2 * i];
        if (irange != 0) x = (uint64_t)ilo + (uint64_t)(((unsigned __int128)x * irange) >> 64);
        memcpy(out + 8 * i, &x, 8);
      }
      break;
    case DT_I32:
    case DT_U32:
      for (int i = 0; i < 4; ++i) {
        uint32_t x = w[i];
        if (irange != 0) x = (uint32_t)ilo + (uint32_t)(((uint64_t)x * irange) >> 32);
        memcpy(out + 4 * i, &x, 4);
      }
      break;
    case DT_I16:
    case DT_U16:
      for (int i = 0; i < 8; ++i) {
        uint32_t x = (w[i / 2] >> (16 * (i & 1))) & 0xFFFFu;
        if (irange != 0) x = (uint32_t)ilo + ((x * (uint32_t)irange) >> 16);
        uint16_t b = (uint16_t)x;
        memcpy(out + 2 * i, &b, 2);
      }
      break;
    case DT_I8:
    case DT_U8:
      for (int i = 0; i < 16; ++i) {
        uint32_t x = (w[i / 4] >> (8 * (i & 3))) & 0xFFu;
        if (irange != 0) x = (uint32_t)ilo + ((x * (uint32_t)irange) >> 8);
        out[i] = (uint8_t)x;
      }
      break;
    default: /* BOOL */
      for (int i = 0; i < 16; ++i) out[i] = (uint8_t)((w[i / 4] >> (8 * (i & 3))) & 1u);
      break;
  }
}

/* mode: 0 random, 1 zero, 2 byte(ilo) -- include/tb200.h tb200_fill_mode */
void oracle_fill(uint8_t* dst, uint64_t nbytes, uint32_t dtype, uint32_t mode, uint64_t seed,
                 uint64_t stream, double lo, double span, int64_t ilo, uint64_t irange) {
  if (mode == 1) { memset(dst, 0, nbytes); return; }
  if (mode == 2) { memset(dst, (int)(ilo & 0xFF), nbytes); return; }
  const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (uint64_t g = 0; g * 16 < nbytes; ++g) {
    const uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t w[4];
    uint8_t grp[16];
    oracle_philox4x32_10(ctr, key, w);
    if (dtype == 13) { /* BYTES: <u32 length><alnum chars> per element, irange = length */
      const uint64_t per = irange + 4;
      for (int i = 0; i < 16; ++i) {
        const uint64_t o = (g * 16 + (uint64_t)i) % per;
        const uint32_t rb = (w[i >> 2] >> (8 * (i & 3))) & 0xFFu;
        const uint32_t k = (rb * 62u) >> 8;
        grp[i] = o < 4 ? (uint8_t)((irange >> (8 * o)) & 0xFF) : (uint8_t)(k < 10 ? 48 + k : (k < 36 ? 55 + k : 61 + k));
      }
    } else
    group_bytes(dtype, w, lo, span, ilo, irange, grp);
    uint64_t left = nbytes - g * 16;
    memcpy(dst + g * 16, grp, left < 16 ? left : 16);
  }
}

/* numpy float32 / float16 arithmetic of image_client.preprocess for one pixel value */
static float scaled_f32(uint8_t px, uint32_t scaling, int c, int ch) {
  float x = (float)px;
  if (scaling == 1) return (x / 127.5f) - 1.0f;
  if (scaling == 2) {
    static const float mean3[3] = {123.0f, 117.0f, 104.0f};
    return x - (c == 1 ? 128.0f : mean3[ch]);
  }
  return x;
}
static uint16_t scaled_f16(uint8_t px, uint32_t scaling, int c, int ch) {
  /* numpy evaluates float16 ops in float32 and rounds after each op */
  float x = half_to_float(half_bits((float)px));
  if (scaling == 1) {
    float q = half_to_float(half_bits(x / 127.5f));
    return half_bits(q - 1.0f);
  }
  if (scaling == 2) {
    static const float mean3[3] = {123.0f, 117.0f, 104.0f};
    return half_bits(x - (c == 1 ? 128.0f : mean3[ch]));
  }
  return half_bits(x);
}

/* layout: 0 NCHW, 1 NHWC; src is uint8 NHWC */
int oracle_pack_image(void* dst, uint32_t dst_dtype, uint32_t layout, const uint8_t* src, int n,
                      int h, int w, int c, uint32_t scaling) {
  const uint64_t hw = (uint64_t)h * w;
  for (int img = 0; img < n; ++img) {
    for (uint64_t p = 0; p < hw; ++p) {
      for (int ch = 0; ch < c; ++ch) {
        const uint8_t px = src[((uint64_t)img * hw + p) * c + ch];
        const uint64_t di = layout == 0 ? ((uint64_t)img * c + ch) * hw + p : ((uint64_t)img * hw + p) * c + ch;
        if (dst_dtype == DT_F32) {
          ((float*)dst)[di] = scaled_f32(px, scaling, c, ch);
        } else if (dst_dtype == DT_F16) {
          ((uint16_t*)dst)[di] = scaled_f16(px, scaling, c, ch);
        } else if (dst_dtype == DT_BF16) {
          ((uint16_t*)dst)[di] = (uint16_t)(float_bits(scaled_f32(px, scaling, c, ch)) >> 16);
        } else {
          return -1;
        }
      }
    }
  }
  return 0;
}

/* sum / xor of little-endian u32 words, trailing bytes zero-extended */
void oracle_checksum(const uint8_t* a, uint64_t nbytes, uint64_t* sum, uint32_t* xor32) {
  uint64_t s = 0;
  uint32_t x = 0;
  for (uint64_t off = 0; off < nbytes; off += 4) {
    uint32_t w = 0;
    uint64_t left = nbytes - off;
    memcpy(&w, a + off, left < 4 ? left : 4);
    s += w;
    x ^= w;
  }
  *sum = s;
  *xor32 = x;
}

uint64_t oracle_count_diff_bytes(const uint8_t* a, const uint8_t* b, uint64_t nbytes) {
  uint64_t d = 0;
  for (uint64_t i = 0; i < nbytes; ++i) d += a[i] != b[i];
  return d;
}

/* OUTPUT0 == INPUT0 + INPUT1 and OUTPUT1 == INPUT0 - INPUT1 (int32, wrapping) */
uint64_t oracle_addsub_mismatches(const int32_t* out0, const int32_t* out1, const int32_t* in0,
                                  const int32_t* in1, uint64_t n) {
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; ++i) {
    bad += (uint32_t)out0[i] != (uint32_t)in0[i] + (uint32_t)in1[i];
    bad += (uint32_t)out1[i] != (uint32_t)in0[i] - (uint32_t)in1[i];
  }
  return bad;
}

/* argmax ignoring NaN, lowest index on ties; nonfinite counts NaN and +-inf.
 * returns 0xFFFFFFFF when every value is NaN (or n == 0) */
uint32_t oracle_top1(const float* v, uint64_t n, float* max_value, uint64_t* nonfinite) {
  uint32_t best = 0xFFFFFFFFu;
  float bv = 0.0f;
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (!isfinite(v[i])) ++bad;
    if (isnan(v[i])) continue;
    if (best == 0xFFFFFFFFu || v[i] > bv) {
      best = (uint32_t)i;
      bv = v[i];
    }
  }
  *max_value = bv;
  *nonfinite = bad;
  return best;
}

/* top-k in the order of numpy.argsort(-x, kind="stable"): value descending, lower index
 * first on ties (-0.0 == +0.0), NaN after every number in index order.  Restates what
 * the server's classification extension returns and image_client.postprocess
 * (src/python/examples/image_client.py:196-216) consumes.  Selection sort, O(k*n). */
void oracle_topk(const float* v, uint64_t n, uint32_t k, float* values, uint32_t* indices) {
  uint64_t taken_prev = (uint64_t)-1; /* index of the previous pick */
  int prev_nan = 0;
  float prev_v = 0.0f;
  for (uint32_t r = 0; r < k; ++r) {
    uint64_t best = (uint64_t)-1;
    for (uint64_t i = 0; i < n; ++i) {
      /* candidate must come strictly after the previous pick in the order */
      if (taken_prev != (uint64_t)-1) {
        const int cn = isnan(v[i]);
        int after;
        if (prev_nan) after = cn && i > taken_prev;
        else if (cn) after = 1;
        else after = (v[i] < prev_v) || (v[i] == prev_v && i > taken_prev);
        if (!after) continue;
      }
      if (best == (uint64_t)-1) { best = i; continue; }
      const int bn = isnan(v[best]), cn = isnan(v[i]);
      if (bn && !cn) best = i;
      else if (!bn && !cn && v[i] > v[best]) best = i;
    }
    if (best == (uint64_t)-1) {
      values[r] = 0.0f;
      indices[r] = 0xFFFFFFFFu;
      taken_prev = n; /* nothing comes after */
      prev_nan = 1;
      continue;
    }
    values[r] = v[best];
    indices[r] = (uint32_t)best;
    taken_prev = best;
    prev_nan = isnan(v[best]);
    prev_v = v[best];
  }
}

/* ---- CPU baseline of the reference marshalling for bench.py (a1/a3 of SURVEY
 * section 8): ndarray.tobytes() then b"".join([json] + raw) == two memcpys.
 * PY/http/_infer_input.py:212 and PY/http/_utils.py:141-151. */
void oracle_marshal_http(uint8_t* body, const uint8_t* json, uint64_t json_len, uint8_t* scratch,
                         const uint8_t* tensor, uint64_t nbytes) {
  memcpy(scratch, tensor, nbytes);          /* set_data_from_numpy: tobytes() */
  memcpy(body, json, json_len);             /* b"".join: header ...           */
  memcpy(body + json_len, scratch, nbytes); /* ... followed by the raw tensor */
}
