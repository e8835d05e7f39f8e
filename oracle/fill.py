"""numpy restatement of the fill contract (second, independent oracle next to
oracle/tb200_oracle.c).  TEST INFRASTRUCTURE.

perf_analyzer's generator is not in the reference: parity of the
random VALUES with perf_analyzer is UNPINNED.  The Philox4x32-10 block function
follows Salmon et al. (SC'11) and is pinned by KAT vectors (tests/test_oracle.py).

Contract (DESIGN.md "fill contract"): tensor bytes are cut into 16-byte groups;
group g = philox4x32_10(ctr=(g.lo, g.hi, stream.lo, stream.hi), key=(seed.lo, seed.hi));
the four words map to elements in little-endian memory order.
"""

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over counters (uint64 arrays holding 32-bit values)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> S32) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> S32) ^ c3 ^ np.uint64(k1)
        c1 = p1 & MASK32
        c3 = p0 & MASK32
        c0, c2 = n0, n2
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def group_words(ngroups, seed, stream, first_group=0):
    """uint32 array [ngroups, 4] of raw Philox words."""
    g = np.arange(first_group, first_group + ngroups, dtype=np.uint64)
    s_lo = np.full(ngroups, stream & 0xFFFFFFFF, dtype=np.uint64)
    s_hi = np.full(ngroups, (stream >> 32) & 0xFFFFFFFF, dtype=np.uint64)
    w = philox4x32_10(g & MASK32, g >> S32, s_lo, s_hi, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(w, axis=1).astype(np.uint32)


def _fma32(u, span, lo):
    """float32 fma via float64: u*span is exact in float64 (24x24 bits); the sum is
    rounded to 53 bits and then to 24, which differs from a true fma only in
    double-rounding corner cases -> tests allow 1 ulp for scaled float fills."""
    return (u.astype(np.float64) * np.float64(np.float32(span)) + np.float64(np.float32(lo))).astype(np.float32)


def fill_bytes(nbytes, datatype, seed=0, stream=0, lo=0.0, span=0.0, ilo=0, irange=0):
    ngroups = (nbytes + 15) // 16
    w = group_words(ngroups, seed, stream)  # [G,4] uint32
    unit = span == 0.0
    if datatype == "FP32":
        u = (w >> np.uint32(9)).astype(np.float32) * np.float32(2.0 ** -23)
        if not unit:
            u = _fma32(u, span, lo)
        raw = u.astype("<f4").tobytes()
    elif datatype in ("FP16", "BF16"):
        x16 = np.stack([w & np.uint32(0xFFFF), w >> np.uint32(16)], axis=2).reshape(ngroups, 8)
        if datatype == "FP16":
            u = (x16 >> np.uint32(6)).astype(np.float32) * np.float32(2.0 ** -10)
        else:
            u = (x16 >> np.uint32(9)).astype(np.float32) * np.float32(2.0 ** -7)
        if not unit:
            u = _fma32(u, span, lo)
        if datatype == "FP16":
            raw = u.astype("<f2").tobytes()
        else:
            raw = (u.astype("<f4").view("<u4") >> np.uint32(16)).astype("<u2").tobytes()
    elif datatype == "FP64":
        lo_w = w[:, 0::2].astype(np.uint64)
        hi_w = w[:, 1::2].astype(np.uint64)
        m = ((hi_w << np.uint64(32)) | lo_w) >> np.uint64(12)
        u = m.astype(np.float64) * (2.0 ** -52)
        if not unit:
            u = u * span + lo  # callers compare with a 1-ulp tolerance
        raw = u.astype("<f8").tobytes()
    elif datatype in ("INT64", "UINT64"):
        x = (w[:, 1::2].astype(np.uint64) << S32) | w[:, 0::2].astype(np.uint64)
        if irange:
            prod = [(int(v) * int(irange)) >> 64 for v in x.reshape(-1)]
            x = (np.array(prod, dtype=np.uint64) + np.uint64(ilo & (2**64 - 1))).reshape(x.shape)
        raw = x.astype("<u8").tobytes()
    elif datatype in ("INT32", "UINT32"):
        x = w.astype(np.uint64)
        if irange:
            x = ((x * np.uint64(irange)) >> S32) + np.uint64(ilo & 0xFFFFFFFF)
        raw = (x & MASK32).astype("<u4").tobytes()
    elif datatype in ("INT16", "UINT16"):
        x = np.stack([w & np.uint32(0xFFFF), w >> np.uint32(16)], axis=2).reshape(ngroups, 8).astype(np.uint64)
        if irange:
            x = ((x * np.uint64(irange)) >> np.uint64(16)) + np.uint64(ilo & 0xFFFF)
        raw = (x & np.uint64(0xFFFF)).astype("<u2").tobytes()
    elif datatype in ("INT8", "UINT8"):
        x = np.stack([(w >> np.uint32(8 * i)) & np.uint32(0xFF) for i in range(4)], axis=2).reshape(ngroups, 16).astype(np.uint64)
        if irange:
            x = ((x * np.uint64(irange)) >> np.uint64(8)) + np.uint64(ilo & 0xFF)
        raw = (x & np.uint64(0xFF)).astype("u1").tobytes()
    elif datatype == "BOOL":
        x = np.stack([(w >> np.uint32(8 * i)) & np.uint32(1) for i in range(4)], axis=2).reshape(ngroups, 16)
        raw = x.astype("u1").tobytes()
    else:
        raise ValueError(datatype)
    return np.frombuffer(raw[:nbytes], dtype=np.uint8)
