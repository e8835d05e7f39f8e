"""numpy restatement of image_client.preprocess arithmetic. TEST INFRASTRUCTURE.

Follows src/python/examples/image_client.py:171-188 literally: ``astype`` to the
model datatype, INCEPTION ``(typed / 127.5) - 1``, VGG ``typed - (123,117,104)``
(``- 128`` for one channel), then the HWC -> CHW transpose.  BF16 has no numpy
type: the value is computed in float32 and truncated like
tritonclient.utils.serialize_bf16_tensor (utils/__init__.py:327-331).
"""

import numpy as np


def preprocess_pixels(img_u8_hwc, npdtype, scaling, nchw=True):
    c = img_u8_hwc.shape[2]
    typed = img_u8_hwc.astype(npdtype)
    if scaling == "INCEPTION":
        scaled = (typed / 127.5) - 1
    elif scaling == "VGG":
        if c == 1:
            scaled = typed - np.asarray((128,), dtype=npdtype)
        else:
            scaled = typed - np.asarray((123, 117, 104), dtype=npdtype)
    else:
        scaled = typed
    return np.transpose(scaled, (2, 0, 1)) if nchw else scaled


def pack_batch(images_u8_nhwc, datatype, layout="NCHW", scaling="NONE"):
    """bytes (uint8 array) of the packed batch."""
    npdtype = np.float16 if datatype == "FP16" else np.float32
    outs = [np.ascontiguousarray(preprocess_pixels(img, npdtype, scaling, layout == "NCHW"))
            for img in images_u8_nhwc]
    batch = np.stack(outs, axis=0)
    assert batch.dtype == npdtype, batch.dtype
    if datatype == "BF16":
        return (batch.astype("<f4").reshape(-1).view("<u4") >> np.uint32(16)).astype("<u2").view(np.uint8)
    return np.frombuffer(batch.tobytes(), dtype=np.uint8)


# ---------------------------------------------------------------------------------------
# Image.resize((w, h), Image.BILINEAR) -- the call image_client.preprocess makes
# (src/python/examples/image_client.py:166).  The arithmetic lives in Pillow
# (third-party; libImaging/Resample.c, pinned here by comparison with the Pillow installed
# in this image, see tests/test_oracle.py): two separable passes, horizontal first, each
# with 8-bit output; triangle filter whose support grows with the down-scale factor
# (antialiasing); coefficients normalised in double precision and quantised to 22-bit
# fixed point.
# ---------------------------------------------------------------------------------------
_PRECISION_BITS = 32 - 8 - 2


def resample_coefficients(in_size, out_size):
    """(bounds int32[out,2] = (first source index, tap count), coeffs int32[out, ksize])."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale >= 1.0 else 1.0
    support = 1.0 * filterscale  # bilinear: filter support 1.0
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            t = (x + xmin - center + 0.5) * ss
            if t < 0.0:
                t = -t
            w = 1.0 - t if t < 1.0 else 0.0
            k[x] = w
            ww += w
        if ww != 0.0:
            k[:xmax] /= ww
        bounds[xx] = (xmin, xmax)
        for x in range(xmax):
            v = k[x] * (1 << _PRECISION_BITS)
            coeffs[xx, x] = int(v - 0.5) if k[x] < 0 else int(v + 0.5)
    return bounds, coeffs


def _resample_axis0(img, bounds, coeffs):
    """One pass along axis 0 of a uint8 array [n, ...] -> [out, ...] (8-bit result)."""
    out = np.empty((bounds.shape[0],) + img.shape[1:], dtype=np.uint8)
    src = img.astype(np.int64)
    for i, (first, count) in enumerate(bounds):
        acc = np.full(img.shape[1:], 1 << (_PRECISION_BITS - 1), dtype=np.int64)
        for t in range(count):
            acc += src[first + t] * int(coeffs[i, t])
        out[i] = np.clip(acc >> _PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_bilinear_resize(img_u8_hwc, out_h, out_w):
    """uint8 [h, w, c] -> uint8 [out_h, out_w, c], bit-identical to
    ``np.array(Image.fromarray(img).resize((out_w, out_h), Image.BILINEAR))``."""
    img = np.ascontiguousarray(img_u8_hwc, dtype=np.uint8)
    h, w = img.shape[:2]
    if h > 100 * w and h != out_h and w != out_w:
        # observed with Pillow 12.2 (and pinned by tests/test_oracle.py): sources taller than
        # 100:1 are resampled vertically first
        bounds, coeffs = resample_coefficients(h, out_h)
        img = _resample_axis0(img, bounds, coeffs)
        h = out_h
    if w != out_w:  # horizontal pass first
        bounds, coeffs = resample_coefficients(w, out_w)
        img = np.ascontiguousarray(np.swapaxes(_resample_axis0(np.ascontiguousarray(np.swapaxes(img, 0, 1)), bounds, coeffs), 0, 1))
    if h != out_h:
        bounds, coeffs = resample_coefficients(h, out_h)
        img = _resample_axis0(img, bounds, coeffs)
    return img
