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
