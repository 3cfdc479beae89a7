"""Oracle restatement of the reference's .npz field preprocessing (test infrastructure only).

``preprocess_fields`` follows /root/reference/augmented_cyclegan/dataloader.py:17-34 line by line.  The resize at
dataloader.py:30 is ``skimage.transform.resize(x, (grid_size, grid_size))`` -- a third-party dependency that is absent
from /root/reference (no requirements file pins it) and from this image.  The reference is Python 2 (``print``
statements, dataloader.py:16), and scikit-image 0.14.x is the last release line that supports Python 2, so the
behaviour restated here is skimage <= 0.14 ``transform/_warps.py:resize`` with its defaults as called there:

* ``order=1`` (bilinear), ``mode=None`` -> ``'constant'`` with ``cval=0`` (the 0.13 / 0.14 default, with a warning that
  0.15 would change it to 'reflect'), ``anti_aliasing=None`` -> no Gaussian pre-filter (0.14; the parameter does not
  exist in 0.13), ``clip=True``, ``preserve_range=False`` (a float image passes through ``img_as_float`` unchanged);
* an [h, w, c] image with a 2-element output shape keeps its channels and is warped channel by channel with the
  affine map input = scale * (output + 0.5) - 0.5 (``resize`` fits it through three corner points);
* ``_warps_cy._warp_fast`` bilinear interpolation: the four neighbours (floor / ceil of the coordinate), each read
  through ``get_pixel2d`` which returns ``cval`` outside the image in 'constant' mode;
* ``clip=True`` (``_clip_warp_output``) clamps the result to the whole input image's [min, max], keeping outputs that
  equal ``cval`` exactly when ``cval`` lies outside that range.  For the loader this is a no-op: every non-constant
  channel has just been scaled to span [-1, 1] and constant ones are 0, so all interpolants (convex combinations of
  those values and cval = 0) already lie inside the range -- which is why the product kernels do not clip.

PARITY UNPINNED for the resize: there is no golden vector and no runnable skimage here; the known-answer tests in
tests/test_oracle.py are hand-computed from the algorithm above.  The scaling / NaN arithmetic (dataloader.py:17-25) is
plain numpy and is checked against numpy itself.
"""
import math

import numpy as np


def skimage014_resize(img, out_hw):
    """img [h, w, c] float -> [oh, ow, c] float64: plain loops, for small cases"""
    img = np.asarray(img, dtype=np.float64)
    h, w, c = img.shape
    oh, ow = out_hw
    rs, cs = float(h) / oh, float(w) / ow
    out = np.zeros((oh, ow, c), dtype=np.float64)

    def px(r, q, ch):
        return img[r, q, ch] if (0 <= r < h and 0 <= q < w) else 0.0

    for ch in range(c):
        for oy in range(oh):
            r = rs * (oy + 0.5) - 0.5
            r0, r1 = int(math.floor(r)), int(math.ceil(r))
            dr = r - r0
            for ox in range(ow):
                q = cs * (ox + 0.5) - 0.5
                q0, q1 = int(math.floor(q)), int(math.ceil(q))
                dq = q - q0
                top = (1 - dq) * px(r0, q0, ch) + dq * px(r0, q1, ch)
                bot = (1 - dq) * px(r1, q0, ch) + dq * px(r1, q1, ch)
                out[oy, ox, ch] = (1 - dr) * top + dr * bot
    # _clip_warp_output: clamp to the WHOLE input image's range; exact cval outputs survive when cval is outside it
    lo, hi = img.min(), img.max()
    keep = (out == 0.0) if not (lo <= 0.0 <= hi) else None
    np.clip(out, lo, hi, out=out)
    if keep is not None:
        out[keep] = 0.0
    return out


def preprocess_fields(arr, grid_size=None):
    """dataloader.py:17-34 (`_load` without the file read)"""
    arr = np.asarray(arr)[..., :3]                                                         # :17
    arr = np.nan_to_num(arr)                                                               # :19
    if arr.ndim == 3:                                                                      # :20-21
        arr = np.expand_dims(arr, axis=2)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):                                   # :24
        lo = arr.min((1, 2))[:, np.newaxis, np.newaxis]
        hi = arr.max((1, 2))[:, np.newaxis, np.newaxis]
        arr = -1 + 2 * (arr - lo) / (hi - lo)
    arr = np.nan_to_num(arr)                                                               # :25
    arr[arr == np.inf] = 0
    arr[arr == -np.inf] = 0
    big = np.finfo(arr.dtype).max
    arr[np.abs(arr) >= big] = 0        # nan_to_num has already turned +-inf into +-max: the two lines above never fire; the
    #                                    intent (and any later numpy) is "non-finite -> 0", which only x/0 with x != 0 produces --
    #                                    impossible here because hi == lo implies arr - lo == 0 everywhere (0/0 = NaN -> 0)
    if grid_size is not None:                                                              # :26-31
        arr = np.stack([skimage014_resize(x, (grid_size, grid_size)) for x in arr])
    arr = np.transpose(arr, (0, 3, 1, 2))                                                  # :33
    return arr.astype('float32')                                                           # :34
