"""Radiance `.hdr` (RGBE) images: the format of the reference's environment maps
(`envmap_filename: textures/gamrig_2k.hdr`, `round_platform_2k.hdr`, python/scene_config.py:102-340).

Reader: `#?RADIANCE` / `#?RGBE` header, `FORMAT=32-bit_rle_rgbe`, resolution line `-Y H +X W` (the
standard orientation), scanlines either flat (4 bytes per pixel) or adaptive run-length encoded per
channel (marker 2 2 hi lo; a count byte > 128 repeats the next byte count-128 times, otherwise
`count` literal bytes follow).  Decoding follows the common convention (OpenCV, Mitsuba's bitmap
reader): value = mantissa x 2^(e - 136), e == 0 -> 0.  Writer: flat scanlines (valid for every
reader), used by tests and for exporting maps.  Checked against OpenCV's codec in tests/test_host.py.
"""
from __future__ import annotations

import re

import numpy as np


def _decode(rgbe: np.ndarray) -> np.ndarray:
    e = rgbe[..., 3].astype(np.int32)
    scale = np.where(e > 0, np.ldexp(np.float32(1.0), e - 136), np.float32(0.0)).astype(np.float32)
    return rgbe[..., :3].astype(np.float32) * scale[..., None]


def read_hdr(path: str) -> np.ndarray:
    """-> float32 (H, W, 3) RGB."""
    with open(path, "rb") as f:
        buf = f.read()
    if not (buf.startswith(b"#?RADIANCE") or buf.startswith(b"#?RGBE")):
        raise ValueError(f"{path}: not a Radiance HDR file")
    end = buf.find(b"\n\n")
    if end < 0:
        raise ValueError(f"{path}: header is not terminated")
    header = buf[:end].decode("latin-1")
    fmt = re.search(r"^FORMAT=(\S+)", header, re.M)
    if fmt and fmt.group(1) != "32-bit_rle_rgbe":
        raise NotImplementedError(f"{path}: pixel format {fmt.group(1)} (only 32-bit_rle_rgbe)")
    nl = buf.index(b"\n", end + 2)
    res = buf[end + 2:nl].decode("latin-1").split()
    if len(res) != 4 or res[0] != "-Y" or res[2] != "+X":
        raise NotImplementedError(f"{path}: orientation '{' '.join(res)}' (only '-Y H +X W')")
    h, w = int(res[1]), int(res[3])
    data = np.frombuffer(buf, dtype=np.uint8, offset=nl + 1)
    out = np.empty((h, w, 4), dtype=np.uint8)
    pos = 0
    for y in range(h):
        rle = 8 <= w < 32768 and pos + 4 <= data.size and data[pos] == 2 and data[pos + 1] == 2 and \
            ((int(data[pos + 2]) << 8) | int(data[pos + 3])) == w
        if not rle:                                            # flat scanline
            if pos + 4 * w > data.size:
                raise ValueError(f"{path}: truncated scanline {y}")
            out[y] = data[pos:pos + 4 * w].reshape(w, 4)
            pos += 4 * w
            continue
        pos += 4
        for c in range(4):
            x = 0
            while x < w:
                if pos >= data.size:
                    raise ValueError(f"{path}: truncated scanline {y}")
                n = int(data[pos])
                if n > 128:                                    # run
                    n -= 128
                    if n == 0 or x + n > w or pos + 1 >= data.size:
                        raise ValueError(f"{path}: bad run in scanline {y}")
                    out[y, x:x + n, c] = data[pos + 1]
                    pos += 2
                else:                                          # literals
                    if n == 0 or x + n > w or pos + 1 + n > data.size:
                        raise ValueError(f"{path}: bad literal block in scanline {y}")
                    out[y, x:x + n, c] = data[pos + 1:pos + 1 + n]
                    pos += 1 + n
                x += n
    return _decode(out)


def write_hdr(path: str, image) -> None:
    """image (H, W, 3) float32 >= 0 -> flat RGBE scanlines (shared exponent of the largest channel)."""
    a = image.detach().cpu().numpy() if hasattr(image, "detach") else np.asarray(image)
    if a.ndim != 3 or a.shape[2] != 3:
        raise ValueError(f"expected an (H, W, 3) image, got {a.shape}")
    a = np.maximum(np.asarray(a, dtype=np.float32), 0.0)
    m = a.max(axis=2)
    mant, e = np.frexp(m)                                      # m = mant * 2^e, mant in [0.5, 1)
    scale = np.where(m > 1e-32, mant * 256.0 / np.maximum(m, 1e-38), 0.0).astype(np.float32)
    rgbe = np.zeros(a.shape[:2] + (4,), dtype=np.uint8)
    rgbe[..., :3] = np.clip(a * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    h, w = a.shape[:2]
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n" + f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())
