"""OpenEXR scanline images: the file format either side of the optimisation loop.

The reference stores its reference renderings as `ref_%06d.exr` (`mi.Bitmap(result).write(fname)`,
optimize.py:53, :57), reads them back with `mi.Bitmap(f)` (optimize.py:78-88) and writes previews
as `opt_<suffix>_%04d.exr` (optimize.py:128-131).  This module reads and writes that container
without OpenEXR: single-part scanline files, float32 / float16 / uint32 channels, compression
NONE, ZIPS (one scanline per block) or ZIP (16 scanlines per block).  PIZ and the lossy codecs
are not implemented and raise NotImplementedError naming the codec.

Layout written (OpenEXR file layout, version 2, no flags):
  magic 0x01312f76, version 2
  attributes: channels (chlist, alphabetical: A? B G R), compression, dataWindow, displayWindow,
              lineOrder (increasing y), pixelAspectRatio, screenWindowCenter, screenWindowWidth; a 0 byte
  offset table: one uint64 per block
  blocks: int32 first scanline, int32 byte count, data; inside a block every scanline holds its
          channels one after the other (all of B, then all of G, ...)
ZIP/ZIPS payload = zlib(deflate) of the block after (1) splitting the bytes into even and odd
halves and (2) replacing every byte by its difference to the previous one plus 128.

Checked in tests/test_host.py against OpenCV's OpenEXR codec (an independent reader / writer).
"""
from __future__ import annotations

import struct
import zlib
from typing import Dict, Tuple

import numpy as np

_MAGIC = 20000630
_PIXEL_TYPES = {0: np.dtype("<u4"), 1: np.dtype("<f2"), 2: np.dtype("<f4")}
_COMPRESSION_NAMES = {0: "NONE", 1: "RLE", 2: "ZIPS", 3: "ZIP", 4: "PIZ", 5: "PXR24", 6: "B44", 7: "B44A", 8: "DWAA", 9: "DWAB"}
_LINES_PER_BLOCK = {0: 1, 2: 1, 3: 16}


def _attr(name: str, type_name: str, payload: bytes) -> bytes:
    return name.encode() + b"\0" + type_name.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def _zip_encode(raw: bytes) -> bytes:
    b = np.frombuffer(raw, dtype=np.uint8)
    t = np.concatenate([b[0::2], b[1::2]]).astype(np.int16)
    d = t.copy()
    d[1:] = t[1:] - t[:-1] + 128
    return zlib.compress((d & 0xFF).astype(np.uint8).tobytes())


def _zip_decode(payload: bytes, size: int) -> bytes:
    d = np.frombuffer(zlib.decompress(payload), dtype=np.uint8).astype(np.int64)
    if d.size != size:
        raise ValueError("EXR: corrupt ZIP block")
    d[1:] -= 128
    t = (np.cumsum(d) & 0xFF).astype(np.uint8)
    half = (size + 1) // 2
    out = np.empty(size, dtype=np.uint8)
    out[0::2] = t[:half]
    out[1::2] = t[half:]
    return out.tobytes()


def write_exr(path: str, image, compression: str = "ZIP") -> None:
    """image: (H, W, 3) RGB or (H, W, 4) RGBA, stored as float32 channels named R, G, B (, A)."""
    a = image.detach().cpu().numpy() if hasattr(image, "detach") else np.asarray(image)
    if a.ndim != 3 or a.shape[2] not in (3, 4):
        raise ValueError(f"expected an (H, W, 3|4) image, got {a.shape}")
    comp = {v: k for k, v in _COMPRESSION_NAMES.items()}.get(compression.upper())
    if comp not in _LINES_PER_BLOCK:
        raise NotImplementedError(f"EXR compression {compression} is not implemented (NONE, ZIPS, ZIP are)")
    a = np.ascontiguousarray(a, dtype="<f4")
    h, w, c = a.shape
    names = sorted("RGBA"[:c])                                   # file order is alphabetical
    planes = {n: a[:, :, "RGBA".index(n)] for n in names}
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", 2, 0, 1, 1) for n in names) + b"\0"
    window = struct.pack("<4i", 0, 0, w - 1, h - 1)
    header = struct.pack("<ii", _MAGIC, 2) + b"".join([
        _attr("channels", "chlist", chlist),
        _attr("compression", "compression", struct.pack("<B", comp)),
        _attr("dataWindow", "box2i", window),
        _attr("displayWindow", "box2i", window),
        _attr("lineOrder", "lineOrder", b"\0"),
        _attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)),
        _attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)),
        _attr("screenWindowWidth", "float", struct.pack("<f", 1.0)),
    ]) + b"\0"
    lines = _LINES_PER_BLOCK[comp]
    blocks = []
    for y0 in range(0, h, lines):
        raw = b"".join(planes[n][y].tobytes() for y in range(y0, min(h, y0 + lines)) for n in names)
        data = raw
        if comp != 0:
            z = _zip_encode(raw)
            data = z if len(z) < len(raw) else raw               # a block that does not shrink is stored raw
        blocks.append(struct.pack("<ii", y0, len(data)) + data)
    offset = len(header) + 8 * len(blocks)
    table = []
    for b in blocks:
        table.append(offset)
        offset += len(b)
    with open(path, "wb") as f:
        f.write(header)
        f.write(struct.pack(f"<{len(table)}Q", *table))
        for b in blocks:
            f.write(b)


def _read_header(buf: bytes) -> Tuple[Dict[str, Tuple[str, bytes]], int]:
    magic, version = struct.unpack_from("<ii", buf, 0)
    if magic != _MAGIC:
        raise ValueError("not an OpenEXR file")
    if version & 0xFF != 2 or version & 0x1A00:                  # tiled (0x200), deep (0x800), multi-part (0x1000)
        raise NotImplementedError(f"EXR version field {version:#x}: only single-part scanline files are supported")
    pos, attrs = 8, {}
    while buf[pos] != 0:
        e = buf.index(b"\0", pos)
        name = buf[pos:e].decode()
        e2 = buf.index(b"\0", e + 1)
        type_name = buf[e + 1:e2].decode()
        (size,) = struct.unpack_from("<i", buf, e2 + 1)
        attrs[name] = (type_name, buf[e2 + 5:e2 + 5 + size])
        pos = e2 + 5 + size
    return attrs, pos + 1


def read_exr(path: str) -> np.ndarray:
    """-> float32 (H, W, C): channels R, G, B (, A) in that order when present, else alphabetical."""
    with open(path, "rb") as f:
        buf = f.read()
    attrs, pos = _read_header(buf)
    comp = attrs["compression"][1][0]
    if comp not in _LINES_PER_BLOCK:
        raise NotImplementedError(f"EXR compression {_COMPRESSION_NAMES.get(comp, comp)} is not implemented "
                                  "(NONE, ZIPS, ZIP are)")
    channels, p, ch = [], 0, attrs["channels"][1]
    while ch[p] != 0:
        e = ch.index(b"\0", p)
        ptype, _, xs, ys = struct.unpack_from("<iB3xii", ch, e + 1)
        if (xs, ys) != (1, 1):
            raise NotImplementedError("EXR: sub-sampled channels are not supported")
        channels.append((ch[p:e].decode(), _PIXEL_TYPES[ptype]))
        p = e + 17
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    lines = _LINES_PER_BLOCK[comp]
    n_blocks = (h + lines - 1) // lines
    offsets = struct.unpack_from(f"<{n_blocks}Q", buf, pos)
    row_bytes = sum(dt.itemsize for _, dt in channels) * w
    planes = {n: np.empty((h, w), dtype=np.float32) for n, _ in channels}
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        rows = min(lines, y1 + 1 - y)
        raw = buf[off + 8:off + 8 + size]
        if comp != 0 and size < rows * row_bytes:
            raw = _zip_decode(raw, rows * row_bytes)
        if len(raw) != rows * row_bytes:
            raise ValueError("EXR: block size does not match the header")
        q = 0
        for r in range(rows):
            for n, dt in channels:
                planes[n][y - y0 + r] = np.frombuffer(raw, dtype=dt, count=w, offset=q).astype(np.float32)
                q += dt.itemsize * w
    names = [n for n, _ in channels]
    order = [n for n in "RGBA" if n in names] if {"R", "G", "B"} <= set(names) else names
    return np.stack([planes[n] for n in order], axis=-1)
