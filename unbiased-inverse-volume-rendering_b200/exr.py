"""OpenEXR scanline images: the file format either side of the optimisation loop.

The reference stores its reference renderings as `ref_%06d.exr` (`mi.Bitmap(result).write(fname)`,
optimize.py:53, :57), reads them back with `mi.Bitmap(f)` (optimize.py:78-88) and writes previews
as `opt_<suffix>_%04d.exr` (optimize.py:128-131).  This module reads and writes that container
without OpenEXR: single-part scanline files, float32 / float16 / uint32 channels, compression
NONE, ZIPS (one scanline per block) or ZIP (16 scanlines per block); PIZ (32 scanlines per block,
wavelet + Huffman: what Mitsuba's `Bitmap.write` and most HDRI libraries produce) is read but not
written.  RLE, PXR24 and the lossy codecs raise NotImplementedError naming the codec.

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
_READ_LINES_PER_BLOCK = {0: 1, 2: 1, 3: 16, 4: 32}       # reading also understands PIZ (4)


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


# ---- PIZ (read only): 16-bit wavelet transform + Huffman coding of a 32-scanline block -------------------
# Layout of a block (OpenEXR ImfPizCompressor): u16 minNonZero, u16 maxNonZero, the bytes
# [minNonZero, maxNonZero] of a 65536-bit "value used" bitmap, i32 length, Huffman stream.  The decoded
# u16 words hold, channel after channel, the wavelet coefficients of the block's pixels (a 32-bit channel
# counts as two interleaved 16-bit planes); after the inverse wavelet a look-up table built from the bitmap
# maps the compacted values back.

def _piz_huffman(buf: bytes, n_out: int) -> np.ndarray:
    im, i_max, _, n_bits = struct.unpack_from("<4I", buf, 0)
    if not (im < 65537 and i_max < 65537):
        raise ValueError("EXR: corrupt PIZ Huffman header")
    pos, c, lc = 20, 0, 0
    lengths = np.zeros(65537, dtype=np.int64)
    k = im
    while k <= i_max:                                        # packed code lengths: 6 bits each, zero runs
        while lc < 6:
            c = (c << 8) | buf[pos]; pos += 1; lc += 8
        lc -= 6
        l = (c >> lc) & 63
        if l == 63:
            while lc < 8:
                c = (c << 8) | buf[pos]; pos += 1; lc += 8
            lc -= 8
            k += ((c >> lc) & 255) + 6
        elif l >= 59:
            k += l - 59 + 2
        else:
            lengths[k] = l
            k += 1
    # canonical codes: within a length in symbol order, the longest codes are numerically smallest
    count = np.bincount(lengths, minlength=59)
    start = [0] * 59
    code = 0
    for l in range(58, 0, -1):
        start[l] = code
        code = (code + int(count[l])) >> 1
    short_sym = np.full(1 << 14, -1, dtype=np.int64)         # 14-bit prefix -> symbol, length
    short_len = np.zeros(1 << 14, dtype=np.int64)
    long_codes = {}
    for sym in np.nonzero(lengths)[0]:
        l = int(lengths[sym])
        cd = start[l]
        start[l] += 1
        if l <= 14:
            lo = cd << (14 - l)
            short_sym[lo:lo + (1 << (14 - l))] = sym
            short_len[lo:lo + (1 << (14 - l))] = l
        else:
            long_codes[(l, cd)] = int(sym)
    sym_l, len_l = short_sym.tolist(), short_len.tolist()
    out = np.empty(n_out, dtype=np.uint16)
    n, c, lc, used, end = 0, 0, 0, 0, len(buf)
    max_long = max((l for l, _ in long_codes), default=0)
    while used < n_bits and n < n_out:
        while lc < 14 and pos < end:
            c = ((c & 0xFFFFFFFFFF) << 8) | buf[pos]; pos += 1; lc += 8
        idx = (c >> (lc - 14)) & 0x3FFF if lc >= 14 else (c << (14 - lc)) & 0x3FFF
        l = len_l[idx]
        if l:
            sym = sym_l[idx]
        else:                                                # a code longer than 14 bits
            sym = -1
            for l in range(15, max_long + 1):
                while lc < l and pos < end:
                    c = ((c & 0xFFFFFFFFFFFFFFFF) << 8) | buf[pos]; pos += 1; lc += 8
                if lc < l:
                    break
                sym = long_codes.get((l, (c >> (lc - l)) & ((1 << l) - 1)), -1)
                if sym >= 0:
                    break
            if sym < 0:
                raise ValueError("EXR: corrupt PIZ Huffman stream")
        lc -= l
        used += l
        if sym == i_max:                                     # run: repeat the previous word
            while lc < 8:
                c = ((c & 0xFFFFFFFFFF) << 8) | buf[pos]; pos += 1; lc += 8
            lc -= 8
            used += 8
            run = (c >> lc) & 255
            if n == 0 or n + run > n_out:
                raise ValueError("EXR: corrupt PIZ run")
            out[n:n + run] = out[n - 1]
            n += run
        else:
            out[n] = sym
            n += 1
    if n != n_out:
        raise ValueError("EXR: PIZ block decodes to the wrong size")
    return out


def _wdec(l, h, w14: bool):
    """Inverse of one wavelet butterfly on arrays of 16-bit words (OpenEXR wdec14 / wdec16)."""
    if w14:
        ls, hs = l.astype(np.int16).astype(np.int32), h.astype(np.int16).astype(np.int32)
        a = ls + (hs & 1) + (hs >> 1)
        return a.astype(np.int16).astype(np.uint16), (a - hs).astype(np.int16).astype(np.uint16)
    m, d = l.astype(np.int32), h.astype(np.int32)
    b = (m - (d >> 1)) & 0xFFFF
    return ((d + b - 0x8000) & 0xFFFF).astype(np.uint16), b.astype(np.uint16)


def _wav2_decode(v: np.ndarray, max_value: int) -> None:
    """In-place inverse 2-D wavelet transform of a (ny, nx) plane of 16-bit words."""
    ny, nx = v.shape
    w14 = max_value < (1 << 14)
    p = 1
    while p <= min(nx, ny):
        p <<= 1
    p >>= 1
    p2, p = p, p >> 1
    while p >= 1:
        ys, xs = slice(0, ny - p2 + 1, p2), slice(0, nx - p2 + 1, p2)
        ysp, xsp = slice(p, p + ny - p2 + 1, p2), slice(p, p + nx - p2 + 1, p2)
        my, mx_ = len(range(0, ny - p2 + 1, p2)), len(range(0, nx - p2 + 1, p2))
        if my and mx_:
            i00, i10 = _wdec(v[ys, xs], v[ysp, xs], w14)
            i01, i11 = _wdec(v[ys, xsp], v[ysp, xsp], w14)
            v[ys, xs], v[ys, xsp] = _wdec(i00, i01, w14)
            v[ysp, xs], v[ysp, xsp] = _wdec(i10, i11, w14)
        if nx & p and my:                                    # odd column left over at this level
            px = mx_ * p2
            v[ys, px], v[ysp, px] = _wdec(v[ys, px], v[ysp, px], w14)
        if ny & p and mx_:                                   # odd row
            py = my * p2
            v[py, xs], v[py, xsp] = _wdec(v[py, xs], v[py, xsp], w14)
        p2, p = p, p >> 1


def _piz_decode(payload: bytes, rows: int, width: int, channels) -> bytes:
    lo, hi = struct.unpack_from("<HH", payload, 0)
    bitmap = np.zeros(8192, dtype=np.uint8)
    pos = 4
    if lo <= hi:
        bitmap[lo:hi + 1] = np.frombuffer(payload, dtype=np.uint8, count=hi - lo + 1, offset=pos)
        pos += hi - lo + 1
    (length,) = struct.unpack_from("<i", payload, pos)
    used = np.unpackbits(bitmap, bitorder="little").astype(bool)
    used[0] = True
    lut = np.zeros(65536, dtype=np.uint16)
    values = np.nonzero(used)[0]
    lut[:values.size] = values
    max_value = values.size - 1
    words = [dt.itemsize // 2 for _, dt in channels]
    total = rows * width * sum(words)
    data = _piz_huffman(payload[pos + 4:pos + 4 + length], total)
    planes, q = [], 0
    for wds in words:
        plane = data[q:q + rows * width * wds].reshape(rows, width * wds)
        q += plane.size
        for j in range(wds):
            sub = plane[:, j::wds].copy()
            _wav2_decode(sub, max_value)
            plane[:, j::wds] = sub
        planes.append(lut[plane])
    return b"".join(pl[r].astype("<u2").tobytes() for r in range(rows) for pl in planes)


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
    try:
        return _read_exr(buf)
    except (struct.error, IndexError, KeyError) as e:         # truncated / damaged container
        raise ValueError(f"{path}: damaged OpenEXR file ({type(e).__name__}: {e})") from None


def _read_exr(buf: bytes) -> np.ndarray:
    attrs, pos = _read_header(buf)
    comp = attrs["compression"][1][0]
    if comp not in _READ_LINES_PER_BLOCK:
        raise NotImplementedError(f"EXR compression {_COMPRESSION_NAMES.get(comp, comp)} is not implemented "
                                  "(NONE, ZIPS, ZIP, PIZ are)")
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
    lines = _READ_LINES_PER_BLOCK[comp]
    n_blocks = (h + lines - 1) // lines
    offsets = struct.unpack_from(f"<{n_blocks}Q", buf, pos)
    row_bytes = sum(dt.itemsize for _, dt in channels) * w
    planes = {n: np.empty((h, w), dtype=np.float32) for n, _ in channels}
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        rows = min(lines, y1 + 1 - y)
        raw = buf[off + 8:off + 8 + size]
        if comp != 0 and size < rows * row_bytes:             # a block that did not shrink is stored raw
            try:
                raw = _piz_decode(raw, rows, w, channels) if comp == 4 else _zip_decode(raw, rows * row_bytes)
            except (IndexError, struct.error, zlib.error, OverflowError) as e:
                raise ValueError(f"EXR: corrupt block at scanline {y}: {e}") from None
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
