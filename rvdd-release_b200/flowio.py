"""Flow / frame files: the on-disk format of the reference's flow cache, without libtiff.

The reference writes every flow as ``iio.write(path, flow.astype(float32))`` with ``flow`` of shape ``(h, w, 2)``
(data/base_dataset.py:180) and reads it back with ``iio.read`` (data/base_dataset.py:118, infer4rec_dataset.py:200).
``iio`` is a libtiff wrapper; what it puts on disk is a little-endian TIFF with PLANARCONFIG_CONTIG, SamplesPerPixel
= channels, 32-bit SAMPLEFORMAT_IEEEFP, PHOTOMETRIC_MINISBLACK for 1 / 2 samples and RGB (+ one extra sample) for
3 / 4, a single strip, LZW below 2000x2000 pixels and uncompressed above (3rdparty/tvl1flow/iio.c:2965-3033).

``write_tif`` produces the uncompressed variant of exactly that layout (any libtiff reader, hence the reference's
loaders, accepts it); ``read_tif`` reads both variants, i.e. also caches written by the reference itself (a small
TIFF-LZW decoder is included for that).  Neither libtiff, iio, tifffile nor imagecodecs exist in this image, and
``cv2.imwrite`` refuses 2-channel images.
"""
import os
import struct

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 16: ("Q", 8)}
_PHOTOMETRIC = {1: 1, 2: 1, 3: 2, 4: 2}


def write_tif(path, arr):
    """Write ``arr`` ((h, w) or (h, w, c), stored as float32) as a single-strip uncompressed baseline TIFF.
    The file appears atomically (tmp + rename), so an interrupted precompute never leaves a truncated flow behind --
    the cache's resume rule is "skip pairs whose file exists" (data/base_dataset.py:170-171)."""
    a = np.ascontiguousarray(arr, dtype="<f4")
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    data = memoryview(a).cast("B")          # no copy: the pixel bytes go to the file straight from the array
    entries = [
        (256, 4, 1, w), (257, 4, 1, h),                     # ImageWidth, ImageLength
        (258, 3, c, [32] * c),                              # BitsPerSample
        (259, 3, 1, 1),                                     # Compression: none
        (262, 3, 1, _PHOTOMETRIC.get(c, 1)),                # Photometric
        (273, 4, 1, 8),                                     # StripOffsets: pixel data right after the header
        (277, 3, 1, c),                                     # SamplesPerPixel
        (278, 4, 1, h),                                     # RowsPerStrip: one strip
        (279, 4, 1, len(data)),                             # StripByteCounts
        (284, 3, 1, 1),                                     # PlanarConfiguration: contig
        (339, 3, c, [3] * c),                               # SampleFormat: IEEE float
    ]
    if c == 4:
        entries.append((338, 3, 1, 2))                      # ExtraSamples: unassociated alpha (iio.c:2980)
    entries.sort()
    ifd_off = 8 + len(data) + (len(data) & 1)
    extra_off = ifd_off + 2 + 12 * len(entries) + 4
    ifd, extra = struct.pack("<H", len(entries)), b""
    for tag, typ, cnt, val in entries:
        fmt, size = _TYPES[typ]
        vals = val if isinstance(val, list) else [val]
        raw = struct.pack("<%d%s" % (cnt, fmt), *vals)
        if len(raw) <= 4:
            ifd += struct.pack("<HHI", tag, typ, cnt) + raw.ljust(4, b"\0")
        else:
            ifd += struct.pack("<HHII", tag, typ, cnt, extra_off + len(extra))
            extra += raw + (b"\0" if len(raw) & 1 else b"")
    ifd += struct.pack("<I", 0)
    tmp = "%s.tmp.%d" % (path, os.getpid())
    with open(tmp, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, ifd_off))
        f.write(data)
        if len(data) & 1:
            f.write(b"\0")
        f.write(ifd)
        f.write(extra)
    os.replace(tmp, path)


def _lzw_decode(buf, expected):
    """TIFF 6.0 LZW (MSB-first codes of 9..12 bits, ClearCode 256, EOI 257, 'early change')."""
    out = bytearray()
    table = [bytes([i]) for i in range(256)] + [b"", b""]
    bits, nbits, width, prev = 0, 0, 9, None
    for byte in buf:
        bits = (bits << 8) | byte
        nbits += 8
        while nbits >= width:
            nbits -= width
            code = (bits >> nbits) & ((1 << width) - 1)
            if code == 256:
                table = table[:258]
                width, prev = 9, None
                continue
            if code == 257:
                return bytes(out[:expected])
            if prev is None:
                entry = table[code]
            else:
                entry = table[code] if code < len(table) else prev + prev[:1]
                table.append(prev + entry[:1])
            out += entry
            prev = entry
            if len(table) >= (1 << width) - 1 and width < 12:
                width += 1
            if len(out) >= expected:
                return bytes(out[:expected])
    return bytes(out[:expected])


def read_tif_into(path, out):
    """Fast path of the precompute readers: if ``path`` is an uncompressed little-endian float32 TIFF whose strips are
    stored back to back (what ``write_tif`` and libtiff write), read its pixels STRAIGHT into ``out`` ((h, w, c)
    float32, C-contiguous -- e.g. a slice of a pinned staging buffer) with one ``readinto`` and return True.  Anything
    else returns False and the caller falls back to ``read_image``."""
    with open(path, "rb") as f:
        head = f.read(8)
        if head[:4] != b"II*\0":
            return False
        off = struct.unpack("<I", head[4:8])[0]
        f.seek(off)
        n = struct.unpack("<H", f.read(2))[0]
        ifd = f.read(12 * n)
        tags = {}
        for i in range(n):
            tag, typ, cnt, val = struct.unpack("<HHII", ifd[12 * i:12 * i + 12])
            if typ == 3 and cnt == 1:
                val &= 0xffff
            tags[tag] = (typ, cnt, val)

        def first(tag, default):
            """first value of a SHORT / LONG tag (values that do not fit the 4-byte field are fetched)"""
            if tag not in tags:
                return default
            typ, cnt, val = tags[tag]
            size = _TYPES.get(typ, ("B", 1))[1]
            if size * cnt <= 4:
                return val & 0xffff if typ == 3 else val
            f.seek(val)
            return struct.unpack("<" + _TYPES[typ][0], f.read(size))[0]

        w, h, c = first(256, 0), first(257, 0), first(277, 1)
        if (first(259, 1) != 1 or first(339, 1) != 3 or first(258, 1) != 32 or (c > 1 and first(284, 1) != 1)
                or first(317, 1) != 1 or out.shape != (h, w, c) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]):
            return False
        typ, cnt, val = tags[273]
        if cnt != 1:                                        # several strips: accept only if they are contiguous
            fmt, size = _TYPES[typ]
            f.seek(val)
            offs = struct.unpack("<%d%s" % (cnt, fmt), f.read(size * cnt))
            t2, c2, v2 = tags[279]
            f.seek(v2)
            cnts = struct.unpack("<%d%s" % (c2, _TYPES[t2][0]), f.read(_TYPES[t2][1] * c2))
            if any(offs[i] + cnts[i] != offs[i + 1] for i in range(cnt - 1)):
                return False
            val = offs[0]
        f.seek(val)
        return f.readinto(memoryview(out).cast("B")) == out.nbytes


def read_tif(path):
    """Read a TIFF written by ``write_tif`` or by the reference's ``iio.write`` -> (h, w, c) array in its stored
    sample type (callers do ``.astype(np.float32)`` like data/base_dataset.py:118)."""
    with open(path, "rb") as f:
        raw = f.read()
    bo = {b"II": "<", b"MM": ">"}.get(raw[:2])
    if bo is None or struct.unpack(bo + "H", raw[2:4])[0] != 42:
        raise ValueError("%s: not a classic TIFF" % path)
    off = struct.unpack(bo + "I", raw[4:8])[0]
    n = struct.unpack(bo + "H", raw[off:off + 2])[0]
    tags = {}
    for i in range(n):
        tag, typ, cnt = struct.unpack(bo + "HHI", raw[off + 2 + 12 * i: off + 10 + 12 * i])
        fmt, size = _TYPES.get(typ, ("B", 1))
        field = raw[off + 10 + 12 * i: off + 14 + 12 * i]
        if size * cnt > 4:
            p = struct.unpack(bo + "I", field)[0]
            field = raw[p: p + size * cnt]
        if typ == 2:
            tags[tag] = field[:cnt]
        elif typ == 5:
            tags[tag] = list(struct.unpack(bo + "%dI" % (2 * cnt), field[:8 * cnt]))
        else:
            tags[tag] = list(struct.unpack(bo + "%d%s" % (cnt, fmt), field[:size * cnt]))
    w, h = tags[256][0], tags[257][0]
    c = tags.get(277, [1])[0]
    bps = tags.get(258, [1])[0]
    fmt = tags.get(339, [1])[0]
    comp = tags.get(259, [1])[0]
    if tags.get(284, [1])[0] != 1 and c > 1:
        raise ValueError("%s: separate planes are not supported" % path)
    if tags.get(317, [1])[0] != 1:
        raise ValueError("%s: TIFF predictor is not supported" % path)
    kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
    if kind is None or bps not in (8, 16, 32, 64):
        raise ValueError("%s: unsupported sample format" % path)
    dt = np.dtype("%s%s%d" % (bo, kind, bps // 8))
    rps = min(tags.get(278, [h])[0], h)
    offsets, counts = tags[273], tags[279]
    rowbytes = w * c * dt.itemsize
    chunks = []
    for s, (o, k) in enumerate(zip(offsets, counts)):
        rows = min(rps, h - s * rps)
        blob = raw[o:o + k]
        if comp == 5:
            blob = _lzw_decode(blob, rows * rowbytes)
        elif comp != 1:
            raise ValueError("%s: unsupported TIFF compression %d" % (path, comp))
        chunks.append(blob[:rows * rowbytes])
    data = b"".join(chunks)
    if len(data) != h * rowbytes:
        raise ValueError("%s: truncated pixel data" % path)
    return np.frombuffer(data, dtype=dt).reshape(h, w, c).astype(dt.newbyteorder("="), copy=True)


def read_image(path):
    """Frame reader used by the precompute driver: float TIFFs through ``read_tif``, ``.npy`` arrays, anything else
    (png, jpg, integer tiff) through OpenCV, always as (h, w, c) in RGB(A) channel order like ``iio.read``."""
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        a = np.load(path)
    elif ext in (".tif", ".tiff"):
        a = read_tif(path)
    else:
        import cv2
        a = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if a is None:
            raise ValueError("cannot read %s" % path)
        if a.ndim == 3 and a.shape[2] >= 3:
            a = a[:, :, [2, 1, 0] + list(range(3, a.shape[2]))]
    return a if a.ndim == 3 else a[:, :, None]
