"""Minimal writer of uncompressed 16-bit single-channel strip TIFFs (classic and BigTIFF) for the tests of the streaming
LDEM reader: the layout of the LOLA LDEM files (NASA PDS: LDEM_128.TIF is an uncompressed BigTIFF-sized raster)."""
import struct

import numpy as np


def write_strip_tiff(path, arr_u16, rows_per_strip, big=False, byteorder="<"):
    a = np.ascontiguousarray(arr_u16).view(np.uint16)
    H, W = a.shape
    bo = byteorder
    n_strips = -(-H // rows_per_strip)
    data = a.astype(np.dtype(np.uint16).newbyteorder(bo)).tobytes()
    counts = [min(rows_per_strip, H - i * rows_per_strip) * W * 2 for i in range(n_strips)]
    hdr = 16 if big else 8
    offs, at = [], hdr
    for c in counts:
        offs.append(at)
        at += c
    T_SHORT, T_LONG, T_LONG8 = 3, 4, 16
    ot = T_LONG8 if big else T_LONG
    tags = [(256, T_LONG, [W]), (257, T_LONG, [H]), (258, T_SHORT, [16]), (259, T_SHORT, [1]), (262, T_SHORT, [1]),
            (273, ot, offs), (277, T_SHORT, [1]), (278, T_LONG, [rows_per_strip]), (279, ot, counts), (339, T_SHORT, [1])]
    size = {T_SHORT: 2, T_LONG: 4, T_LONG8: 8}
    fmt = {T_SHORT: "H", T_LONG: "I", T_LONG8: "Q"}
    ifd_at = at + (at & 1)
    n = len(tags)
    entry = 20 if big else 12
    ifd_len = (8 if big else 2) + n * entry + (8 if big else 4)
    extra_at = ifd_at + ifd_len
    extra = b""
    ifd = struct.pack(bo + ("Q" if big else "H"), n)
    for tag, typ, vals in tags:
        raw = struct.pack(bo + fmt[typ] * len(vals), *vals)
        inline = 8 if big else 4
        if len(raw) <= inline:
            field = raw.ljust(inline, b"\0")
        else:
            field = struct.pack(bo + ("Q" if big else "I"), extra_at + len(extra))
            extra += raw + (b"\0" if len(raw) & 1 else b"")
        ifd += struct.pack(bo + "HH" + ("Q" if big else "I"), tag, typ, len(vals)) + field
    ifd += struct.pack(bo + ("Q" if big else "I"), 0)
    mark = b"II" if bo == "<" else b"MM"
    with open(path, "wb") as f:
        if big:
            f.write(mark + struct.pack(bo + "HHHQ", 43, 8, 0, ifd_at))
        else:
            f.write(mark + struct.pack(bo + "HI", 42, ifd_at))
        f.write(data)
        f.write(b"\0" * (ifd_at - at))
        f.write(ifd)
        f.write(extra)
