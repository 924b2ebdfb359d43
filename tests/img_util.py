"""Tiny PNG encoder for the image-format tests (every scanline filter type in turn)."""
import numpy as np


def write_png(path, img, ctype, depth=8, filters=(0, 1, 2, 3, 4)):
    import struct
    import zlib
    img = np.asarray(img)
    h, w = img.shape[:2]
    ch = 1 if img.ndim == 2 else img.shape[2]
    if depth == 16:
        rows = img.astype(">u2").reshape(h, -1).view(np.uint8)
    elif depth == 8:
        rows = img.astype(np.uint8).reshape(h, -1)
    else:
        bits = ((img.reshape(h, w, 1).astype(np.int64) >> np.arange(depth - 1, -1, -1)) & 1).astype(np.uint8).reshape(h, -1)
        rows = np.packbits(bits, axis=1)
    bpp = max(1, ch * depth // 8)
    raw = bytearray()
    prev = np.zeros(rows.shape[1], np.int64)
    for y in range(h):
        cur = rows[y].astype(np.int64)
        ft = filters[y % len(filters)]
        left = np.concatenate([np.zeros(bpp, np.int64), cur[:-bpp]]) if rows.shape[1] > bpp else np.zeros_like(cur)
        ul = np.concatenate([np.zeros(bpp, np.int64), prev[:-bpp]]) if rows.shape[1] > bpp else np.zeros_like(cur)
        if ft == 0:
            out = cur
        elif ft == 1:
            out = cur - left
        elif ft == 2:
            out = cur - prev
        elif ft == 3:
            out = cur - ((left + prev) >> 1)
        else:
            pa, pb, pc = np.abs(prev - ul), np.abs(left - ul), np.abs(left + prev - 2 * ul)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, ul))
            out = cur - pred
        raw.append(ft)
        raw += bytes((out & 255).astype(np.uint8))
        prev = cur

    def chunk(kind, body):
        return struct.pack(">I", len(body)) + kind + body + struct.pack(">I", zlib.crc32(kind + body))

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(bytes(raw))) + chunk(b"IEND", b""))
