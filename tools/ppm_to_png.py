"""PPM (P6) -> PNG with nothing but zlib: python tools/ppm_to_png.py in.ppm out.png"""
import struct
import sys
import zlib


def main(src, dst):
    raw = open(src, "rb").read()
    assert raw[:2] == b"P6"
    parts = raw.split(b"\n", 3)
    w, h = (int(v) for v in parts[1].split())
    pix = parts[3]
    rows = b"".join(b"\x00" + pix[3 * w * y:3 * w * (y + 1)] for y in range(h))

    def chunk(kind, data):
        c = kind + data
        return struct.pack(">I", len(data)) + c + struct.pack(">I", zlib.crc32(c) & 0xFFFFFFFF)

    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(rows, 9)) + chunk(b"IEND", b"")
    open(dst, "wb").write(png)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
