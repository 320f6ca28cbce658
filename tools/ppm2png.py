#!/usr/bin/env python
"""ppm2png.py in.ppm out.png -- tiny P6 -> PNG converter (no PIL needed) for eyeballing renders."""
import struct, sys, zlib
import numpy as np
d = open(sys.argv[1], 'rb').read()
hdr = d.split(b'\n', 3)
W, H = map(int, hdr[1].split())
img = np.frombuffer(hdr[3], dtype=np.uint8).reshape(H, W, 3)
raw = b''.join(b'\x00' + img[y].tobytes() for y in range(H))
def chunk(t, b):
    c = struct.pack('>I', len(b)) + t + b
    return c + struct.pack('>I', zlib.crc32(t + b) & 0xffffffff)
open(sys.argv[2], 'wb').write(b'\x89PNG\r\n\x1a\n' + chunk(b'IHDR', struct.pack('>IIBBBBB', W, H, 8, 2, 0, 0, 0)) + chunk(b'IDAT', zlib.compress(raw)) + chunk(b'IEND', b''))
