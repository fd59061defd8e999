"""The gseg CLI's image readers/writers (csrc/gseg_imageio.hpp) -- no GPU needed:
`gseg --convert in out` only decodes and encodes.  PNGs are produced here with zlib for every colour
type, bit depth and row filter the decoder claims to handle."""
import importlib
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

gseg = importlib.import_module("graph-algorithm-image-segmentation-gpgpu_b200")


@pytest.fixture(scope="module")
def cli():
    gseg.build()
    assert os.path.exists(gseg.CLI_PATH)
    return gseg.CLI_PATH


def _chunk(t, d):
    return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d))


def _filter_row(ft, cur, up, bpp):
    out = bytearray(len(cur))
    for i in range(len(cur)):
        a = cur[i - bpp] if i >= bpp else 0
        b = up[i]
        c = up[i - bpp] if i >= bpp else 0
        if ft == 0: pred = 0
        elif ft == 1: pred = a
        elif ft == 2: pred = b
        elif ft == 3: pred = (a + b) >> 1
        else:
            p = a + b - c
            pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
            pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
        out[i] = (cur[i] - pred) & 255
    return bytes(out)


def make_png(rows, w, h, depth, ctype, nch, filters, plte=None, idat_split=1):
    """rows: list of h byte strings of packed samples."""
    bpp = max(1, nch * depth // 8)
    raw = b""
    up = bytes(len(rows[0]))
    for y, r in enumerate(rows):
        ft = filters[y % len(filters)]
        raw += bytes([ft]) + _filter_row(ft, r, up, bpp)
        up = r
    z = zlib.compress(raw, 6)
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    png += _chunk(b"tEXt", b"Comment\0ancillary chunk to skip")
    if plte is not None:
        png += _chunk(b"PLTE", plte)
    step = (len(z) + idat_split - 1) // idat_split
    for i in range(0, len(z), step):
        png += _chunk(b"IDAT", z[i:i + step])
    return png + _chunk(b"IEND", b"")


def read_ppm(path):
    data = open(path, "rb").read()
    assert data[:3] == b"P6\n"
    hdr, rest = data[3:].split(b"\n255\n", 1)
    w, h = map(int, hdr.split())
    return np.frombuffer(rest, np.uint8).reshape(h, w, 3)


def convert(cli, src, dst):
    r = subprocess.run([cli, "--convert", str(src), str(dst)], capture_output=True, text=True)
    return r


@pytest.mark.parametrize("ctype,depth", [(2, 8), (2, 16), (6, 8), (6, 16), (0, 8), (0, 16), (4, 8), (0, 1), (0, 2), (0, 4),
                                         (3, 1), (3, 2), (3, 4), (3, 8)])
def test_png_decode_all_types_and_filters(cli, tmp_path, ctype, depth):
    rng = np.random.default_rng(ctype * 100 + depth)
    w, h = 37, 11
    nch = {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    plte = None
    if depth >= 8:
        samp = rng.integers(0, 256, (h, w, nch), dtype=np.uint8)      # the high (or only) byte of every sample
        if depth == 16:
            lo = rng.integers(0, 256, (h, w, nch), dtype=np.uint8)
            rows = [np.stack([samp[y], lo[y]], -1).tobytes() for y in range(h)]
        else:
            rows = [samp[y].tobytes() for y in range(h)]
        vals = samp
    else:
        v = rng.integers(0, 1 << depth, (h, w), dtype=np.uint8)
        rows = []
        for y in range(h):
            bits = "".join(format(int(x), "0%db" % depth) for x in v[y])
            bits += "0" * (-len(bits) % 8)
            rows.append(bytes(int(bits[i:i + 8], 2) for i in range(0, len(bits), 8)))
        vals = v[..., None]
    if ctype == 3:
        pal = rng.integers(0, 256, (1 << depth, 3), dtype=np.uint8)
        plte = pal.tobytes()
        want = pal[vals[..., 0]]
    elif nch <= 2:
        g = vals[..., 0].astype(np.int32)
        if depth < 8:
            g = g * 255 // ((1 << depth) - 1)
        want = np.repeat(g.astype(np.uint8)[..., None], 3, -1)
    else:
        want = vals[..., :3]
    png = make_png(rows, w, h, depth, ctype, nch, filters=[0, 1, 2, 3, 4], plte=plte, idat_split=3)
    src, dst = tmp_path / "in.png", tmp_path / "out.ppm"
    src.write_bytes(png)
    r = convert(cli, src, dst)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == "%dx%d" % (w, h)
    assert np.array_equal(read_ppm(dst), want)


def test_png_write_roundtrip_and_third_party_decoder(cli, tmp_path):
    rng = np.random.default_rng(5)
    w, h = 123, 45
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[10:30, 20:90] = (7, 200, 31)                                   # a flat region, like a segment colour
    ppm, png, back = tmp_path / "a.ppm", tmp_path / "a.PNG", tmp_path / "b.ppm"
    ppm.write_bytes(b"P6\n# comment\n%d %d\n255\n" % (w, h) + img.tobytes())
    assert convert(cli, ppm, png).returncode == 0
    data = png.read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    assert convert(cli, png, back).returncode == 0
    assert np.array_equal(read_ppm(back), img)
    cv2 = pytest.importorskip("cv2")                                    # an independent decoder, when present
    dec = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
    assert np.array_equal(dec[..., ::-1], img)


def test_pgm_input_and_errors(cli, tmp_path):
    w, h = 9, 4
    g = np.arange(w * h, dtype=np.uint8).reshape(h, w)
    pgm, out = tmp_path / "g.pgm", tmp_path / "g.ppm"
    pgm.write_bytes(b"P5 %d %d 255\n" % (w, h) + g.tobytes())
    assert convert(cli, pgm, out).returncode == 0
    assert np.array_equal(read_ppm(out), np.repeat(g[..., None], 3, -1))
    # corrupt CRC, truncated data, interlaced, JPEG magic, missing file: an error message, never a crash
    rows = [bytes(3 * w) for _ in range(h)]
    good = make_png(rows, w, h, 8, 2, 3, [0])
    bad_crc = bytearray(good); bad_crc[30] ^= 1
    inter = good.replace(struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0), struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 1))
    ihdr = struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 1)
    inter = inter.replace(struct.pack(">I", zlib.crc32(b"IHDR" + struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))),
                          struct.pack(">I", zlib.crc32(b"IHDR" + ihdr)))
    short = make_png(rows[:-1], w, h, 8, 2, 3, [0])                    # one row missing from the image data
    cases = {"crc.png": (bytes(bad_crc), "CRC"), "cut.png": (good[:40], "PNG"), "inter.png": (inter, "interlaced"),
             "short.png": (short, "inflate"), "x.jpg": (b"\xff\xd8\xff\xe0" + bytes(64), "JPEG"), "empty.ppm": (b"", "PPM")}
    for name, (blob, word) in cases.items():
        f = tmp_path / name
        f.write_bytes(blob)
        r = convert(cli, f, out)
        assert r.returncode == 1 and word in r.stderr, (name, r.stderr)
    assert convert(cli, tmp_path / "missing.png", out).returncode == 1
