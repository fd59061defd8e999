"""In-house baseline-JPEG decoder (csrc/gseg_jpeg_core.h, gseg_jpeg.hpp, gseg_jpeg.cuh).

CPU part: the decoder's arithmetic is written once as __host__ __device__ functions; tests/jpeg_host.cpp drives them
with loops (one iteration per GPU thread) and the result is compared bit for bit with libjpeg (cv2.imdecode) -- the
decoder the reference's cv::imread uses (README.md:26).  GPU part (-m gpu): the kernels, through the C-ABI
(gseg_segment_jpeg / gseg_input_rgb / pool jobs), against the same libjpeg pixels, and the partition against the
oracle on those pixels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host():
    """tests/jpeg_host.cpp -> tests/_build/libjpeghost.so (test infrastructure; the product never loads it)."""
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libjpeghost.so")
    src = os.path.join(HERE, "jpeg_host.cpp")
    csrc = os.path.join(HERE, "..", "graph-algorithm-image-segmentation-gpgpu_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("gseg_jpeg_core.h", "gseg_jpeg.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    L = C.CDLL(so)
    L.jpeg_host_decode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.jpeg_host_why.restype = C.c_char_p
    L.jpeg_host_why.argtypes = [C.c_char_p, C.c_size_t]
    L.jpeg_host_info.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.jpeg_host_decode_sync.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]

    def decode(data):
        w, h, ni = C.c_int(), C.c_int(), C.c_int()
        out = np.zeros(3 << 22, np.uint8)
        rc = L.jpeg_host_decode(data, len(data), out.ctypes.data, out.size, C.byref(w), C.byref(h), C.byref(ni))
        if rc:
            return rc, L.jpeg_host_why(data, len(data)).decode()
        return out[:3 * w.value * h.value].reshape(h.value, w.value, 3).copy(), ni.value
    def decode_sync(data, sub_bytes):
        """the marker-less path: self-synchronising sub-sequences of sub_bytes bytes; returns (pixels | rc, rounds)"""
        w, h, r = C.c_int(), C.c_int(), C.c_int()
        out = np.zeros(3 << 22, np.uint8)
        rc = L.jpeg_host_decode_sync(data, len(data), out.ctypes.data, out.size, C.byref(w), C.byref(h), sub_bytes, C.byref(r))
        if rc:
            return rc, r.value
        return out[:3 * w.value * h.value].reshape(h.value, w.value, 3).copy(), r.value
    decode.lib = L
    decode.sync = decode_sync
    return decode


SAMPLING = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420,
            "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
            "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}


def encode(img, quality=90, sampling="420", rst=0, optimize=0, extra=()):
    bgr = np.ascontiguousarray(img[..., ::-1]) if img.ndim == 3 else np.ascontiguousarray(img)
    ok, enc = cv2.imencode(".jpg", bgr, [cv2.IMWRITE_JPEG_QUALITY, int(quality), cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SAMPLING[sampling],
                                        cv2.IMWRITE_JPEG_RST_INTERVAL, int(rst), cv2.IMWRITE_JPEG_OPTIMIZE, int(optimize), *extra])
    assert ok
    return enc


def libjpeg(enc):
    return np.ascontiguousarray(cv2.imdecode(enc, cv2.IMREAD_COLOR)[..., ::-1])


@pytest.mark.parametrize("sampling", sorted(SAMPLING))
def test_host_build_matches_libjpeg_bit_for_bit(host, oracle, sampling):
    """Every sampling layout x sizes that are / are not multiples of the MCU x qualities x restart intervals x default
    and optimised Huffman tables: identical pixels to libjpeg (islow IDCT, fancy upsampling, fixed-point YCbCr)."""
    n = 0
    for (w, h) in [(64, 48), (65, 47), (17, 9), (8, 8), (1, 1), (3, 5), (2, 33), (161, 120)]:
        img = oracle.synth(w, h, 7 + w)
        for q in (35, 90, 100):
            for rst in (0, 1, 5):
                enc = encode(img, q, sampling, rst, optimize=(q == 90))
                got, nint = host(enc.tobytes())
                assert not isinstance(got, int), (w, h, q, rst, nint)
                assert np.array_equal(got, libjpeg(enc)), (w, h, q, rst)
                n += 1
    assert n == 72


def test_host_restart_intervals_and_grey(host, oracle):
    img = oracle.synth(320, 240, 3)
    for rst, want in ((0, 1), (1, 300), (7, 43), (20, 15)):          # 4:2:0: 20 x 15 MCUs
        got, nint = host(encode(img, 85, "420", rst).tobytes())
        assert nint == want
    enc = encode(img[..., 1], 90, "444", 4)                              # single component: 40 x 30 blocks
    got, nint = host(enc.tobytes())
    assert nint == 300 and np.array_equal(got, libjpeg(enc))
    assert np.array_equal(got[..., 0], got[..., 1]) and np.array_equal(got[..., 0], got[..., 2])
    # flat and extreme content (long zero runs, ZRL symbols, saturated samples)
    for fill in (0, 255):
        enc = encode(np.full((40, 56, 3), fill, np.uint8), 75, "420", 2)
        assert np.array_equal(host(enc.tobytes())[0], libjpeg(enc))
    rng = np.random.default_rng(5)
    noise = rng.integers(0, 256, (72, 88, 3), dtype=np.uint8)
    for q in (10, 100):
        enc = encode(noise, q, "444", 3)
        assert np.array_equal(host(enc.tobytes())[0], libjpeg(enc))


@pytest.mark.parametrize("sampling", sorted(SAMPLING))
def test_host_self_synchronising_decode_matches_libjpeg(host, oracle, sampling):
    """Files WITHOUT restart markers: sub-sequences decoded from guessed states, iterated until every entry state equals
    its predecessor's exit state, then written (DC as differences + prefix sums).  Identical pixels to libjpeg whatever
    the sub-sequence size; the number of rounds stays small for ordinary qualities."""
    worst = 0
    for (w, h) in [(160, 120), (161, 123), (17, 9), (8, 8), (1, 1)]:
        img = oracle.synth(w, h, 11 + w)
        for q in (50, 90, 100):
            enc = encode(img, q, sampling, 0, optimize=(q == 50))
            ref = libjpeg(enc)
            for sub in (16, 64, 128):
                got, rounds = host.sync(enc.tobytes(), sub)
                assert not isinstance(got, int), (w, h, q, sub, got)
                assert np.array_equal(got, ref), (w, h, q, sub)
                if q <= 90 and sub == 128:
                    worst = max(worst, rounds)
    assert worst <= 16
    # files WITH restart markers through the same path: every marker re-synchronises (the padding in front of it is
    # recognised, decoding goes on behind it, the DC predictors restart in the prefix sums)
    img = oracle.synth(200, 120, 5)
    for rst in (1, 3, 20, 97):
        for q in (50, 95):
            enc = encode(img, q, sampling, rst)
            ref = libjpeg(enc)
            for sub in (16, 128, 4096):
                got, _ = host.sync(enc.tobytes(), sub)
                assert not isinstance(got, int) and np.array_equal(got, ref), (rst, q, sub)


def test_host_self_synchronising_decode_of_stuffing_heavy_data(host):
    """Noise at quality 100: one 0xFF in ~100 bytes, i.e. sub-sequence borders that fall on and next to stuffed bytes."""
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    for sampling in ("444", "420"):
        enc = encode(noise, 100, sampling, 0)
        assert enc.tobytes().count(b"\xff\x00") > 100
        ref = libjpeg(enc)
        for sub in (8, 9, 31, 128):
            got, _ = host.sync(enc.tobytes(), sub)
            assert not isinstance(got, int) and np.array_equal(got, ref), (sampling, sub)


def _segments(data):
    """[(marker, start, end)] of the header segments in front of the scan data."""
    out, p = [], 2
    while data[p + 1] != 0xDA:
        n = (data[p + 2] << 8) | data[p + 3]
        out.append((data[p + 1], p, p + 2 + n))
        p += 2 + n
    out.append((0xDA, p, p + 2 + ((data[p + 2] << 8) | data[p + 3])))
    return out


def test_host_parser_header_variants(host, oracle):
    """Header forms cv2's encoder never writes, made by rewriting its files: 16-bit quantisation tables, SOF1, comment and
    application segments, fill bytes in front of markers, one table per DHT / DQT segment, a grey image that declares
    2x2 sampling.  Same pixels as libjpeg decodes from the rewritten file (and as from the original)."""
    img = oracle.synth(120, 88, 4)
    for sampling, rst in (("420", 0), ("444", 5)):
        orig = encode(img, 85, sampling, rst).tobytes()
        ref = libjpeg(np.frombuffer(orig, np.uint8))
        segs = _segments(orig)
        out = bytearray(orig[:2])
        for m, a, b in segs:
            body = orig[a + 4:b]
            if m == 0xDB:                                   # DQT: split into one table per segment, 16-bit entries
                i = 0
                while i < len(body):
                    tq = body[i] & 15
                    vals = body[i + 1:i + 65]
                    wide = bytes([0x10 | tq]) + b"".join(bytes([0, v]) for v in vals)
                    out += b"\xff\xff\xff\xdb" + (len(wide) + 2).to_bytes(2, "big") + wide   # with fill bytes in front
                    i += 65
            elif m == 0xC4:                                 # DHT: one table per segment
                i = 0
                while i < len(body):
                    ns = sum(body[i + 1:i + 17])
                    tab = body[i:i + 17 + ns]
                    out += b"\xff\xc4" + (len(tab) + 2).to_bytes(2, "big") + tab
                    i += 17 + ns
            elif m == 0xC0:                                 # baseline -> extended sequential (same coding)
                out += b"\xff\xfe" + (2 + 11).to_bytes(2, "big") + b"hello gseg!"                # a comment
                out += b"\xff\xe5" + (2 + 4).to_bytes(2, "big") + b"\x00\x01\x02\x03"           # an application segment
                out += b"\xff\xc1" + orig[a + 2:b]
            else:
                out += orig[a:b]
        out += orig[segs[-1][2]:]
        data = bytes(out)
        dec = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        assert dec is not None and np.array_equal(dec[..., ::-1], ref)        # libjpeg takes the rewritten file
        got, _ = host(data)
        assert not isinstance(got, int) and np.array_equal(got, ref), sampling
        got, _ = host.sync(data, 64)
        assert not isinstance(got, int) and np.array_equal(got, ref), sampling
    # single component with declared sampling factors 2x2: still one block per MCU (T.81 A.2.2)
    g = encode(img[..., 1], 90, "444", 3).tobytes()
    segs = _segments(g)
    sof = [s for s in segs if s[0] == 0xC0][0]
    patched = bytearray(g)
    assert patched[sof[1] + 9] == 1 and patched[sof[1] + 11] == 0x11
    patched[sof[1] + 11] = 0x22
    ref = cv2.imdecode(np.frombuffer(bytes(patched), np.uint8), cv2.IMREAD_COLOR)
    assert ref is not None
    got, _ = host(bytes(patched))
    assert not isinstance(got, int) and np.array_equal(got, ref)


def test_host_parser_rejects_what_it_does_not_decode(host, oracle):
    img = oracle.synth(64, 48, 9)
    enc = encode(img, 90, "420", 0, extra=(cv2.IMWRITE_JPEG_PROGRESSIVE, 1)).tobytes()
    assert host(enc) == (2, "progressive, lossless, hierarchical or arithmetic-coded frame")
    assert host(b"\xff\xd8\xff\xe0 this is not a jpeg" + bytes(64))[0] == 1
    assert host(b"GIF89a" + bytes(64))[0] == 1
    good = encode(img, 90, "420", 4).tobytes()
    assert host(good[:200])[0] == 1                                   # truncated inside the headers
    w, h = C.c_int(), C.c_int()
    assert host.lib.jpeg_host_info(good, len(good), C.byref(w), C.byref(h)) == 0 and (w.value, h.value) == (64, 48)
    assert host.lib.jpeg_host_info(enc, len(enc), C.byref(w), C.byref(h)) == 0 and (w.value, h.value) == (64, 48)  # size of a progressive file too
    # entropy-coded data cut short or overwritten: decodes to *something* without reading out of bounds, flags garbage
    cut = good[:len(good) // 2]
    r = host(cut)
    assert r[0] == 1 or isinstance(r[0], np.ndarray) or r[0] == 4
    bad = bytearray(good)
    off = bad.rfind(b"\xff\xda") + 14
    for i in range(off, len(bad) - 2):
        bad[i] = 0xFE if bad[i] != 0xFF and bad[i - 1] != 0xFF else bad[i]
    r = host(bytes(bad))
    assert r[0] == 4 or isinstance(r[0], np.ndarray)


def test_host_random_streams_never_crash(host):
    """Fuzz: random bytes behind valid headers (what a corrupt file looks like to the Huffman threads)."""
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    for t in range(40):
        enc = bytearray(encode(img, 80, ("420", "444", "422")[t % 3], (0, 3)[t % 2]).tobytes())
        sos = enc.rfind(b"\xff\xda") + 14
        k = rng.integers(sos, len(enc) - 2, 24)
        for i in k:
            enc[i] = int(rng.integers(0, 255))                      # never 0xFF: keeps the marker structure
        r = host(bytes(enc))
        assert isinstance(r[0], np.ndarray) or r[0] in (1, 4)


def test_host_sanitizer_fuzz(oracle, tmp_path):
    """The parser reads untrusted files inside the product, the decoder's arithmetic runs on whatever they contain: a
    mutation fuzzer (tests/jpeg_fuzz_host.cpp) built with AddressSanitizer + UBSan, both decode paths on exactly sized
    buffers.  Any wild access, out-of-range shift or signed overflow aborts it."""
    exe = os.path.join(HERE, "_build", "jpeg_fuzz_host")
    src = os.path.join(HERE, "jpeg_fuzz_host.cpp")
    csrc = os.path.join(HERE, "..", "graph-algorithm-image-segmentation-gpgpu_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in ("gseg_jpeg_core.h", "gseg_jpeg.hpp")]
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                            "-o", exe, src], capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("no sanitizer runtime for g++ here: " + r.stderr[-200:])
    files = []
    for i, (w, h) in enumerate([(64, 48), (33, 17), (120, 80), (8, 8)]):
        img = oracle.synth(w, h, i + 1)
        for j, sampling in enumerate(sorted(SAMPLING)):
            for rst in (0, 3):
                f = tmp_path / ("s%d_%d_%d.jpg" % (i, j, rst))
                f.write_bytes(encode(img, (50, 90, 100)[(i + j) % 3], sampling, rst, optimize=j & 1).tobytes())
                files.append(str(f))
    g = tmp_path / "grey.jpg"
    g.write_bytes(encode(oracle.synth(40, 40, 9)[..., 0], 90, "444", 0).tobytes())
    files.append(str(g))
    r = subprocess.run([exe, "6000", "5"] + files, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-3000:])
    n = [int(x) for x in r.stdout.split() if x.isdigit()]
    assert n[0] == 6000 and n[1] > 500 and n[3] > 500, r.stdout   # both outcomes are exercised: rejected and decoded


# ---------------------------------------------------------------- GPU ----------------------------------------------
def same_partition(oracle, a, b):
    ca, na = oracle.canon(a)
    cb, nb = oracle.canon(b)
    return na == nb and np.array_equal(ca, cb)


@pytest.mark.gpu
def test_gpu_decode_identical_to_libjpeg_and_partition_to_oracle(gseg, oracle):
    seg = gseg.Segmenter(400, 300)
    seg.set_jpeg_backend(gseg.JPEG_OWN)
    cases = [(320, 240, "420", 90, 4), (321, 243, "444", 75, 1), (200, 150, "422", 85, 7), (64, 300, "440", 95, 2),
             (400, 96, "411", 60, 3), (17, 9, "420", 90, 0), (8, 8, "444", 100, 0), (1, 1, "420", 90, 0),
             (400, 300, "420", 90, 25), (333, 222, "444", 80, 42), (400, 300, "422", 95, 100)]   # long intervals: the sub-sequence path, with markers
    for i, (w, h, sampling, q, rst) in enumerate(cases):
        img = oracle.synth(w, h, 50 + i)
        enc = encode(img, q, sampling, rst, optimize=i & 1)
        wh = seg.segment_jpeg(enc.tobytes(), sigma=0.8, k=300.0, min_size=20, connectivity=8, variant=0)
        assert wh == (w, h) and seg.jpeg_backend_used() == gseg.JPEG_OWN
        rgb = seg.input_rgb()
        assert np.array_equal(rgb, libjpeg(enc)), (w, h, sampling, q, rst)
        ref = oracle.pipeline(rgb, 0.8, 300.0, 20, 8, oracle.FELZ)
        assert seg.num_components() == ref["n"] and same_partition(oracle, seg.labels(), ref["labels"])
    # both schedules and the other variants on one decoded image
    img = oracle.synth(320, 240, 77)
    enc = encode(img, 90, "420", 8)
    dec = libjpeg(enc)
    for variant, conn in ((gseg.HIER, 8), (gseg.SUPERPIX, 4)):
        for flags in (0, gseg.FLAG_HOST_LOOP):
            seg.segment_jpeg(enc.tobytes(), sigma=0.8, k=300.0, min_size=20, connectivity=conn, variant=variant, flags=flags)
            assert np.array_equal(seg.input_rgb(), dec)
            ref = oracle.pipeline(dec, 0.8, 300.0, 20, conn, variant)
            assert same_partition(oracle, seg.labels(), ref["labels"])
    # grey-scale
    enc = encode(img[..., 0], 90, "444", 5)
    seg.segment_jpeg(enc.tobytes(), sigma=0.5, k=200.0, min_size=10, connectivity=4, variant=0)
    assert np.array_equal(seg.input_rgb(), libjpeg(enc))
    seg.close()


@pytest.mark.gpu
def test_gpu_files_without_restart_markers(gseg, oracle, monkeypatch):
    """k_jpeg_sync + k_jpeg_dcscan: what cv::imwrite / a camera writes (no DRI).  Pixels == libjpeg, partition == oracle;
    sub-sequence sizes that give every thread one sub-sequence and several."""
    for sub in ("128", "32"):
        monkeypatch.setenv("GSEG_JPEG_SUB", sub)
        seg = gseg.Segmenter(640, 480)
        seg.set_jpeg_backend(gseg.JPEG_OWN)
        cases = [(640, 480, "420", 90), (321, 243, "444", 75), (200, 150, "422", 100), (64, 300, "440", 50), (400, 96, "411", 95),
                 (17, 9, "420", 90), (8, 8, "444", 100), (1, 1, "420", 90)]
        for i, (w, h, sampling, q) in enumerate(cases):
            img = oracle.synth(w, h, 150 + i)
            enc = encode(img, q, sampling, 0, optimize=i & 1)
            assert seg.segment_jpeg(enc.tobytes(), sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0) == (w, h)
            assert seg.jpeg_backend_used() == gseg.JPEG_OWN
            rgb = seg.input_rgb()
            assert np.array_equal(rgb, libjpeg(enc)), (sub, w, h, sampling, q)
            if i < 3:
                ref = oracle.pipeline(rgb, 0.8, 300.0, 20, 4, oracle.FELZ)
                assert seg.num_components() == ref["n"] and same_partition(oracle, seg.labels(), ref["labels"])
        # alternating with a restart-marker file on the same context (the coefficient array is shared and must be clean)
        a = encode(oracle.synth(320, 240, 5), 90, "420", 0)
        b = encode(oracle.synth(320, 240, 6), 90, "444", 3)
        for enc in (a, b, a, b):
            seg.segment_jpeg(enc.tobytes(), sigma=0.8, k=300.0, min_size=20)
            assert np.array_equal(seg.input_rgb(), libjpeg(enc))
        # stuffing-heavy data, and corrupt data: an error or an image, never a crash
        rng = np.random.default_rng(4)
        enc = encode(rng.integers(0, 256, (240, 320, 3), dtype=np.uint8), 100, "420", 0)
        seg.segment_jpeg(enc.tobytes(), sigma=0.8, k=300.0, min_size=20)
        assert np.array_equal(seg.input_rgb(), libjpeg(enc))
        bad = bytearray(a.tobytes())
        sos = bad.rfind(b"\xff\xda") + 14
        for i in range(sos + 200, len(bad) - 2, 7):
            if bad[i] != 0xFF and bad[i - 1] != 0xFF and bad[i] != 0:
                bad[i] ^= 0x5A if (bad[i] ^ 0x5A) not in (0xFF, 0x00) else 0x11
        try:
            seg.segment_jpeg(bytes(bad), sigma=0.8, k=300.0, min_size=20)
        except gseg.GsegError as e:
            assert "corrupt JPEG" in str(e)
        seg.segment_jpeg(a.tobytes(), sigma=0.8, k=300.0, min_size=20)
        assert np.array_equal(seg.input_rgb(), libjpeg(a))
        seg.close()


@pytest.mark.gpu
def test_gpu_backends_and_errors(gseg, oracle):
    seg = gseg.Segmenter(400, 300)
    img = oracle.synth(320, 240, 21)
    # automatic choice: baseline files -> in-house, with restart markers (one thread per interval) or without (self-
    # synchronising sub-sequences); progressive -> nvJPEG when it is loadable
    seg.segment_jpeg(encode(img, 90, "420", 4).tobytes(), sigma=0.8, k=300.0, min_size=20)
    assert seg.jpeg_backend_used() == gseg.JPEG_OWN
    big = oracle.synth(400, 300, 22)
    seg.segment_jpeg(encode(big, 90, "444", 0).tobytes(), sigma=0.8, k=300.0, min_size=20)   # 1900 MCUs, no restart markers
    assert seg.jpeg_backend_used() == gseg.JPEG_OWN
    prog = encode(img, 90, "420", 0, extra=(cv2.IMWRITE_JPEG_PROGRESSIVE, 1))
    try:
        seg.segment_jpeg(prog.tobytes(), sigma=0.8, k=300.0, min_size=20)
        assert seg.jpeg_backend_used() == gseg.JPEG_NVJPEG
        assert np.abs(seg.input_rgb().astype(np.int32) - libjpeg(prog).astype(np.int32)).mean() < 8.0
    except gseg.GsegError as e:
        assert "optional dependency" in str(e)
    # forced in-house: a progressive file is refused
    seg.set_jpeg_backend(gseg.JPEG_OWN)
    with pytest.raises(gseg.GsegError) as e:
        seg.segment_jpeg(encode(img, 90, "420", 0, extra=(cv2.IMWRITE_JPEG_PROGRESSIVE, 1)).tobytes(), sigma=0.8, k=300.0, min_size=20)
    assert "progressive" in str(e.value)
    enc = encode(img, 90, "420", 0)
    seg.segment_jpeg(enc.tobytes(), sigma=0.8, k=300.0, min_size=20)
    assert seg.jpeg_backend_used() == gseg.JPEG_OWN and np.array_equal(seg.input_rgb(), libjpeg(enc))
    # not a JPEG, too large for the context: errors; the context stays usable
    with pytest.raises(gseg.GsegError):
        seg.segment_jpeg(b"\xff\xd8\xff\xe0 this is not a jpeg" + bytes(64), sigma=0.8, k=300.0, min_size=20)
    with pytest.raises(gseg.GsegError):
        seg.segment_jpeg(encode(np.zeros((400, 600, 3), np.uint8), 90, "420", 4).tobytes(), sigma=0.8, k=300.0, min_size=20)
    # corrupt entropy-coded data: reported by the wait (GSEG_E_ARG), never a crash; then a good image again
    bad = bytearray(encode(img, 90, "420", 4).tobytes())
    sos = bad.rfind(b"\xff\xda") + 14
    for i in range(sos, len(bad) - 2):
        if bad[i] != 0xFF and bad[i - 1] != 0xFF and bad[i] != 0:
            bad[i] = 0xFE
    try:
        seg.segment_jpeg(bytes(bad), sigma=0.8, k=300.0, min_size=20)
        flagged = False
    except gseg.GsegError as e2:
        flagged = "corrupt JPEG" in str(e2)
    assert flagged
    seg.segment(img, sigma=0.8, k=300.0, min_size=20, connectivity=8, variant=0)
    assert same_partition(oracle, seg.labels(), oracle.pipeline(img, 0.8, 300.0, 20, 8, oracle.FELZ)["labels"])
    seg.close()


@pytest.mark.gpu
def test_gpu_pool_decodes_jpeg_jobs_ahead(gseg, oracle):
    """gseg_pool_*: JPEG jobs are decoded by the in-house kernels on the context's copy stream into the free staging
    buffer while the context's previous job runs; mixed with raw jobs, results in submission order."""
    from importlib import import_module
    batch = import_module(gseg.__name__ + ".batch")
    w, h = 240, 160
    items, decs = [], []
    for i in range(13):
        img = oracle.synth(w, h, 900 + i)
        if i % 4 == 3:
            items.append(img); decs.append(img)
        else:
            enc = encode(img, 90, ("420", "444", "422")[i % 3], 1 + i)
            decs.append(libjpeg(enc))
            items.append(batch.Jpeg(np.frombuffer(enc.tobytes(), np.uint8).copy()) if i % 2 else enc.tobytes())
    pool = batch.Pool(gseg, w, h, contexts=3, caps=gseg.CAP_JPEG)
    try:
        outs = [np.zeros((h, w), np.int32) for _ in items]
        jobs = pool.jobs(items, outs, elem_bytes=4, sigma=0.8, k=300.0, min_size=20, connectivity=4, variant=0)
        res = pool.run(jobs)
        for i in range(len(items)):
            assert res[i].status == 0 and (res[i].w, res[i].h) == (w, h)
            ref = oracle.pipeline(decs[i], 0.8, 300.0, 20, 4, oracle.FELZ)
            assert res[i].n_components == ref["n"] and same_partition(oracle, outs[i], ref["labels"]), i
    finally:
        pool.close()
