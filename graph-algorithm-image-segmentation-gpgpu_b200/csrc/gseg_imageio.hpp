// gseg_imageio.hpp -- image files for the gseg CLI: binary PPM/PGM and PNG, in and out (host only).
//
// the reference's executables read their input with OpenCV (`cv::imread`, any
// format) and write a random-colour image (Report.pdf p2 Fig.1, p4 s3.2.3); upstream `segment` reads
// and writes PPM.  OpenCV's C++ headers are not in this image, zlib is, so PNG is done here directly:
// decoder for every non-interlaced PNG colour type / bit depth (grey, RGB, palette, with or without
// alpha; 16-bit samples keep their high byte, alpha is dropped), encoder for 8-bit RGB.  JPEG files are
// not decoded here: the CLI hands their bytes to gseg_segment_jpeg / a pool job, which decodes them on the GPU (the kernels of
// gseg_jpeg.cuh; nvJPEG for progressive files).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <zlib.h>

namespace gsegio {

inline bool read_file(const char *path, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return true;
}

// ---- PNM (P6 colour, P5 grey; maxval < 256) ------------------------------------------------------
inline bool decode_pnm(const std::vector<uint8_t> &buf, std::vector<uint8_t> &rgb, int &w, int &h, std::string &err) {
    size_t pos = 0;
    auto token = [&](char *out, size_t n) -> bool { // next whitespace-delimited header token, '#' comments skipped
        for (;;) {
            while (pos < buf.size() && (buf[pos] == ' ' || buf[pos] == '\t' || buf[pos] == '\n' || buf[pos] == '\r')) ++pos;
            if (pos < buf.size() && buf[pos] == '#') { while (pos < buf.size() && buf[pos] != '\n') ++pos; continue; }
            break;
        }
        size_t i = 0;
        while (pos < buf.size() && !(buf[pos] == ' ' || buf[pos] == '\t' || buf[pos] == '\n' || buf[pos] == '\r') && i + 1 < n)
            out[i++] = (char)buf[pos++];
        out[i] = 0;
        return i > 0;
    };
    char t[64];
    if (!token(t, sizeof t) || (strcmp(t, "P6") && strcmp(t, "P5"))) { err = "not a binary PPM/PGM"; return false; }
    const int ch = t[1] == '6' ? 3 : 1;
    int maxv = 0;
    bool ok = token(t, sizeof t) && (w = atoi(t)) > 0;
    ok = ok && token(t, sizeof t) && (h = atoi(t)) > 0;
    ok = ok && token(t, sizeof t) && (maxv = atoi(t)) > 0 && maxv < 256;
    if (!ok) { err = "bad PNM header (need maxval < 256)"; return false; }
    ++pos; // the single whitespace byte after maxval
    const size_t n = (size_t)w * h;
    if (buf.size() < pos + n * ch) { err = "PNM data truncated"; return false; }
    rgb.resize(n * 3);
    if (ch == 3) memcpy(rgb.data(), buf.data() + pos, n * 3);
    else
        for (size_t i = 0; i < n; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = buf[pos + i];
    return true;
}

inline bool write_ppm(const char *path, const uint8_t *rgb, int w, int h) {
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    const size_t n = (size_t)w * h * 3;
    const bool ok = fwrite(rgb, 1, n, f) == n;
    return fclose(f) == 0 && ok;
}

// ---- PNG ----------------------------------------------------------------------------------------
inline uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline void put32(std::vector<uint8_t> &v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }

static const uint8_t kPngSig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};

inline bool is_png(const std::vector<uint8_t> &buf) { return buf.size() >= 8 && !memcmp(buf.data(), kPngSig, 8); }

inline bool decode_png(const std::vector<uint8_t> &buf, std::vector<uint8_t> &rgb, int &w, int &h, std::string &err) {
    if (!is_png(buf)) { err = "not a PNG"; return false; }
    size_t pos = 8;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte;
    bool seen_end = false;
    while (pos + 12 <= buf.size() && !seen_end) {
        const uint32_t len = be32(&buf[pos]);
        if (len > buf.size() || pos + 12 + (size_t)len > buf.size()) { err = "PNG chunk truncated"; return false; }
        const uint8_t *type = &buf[pos + 4], *data = &buf[pos + 8];
        if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), type, 4 + len) != be32(data + len)) { err = "PNG chunk CRC mismatch"; return false; }
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13) { err = "bad IHDR"; return false; }
            w = (int)be32(data); h = (int)be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
            if (w <= 0 || h <= 0 || data[10] != 0 || data[11] != 0) { err = "bad IHDR"; return false; }
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!memcmp(type, "IEND", 4)) seen_end = true;
        pos += 12 + (size_t)len;
    }
    if (ctype < 0 || !seen_end) { err = "PNG without IHDR/IEND"; return false; }
    if (interlace) { err = "interlaced (Adam7) PNG is not supported"; return false; }
    int nch;
    switch (ctype) {
        case 0: nch = 1; break;
        case 2: nch = 3; break;
        case 3: nch = 1; break;
        case 4: nch = 2; break;
        case 6: nch = 4; break;
        default: err = "bad PNG colour type"; return false;
    }
    const bool depth_ok = ctype == 0 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)
                        : ctype == 3 ? (depth == 1 || depth == 2 || depth == 4 || depth == 8)
                                     : (depth == 8 || depth == 16);
    if (!depth_ok) { err = "bad PNG bit depth"; return false; }
    if (ctype == 3 && plte.size() < 3) { err = "palette PNG without PLTE"; return false; }
    const size_t bpp_bits = (size_t)nch * depth, rowbytes = ((size_t)w * bpp_bits + 7) / 8;
    const size_t bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1; // filter distance in bytes
    std::vector<uint8_t> raw((rowbytes + 1) * (size_t)h);
    uLongf rawlen = (uLongf)raw.size();
    const int zr = uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || rawlen != raw.size()) { err = "PNG image data does not inflate to the declared size"; return false; }
    // unfilter in place
    std::vector<uint8_t> zero(rowbytes, 0);
    for (int y = 0; y < h; ++y) {
        uint8_t *row = &raw[(rowbytes + 1) * (size_t)y];
        const int ft = row[0];
        uint8_t *cur = row + 1;
        const uint8_t *up = y ? &raw[(rowbytes + 1) * (size_t)(y - 1) + 1] : zero.data(); // previous row, already unfiltered
        for (size_t i = 0; i < rowbytes; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = up[i], c = i >= bpp ? up[i - bpp] : 0;
            int pred;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: err = "bad PNG filter type"; return false;
            }
            cur[i] = (uint8_t)(cur[i] + pred);
        }
    }
    rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = &raw[(rowbytes + 1) * (size_t)y + 1];
        uint8_t *out = &rgb[(size_t)y * w * 3];
        for (int x = 0; x < w; ++x) {
            auto sample = [&](int ch) -> int { // sample `ch` of pixel x, scaled to 8 bits (palette: the index)
                if (depth == 8) return row[(size_t)x * nch + ch];
                if (depth == 16) return row[((size_t)x * nch + ch) * 2];
                const size_t bit = (size_t)x * depth; // nch == 1 for sub-byte depths
                const int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                return ctype == 3 ? v : v * 255 / ((1 << depth) - 1);
            };
            if (ctype == 3) {
                size_t idx = (size_t)sample(0);
                if (3 * idx + 2 >= plte.size()) idx = 0;
                out[3 * x] = plte[3 * idx]; out[3 * x + 1] = plte[3 * idx + 1]; out[3 * x + 2] = plte[3 * idx + 2];
            } else if (nch <= 2) {
                out[3 * x] = out[3 * x + 1] = out[3 * x + 2] = (uint8_t)sample(0);
            } else {
                out[3 * x] = (uint8_t)sample(0); out[3 * x + 1] = (uint8_t)sample(1); out[3 * x + 2] = (uint8_t)sample(2);
            }
        }
    }
    return true;
}

inline bool write_png(const char *path, const uint8_t *rgb, int w, int h) {
    const size_t rowbytes = (size_t)w * 3;
    std::vector<uint8_t> raw((rowbytes + 1) * (size_t)h);
    for (int y = 0; y < h; ++y) { // Up filter on every row but the first: flat segment colours compress to nothing
        uint8_t *row = &raw[(rowbytes + 1) * (size_t)y];
        const uint8_t *src = rgb + rowbytes * (size_t)y;
        row[0] = y ? 2 : 0;
        if (!y) memcpy(row + 1, src, rowbytes);
        else
            for (size_t i = 0; i < rowbytes; ++i) row[1 + i] = (uint8_t)(src[i] - src[i - rowbytes]);
    }
    uLongf zlen = compressBound((uLong)raw.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return false;
    std::vector<uint8_t> out(kPngSig, kPngSig + 8);
    auto chunk = [&](const char *type, const uint8_t *data, size_t len) {
        put32(out, (uint32_t)len);
        const size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), data, data + len);
        put32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), &out[start], (uInt)(4 + len)));
    };
    std::vector<uint8_t> ihdr;
    put32(ihdr, (uint32_t)w); put32(ihdr, (uint32_t)h);
    const uint8_t tail[5] = {8, 2, 0, 0, 0};
    ihdr.insert(ihdr.end(), tail, tail + 5);
    chunk("IHDR", ihdr.data(), ihdr.size());
    chunk("IDAT", z.data(), zlen);
    chunk("IEND", nullptr, 0);
    FILE *f = fopen(path, "wb");
    if (!f) return false;
    const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    return fclose(f) == 0 && ok;
}

// ---- by content (read) / by extension (write) -----------------------------------------------------
inline bool read_image(const char *path, std::vector<uint8_t> &rgb, int &w, int &h, std::string &err) {
    std::vector<uint8_t> buf;
    if (!read_file(path, buf)) { err = "cannot open file"; return false; }
    if (is_png(buf)) return decode_png(buf, rgb, w, h, err);
    if (buf.size() >= 2 && buf[0] == 0xFF && buf[1] == 0xD8) { err = "JPEG is decoded on the GPU (gseg_segment_jpeg), not by this host reader"; return false; }
    return decode_pnm(buf, rgb, w, h, err);
}

inline bool has_suffix(const std::string &s, const char *suf) {
    const size_t n = strlen(suf);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (tolower((unsigned char)s[s.size() - n + i]) != suf[i]) return false;
    return true;
}

inline bool write_image(const char *path, const uint8_t *rgb, int w, int h) {
    return has_suffix(path, ".png") ? write_png(path, rgb, w, h) : write_ppm(path, rgb, w, h);
}

} // namespace gsegio
