#include "image_io.h"

#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace faldoi_host {

namespace {

// The file's bytes, in a buffer that the calling thread reuses from file to file (a sequence decodes thousands of
// frames on a few pool threads: fresh multi-megabyte allocations per file cost more in page faults than the decode).
// max_bytes > 0 reads only the head of the file (header probes).
const std::vector<uint8_t> &slurp(const std::string &path, size_t max_bytes = 0) {
    static thread_local std::vector<uint8_t> buf;
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open '" + path + "'");
    buf.clear();
    if (max_bytes) {
        buf.resize(max_bytes);
        buf.resize(std::fread(buf.data(), 1, max_bytes, f));
    } else if (std::fseek(f, 0, SEEK_END) == 0) {
        const long n = std::ftell(f);
        std::rewind(f);
        if (n > 0) {
            buf.resize((size_t)n);
            buf.resize(std::fread(buf.data(), 1, (size_t)n, f));
        }
    } else {  // not seekable: read in pieces
        uint8_t tmp[65536];
        size_t got;
        while ((got = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    }
    std::fclose(f);
    return buf;
}

// the image's sample storage: the caller's sink when it is large enough, else the image's own vector
float *attach(Image &im, const Sink &sink) {
    const size_t n = im.count();
    if (sink.ptr && sink.cap >= n) {
        im.ext = sink.ptr;
        return im.ext;
    }
    im.ext = nullptr;
    im.data.resize(n);
    return im.data.data();
}

uint32_t be32(const uint8_t *p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }

// ---------------------------------------------------------------- PNG
inline int paeth(int a, int b, int c) {
    // p = a + b - c; |p - a| = |b - c|, |p - b| = |a - c|, |p - c| = |a + b - 2c|; selects instead of branches
    const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
    const int bc = pb <= pc ? b : c;
    return (pa <= pb && pa <= pc) ? a : bc;
}

Image read_png(const std::vector<uint8_t> &buf, const std::string &path, const Sink &sink, bool header_only) {
    auto bad = [&](const char *why) { return std::runtime_error("PNG '" + path + "': " + why); };
    if (buf.size() < 33) throw bad("truncated");
    size_t pos = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    static thread_local std::vector<uint8_t> idat, raw, pix, zero_row;  // reused from file to file by this thread
    std::vector<uint8_t> plte;
    idat.clear();
    bool have_ihdr = false;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const char *type = (const char *)&buf[pos + 4];
        if (pos + 12 + len > buf.size()) {
            if (header_only && have_ihdr) break;
            throw bad("truncated chunk");
        }
        const uint8_t *d = &buf[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) throw bad("bad IHDR");
            w = be32(d);
            h = be32(d + 4);
            depth = d[8];
            ctype = d[9];
            interlace = d[12];
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) {
            plte.assign(d, d + len);
        } else if (!memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), d, d + len);
        } else if (!memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (!have_ihdr || w == 0 || h == 0) throw bad("no IHDR");
    if (interlace) throw bad("interlaced PNG is not supported");
    // the header of an untrusted file: sizes that fit an int (and a sane pixel count), and only the bit depths
    // the PNG specification allows for the colour type -- before any allocation or shift uses them
    if (w > 0x7fffffffu || h > 0x7fffffffu || (uint64_t)w * h > (1ull << 30)) throw bad("image too large");
    int channels;
    bool depth_ok;
    switch (ctype) {
        case 0: channels = 1, depth_ok = (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16); break;
        case 2: channels = 3, depth_ok = (depth == 8 || depth == 16); break;
        case 3: channels = 1, depth_ok = (depth == 1 || depth == 2 || depth == 4 || depth == 8); break;
        case 4: channels = 2, depth_ok = (depth == 8 || depth == 16); break;
        case 6: channels = 4, depth_ok = (depth == 8 || depth == 16); break;
        default: throw bad("unknown colour type");
    }
    if (!depth_ok) throw bad("bit depth not allowed for this colour type");
    if (header_only) {
        Image hdr;
        hdr.w = (int)w, hdr.h = (int)h, hdr.pd = (ctype == 3) ? 3 : channels;
        return hdr;
    }
    const size_t bpp_bits = (size_t)channels * depth;
    const size_t stride = (w * bpp_bits + 7) / 8;
    const size_t bpp = (bpp_bits + 7) / 8;  // filter unit in bytes
    raw.resize((stride + 1) * h);
    uLongf outlen = raw.size();
    if (uncompress(raw.data(), &outlen, idat.data(), idat.size()) != Z_OK || outlen != raw.size())
        throw bad("zlib inflate failed");
    // undo the scanline filters (one specialised loop per filter type; the first `bpp` bytes of a row have
    // no left neighbour)
    pix.resize(stride * h);
    zero_row.assign(stride, 0);
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t ft = raw[y * (stride + 1)];
        const uint8_t *in = &raw[y * (stride + 1) + 1];
        uint8_t *cur = &pix[y * stride];
        const uint8_t *up = y ? &pix[(y - 1) * stride] : zero_row.data();
        const size_t head = std::min(bpp, stride);
        switch (ft) {
            case 0: memcpy(cur, in, stride); break;
            case 1:
                for (size_t i = 0; i < head; i++) cur[i] = in[i];
                for (size_t i = head; i < stride; i++) cur[i] = (uint8_t)(in[i] + cur[i - bpp]);
                break;
            case 2:
                for (size_t i = 0; i < stride; i++) cur[i] = (uint8_t)(in[i] + up[i]);
                break;
            case 3:
                for (size_t i = 0; i < head; i++) cur[i] = (uint8_t)(in[i] + up[i] / 2);
                for (size_t i = head; i < stride; i++) cur[i] = (uint8_t)(in[i] + (cur[i - bpp] + up[i]) / 2);
                break;
            case 4:
                for (size_t i = 0; i < head; i++) cur[i] = (uint8_t)(in[i] + paeth(0, up[i], 0));
                for (size_t i = head; i < stride; i++) cur[i] = (uint8_t)(in[i] + paeth(cur[i - bpp], up[i], up[i - bpp]));
                break;
            default: throw bad("bad filter type");
        }
    }
    Image im;
    im.w = (int)w;
    im.h = (int)h;
    const bool pal = (ctype == 3);
    im.pd = pal ? 3 : channels;
    const size_t n = (size_t)w * h;
    float *out = attach(im, sink);
    if (depth == 8 && !pal) {  // the common case (8-bit gray / RGB / RGBA): straight de-interleave
        for (int c = 0; c < channels; c++) {
            float *dst = out + c * n;
            for (uint32_t y = 0; y < h; y++) {
                const uint8_t *row = &pix[y * stride] + c;
                float *d = dst + (size_t)y * w;
                for (uint32_t x = 0; x < w; x++) d[x] = (float)row[(size_t)x * channels];
            }
        }
        return im;
    }
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t *row = &pix[y * stride];
        for (uint32_t x = 0; x < w; x++)
            for (int c = 0; c < channels; c++) {
                const size_t s = (size_t)x * channels + c;  // sample index in the row
                unsigned v;
                if (depth == 8)
                    v = row[s];
                else if (depth == 16)
                    v = (unsigned)row[2 * s] << 8 | row[2 * s + 1];
                else {
                    const size_t bit = s * depth;
                    v = (row[bit / 8] >> (8 - depth - bit % 8)) & ((1u << depth) - 1);
                }
                if (pal) {
                    if (3 * v + 2 >= plte.size()) throw bad("palette index out of range");
                    for (int k = 0; k < 3; k++) out[k * n + (size_t)y * w + x] = plte[3 * v + k];
                } else {
                    out[c * n + (size_t)y * w + x] = (float)v;
                }
            }
    }
    return im;
}

// ---------------------------------------------------------------- PNM
Image read_pnm(const std::vector<uint8_t> &buf, const std::string &path, const Sink &sink, bool header_only) {
    auto bad = [&](const char *why) { return std::runtime_error("PNM '" + path + "': " + why); };
    const int kind = buf[1] - '0';
    if (kind < 1 || kind > 6) throw bad("unsupported magic");
    size_t pos = 2;
    auto next_int = [&]() -> long {
        for (;;) {
            while (pos < buf.size() && isspace(buf[pos])) pos++;
            if (pos < buf.size() && buf[pos] == '#') {
                while (pos < buf.size() && buf[pos] != '\n') pos++;
                continue;
            }
            break;
        }
        if (pos >= buf.size() || !isdigit(buf[pos])) throw bad("bad header");
        long v = 0;
        while (pos < buf.size() && isdigit(buf[pos])) v = v * 10 + (buf[pos++] - '0');
        return v;
    };
    Image im;
    im.w = (int)next_int();
    im.h = (int)next_int();
    const bool bitmap = (kind == 1 || kind == 4);
    const long maxv = bitmap ? 1 : next_int();
    im.pd = (kind == 3 || kind == 6) ? 3 : 1;
    if (im.w <= 0 || im.h <= 0 || maxv <= 0 || maxv > 65535) throw bad("bad header");
    if (header_only) return im;
    const size_t n = (size_t)im.w * im.h;
    float *out = attach(im, sink);
    if (kind <= 3) {
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < im.pd; c++) out[c * n + i] = (float)(bitmap ? 1 - next_int() : next_int());
        return im;
    }
    pos++;  // single whitespace after the header
    if (bitmap) {
        const size_t stride = (im.w + 7) / 8;
        if (pos + stride * im.h > buf.size()) throw bad("truncated");
        for (int y = 0; y < im.h; y++)
            for (int x = 0; x < im.w; x++) out[(size_t)y * im.w + x] = 1.f - ((buf[pos + y * stride + x / 8] >> (7 - x % 8)) & 1);
        return im;
    }
    const int bps = maxv > 255 ? 2 : 1;
    if (pos + n * im.pd * bps > buf.size()) throw bad("truncated");
    if (bps == 1 && im.pd == 3) {  // the common case, binary 8-bit RGB: one pass per plane keeps the stores sequential
        const uint8_t *src = &buf[pos];
        for (int c = 0; c < 3; c++) {
            float *dst = out + c * n;
            for (size_t i = 0; i < n; i++) dst[i] = (float)src[3 * i + c];
        }
        return im;
    }
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < im.pd; c++) {
            const uint8_t *p = &buf[pos + (i * im.pd + c) * bps];
            out[c * n + i] = (float)(bps == 2 ? (p[0] << 8 | p[1]) : p[0]);
        }
    return im;
}

// ---------------------------------------------------------------- .flo ("PIEH", w, h, interleaved u,v)
Image read_flo(const std::vector<uint8_t> &buf, const std::string &path, const Sink &sink, bool header_only) {
    int32_t wh[2];
    memcpy(wh, &buf[4], 8);
    Image im;
    im.w = wh[0];
    im.h = wh[1];
    im.pd = 2;
    const size_t n = (size_t)im.w * im.h;
    if (im.w <= 0 || im.h <= 0) throw std::runtime_error(".flo '" + path + "': bad header");
    if (header_only) return im;
    if (buf.size() < 12 + n * 8) throw std::runtime_error(".flo '" + path + "': truncated");
    float *out = attach(im, sink);
    const float *f = (const float *)&buf[12];
    for (size_t i = 0; i < n; i++) {
        float uv[2];
        memcpy(uv, f + 2 * i, 8);
        out[i] = uv[0];
        out[n + i] = uv[1];
    }
    return im;
}

uint32_t crc_chunk(const char *type, const std::vector<uint8_t> &d) {
    uint32_t c = crc32(0L, (const Bytef *)type, 4);
    if (!d.empty()) c = crc32(c, d.data(), (uInt)d.size());
    return c;
}

void put_chunk(std::ofstream &f, const char *type, const std::vector<uint8_t> &d) {
    uint8_t b[4];
    auto be = [&](uint32_t v) {
        b[0] = v >> 24, b[1] = v >> 16, b[2] = v >> 8, b[3] = v;
        f.write((const char *)b, 4);
    };
    be((uint32_t)d.size());
    f.write(type, 4);
    if (!d.empty()) f.write((const char *)d.data(), d.size());
    be(crc_chunk(type, d));
}

}  // namespace

static Image read_any(const std::string &path, const Sink &sink, bool header_only) {
    const std::vector<uint8_t> &buf = slurp(path, header_only ? 4096 : 0);
    static const uint8_t png_magic[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (buf.size() >= 8 && !memcmp(buf.data(), png_magic, 8)) return read_png(buf, path, sink, header_only);
    if (buf.size() >= 12 && !memcmp(buf.data(), "PIEH", 4)) return read_flo(buf, path, sink, header_only);
    if (buf.size() >= 7 && buf[0] == 'P' && buf[1] >= '1' && buf[1] <= '6') return read_pnm(buf, path, sink, header_only);
    throw std::runtime_error("'" + path + "': unsupported image format (PNG, PNM and .flo are supported)");
}

Image read_image_split(const std::string &path, Sink sink) { return read_any(path, sink, false); }

void probe_image(const std::string &path, int *w, int *h, int *pd) {
    const Image im = read_any(path, Sink(), true);
    *w = im.w, *h = im.h, *pd = im.pd;
}

void write_flo(const std::string &path, const float *u1, const float *u2, int w, int h) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    const float magic = 202021.25f;
    const int32_t wh[2] = {w, h};
    f.write((const char *)&magic, 4);
    f.write((const char *)wh, 8);
    std::vector<float> row(2 * (size_t)w);
    for (int j = 0; j < h; j++) {
        for (int i = 0; i < w; i++) {
            row[2 * i] = u1[(size_t)j * w + i];
            row[2 * i + 1] = u2[(size_t)j * w + i];
        }
        f.write((const char *)row.data(), row.size() * sizeof(float));
    }
    f.close();  // a full disk or an I/O error must not leave a truncated flow behind a successful exit
    if (!f) throw std::runtime_error("error while writing '" + path + "'");
}

void write_image_float_split(const std::string &path, const float *planes, int w, int h, int pd) {
    if (pd != 2) throw std::runtime_error("write_image_float_split: only 2-channel flow (.flo) output is supported");
    write_flo(path, planes, planes + (size_t)w * h, w, h);
}

void write_png_gray8(const std::string &path, const int *values, int w, int h) {
    std::vector<uint8_t> raw((size_t)(w + 1) * h);
    for (int j = 0; j < h; j++) {
        raw[(size_t)j * (w + 1)] = 0;
        for (int i = 0; i < w; i++) {
            const int v = values[(size_t)j * w + i];
            raw[(size_t)j * (w + 1) + 1 + i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
    uLongf clen = compressBound(raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), raw.size(), 6) != Z_OK) throw std::runtime_error("zlib deflate failed");
    comp.resize(clen);
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot write '" + path + "'");
    static const uint8_t magic[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    f.write((const char *)magic, 8);
    std::vector<uint8_t> ihdr(13);
    ihdr[0] = w >> 24, ihdr[1] = w >> 16, ihdr[2] = w >> 8, ihdr[3] = w;
    ihdr[4] = h >> 24, ihdr[5] = h >> 16, ihdr[6] = h >> 8, ihdr[7] = h;
    ihdr[8] = 8, ihdr[9] = 0, ihdr[10] = 0, ihdr[11] = 0, ihdr[12] = 0;
    put_chunk(f, "IHDR", ihdr);
    put_chunk(f, "IDAT", comp);
    put_chunk(f, "IEND", {});
    f.close();
    if (!f) throw std::runtime_error("error while writing '" + path + "'");
}

}  // namespace faldoi_host
