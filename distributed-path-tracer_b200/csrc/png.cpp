// png.cpp — PNG encode/decode on top of zlib (load-time / output-time only).
//
// write_png_rgba8 stands in for image::save_to_memory_png (stb_image_write,
// LIB/image/image.cpp:111-122): RGBA8, non-interlaced.  read_png stands in for
// image::load via stb_image (LIB/image/image.cpp:23-54) for the texture formats
// the bundled scenes use: the decoded channel count follows stb's rules
// (palette → RGB/RGBA, tRNS adds alpha, 16-bit keeps the high byte).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>
#include <zlib.h>

#include "errors.hpp"
#include "scene.hpp"

namespace ptb {

namespace {

void put32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(uint8_t(x >> 24));
    v.push_back(uint8_t(x >> 16));
    v.push_back(uint8_t(x >> 8));
    v.push_back(uint8_t(x));
}

void chunk(std::vector<uint8_t>& out, const char* type, const uint8_t* data, size_t n) {
    put32(out, static_cast<uint32_t>(n));
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    const uint32_t crc = static_cast<uint32_t>(crc32(0L, out.data() + start, static_cast<uInt>(n + 4)));
    put32(out, crc);
}

uint32_t get32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }

int paeth(int a, int b, int c) {
    const int p = a + b - c;
    const int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

} // namespace

void write_png_rgba8(const std::string& path, const uint8_t* rgba8, uint32_t w, uint32_t h) {
    std::vector<uint8_t> raw;
    raw.reserve(size_t(h) * (size_t(w) * 4 + 1));
    for (uint32_t y = 0; y < h; y++) {
        raw.push_back(0); // filter: none
        raw.insert(raw.end(), rgba8 + size_t(y) * w * 4, rgba8 + size_t(y + 1) * w * 4);
    }
    uLongf zlen = compressBound(static_cast<uLong>(raw.size()));
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, raw.data(), static_cast<uLong>(raw.size()), 6) != Z_OK)
        throw Error(PTB_E_IO, "PNG: zlib compress failed");
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    put32(ihdr, w);
    put32(ihdr, h);
    ihdr.push_back(8); // bit depth
    ihdr.push_back(6); // RGBA
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    chunk(out, "IHDR", ihdr.data(), ihdr.size());
    chunk(out, "IDAT", z.data(), zlen);
    chunk(out, "IEND", nullptr, 0);
    std::ofstream f(path, std::ios::binary);
    if (!f) throw Error(PTB_E_IO, "cannot open " + path + " for writing");
    f.write(reinterpret_cast<const char*>(out.data()), static_cast<std::streamsize>(out.size()));
    if (!f) throw Error(PTB_E_IO, "short write to " + path);
}

void read_png(const std::string& path, OwnedTexture& tex) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(PTB_E_IO, "cannot open texture " + path);
    std::ostringstream ss;
    ss << f.rdbuf();
    const std::string file = ss.str();
    const uint8_t* d = reinterpret_cast<const uint8_t*>(file.data());
    const size_t n = file.size();
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (n < 8 || std::memcmp(d, sig, 8) != 0)
        throw Error(PTB_E_IO, "texture " + path + " is not a PNG (only PNG textures are supported)");
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, palette, trns;
    size_t p = 8;
    bool end = false;
    while (!end && p + 12 <= n) {
        const uint32_t len = get32(d + p);
        if (p + 12 + size_t(len) > n) throw Error(PTB_E_IO, "PNG: truncated chunk in " + path);
        const char* type = reinterpret_cast<const char*>(d + p + 4);
        const uint8_t* body = d + p + 8;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) throw Error(PTB_E_IO, "PNG: bad IHDR");
            w = get32(body);
            h = get32(body + 4);
            depth = body[8];
            ctype = body[9];
            interlace = body[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            palette.assign(body, body + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(body, body + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            end = true;
        }
        p += 12 + size_t(len);
    }
    if (!w || !h || w > (1u << 15) || h > (1u << 15)) throw Error(PTB_E_IO, "PNG: bad dimensions in " + path);
    if (interlace) throw Error(PTB_E_IO, "PNG: interlaced images are not supported: " + path);
    int src_ch = 0;
    switch (ctype) {
    case 0: src_ch = 1; break;
    case 2: src_ch = 3; break;
    case 3: src_ch = 1; break;
    case 4: src_ch = 2; break;
    case 6: src_ch = 4; break;
    default: throw Error(PTB_E_IO, "PNG: bad colour type in " + path);
    }
    if (!(depth == 8 || depth == 16 || ((ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4))))
        throw Error(PTB_E_IO, "PNG: unsupported bit depth in " + path);
    const size_t bpp_bits = size_t(src_ch) * depth;
    const size_t row_bytes = (size_t(w) * bpp_bits + 7) / 8;
    const size_t fbpp = std::max<size_t>(1, bpp_bits / 8); // filter byte distance
    std::vector<uint8_t> raw(size_t(h) * (row_bytes + 1));
    uLongf raw_len = static_cast<uLongf>(raw.size());
    if (uncompress(raw.data(), &raw_len, idat.data(), static_cast<uLong>(idat.size())) != Z_OK || raw_len != raw.size())
        throw Error(PTB_E_IO, "PNG: inflate failed for " + path);
    // unfilter in place
    std::vector<uint8_t> zero(row_bytes, 0);
    for (uint32_t y = 0; y < h; y++) {
        uint8_t* row = raw.data() + size_t(y) * (row_bytes + 1);
        const int filter = row[0];
        uint8_t* cur = row + 1;
        const uint8_t* prev = y ? raw.data() + size_t(y - 1) * (row_bytes + 1) + 1 : zero.data();
        for (size_t i = 0; i < row_bytes; i++) {
            const int a = i >= fbpp ? cur[i - fbpp] : 0, b = prev[i], c = i >= fbpp ? prev[i - fbpp] : 0;
            int v = cur[i];
            switch (filter) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: throw Error(PTB_E_IO, "PNG: bad filter in " + path);
            }
            cur[i] = static_cast<uint8_t>(v);
        }
    }
    // expand to 8-bit channels following stb_image's conventions
    int out_ch = src_ch;
    if (ctype == 3) out_ch = trns.empty() ? 3 : 4;
    else if (!trns.empty() && (ctype == 0 || ctype == 2)) out_ch = src_ch + 1;
    tex.width = w;
    tex.height = h;
    tex.channels = static_cast<uint32_t>(out_ch);
    tex.is_float = 0;
    tex.pixels.assign(size_t(w) * h * out_ch, 255);
    auto sample = [&](const uint8_t* row, size_t idx) -> uint32_t { // idx-th sample of the row, raw value
        if (depth == 8) return row[idx];
        if (depth == 16) return (uint32_t(row[2 * idx]) << 8) | row[2 * idx + 1];
        const size_t bit = idx * depth;
        return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
    };
    const uint32_t grey_scale = depth < 8 ? 255u / ((1u << depth) - 1u) : 1u;
    for (uint32_t y = 0; y < h; y++) {
        const uint8_t* row = raw.data() + size_t(y) * (row_bytes + 1) + 1;
        uint8_t* o = tex.pixels.data() + size_t(y) * w * out_ch;
        for (uint32_t x = 0; x < w; x++) {
            if (ctype == 3) {
                const uint32_t idx = sample(row, x);
                for (int c = 0; c < 3; c++) o[x * out_ch + c] = idx * 3 + c < palette.size() ? palette[idx * 3 + c] : 0;
                if (out_ch == 4) o[x * 4 + 3] = idx < trns.size() ? trns[idx] : 255;
                continue;
            }
            bool transparent = !trns.empty();
            for (int c = 0; c < src_ch; c++) {
                const uint32_t v = sample(row, size_t(x) * src_ch + c);
                if (!trns.empty() && (ctype == 0 || ctype == 2)) {
                    const uint32_t key = (uint32_t(trns[2 * c]) << 8) | trns[2 * c + 1];
                    if (v != key) transparent = false;
                }
                o[x * out_ch + c] = depth == 16 ? uint8_t(v >> 8) : uint8_t(v * grey_scale);
            }
            if (out_ch == src_ch + 1 && (ctype == 0 || ctype == 2)) o[x * out_ch + src_ch] = transparent ? 0 : 255;
        }
    }
}

} // namespace ptb
