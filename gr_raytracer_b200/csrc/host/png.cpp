// png.cpp -- PNG decode/encode and Radiance HDR encode for the host side of libgrrt.
//
// The reference uses the `image` crate (0.25) for all three: texture decode via
// ImageReader::open(..).decode() + get_pixel -> Rgba<u8> (reference src/rendering/texture.rs:47-59,
// 76-79), PNG / HDR output via save / save_with_format (src/rendering/raytracer.rs:468-492).  PNG is
// lossless, so any conforming decoder yields the same texels; zlib does inflate/deflate, the rest
// (chunks, filters, palettes, bit depths, CRC) is implemented here.  Interlaced PNGs are rejected.
#include "png.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>

namespace grrt_host {
namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}
int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    return pb <= pc ? b : c;
}
bool read_file(const std::string& path, std::vector<uint8_t>& out, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) {
        err = "cannot open '" + path + "'";
        return false;
    }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = out.empty() ? 0 : std::fread(out.data(), 1, out.size(), f);
    std::fclose(f);
    if (got != out.size()) {
        err = "short read on '" + path + "'";
        return false;
    }
    return true;
}
// 16-bit sample -> 8-bit the way image's DynamicImage::to_rgba8 does (rounded division by 257)
uint8_t down16(unsigned v) { return (uint8_t)((v + 128u) / 257u); }

}  // namespace

bool decode_png_rgba8(const std::string& path, std::vector<uint8_t>& rgba, uint32_t& width, uint32_t& height, std::string& err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file, err)) return false;
    static const uint8_t SIG[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (file.size() < 8 || std::memcmp(file.data(), SIG, 8) != 0) {
        err = "'" + path + "' is not a PNG file";
        return false;
    }
    size_t p = 8;
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool seen_end = false;
    while (p + 12 <= file.size() && !seen_end) {
        uint32_t len = be32(&file[p]);
        const uint8_t* type = &file[p + 4];
        if (p + 12 + (size_t)len > file.size()) {
            err = "truncated PNG chunk";
            return false;
        }
        const uint8_t* data = &file[p + 8];
        uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), type, 4 + len);
        if (crc != be32(&file[p + 8 + len])) {
            err = "PNG chunk CRC mismatch";
            return false;
        }
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) { err = "bad IHDR"; return false; }
            w = be32(data);
            h = be32(data + 4);
            depth = data[8];
            ctype = data[9];
            interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            seen_end = true;
        }
        p += 12 + (size_t)len;
    }
    if (w == 0 || h == 0 || ctype < 0) { err = "PNG without IHDR"; return false; }
    if (interlace != 0) { err = "interlaced PNG is not supported"; return false; }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "unknown PNG colour type"; return false;
    }
    if (!(depth == 8 || depth == 16 || (depth < 8 && (ctype == 0 || ctype == 3) && (depth == 1 || depth == 2 || depth == 4)))) {
        err = "unsupported PNG bit depth";
        return false;
    }
    const size_t bits_pp = (size_t)channels * depth;
    const size_t stride = ((size_t)w * bits_pp + 7) / 8;
    const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;  // filter distance in bytes
    std::vector<uint8_t> raw((stride + 1) * (size_t)h);
    uLongf raw_len = (uLongf)raw.size();
    int zr = uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size());
    if (zr != Z_OK || raw_len != raw.size()) {
        err = "PNG inflate failed";
        return false;
    }
    // defilter in place
    std::vector<uint8_t> prev(stride, 0), cur(stride);
    rgba.assign((size_t)w * h * 4, 255);
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* line = &raw[(stride + 1) * (size_t)y];
        int filter = line[0];
        const uint8_t* src = line + 1;
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int v;
            switch (filter) {
                case 0: v = src[i]; break;
                case 1: v = src[i] + a; break;
                case 2: v = src[i] + b; break;
                case 3: v = src[i] + ((a + b) >> 1); break;
                case 4: v = src[i] + paeth(a, b, c); break;
                default: err = "bad PNG filter type"; return false;
            }
            cur[i] = (uint8_t)v;
        }
        uint8_t* dst = &rgba[(size_t)y * w * 4];
        for (uint32_t x = 0; x < w; ++x) {
            unsigned s[4] = {0, 0, 0, 0};
            if (depth == 8) {
                for (int ch = 0; ch < channels; ++ch) s[ch] = cur[(size_t)x * channels + ch];
            } else if (depth == 16) {
                for (int ch = 0; ch < channels; ++ch) {
                    size_t o = ((size_t)x * channels + ch) * 2;
                    s[ch] = ((unsigned)cur[o] << 8) | cur[o + 1];
                }
            } else {
                size_t bit = (size_t)x * depth;
                s[0] = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1u);
            }
            uint8_t r, g, b, a = 255;
            if (ctype == 3) {
                unsigned idx = s[0];
                if ((size_t)idx * 3 + 2 >= plte.size()) { err = "PNG palette index out of range"; return false; }
                r = plte[idx * 3];
                g = plte[idx * 3 + 1];
                b = plte[idx * 3 + 2];
                if (idx < trns.size()) a = trns[idx];
            } else {
                auto to8 = [&](unsigned v) -> uint8_t {
                    if (depth == 16) return down16(v);
                    if (depth == 8) return (uint8_t)v;
                    return (uint8_t)(v * 255u / ((1u << depth) - 1u));
                };
                if (ctype == 0 || ctype == 4) {
                    r = g = b = to8(s[0]);
                    if (ctype == 4) a = to8(s[1]);
                    else if (trns.size() >= 2 && s[0] == (((unsigned)trns[0] << 8) | trns[1])) a = 0;
                } else {
                    r = to8(s[0]);
                    g = to8(s[1]);
                    b = to8(s[2]);
                    if (ctype == 6) a = to8(s[3]);
                    else if (trns.size() >= 6 && s[0] == (((unsigned)trns[0] << 8) | trns[1]) &&
                             s[1] == (((unsigned)trns[2] << 8) | trns[3]) && s[2] == (((unsigned)trns[4] << 8) | trns[5]))
                        a = 0;
                }
            }
            dst[4 * x] = r;
            dst[4 * x + 1] = g;
            dst[4 * x + 2] = b;
            dst[4 * x + 3] = a;
        }
        prev.swap(cur);
    }
    width = w;
    height = h;
    return true;
}

bool encode_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t width, uint32_t height, std::string& err) {
    std::vector<uint8_t> raw;
    raw.reserve(((size_t)width * 3 + 1) * height);
    for (uint32_t y = 0; y < height; ++y) {
        raw.push_back(0);  // filter type None
        raw.insert(raw.end(), rgb + (size_t)y * width * 3, rgb + (size_t)(y + 1) * width * 3);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) {
        err = "deflate failed";
        return false;
    }
    comp.resize(clen);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    auto chunk = [&](const char* type, const std::vector<uint8_t>& data) {
        put_be32(out, (uint32_t)data.size());
        size_t start = out.size();
        out.insert(out.end(), type, type + 4);
        out.insert(out.end(), data.begin(), data.end());
        put_be32(out, (uint32_t)crc32(crc32(0L, Z_NULL, 0), &out[start], (uInt)(4 + data.size())));
    };
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width);
    put_be32(ihdr, height);
    ihdr.push_back(8);
    ihdr.push_back(2);
    ihdr.push_back(0);
    ihdr.push_back(0);
    ihdr.push_back(0);
    chunk("IHDR", ihdr);
    chunk("IDAT", comp);
    chunk("IEND", {});
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        err = "cannot create '" + path + "'";
        return false;
    }
    bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) err = "short write on '" + path + "'";
    return ok;
}

// Radiance RGBE, flat (non-RLE) scanlines; pixel conversion as image::codecs::hdr::to_rgbe8.
// The reference hands the encoder XYZ as f32 "RGB" untouched (raytracer.rs:471-480).
bool encode_hdr_xyz(const std::string& path, const double* xyza, uint32_t width, uint32_t height, std::string& err) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        err = "cannot create '" + path + "'";
        return false;
    }
    std::fprintf(f, "#?RADIANCE\nSOFTWARE=grrt-b200\nFORMAT=32-bit_rle_rgbe\n\n-Y %u +X %u\n", height, width);
    std::vector<uint8_t> line((size_t)width * 4);
    bool ok = true;
    for (uint32_t y = 0; y < height && ok; ++y) {
        for (uint32_t x = 0; x < width; ++x) {
            const double* c = xyza + 4 * ((size_t)y * width + x);
            float pix[3] = {(float)c[0], (float)c[1], (float)c[2]};
            float mx = std::fmax(pix[0], std::fmax(pix[1], pix[2]));
            uint8_t* o = &line[4 * (size_t)x];
            if (!(mx > 0.0f)) {
                o[0] = o[1] = o[2] = o[3] = 0;
            } else {
                int exp = (int)std::floor(std::log2(mx)) + 1;
                float mul = std::ldexp(1.0f, exp);
                for (int k = 0; k < 3; ++k) {
                    float v = std::trunc(pix[k] / mul * 256.0f);
                    o[k] = (uint8_t)(v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v));
                }
                o[3] = (uint8_t)(exp + 128);
            }
        }
        ok = std::fwrite(line.data(), 1, line.size(), f) == line.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    if (!ok) err = "short write on '" + path + "'";
    return ok;
}

}  // namespace grrt_host
