// png.hpp -- image file I/O of the host side (see png.cpp)
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace grrt_host {
bool decode_png_rgba8(const std::string& path, std::vector<uint8_t>& rgba, uint32_t& width, uint32_t& height, std::string& err);
bool encode_png_rgb8(const std::string& path, const uint8_t* rgb, uint32_t width, uint32_t height, std::string& err);
bool encode_hdr_xyz(const std::string& path, const double* xyza, uint32_t width, uint32_t height, std::string& err);
}  // namespace grrt_host
