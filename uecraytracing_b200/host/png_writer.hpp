// png_writer.hpp -- 8-bit RGB PNG encoder behind an stbi_write_png-shaped call.
// The reference writes its image with the vendored stb_image_write.h (source.cpp:224-226);
// third-party sources are not copied into this repo, so this is a fresh, minimal encoder:
// filter type 0 on every row, one zlib stream (system zlib) in a single IDAT chunk.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <vector>

namespace uecrt_host {

inline void png_chunk(std::vector<unsigned char>& out, const char type[4], const unsigned char* data, std::size_t n) {
  auto be32 = [&](std::uint32_t v) { out.push_back(v >> 24); out.push_back(v >> 16); out.push_back(v >> 8); out.push_back(v); };
  be32((std::uint32_t)n);
  const std::size_t start = out.size();
  out.insert(out.end(), type, type + 4);
  if (n) out.insert(out.end(), data, data + n);
  be32((std::uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start)));
}

// returns non-zero on success, 0 on failure (same convention as stbi_write_png)
inline int write_png(const char* filename, int w, int h, int comp, const void* data, int stride_bytes) {
  if (comp != 3 || w <= 0 || h <= 0 || !data) return 0;
  std::vector<unsigned char> raw((std::size_t)h * (1 + (std::size_t)w * 3));
  for (int y = 0; y < h; ++y) {
    unsigned char* row = &raw[(std::size_t)y * (1 + (std::size_t)w * 3)];
    row[0] = 0;
    const unsigned char* src = static_cast<const unsigned char*>(data) + (std::size_t)y * stride_bytes;
    for (std::size_t i = 0; i < (std::size_t)w * 3; ++i) row[1 + i] = src[i];
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return 0;
  std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  unsigned char ihdr[13] = {(unsigned char)(w >> 24), (unsigned char)(w >> 16), (unsigned char)(w >> 8), (unsigned char)w,
                            (unsigned char)(h >> 24), (unsigned char)(h >> 16), (unsigned char)(h >> 8), (unsigned char)h,
                            8, 2, 0, 0, 0};
  png_chunk(out, "IHDR", ihdr, 13);
  png_chunk(out, "IDAT", z.data(), zlen);
  png_chunk(out, "IEND", nullptr, 0);
  std::FILE* f = std::fopen(filename, "wb");
  if (!f) return 0;
  const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
  return (std::fclose(f) == 0 && ok) ? 1 : 0;
}

// PFM ("PF": 3 x float32 per pixel, little-endian, BOTTOM row first) of sum/spp from a W*H*4 buffer
inline int write_pfm(const char* filename, int w, int h, const float* accum4, unsigned spp) {
  if (w <= 0 || h <= 0 || !accum4 || spp == 0) return 0;
  std::FILE* f = std::fopen(filename, "wb");
  if (!f) return 0;
  std::fprintf(f, "PF\n%d %d\n-1.0\n", w, h);
  std::vector<float> row((std::size_t)w * 3);
  const float inv = 1.0f / (float)spp;
  bool ok = true;
  for (int y = h - 1; y >= 0 && ok; --y) {
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < 3; ++c) row[(std::size_t)x * 3 + c] = accum4[((std::size_t)y * w + x) * 4 + c] * inv;
    ok = std::fwrite(row.data(), sizeof(float), row.size(), f) == row.size();
  }
  return (std::fclose(f) == 0 && ok) ? 1 : 0;
}

}  // namespace uecrt_host
