// raster_io.cc - ".euf" float rasters: b"EUF1", int32 width, height, channels, then float32
// pixels, row-major, interleaved (the memory layout envutil hands to OIIO, envutil_basic.h:760-775).
// Stands in for OpenImageIO, which this image does not have.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "envutil_host.h"

namespace eu_host {

static FILE* open_header(const std::string& fn, int& w, int& h, int& c) {
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return nullptr;
  char magic[4];
  int32_t hdr[3];
  if (std::fread(magic, 1, 4, f) != 4 || std::memcmp(magic, "EUF1", 4) != 0 || std::fread(hdr, 4, 3, f) != 3 ||
      hdr[0] <= 0 || hdr[1] <= 0 || hdr[2] <= 0 || hdr[2] > 4) {
    std::fclose(f);
    return nullptr;
  }
  w = hdr[0];
  h = hdr[1];
  c = hdr[2];
  return f;
}

bool read_raster_header(const std::string& fn, int& w, int& h, int& c) {
  FILE* f = open_header(fn, w, h, c);
  if (!f) return false;
  std::fclose(f);
  return true;
}

bool read_raster(const std::string& fn, int& w, int& h, int& c, std::vector<float>& px) {
  FILE* f = open_header(fn, w, h, c);
  if (!f) return false;
  px.resize((size_t)w * h * c);
  bool ok = std::fread(px.data(), 4, px.size(), f) == px.size();
  std::fclose(f);
  return ok;
}

bool write_raster(const std::string& fn, int w, int h, int c, const float* px) {
  FILE* f = std::fopen(fn.c_str(), "wb");
  if (!f) return false;
  int32_t hdr[3] = {w, h, c};
  size_t n = (size_t)w * h * c;
  bool ok = std::fwrite("EUF1", 1, 4, f) == 4 && std::fwrite(hdr, 4, 3, f) == 3 && std::fwrite(px, 4, n, f) == n;
  std::fclose(f);
  return ok;
}

}  // namespace eu_host
