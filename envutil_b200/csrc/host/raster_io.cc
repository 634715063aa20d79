// raster_io.cc - ".euf" float rasters: b"EUF1", int32 width, height, channels, then float32
// pixels, row-major, interleaved (the memory layout envutil hands to OIIO, envutil_basic.h:760-775).
// Stands in for OpenImageIO, which this image does not have.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

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

// cubeface_series, envutil_basic.h:267-356: a name with exactly one '%' is a printf format taking the
// six canonical directions - a cubemap stored as six square images
static const char* const face_name[6] = {"left", "right", "top", "bottom", "front", "back"};
static bool series_names(const std::string& fn, std::string out[6]) {
  int percent = 0;
  for (char ch : fn) percent += ch == '%';
  if (percent != 1) return false;
  std::vector<char> buffer(fn.size() + 16);
  for (int i = 0; i < 6; i++) {
    std::snprintf(buffer.data(), buffer.size(), fn.c_str(), face_name[i]);
    out[i] = buffer.data();
  }
  return true;
}

bool read_raster_header(const std::string& fn, int& w, int& h, int& c) {
  FILE* f = open_header(fn, w, h, c);
  std::string faces[6];
  if (!f && series_names(fn, faces)) {  // the series is reported as the 1:6 stripe it stands for
    f = open_header(faces[0], w, h, c);
    if (f && w != h) {
      std::fclose(f);
      return false;
    }
    h = 6 * w;
  }
  if (!f) return false;
  std::fclose(f);
  return true;
}

bool read_raster(const std::string& fn, int& w, int& h, int& c, std::vector<float>& px) {
  FILE* f = open_header(fn, w, h, c);
  std::string faces[6];
  if (!f && series_names(fn, faces)) {  // cubemap_t::load, cubemap.h:1172-1200: six faces, stacked
    for (int i = 0; i < 6; i++) {
      int fw, fh, fc;
      std::vector<float> face;
      if (!read_raster(faces[i], fw, fh, fc, face) || fw != fh) return false;
      if (i == 0) {
        w = fw;
        c = fc;
        px.resize((size_t)w * w * c * 6);
      } else if (fw != w || fc != c) {
        return false;
      }
      std::memcpy(px.data() + (size_t)i * w * w * c, face.data(), face.size() * sizeof(float));
    }
    h = 6 * w;
    return true;
  }
  if (!f) return false;
  px.resize((size_t)w * h * c);
  bool ok = std::fread(px.data(), 4, px.size(), f) == px.size();
  std::fclose(f);
  return ok;
}

bool write_raster(const std::string& fn, int w, int h, int c, const float* px) {
  std::string faces[6];
  if (h == 6 * w && series_names(fn, faces)) {  // save_array, envutil_basic.h:722-755: six sub-arrays
    for (int i = 0; i < 6; i++)
      if (!write_raster(faces[i], w, w, c, px + (size_t)i * w * w * c)) return false;
    return true;
  }
  FILE* f = std::fopen(fn.c_str(), "wb");
  if (!f) return false;
  int32_t hdr[3] = {w, h, c};
  size_t n = (size_t)w * h * c;
  bool ok = std::fwrite("EUF1", 1, 4, f) == 4 && std::fwrite(hdr, 4, 3, f) == 3 && std::fwrite(px, 4, n, f) == n;
  std::fclose(f);
  return ok;
}

}  // namespace eu_host
