// Minimal stand-in for <OpenImageIO/imageio.h> (test infrastructure only).
// Lets the reference's own translation units compile and run without OpenImageIO.
// Images are plain float32 rasters in the ".euf" container:
//   bytes 0-3 "EUF1", int32 width, int32 height, int32 nchannels, then
//   height*width*nchannels little-endian float32, row-major, top row first, interleaved.
// Touch points replaced: envutil_basic.h:545-630 (get_image_metrics), :822-986
// (read_image_data), :710-817 (save_array), cubemap.h:978-1140.
#pragma once
#include <cassert>
#include <cstdint>
#include <functional>
#include <limits>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
#include <chrono>
#include <iostream>

namespace OIIO {

typedef std::ptrdiff_t stride_t;
const stride_t AutoStride = std::numeric_limits<stride_t>::min();

struct TypeDesc {
  enum BASETYPE { UNKNOWN, NONE, UINT8, INT8, UINT16, INT16, UINT32, INT32, UINT64, INT64, HALF, FLOAT, DOUBLE, STRING };
  BASETYPE basetype;
  TypeDesc(BASETYPE b = UNKNOWN) : basetype(b) {}
  TypeDesc(const std::string& s) : basetype(UNKNOWN) {
    if (s == "float") basetype = FLOAT;
    else if (s == "half") basetype = HALF;
    else if (s == "int") basetype = INT32;
    else if (s == "string") basetype = STRING;
  }
  TypeDesc(const char* s) : TypeDesc(std::string(s)) {}
  bool operator==(const TypeDesc& o) const { return basetype == o.basetype; }
  bool operator!=(const TypeDesc& o) const { return basetype != o.basetype; }
  bool operator==(BASETYPE b) const { return basetype == b; }
};
static const TypeDesc TypeFloat(TypeDesc::FLOAT);
static const TypeDesc TypeHalf(TypeDesc::HALF);
static const TypeDesc TypeInt(TypeDesc::INT32);
static const TypeDesc TypeString(TypeDesc::STRING);

struct ROI {
  ROI() {}
};

inline std::string geterror(bool = true) { return std::string(); }

class ImageSpec {
 public:
  int width = 0, height = 0, nchannels = 0;
  TypeDesc format;
  std::map<std::string, std::string> attrs;

  ImageSpec() {}
  ImageSpec(int w, int h, int c, TypeDesc f = TypeDesc::FLOAT) : width(w), height(h), nchannels(c), format(f) {}

  struct AttrRef {
    ImageSpec* spec;
    const ImageSpec* cspec;
    std::string name;
    template <class T> const T& operator=(const T& v) {
      spec->attrs[name] = to_str(v);
      return v;
    }
    operator std::string() const {
      auto it = cspec->attrs.find(name);
      return it == cspec->attrs.end() ? std::string() : it->second;
    }
    static std::string to_str(const std::string& s) { return s; }
    static std::string to_str(const char* s) { return s; }
    template <class T> static std::string to_str(const T& v) { return std::to_string(v); }
  };
  AttrRef operator[](const std::string& n) { return AttrRef{this, this, n}; }
  AttrRef operator[](const std::string& n) const { return AttrRef{nullptr, this, n}; }

  void attribute(const std::string& n, TypeDesc, const std::string& v) { attrs[n] = v; }
  void attribute(const std::string& n, const std::string& v) { attrs[n] = v; }
  bool getattribute(const std::string& n, TypeDesc t, void* out) const {
    auto it = attrs.find(n);
    if (it == attrs.end()) return false;
    if (t == TypeDesc::FLOAT) *(float*)out = std::stof(it->second);
    else if (t == TypeDesc::INT32) *(int*)out = std::stoi(it->second);
    else return false;
    return true;
  }
  std::string get_string_attribute(const std::string& n, const std::string& dflt = std::string()) const {
    auto it = attrs.find(n);
    return it == attrs.end() ? dflt : it->second;
  }
};

// ---- .euf file helpers ------------------------------------------------------------------
namespace eushim {
struct Raster {
  int w = 0, h = 0, c = 0;
  std::vector<float> px;
};
inline bool read_header(const std::string& fn, int& w, int& h, int& c) {
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return false;
  char magic[4];
  int32_t hdr[3];
  bool ok = std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "EUF1", 4) == 0 && std::fread(hdr, 4, 3, f) == 3;
  std::fclose(f);
  if (!ok) return false;
  w = hdr[0]; h = hdr[1]; c = hdr[2];
  return true;
}
// accumulated wall time spent reading input files / writing output files, so that a harness can
// subtract file I/O from the reference's own timers (its "frame rendering time" includes
// save_array, envutil_payload.cc:476-555)
inline double& read_ms() { static double ms = 0; return ms; }
inline double& write_ms() { static double ms = 0; return ms; }
struct ReadTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~ReadTimer() { read_ms() += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};
inline bool read_raster(const std::string& fn, Raster& r) {
  ReadTimer timer;
  FILE* f = std::fopen(fn.c_str(), "rb");
  if (!f) return false;
  char magic[4];
  int32_t hdr[3];
  bool ok = std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "EUF1", 4) == 0 && std::fread(hdr, 4, 3, f) == 3;
  if (ok) {
    r.w = hdr[0]; r.h = hdr[1]; r.c = hdr[2];
    r.px.resize(std::size_t(r.w) * r.h * r.c);
    ok = std::fread(r.px.data(), 4, r.px.size(), f) == r.px.size();
  }
  std::fclose(f);
  return ok;
}
}  // namespace eushim

class ImageInput {
 public:
  ImageSpec m_spec;
  std::string m_name;
  static std::unique_ptr<ImageInput> open(const std::string& fn, const ImageSpec* = nullptr) {
    int w, h, c;
    if (!eushim::read_header(fn, w, h, c)) return nullptr;
    std::unique_ptr<ImageInput> p(new ImageInput);
    p->m_spec = ImageSpec(w, h, c, TypeDesc::FLOAT);
    p->m_name = fn;
    return p;
  }
  const ImageSpec& spec() const { return m_spec; }
  bool close() { return true; }
  bool supports(const std::string& what) const { return what == "scanlines"; }
  bool read_scanlines(int, int, int ybegin, int yend, int, int chbegin, int chend, TypeDesc, void* data,
                      stride_t xstride = AutoStride, stride_t ystride = AutoStride) {
    eushim::Raster r;
    if (!eushim::read_raster(m_name, r)) return false;
    int nc = chend - chbegin;
    if (xstride == AutoStride) xstride = nc * 4;
    if (ystride == AutoStride) ystride = xstride * r.w;
    for (int y = ybegin; y < yend; y++)
      for (int x = 0; x < r.w; x++)
        for (int ch = 0; ch < nc; ch++)
          *(float*)((char*)data + (y - ybegin) * ystride + x * xstride + ch * 4) =
              r.px[(std::size_t(y) * r.w + x) * r.c + chbegin + ch];
    return true;
  }
  bool read_image(int s, int m, int chbegin, int chend, TypeDesc t, void* data, stride_t xstride = AutoStride,
                  stride_t ystride = AutoStride, stride_t = AutoStride) {
    return read_scanlines(s, m, 0, m_spec.height, 0, chbegin, chend, t, data, xstride, ystride);
  }
};

class ImageOutput {
 public:
  static std::unique_ptr<ImageOutput> create(const std::string&) { return nullptr; }
};

class ImageCache;

}  // namespace OIIO
