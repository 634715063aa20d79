// Minimal stand-in for <OpenImageIO/imagebuf.h> (test infrastructure only).
#pragma once
#include "imageio.h"
#include <cstdlib>
namespace OIIO {
class ImageBuf {
 public:
  ImageSpec m_spec;
  char* m_wrapped = nullptr;  // application memory (not owned)
  stride_t m_xs = 0, m_ys = 0;
  eushim::Raster m_own;  // pixel data read from a file
  bool m_from_file = false;

  // wrap application memory (envutil_basic.h:771-774, :917-921)
  ImageBuf(const ImageSpec& spec, void* buffer, stride_t xstride = AutoStride, stride_t ystride = AutoStride)
      : m_spec(spec), m_wrapped((char*)buffer), m_xs(xstride), m_ys(ystride) {}
  // read from file (envutil_basic.h:928)
  ImageBuf(const std::string& name, int = 0, int = 0, ImageCache* = nullptr, const ImageSpec* = nullptr) {
    m_from_file = true;
    if (!eushim::read_raster(name, m_own)) {
      std::cerr << "eushim: cannot read '" << name << "'" << std::endl;
      std::exit(-1);
    }
    m_spec = ImageSpec(m_own.w, m_own.h, m_own.c, TypeDesc::FLOAT);
  }
  bool init_spec(const std::string& name, int, int) {
    int w, h, c;
    if (!eushim::read_header(name, w, h, c)) return false;
    m_spec.width = w; m_spec.height = h; m_spec.nchannels = c; m_spec.format = TypeDesc::FLOAT;
    return true;
  }
  const ImageSpec& spec() const { return m_spec; }
  int nchannels() const { return m_spec.nchannels; }
  // copy pixels of src into the wrapped memory; channels beyond the memory's pixel stride are dropped
  bool copy(const ImageBuf& src, TypeDesc = TypeDesc::UNKNOWN) {
    if (!m_wrapped || !src.m_from_file) return false;
    int w = src.m_own.w, h = src.m_own.h, c = src.m_own.c;
    int cmax = int(m_xs / 4) < c ? int(m_xs / 4) : c;
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++)
        for (int ch = 0; ch < cmax; ch++)
          *(float*)(m_wrapped + y * m_ys + x * m_xs + ch * 4) = src.m_own.px[(std::size_t(y) * w + x) * c + ch];
    m_spec.width = w; m_spec.height = h; m_spec.nchannels = c; m_spec.format = TypeDesc::FLOAT;
    return true;
  }
  bool get_pixels(ROI, TypeDesc, void* result, stride_t xstride = AutoStride, stride_t ystride = AutoStride) const {
    if ((char*)result == m_wrapped && xstride == m_xs && ystride == m_ys) return true;  // in place
    return false;
  }
  bool write(const std::string& filename) const {
    auto t0 = std::chrono::steady_clock::now();
    bool ok = true;
    if (!std::getenv("EUSHIM_NOWRITE")) {
      FILE* f = std::fopen(filename.c_str(), "wb");
      if (!f) return false;
      int32_t hdr[3] = {m_spec.width, m_spec.height, m_spec.nchannels};
      std::fwrite("EUF1", 1, 4, f);
      std::fwrite(hdr, 4, 3, f);
      std::size_t rowb = std::size_t(m_spec.width) * m_spec.nchannels * 4;
      if (m_xs == stride_t(m_spec.nchannels * 4)) {
        for (int y = 0; y < m_spec.height; y++) ok = ok && std::fwrite(m_wrapped + y * m_ys, 1, rowb, f) == rowb;
      } else {
        for (int y = 0; y < m_spec.height; y++)
          for (int x = 0; x < m_spec.width; x++)
            ok = ok && std::fwrite(m_wrapped + y * m_ys + x * m_xs, 4, m_spec.nchannels, f) == std::size_t(m_spec.nchannels);
      }
      std::fclose(f);
    }
    eushim::write_ms() += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "eushim: write time " << eushim::write_ms() << " ms (cumulated)" << std::endl;
    std::cout << "eushim: read time " << eushim::read_ms() << " ms (cumulated)" << std::endl;
    return ok;
  }
};
}  // namespace OIIO
