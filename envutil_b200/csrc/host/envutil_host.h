// envutil_host.h - C++ host above the C ABI: envutil's command-line / PTO surface and its
// dispatch boundary, restated for the B200 back-end.
//
// Mirrors, with the same names and the same argument meaning:
//   struct arguments (+ facet_spec)          reference envutil_basic.h:432-705
//   arguments::init / arguments::twine_setup reference envutil_main.cc:178-1251,1405-1616
//   struct dispatch_base, get_dispatch()      reference envutil_dispatch.h:50-74
//   core(), pipe mode                         reference envutil_main.cc:1634-1733,1948-1982
// The per-pixel work behind payload() is not here: cuda_dispatch::payload marshals `args` into
// the POD structs of include/envutil_b200.h and calls libenvutil_b200.so.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "envutil_b200.h"

namespace eu_host {

constexpr int EU_HOST_MAX_TAPS = 1024;
extern const char* const projection_name[];  // envutil_basic.h:111-121

// facet_spec: the C-ABI POD plus what only the host needs
struct facet_spec {
  eu_facet_t f{};
  int facet_no = 0;
  std::string filename, asset_key, projection_str;
  float brighten = 0.0f;  // Eev while parsing, linear gain after init (envutil_main.cc:1030-1061)
  bool has_lens_crop = false, has_pto_mask = false;
  int crop_x0 = 0, crop_x1 = 0, crop_y0 = 0, crop_y1 = 0;  // i-line S clause
  std::vector<std::vector<float>> mask_xy;                  // k-lines t0: x y x y ... per polygon
  int native_nchannels = 0;                                 // channels of the file (f.nchannels may be +1)
};

struct arguments {
  // target (facet_base part of the reference's `arguments`)
  eu_target_t t{};
  std::string projection_str;
  // job
  bool verbose = false;
  std::string output, pto_file, synopsis = "panorama", twf_file, split;
  int prefilter_degree = -1, spline_degree = 1, twine = -1, twine_max = 8;
  bool twine_normalize = false, twine_precise = false;
  double twine_width = 1.0, twine_density = 1.0, twine_sigma = 0.0, twine_threshold = 0.0;
  std::vector<eu_tap_t> twine_spread;
  int support_min = 8, tile_size = 64;
  int nchannels = 0, nfacets = 0, solo = -1, single = -1, mask_for = -1;
  bool store_cropped = false;  // p-line S clause (envutil_basic.h:684-687)
  int p_crop_x0 = 0, p_crop_x1 = 0, p_crop_y0 = 0, p_crop_y1 = 0;
  float brighten = 1.0f;
  std::vector<facet_spec> facet_spec_v;
  std::vector<std::string> addenda;
  // back-end options (not in the reference)
  // back-end options (not part of envutil's surface): --padded / --plain_texels force the 16- / 12-byte RGB texel layout
  // (default: the library's rule), --no_tiles the direct-gather kernels, --contracted the fused multiply-add arithmetic
  bool padded = false, plain_texels = false, no_tiles = false, contracted = false, dry_run = false;
  int device = 0;
  // tethered output (reference envutil_basic.h:637-638, envutil_main.cc:1650,1791): payload() stores one uint32 sRGBA
  // value per pixel into p_screen_data (the viewer's frame buffer) instead of writing an image file. --screen_out FILE
  // (back-end option) runs a job that way without a viewer and writes the buffer to FILE as raw little-endian uint32.
  bool tethered = false;
  std::uint32_t* p_screen_data = nullptr;
  std::string screen_out;

  // parse the command line the reference's way; returns 0 or a negative eu_status_t, with the
  // reason in `error` (the reference asserts / exit(-1)s instead)
  int init(int argc, const char** argv);
  int twine_setup();
  // (facet_base&) args = facet_spec_v[i]; args.single = i - the target takes over facet i's geometry
  // (envutil_main.cc:1157-1178 and the --split loop, :1679-1721)
  int take_single(int i);
  std::string error;
};

extern arguments args;  // the job description is a global, as in the reference (envutil_main.cc:176)

struct dispatch_base {  // reference envutil_dispatch.h:50-66
  const char* hwy_target_name = "sm_100a";
  virtual int payload(int nchannels, int ninputs, int projection) const = 0;
  virtual ~dispatch_base() {}
};
const dispatch_base* get_dispatch();  // reference envutil_dispatch.h:71-74

int core(int argc, const char** argv);              // one job        (envutil_main.cc:1634)
std::vector<std::string> tokenize(const std::string& line);  // pipe mode (envutil_basic.cc:329)

// raster files: the ".euf" float container (see envutil_b200/euf.py); OpenImageIO is not linked
bool read_raster_header(const std::string& fn, int& w, int& h, int& c);
bool read_raster(const std::string& fn, int& w, int& h, int& c, std::vector<float>& px);
bool write_raster(const std::string& fn, int w, int h, int c, const float* px);

// PTO subset (reference pto.h:82-240)
struct pto_line {
  char head = 0;
  std::vector<std::pair<std::string, std::string>> fields;
  const std::string& get(const std::string& key) const;
};
bool parse_pto_line(const std::string& s, std::vector<pto_line>& lines);
bool read_pto_file(const std::string& fn, const std::vector<std::string>& addenda, std::vector<pto_line>& lines,
                   std::string& err);

}  // namespace eu_host
