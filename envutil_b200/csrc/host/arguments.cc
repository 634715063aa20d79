// arguments.cc - arguments::init and arguments::twine_setup for the B200 host: the reference's
// option table (envutil_main.cc:190-372), PTO ingestion (:522-905), free facets (:935-976),
// Eev -> brighten (:1006-1061), channel-count rule (:1063-1154), target set-up (:1171-1232).
// Arithmetic follows the reference's types: command-line angles and hfov are parsed as FLOAT
// (ap[...].get<float>) and converted to radians in double; PTO values are parsed as double.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

#include "envutil_host.h"

namespace eu_host {

const char* const projection_name[] = {"spherical", "cylindrical", "rectilinear", "stereographic",
                                       "fisheye",   "cubemap",     "biatan6",     "none"};
arguments args;

namespace {

struct option {
  const char* name;
  int nvalues;  // 0 = flag
};
// the reference's options (envutil_main.cc:190-372) plus the two README aliases
// and the back-end switches
const option option_table[] = {
    {"-v", 0}, {"--output", 1}, {"--projection", 1}, {"--hfov", 1}, {"--width", 1}, {"--height", 1},
    {"--support_min", 1}, {"--tile_size", 1}, {"--synopsis", 1}, {"--working_colour_space", 1},
    {"--output_colour_space", 1}, {"--single", 1}, {"--split", 1}, {"--yaw", 1}, {"--pitch", 1}, {"--roll", 1},
    {"--x0", 1}, {"--x1", 1}, {"--y0", 1}, {"--y1", 1}, {"--brighten", 1}, {"--prefilter", 1}, {"--degree", 1},
    {"--spline_degree", 1}, {"--twine", 1}, {"--twf_file", 1}, {"--twine_normalize", 0}, {"--twine_precise", 0},
    {"--twine_width", 1}, {"--twine_density", 1}, {"--twine_sigma", 1}, {"--twine_threshold", 1},
    {"--twine_max", 1}, {"--photo", 1}, {"--facet", 6}, {"--input", 1}, {"--oiio", 1},
    {"--input_colour_space", 1}, {"--pto", 1}, {"--pto_line", 1}, {"--solo", 1}, {"--mask_for", 1},
    {"--nchannels", 1},
    // back-end
    {"--device", 1}, {"--padded", 0}, {"--plain_texels", 0}, {"--no_tiles", 0}, {"--contracted", 0}, {"--screen_out", 1}, {"--dry_run", 0}};

double glean(const std::string& s) { return s.empty() ? 0.0 : std::stod(s); }
int iglean(const std::string& s) { return s.empty() ? 0 : std::stoi(s); }

int projection_from_name(const std::string& s) {
  int prj = 0;
  for (; prj < 7; prj++)
    if (s == projection_name[prj]) break;
  return prj;  // 7 = PRJ_NONE, as the reference's loop leaves it
}

}  // namespace

int arguments::init(int argc, const char** argv) {
  *this = arguments();
  std::map<std::string, std::string> one;        // last value of single-valued options
  std::vector<std::vector<std::string>> facets;  // --facet IMAGE PROJECTION HFOV YAW PITCH ROLL
  std::vector<std::string> inputs, photos;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    const option* o = nullptr;
    for (const auto& cand : option_table)
      if (a == cand.name) o = &cand;
    if (!o) {
      error = "unknown argument '" + a + "'";
      return EU_ERR_ARGUMENT;
    }
    if (i + o->nvalues > argc - 1) {
      error = "argument '" + a + "' needs " + std::to_string(o->nvalues) + " value(s)";
      return EU_ERR_ARGUMENT;
    }
    if (o->nvalues == 0) {
      one[a] = "1";
    } else if (a == "--facet") {
      facets.emplace_back(argv + i + 1, argv + i + 7);
    } else if (a == "--pto_line") {
      addenda.push_back(argv[i + 1]);
    } else if (a == "--input") {
      inputs.push_back(argv[i + 1]);
    } else if (a == "--photo") {
      photos.push_back(argv[i + 1]);
    } else {
      one[a] = argv[i + 1];
    }
    i += o->nvalues;
  }
  auto has = [&](const char* k) { return one.count(k) != 0; };
  auto str = [&](const char* k, const char* d) { return has(k) ? one[k] : std::string(d); };
  auto getf = [&](const char* k, float d) { return has(k) ? std::strtof(one[k].c_str(), nullptr) : d; };
  auto geti = [&](const char* k, int d) { return has(k) ? std::atoi(one[k].c_str()) : d; };

  verbose = has("-v");
  output = str("--output", "");
  pto_file = str("--pto", "");
  twf_file = str("--twf_file", "");
  split = str("--split", "");
  synopsis = str("--synopsis", "panorama");
  prefilter_degree = geti("--prefilter", -1);
  spline_degree = has("--spline_degree") ? geti("--spline_degree", 1) : geti("--degree", 1);
  twine = geti("--twine", -1);
  twine_width = getf("--twine_width", 1.0f);
  twine_density = getf("--twine_density", 1.0f);
  twine_sigma = getf("--twine_sigma", 0.0f);
  twine_threshold = getf("--twine_threshold", 0.0f);
  twine_max = geti("--twine_max", 8);
  twine_normalize = has("--twine_normalize");
  twine_precise = has("--twine_precise");
  t.x0 = getf("--x0", 0.0f);
  t.x1 = getf("--x1", 0.0f);
  t.y0 = getf("--y0", 0.0f);
  t.y1 = getf("--y1", 0.0f);
  t.width = geti("--width", 0);
  t.height = geti("--height", 0);
  t.hfov = getf("--hfov", 90.0f);
  tile_size = geti("--tile_size", 64);
  support_min = geti("--support_min", 8);
  if (t.hfov != 0.0) t.x0 = t.x1 = t.y0 = t.y1 = 0;
  t.yaw = getf("--yaw", 0.0f);
  t.pitch = getf("--pitch", 0.0f);
  t.roll = getf("--roll", 0.0f);
  brighten = getf("--brighten", 1.0f);
  projection_str = str("--projection", "rectilinear");
  device = geti("--device", 0);
  padded = has("--padded");
  no_tiles = has("--no_tiles");
  plain_texels = has("--plain_texels");
  contracted = has("--contracted");
  screen_out = str("--screen_out", "");
  dry_run = has("--dry_run");
  if (prefilter_degree < 0) prefilter_degree = spline_degree;
  t.projection = projection_from_name(projection_str);
  if (t.projection >= EU_PRJ_NONE) {
    error = "unknown projection '" + projection_str + "'";
    return EU_ERR_ARGUMENT;
  }
  if (pto_file.empty() && addenda.empty() && facets.empty() && inputs.empty() && photos.empty()) {
    error = "no facets: give --facet, --input, --pto or --pto_line";
    return EU_ERR_ARGUMENT;
  }
  if (output.empty() && split.empty()) {
    error = "--output is mandatory";
    return EU_ERR_ARGUMENT;
  }
  // --twine_precise is accepted and ignored, as in the reference: only environment9 reads it
  // (environment.h:1997), which dispatch::payload never instantiates
  mask_for = geti("--mask_for", -1);  // envutil_main.cc:999-1001

  bool ignore_p_line = false;
  solo = -1;
  if (t.width == 0) t.width = 1024;
  else ignore_p_line = true;
  if (t.projection == EU_CUBEMAP || t.projection == EU_BIATAN6) {
    t.height = 6 * t.width;
    if (!(t.hfov >= 90.0)) {  // assert ( hfov >= 90.0 ), still in degrees here
      error = "cubemap targets need hfov >= 90";
      return EU_ERR_ARGUMENT;
    }
  }
  if (t.projection == EU_SPHERICAL && t.height == 0) {
    if (t.width & 1) ++t.width;
    t.height = t.width / 2;
  }
  if (t.height == 0) t.height = t.width;

  bool p_line_present = false;
  int p_line_projection = EU_PRJ_NONE, p_line_width = 0, p_line_height = 0;
  double p_line_hfov = 0.0, p_line_eev = 0.0;
  float eev_sum = 0.0f;
  int eev_count = 0;
  nfacets = 0;

  // ---- PTO (envutil_main.cc:522-905) ----------------------------------------------------
  if (!pto_file.empty() || !addenda.empty()) {
    std::vector<pto_line> lines;
    if (!read_pto_file(pto_file, addenda, lines, error)) return EU_ERR_ARGUMENT;
    if (!ignore_p_line) {
      for (const auto& ln : lines) {
        if (ln.head != 'p') continue;
        p_line_present = true;
        int prj = std::stoi(ln.get("f").empty() ? "0" : ln.get("f"));
        static const int map_p[5] = {EU_RECTILINEAR, EU_CYLINDRICAL, EU_SPHERICAL, EU_FISHEYE, EU_STEREOGRAPHIC};
        p_line_projection = (prj >= 0 && prj <= 4) ? map_p[prj] : EU_PRJ_NONE;
        p_line_width = iglean(ln.get("w"));
        p_line_height = iglean(ln.get("h"));
        p_line_hfov = (M_PI / 180.0) * glean(ln.get("v"));
        p_line_eev = glean(ln.get("Eev"));
        if (!ln.get("S").empty()) {  // store_cropped, envutil_main.cc:615-627
          int v[4];
          if (sscanf(ln.get("S").c_str(), "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) != 4) {
            error = "bad S clause '" + ln.get("S") + "' in the p-line";
            return EU_ERR_ARGUMENT;
          }
          store_cropped = true;
          p_crop_x0 = v[0]; p_crop_x1 = v[1]; p_crop_y0 = v[2]; p_crop_y1 = v[3];
        }
        break;  // additional p-lines are ignored
      }
    }
    for (const auto& ln : lines) {
      if (ln.head != 'i') continue;
      facet_spec fs;
      fs.facet_no = nfacets++;
      if (!ln.get("Pano").empty()) {
        // 'Pano' clause, envutil_main.cc:673-712: this facet IS the stitched panorama - it takes the p-line's
        // projection, field of view and crop, and becomes the solo facet the others are re-created from
        // (--split, "unstitching")
        if (!p_line_present) {
          error = "an i-line with a Pano clause needs a p-line";
          return EU_ERR_ARGUMENT;
        }
        fs.filename = ln.get("Pano");
        if (!fs.filename.empty() && fs.filename[0] == '"') fs.filename = fs.filename.substr(1, fs.filename.size() - 2);
        fs.asset_key = fs.filename;
        fs.f.projection = p_line_projection;
        fs.f.hfov = p_line_hfov;
        int w, h, c;
        if (!read_raster_header(fs.filename, w, h, c)) {
          error = "failed to open facet image '" + fs.filename + "'";
          return EU_ERR_ARGUMENT;
        }
        fs.f.width = w;
        fs.f.height = h;
        fs.f.nchannels = c;
        if (store_cropped) {  // the file holds the crop window of the p-line's panorama
          if (p_crop_x1 - p_crop_x0 != w || p_crop_y1 - p_crop_y0 != h) {
            error = "the Pano image must have the size of the p-line's crop window";
            return EU_ERR_ARGUMENT;
          }
          fs.f.width = p_line_width;
          fs.f.height = p_line_height;
          fs.f.window_x_offset = p_crop_x0;
          fs.f.window_y_offset = p_crop_y0;
          fs.f.window_width = w;
          fs.f.window_height = h;
          if (fs.f.width < p_crop_x1 || fs.f.height < p_crop_y1) {
            error = "the p-line's crop window does not lie inside its panorama";
            return EU_ERR_ARGUMENT;
          }
        }
        solo = fs.facet_no;
      } else {
        fs.filename = ln.get("n");
        if (!fs.filename.empty() && fs.filename[0] == '"') fs.filename = fs.filename.substr(1, fs.filename.size() - 2);
        fs.asset_key = fs.filename;
        int prj = iglean(ln.get("f"));
        if (prj == 0) fs.f.projection = EU_RECTILINEAR;
        else if (prj == 1) fs.f.projection = EU_CYLINDRICAL;
        else if (prj == 2 || prj == 3) fs.f.projection = EU_FISHEYE;
        else if (prj == 4) fs.f.projection = EU_SPHERICAL;
        else if (prj == 10) fs.f.projection = EU_STEREOGRAPHIC;
        else {
          error = "can't handle PTO projection code " + std::to_string(prj) + " in i-line";
          return EU_ERR_ARGUMENT;
        }
        int w, h, c;
        if (!read_raster_header(fs.filename, w, h, c)) {
          error = "failed to open facet image '" + fs.filename + "'";
          return EU_ERR_ARGUMENT;
        }
        fs.f.width = w;
        fs.f.height = h;
        fs.f.nchannels = c;
        fs.f.hfov = (M_PI / 180.0) * std::stod(ln.get("v"));
        {  // 'W x0,x1,y0,y1': the file holds a window of an image of w x h pixels (envutil_main.cc:754-786)
          const std::string& win = ln.get("W");
          if (!win.empty()) {
            int v[4];
            if (sscanf(win.c_str(), "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) != 4) {
              error = "bad W clause '" + win + "'";
              return EU_ERR_ARGUMENT;
            }
            fs.f.window_x_offset = v[0];
            fs.f.window_y_offset = v[2];
            fs.f.window_width = v[1] - v[0];
            fs.f.window_height = v[3] - v[2];
            if (fs.f.window_width != w || fs.f.window_height != h) {
              error = "the W window must have the size of the image file";
              return EU_ERR_ARGUMENT;
            }
            fs.f.width = iglean(ln.get("w"));
            fs.f.height = iglean(ln.get("h"));
            if (fs.f.width == 0 || fs.f.height == 0) {
              error = "a W window needs the total size (w, h)";
              return EU_ERR_ARGUMENT;
            }
          }
        }
      }
      fs.projection_str = projection_name[fs.f.projection];
      fs.f.yaw = (M_PI / 180.0) * glean(ln.get("y"));
      fs.f.pitch = (M_PI / 180.0) * glean(ln.get("p"));
      fs.f.roll = (M_PI / 180.0) * glean(ln.get("r"));
      fs.f.tr_x = glean(ln.get("TrX"));
      fs.f.tr_y = glean(ln.get("TrY"));
      fs.f.tr_z = -glean(ln.get("TrZ"));
      fs.f.tp_y = (M_PI / 180.0) * glean(ln.get("Tpy"));
      fs.f.tp_p = (M_PI / 180.0) * glean(ln.get("Tpp"));
      fs.f.tp_r = 0.0;
      fs.f.shear_g = glean(ln.get("g")) / fs.f.height;
      fs.f.shear_t = glean(ln.get("t")) / fs.f.width;
      fs.f.a = glean(ln.get("a"));
      fs.f.b = glean(ln.get("b"));
      fs.f.c = glean(ln.get("c"));
      fs.f.h = glean(ln.get("d"));
      fs.f.v = glean(ln.get("e"));
      int rc = eu_facet_prepare(&fs.f);  // get_step, get_extent, process_geometry
      if (rc) {
        error = "bad facet geometry in i-line " + std::to_string(fs.facet_no);
        return rc;
      }
      fs.brighten = (float)glean(ln.get("Eev"));
      if (fs.brighten != 0.0f) {
        eev_sum += fs.brighten;
        eev_count++;
      }
      fs.native_nchannels = fs.f.nchannels;
      {  // lens crop "S x0,x1,y0,y1" (envutil_main.cc:811-822)
        const std::string& crop = ln.get("S");
        if (!crop.empty()) {
          int v[4];
          if (sscanf(crop.c_str(), "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) != 4) {
            error = "bad S clause '" + crop + "'";
            return EU_ERR_ARGUMENT;
          }
          fs.has_lens_crop = true;
          fs.crop_x0 = v[0]; fs.crop_x1 = v[1]; fs.crop_y0 = v[2]; fs.crop_y1 = v[3];
        }
      }
      facet_spec_v.push_back(fs);
    }
    // k-lines: exclude masks "k iN t0 p"x y x y ..."" (envutil_main.cc:826-904)
    int mask_no = 0;
    for (const auto& ln : lines) {
      if (ln.head != 'k') continue;
      int image = iglean(ln.get("i")), variant = iglean(ln.get("t"));
      if (image < 0 || image >= (int)facet_spec_v.size()) {
        error = "k-line refers to image " + std::to_string(image);
        return EU_ERR_ARGUMENT;
      }
      facet_spec& fct = facet_spec_v[image];
      fct.has_pto_mask = true;
      std::vector<float> xy;
      {  // pairs "number<one blank>number", scanned left to right without overlap
        const std::string& vl = ln.get("p");
        const char* p = vl.c_str();
        auto is_num = [](char ch) { return (ch >= '0' && ch <= '9') || ch == '.' || ch == '+' || ch == '-'; };
        while (*p) {
          while (*p && !is_num(*p)) p++;
          if (!*p) break;
          char* e1;
          double x = strtod(p, &e1);
          if (e1 == p) { p++; continue; }
          if (!isspace((unsigned char)*e1) || !is_num(e1[1])) { p = e1; continue; }
          char* e2;
          double y = strtod(e1 + 1, &e2);
          if (e2 == e1 + 1) { p = e1; continue; }
          xy.push_back((float)x);
          xy.push_back((float)y);
          p = e2;
        }
      }
      if (variant != 0) {
        fprintf(stderr, "warning: mask type not implemented: %d this mask will be ignored\n", variant);
      } else {
        fct.mask_xy.push_back(xy);
      }
      if (fct.filename == fct.asset_key) fct.asset_key += "." + pto_file + ".";
      else fct.asset_key += ".";
      fct.asset_key += std::to_string(mask_no++);
    }
  }

  // ---- free facets (envutil_main.cc:935-976); --input = one facet, projection from the aspect
  for (const auto& in : inputs) {
    int w, h, c;
    if (!read_raster_header(in, w, h, c)) {
      error = "failed to open facet image '" + in + "'";
      return EU_ERR_ARGUMENT;
    }
    if (h == 6 * w) facets.push_back({in, "cubemap", "90", "0", "0", "0"});
    else if (w == 2 * h) facets.push_back({in, "spherical", "360", "0", "0", "0"});
    else {
      error = "--input needs a 2:1 lat/lon or 1:6 cubemap image; use --facet for '" + in + "'";
      return EU_ERR_ARGUMENT;
    }
  }
  // --photo IMAGE (envutil_main.cc:916-927): projection and hfov from the image's metadata, and where
  // there is none - .euf rasters carry none - "rectilinear" and 65 degrees (envutil_basic.h:596-627)
  for (const auto& ph : photos) facets.push_back({ph, "rectilinear", "65", "0", "0", "0"});
  for (const auto& fv : facets) {
    facet_spec fs;
    fs.filename = fv[0];
    fs.projection_str = fv[1];
    fs.f.projection = projection_from_name(fv[1]);
    if (fs.f.projection >= EU_PRJ_NONE) {
      error = "unknown facet projection '" + fv[1] + "'";
      return EU_ERR_ARGUMENT;
    }
    // "%F" fields of facet_spec::init are doubles (envutil_main.cc:113)
    fs.f.hfov = std::strtod(fv[2].c_str(), nullptr);
    fs.f.yaw = std::strtod(fv[3].c_str(), nullptr);
    fs.f.pitch = std::strtod(fv[4].c_str(), nullptr);
    fs.f.roll = std::strtod(fv[5].c_str(), nullptr);
    int w, h, c;
    if (!read_raster_header(fs.filename, w, h, c)) {
      error = "failed to open facet image '" + fs.filename + "'";
      return EU_ERR_ARGUMENT;
    }
    fs.f.width = w;
    fs.f.height = h;
    fs.f.nchannels = c;
    fs.facet_no = nfacets++;
    fs.f.hfov *= M_PI / 180.0;
    fs.f.yaw *= M_PI / 180.0;
    fs.f.pitch *= M_PI / 180.0;
    fs.f.roll *= M_PI / 180.0;
    int rc = eu_facet_prepare(&fs.f);
    if (rc) {
      error = "bad facet geometry for '" + fs.filename + "'";
      return rc;
    }
    fs.asset_key = fs.filename;
    fs.brighten = 0.0f;
    facet_spec_v.push_back(fs);
  }
  if (nfacets == 0) {
    error = "no facets";
    return EU_ERR_ARGUMENT;
  }
  if (solo == -1) solo = geti("--solo", -1);
  if (solo != -1 && solo >= nfacets) {
    error = "--solo is beyond the facet count";
    return EU_ERR_ARGUMENT;
  }
  if (nfacets == 1) solo = 0;
  single = geti("--single", -1);
  if (single != -1 && (single < 0 || single >= nfacets)) {
    error = "--single is beyond the facet count";
    return EU_ERR_ARGUMENT;
  }

  // ---- Eev -> brighten, channel count (envutil_main.cc:1003-1154) -----------------------
  nchannels = 1;
  bool alpha_seen = false;
  if (eev_count > 0) eev_sum /= eev_count;
  if (p_line_eev != 0.0) eev_sum = (float)p_line_eev;
  for (auto& m : facet_spec_v) {
    if (eev_count) {
      if (m.brighten == 0.0f) m.brighten = 1.0f;
      else m.brighten = (float)pow(2.0, m.brighten - eev_sum);
    } else {
      m.brighten = 1.0f;
    }
    if (brighten != 1.0) m.brighten *= brighten;
    m.f.brighten = m.brighten;
    if (m.native_nchannels == 0) m.native_nchannels = m.f.nchannels;
    if (m.has_pto_mask || m.has_lens_crop) {  // masks and crops act through alpha (envutil_main.cc:1065-1069)
      if (m.f.nchannels == 1 || m.f.nchannels == 3) m.f.nchannels++;
    }
    if (m.f.nchannels == 2 || m.f.nchannels == 4) alpha_seen = true;
    if (m.f.nchannels > nchannels) nchannels = m.f.nchannels;
    // --mask_for: facet_spec::masked 1 (white) for that facet, 0 (black) for the others, -1 without the option
    // (envutil_main.cc:1077-1091); eu_facet_t.masked holds it as 2 / 1 / 0
    m.f.masked = mask_for == -1 ? 0 : (m.facet_no == mask_for ? 2 : 1);
  }
  if (mask_for != -1 && mask_for >= (int)facet_spec_v.size()) {
    error = "--mask_for: no such facet";  // the reference asserts (envutil_main.cc:1001)
    return EU_ERR_ARGUMENT;
  }
  if (alpha_seen && nchannels == 3) nchannels = 4;
  int nch = geti("--nchannels", 0);
  if (nch > 0) nchannels = nch;

  // ---- target (envutil_main.cc:1180-1232) -----------------------------------------------
  if (single >= 0) {
    // 'single': the target takes over the facet's geometry (envutil_main.cc:1157-1178)
    // (extent and step follow below, from the same get_extent call)
    int rc2 = take_single(single);
    if (rc2) return rc2;
  } else if (p_line_present) {
    t.hfov = p_line_hfov;
    t.projection = p_line_projection;
    projection_str = projection_name[t.projection];
    t.width = p_line_width;
    t.height = p_line_height;
  } else {
    t.hfov *= M_PI / 180.0;
    t.yaw *= M_PI / 180.0;
    t.pitch *= M_PI / 180.0;
    t.roll *= M_PI / 180.0;
  }
  t.nchannels = nchannels;
  t.step = 0.0;
  if (t.hfov != 0.0) {
    double e[4];
    eu_get_extent(t.projection, t.width, t.height, t.hfov, e);
    t.x0 = e[0];
    t.x1 = e[1];
    t.y0 = e[2];
    t.y1 = e[3];
  }
  if (!(t.x0 <= t.x1) || !(t.y0 <= t.y1) || t.width <= 0 || t.height <= 0) {
    error = "empty target extent";
    return EU_ERR_ARGUMENT;
  }
  t.step = (t.x1 - t.x0) / t.width;
  // a 'single' job stores the whole facet geometry (core(): args.store_cropped = false, envutil_main.cc:1714,1726)
  if (store_cropped && single < 0) {
    t.crop_x0 = p_crop_x0;
    t.crop_y0 = p_crop_y0;
    t.crop_width = p_crop_x1 - p_crop_x0;
    t.crop_height = p_crop_y1 - p_crop_y0;
    if (t.crop_width <= 0 || t.crop_height <= 0 || t.crop_x0 < 0 || t.crop_y0 < 0 || p_crop_x1 > t.width ||
        p_crop_y1 > t.height) {
      error = "p-line crop does not lie inside the target";
      return EU_ERR_ARGUMENT;
    }
  }
  return EU_OK;
}

// arguments::twine_setup, envutil_main.cc:1405-1616
int arguments::twine_setup() {
  twine_spread.clear();
  if (!twf_file.empty()) twine = 1;
  std::vector<eu_facet_t> fv;
  for (const auto& f : facet_spec_v) fv.push_back(f.f);
  eu_opts_t o{};
  o.spline_degree = spline_degree;
  o.solo = solo;
  std::vector<eu_tap_t> taps(EU_HOST_MAX_TAPS);
  int tw = 0;
  int n = eu_make_spread(&t, &o, nfacets, fv.data(), twine, twine_width, twine_density, twine_sigma, twine_threshold,
                         twine_max, taps.data(), (int)taps.size(), &tw);
  if (n < 0) {
    error = "bad twining parameters";
    return n;
  }
  twine = tw;
  if (twf_file.empty()) {
    twine_spread.assign(taps.begin(), taps.begin() + n);
  } else {  // read_twf_file, envutil_main.cc:1360-1403 (twine_width scales the offsets)
    std::ifstream ifs(twf_file);
    if (!ifs.good()) {
      error = "cannot read twf file " + twf_file;
      return EU_ERR_ARGUMENT;
    }
    double sum = 0.0;
    eu_tap_t c;
    while (ifs.good()) {
      ifs >> c.x >> c.y >> c.w;
      if (ifs.eof() && ifs.fail()) break;
      twine_spread.push_back(c);
      sum += c.w;
      if (ifs.eof()) break;
    }
    for (auto& k : twine_spread) {
      k.x *= twine_width;
      k.y *= twine_width;
      if (twine_normalize) k.w /= sum;
    }
  }
  if (twine && twine_spread.empty()) {
    error = "twining is on but the filter is empty";
    return EU_ERR_ARGUMENT;
  }
  if (verbose) {
    printf("final twining filter kernel:\n");
    int ord = 0;
    for (const auto& c : twine_spread) printf("%d\tx:\t%g\ty:\t%g\tw:\t%g\n", ord++, c.x, c.y, c.w);
  }
  return EU_OK;
}

int arguments::take_single(int i) {
  if (i < 0 || i >= (int)facet_spec_v.size()) {
    error = "single facet index out of range";
    return EU_ERR_ARGUMENT;
  }
  const facet_spec& fs = facet_spec_v[i];
  single = i;
  t.single = i + 1;  // the kernels invert the facet's lens correction / translation (tf_ex_facet)
  t.projection = fs.f.projection;
  projection_str = projection_name[t.projection];
  t.width = fs.f.width;
  t.height = fs.f.height;
  t.hfov = fs.f.hfov;
  t.yaw = fs.f.yaw;
  t.pitch = fs.f.pitch;
  t.roll = fs.f.roll;
  t.gain = 0.0;
  if (fs.brighten != 1.0) {  // work(): float unbrighten = 1.0 / fct.brighten
    float unbrighten = 1.0 / fs.brighten;
    t.gain = unbrighten;
  }
  // a 'single' job stores the whole facet (args.store_cropped = false, envutil_main.cc:1714,1726)
  t.crop_x0 = t.crop_y0 = t.crop_width = t.crop_height = 0;
  double e[4];
  eu_get_extent(t.projection, t.width, t.height, t.hfov, e);
  t.x0 = e[0];
  t.x1 = e[1];
  t.y0 = e[2];
  t.y1 = e[3];
  t.step = (t.x1 - t.x0) / t.width;
  return EU_OK;
}

// tokenize, envutil_basic.cc:329-411: blanks separate, single or double quotes group, a
// backslash escapes the active quote character inside a quoted run
std::vector<std::string> tokenize(const std::string& input) {
  std::vector<std::string> result;
  enum { NO_TOKEN, IN_TOKEN, IN_Q } state = NO_TOKEN;
  std::string token;
  char quote = 0;
  for (size_t i = 0; i < input.size(); i++) {
    char ch = input[i];
    bool blank = ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r';
    switch (state) {
      case NO_TOKEN:
        if (blank) break;
        if (ch == '"' || ch == '\'') { state = IN_Q; quote = ch; }
        else { state = IN_TOKEN; token += ch; }
        break;
      case IN_TOKEN:
        if (blank) { result.push_back(token); token.clear(); state = NO_TOKEN; }
        else if (ch == '"' || ch == '\'') { state = IN_Q; quote = ch; }
        else token += ch;
        break;
      case IN_Q:
        if (ch == quote) { state = IN_TOKEN; break; }
        if (ch == '\\' && i + 1 < input.size() && input[i + 1] == quote) ch = input[++i];
        token += ch;
        break;
    }
  }
  if (!token.empty()) result.push_back(token);
  return result;
}

}  // namespace eu_host
