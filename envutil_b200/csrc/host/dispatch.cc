// dispatch.cc - cuda_dispatch: the dispatch_base whose payload() runs on the B200.
// Replaces dispatch::payload -> roll_out -> fuse -> work (reference envutil_payload.cc:2408,
// 2336,1885,425): per facet find-or-stage the source (the asset cache lives in the library),
// render, write the raster, conclude the cycle.
#include <chrono>
#include <cstdint>
#include <cstdio>

#include <algorithm>
#include <vector>

#include "envutil_host.h"

namespace eu_host {

namespace {

struct cuda_dispatch : public dispatch_base {
  int payload(int nchannels, int ninputs, int projection) const override {
    arguments& a = args;
    int rc = eu_init(a.device);
    if (rc) {
      fprintf(stderr, "envutil_b200: %s\n", eu_last_error());
      return rc;
    }
    eu_target_t t = a.t;
    t.projection = projection;
    t.nchannels = nchannels;
    eu_opts_t o{};
    o.spline_degree = a.spline_degree;
    o.prefilter_degree = a.prefilter_degree;
    o.synopsis = a.synopsis == "hdr_merge" ? EU_SYN_HDR_MERGE : EU_SYN_PANORAMA;
    o.solo = a.solo;
    o.support_min = a.support_min;
    o.tile_size = a.tile_size;
    o.reserved[0] = a.padded ? 1 : (a.plain_texels ? 2 : 0);
    o.reserved[1] = (a.no_tiles ? EU_OPT_NO_TILES : 0) | (a.contracted ? EU_OPT_CONTRACTED : 0);
    std::vector<eu_facet_t> fv;
    std::vector<eu_source_h> sv;
    float stage_ms = 0, h2d_ms = 0;
    for (const auto& fs : a.facet_spec_v) {
      fv.push_back(fs.f);
      // one staged source per (file, degree, prefilter, layout): the reference's asset_key is the
      // file name because a process run has one degree; a library that outlives jobs needs more
      std::string key = fs.asset_key + "|" + std::to_string(o.spline_degree) + "|" + std::to_string(o.prefilter_degree) +
                        "|" + std::to_string(o.reserved[0]) + "|" + std::to_string(o.support_min) + "|" +
                        std::to_string(o.tile_size);
      eu_source_h h = eu_source_find(key.c_str());
      if (!h) {
        int w, hh, c;
        std::vector<float> px;
        if (!read_raster(fs.filename, w, hh, c, px) || w != fs.f.window_width || hh != fs.f.window_height ||
            c != fs.native_nchannels) {
          fprintf(stderr, "envutil_b200: cannot read facet image '%s'\n", fs.filename.c_str());
          return EU_ERR_ARGUMENT;
        }
        eu_timing_t tm{};
        if (fs.has_pto_mask || fs.has_lens_crop) {
          std::vector<int32_t> sizes;
          std::vector<float> xy;
          for (const auto& m : fs.mask_xy) {
            sizes.push_back((int32_t)(m.size() / 2));
            xy.insert(xy.end(), m.begin(), m.end());
          }
          eu_alpha_spec_t as{};
          as.native_nchannels = fs.native_nchannels;
          as.has_crop = fs.has_lens_crop ? 1 : 0;
          as.crop_x0 = fs.crop_x0; as.crop_x1 = fs.crop_x1; as.crop_y0 = fs.crop_y0; as.crop_y1 = fs.crop_y1;
          as.n_masks = (int32_t)sizes.size();
          as.mask_sizes = sizes.data();
          as.mask_xy = xy.data();
          rc = eu_source_upload_alpha(key.c_str(), &fv.back(), &o, px.data(), &as, &h, &tm);
        } else {
          rc = eu_source_upload(key.c_str(), &fv.back(), &o, px.data(), &h, &tm);
        }
        if (rc) {
          fprintf(stderr, "envutil_b200: %s\n", eu_last_error());
          return rc;
        }
        stage_ms += tm.render_ms;
        h2d_ms += tm.h2d_ms;
      }
      sv.push_back(h);
    }
    const std::vector<eu_tap_t>& taps = a.twine_spread;
    int n_taps = ninputs == 9 ? (int)taps.size() : 0;  // ninputs == 9 <=> twining (envutil_main.cc:1673)
    // a cropped output (p-line S) has the crop's size (envutil_payload.cc:440-443)
    const int ow = t.crop_width > 0 ? t.crop_width : t.width, oh = t.crop_width > 0 ? t.crop_height : t.height;
    eu_timing_t tm{};
    if (a.tethered || !a.screen_out.empty()) {
      // work()'s tethered branch (envutil_payload.cc:524-531): act + to_screen_t into the viewer's frame buffer
      std::vector<std::uint32_t> own;
      std::uint32_t* frame = a.p_screen_data;
      if (!frame) {
        own.resize((size_t)ow * oh);
        frame = own.data();
      }
      rc = eu_render_screen(&t, &o, (int)fv.size(), fv.data(), sv.data(), taps.data(), n_taps, frame, &tm);
      if (rc) {
        fprintf(stderr, "envutil_b200: %s\n", eu_last_error());
        return rc;
      }
      if (a.verbose)
        printf("frame rendering time: %.3f ms (device), staging %.3f ms, h2d %.3f ms, d2h %.3f ms, %d launches\n", tm.render_ms,
               stage_ms, h2d_ms, tm.d2h_ms, tm.launches);
      if (!a.screen_out.empty()) {
        FILE* fp = fopen(a.screen_out.c_str(), "wb");
        if (!fp || fwrite(frame, sizeof(std::uint32_t), (size_t)ow * oh, fp) != (size_t)ow * oh) {
          fprintf(stderr, "envutil_b200: cannot write '%s'\n", a.screen_out.c_str());
          if (fp) fclose(fp);
          return EU_ERR_ARGUMENT;
        }
        fclose(fp);
      }
      eu_cycle();
      return 0;
    }
    std::vector<float> out((size_t)ow * oh * nchannels);
    rc = eu_render(&t, &o, (int)fv.size(), fv.data(), sv.data(), taps.data(), n_taps, out.data(), &tm);
    if (rc) {
      fprintf(stderr, "envutil_b200: %s\n", eu_last_error());
      return rc;
    }
    if (a.verbose) {
      // the reference prints wall-clock "frame rendering time" (envutil_payload.cc:555)
      printf("frame rendering time: %.3f ms (device), staging %.3f ms, h2d %.3f ms, d2h %.3f ms, %d launches\n",
             tm.render_ms, stage_ms, h2d_ms, tm.d2h_ms, tm.launches);
    }
    if (!write_raster(a.output, ow, oh, nchannels, out.data())) {
      fprintf(stderr, "envutil_b200: cannot write '%s'\n", a.output.c_str());
      return EU_ERR_ARGUMENT;
    }
    eu_cycle();  // conclude_cycle(), envutil_payload.cc:2433
    return 0;
  }
};

}  // namespace

const dispatch_base* get_dispatch() {
  static cuda_dispatch d;
  return &d;
}

// core(), envutil_main.cc:1634-1733
int core(int argc, const char** argv) {
  int rc = args.init(argc, argv);
  if (rc) {
    fprintf(stderr, "envutil_b200: %s\n", args.error.c_str());
    return rc;
  }
  const dispatch_base* dp = get_dispatch();
  if (args.verbose) printf("using %s ISA\n", dp->hwy_target_name);
  rc = args.twine_setup();
  if (rc) {
    fprintf(stderr, "envutil_b200: %s\n", args.error.c_str());
    return rc;
  }
  int nch = args.nchannels;
  int ninp = (args.twine == 0) ? 3 : 9;
  if (args.dry_run) {  // print what the kernels would be given; no GPU needed
    const eu_target_t& t = args.t;
    printf("target %s %dx%d nch %d hfov %.17g yaw %.17g pitch %.17g roll %.17g\n", projection_name[t.projection], t.width,
           t.height, nch, t.hfov, t.yaw, t.pitch, t.roll);
    printf("extent %.17g %.17g %.17g %.17g step %.17g\n", t.x0, t.x1, t.y0, t.y1, t.step);
    if (t.crop_width > 0) printf("crop %dx%d+%d+%d\n", t.crop_width, t.crop_height, t.crop_x0, t.crop_y0);
    printf("degree %d prefilter %d twine %d ninputs %d synopsis %s solo %d\n", args.spline_degree, args.prefilter_degree,
           args.twine, ninp, args.synopsis.c_str(), args.solo);
    for (const auto& f : args.facet_spec_v)
      printf("facet %d %s %s %dx%dx%d hfov %.17g ypr %.17g %.17g %.17g step %.17g brighten %.9g lcp %d shift %.17g %.17g "
             "shear %.17g %.17g masked %d\n",
             f.facet_no, f.filename.c_str(), projection_name[f.f.projection], f.f.width, f.f.height, f.f.nchannels,
             f.f.hfov, f.f.yaw, f.f.pitch, f.f.roll, f.f.step, f.brighten, f.f.has_lcp, f.f.shift_h, f.f.shift_v,
             f.f.shear_g, f.f.shear_t, f.f.masked);
    for (const auto& c : args.twine_spread) printf("tap %.9g %.9g %.9g\n", c.x, c.y, c.w);
    return 0;
  }
  if (!args.split.empty()) {  // one 'single' job per facet, envutil_main.cc:1679-1721
    // image_series (envutil_basic.h:212-263): one '%' makes the name a printf format for the facet number
    const bool is_format = std::count(args.split.begin(), args.split.end(), '%') == 1;
    for (int i = 0; i < args.nfacets; i++) {
      if (i == args.solo) continue;  // the solo facet is what the others are re-created from
      rc = args.take_single(i);
      if (rc) {
        fprintf(stderr, "envutil_b200: %s\n", args.error.c_str());
        return rc;
      }
      if (is_format) {
        std::vector<char> buffer(args.split.size() + 16);
        snprintf(buffer.data(), buffer.size(), args.split.c_str(), i);
        args.output = buffer.data();
      } else {
        args.output = args.split;
      }
      rc = dp->payload(nch, ninp, args.t.projection);
      if (rc) return rc;
    }
    return 0;
  }
  return dp->payload(nch, ninp, args.t.projection);
}

}  // namespace eu_host
