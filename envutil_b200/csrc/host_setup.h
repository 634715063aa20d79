// host_setup.h - internal declarations shared by the host-side files of the back-end.
#pragma once
#include "envutil_b200.h"

struct eu_cubemap_metrics_t {  // the members of metrics_t (reference cubemap.h:84-140) we use
  int face_px, n_tiles, section_px, left_frame_px, right_frame_px;
  double model_to_px, px_to_model, section_md, refc_md;
};

int eu_compute_cubemap_metrics(int face_px, double face_fov, int support_min, int tile_px,
                               eu_cubemap_metrics_t* m);
int eu_make_spread_ex(const eu_target_t* t, int n_facets, const eu_facet_t* facets, int spline_degree,
                      int solo, int twine, double twine_width, double twine_density, double twine_sigma,
                      double twine_threshold, int twine_max, eu_tap_t* taps, int max_taps, int* twine_out);

// 0/1 plane of a facet with PTO exclude masks / lens crop, before feathering (reference
// environment.h:711-790, fill_polygon envutil_basic.cc:236-321): 1 = keep, 0 = excluded
void eu_build_alpha_mask(const eu_facet_t* f, const eu_alpha_spec_t* a, unsigned char* plane);
