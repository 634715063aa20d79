/* envutil_b200.h - C ABI of the B200 back-end for envutil's per-pixel reprojection path.
 *
 * Drop-in boundary: the reference enters its hot path through
 *     virtual int dispatch_base::payload(int nchannels, int ninputs, projection_t) const
 * (reference envutil_dispatch.h:63-65, called from core(), envutil_main.cc:1720,1727), with
 * the job description in the global `arguments args` (envutil_basic.h:633-705). This header
 * is what a `cuda_dispatch : dispatch_base` adapter marshals `args` into (INTEGRATION.md
 * shows the adapter). Plain C: pointers, sizes and POD structs only.
 *
 * All angles are radians (the reference converts degrees right after parsing,
 * envutil_main.cc:948-951,1199-1202). Rasters are interleaved float32, row-major, top row
 * first - the layout of the reference's zimt::array_t targets (envutil_payload.cc:539) and of
 * the buffers it hands to OIIO (envutil_basic.h:760-775).
 */
#ifndef ENVUTIL_B200_H
#define ENVUTIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* projection_t, reference envutil_basic.h:99-109 (same numeric values) */
typedef enum {
  EU_SPHERICAL = 0,
  EU_CYLINDRICAL = 1,
  EU_RECTILINEAR = 2,
  EU_STEREOGRAPHIC = 3,
  EU_FISHEYE = 4,
  EU_CUBEMAP = 5,
  EU_BIATAN6 = 6,
  EU_PRJ_NONE = 7
} eu_projection_t;

/* --synopsis, reference envutil_payload.cc:2302-2315 */
typedef enum { EU_SYN_PANORAMA = 0, EU_SYN_HDR_MERGE = 1 } eu_synopsis_t;

/* status codes (the reference asserts/exit(-1)s instead; the library never exits) */
typedef enum {
  EU_OK = 0,
  EU_ERR_ARGUMENT = -1,
  EU_ERR_UNSUPPORTED = -2,
  EU_ERR_CUDA = -3,
  EU_ERR_NO_DEVICE = -4,
  EU_ERR_STATE = -5
} eu_status_t;

/* Mirror of facet_base + the facet_spec members the hot path reads (reference
 * envutil_basic.h:432-543). The first block is input; the second block is DERIVED and is
 * filled by eu_facet_prepare()/eu_target_prepare() with the reference's own arithmetic
 * (get_extent/get_step envutil_basic.cc:112-226, process_geometry envutil_basic.h:499-543). */
typedef struct eu_facet {
  /* input */
  int32_t projection;     /* eu_projection_t */
  int32_t width, height;  /* total size in pixels */
  int32_t nchannels;      /* channels of the raster (1..4); RGB = 3 */
  double hfov;            /* horizontal field of view */
  double yaw, pitch, roll;
  double tr_x, tr_y, tr_z; /* PanoTools translation (TrZ already negated, envutil_main.cc:787) */
  double tp_y, tp_p, tp_r; /* translation plane orientation */
  double shear_g, shear_t;
  double a, b, c;          /* lens polynomial */
  double h, v;             /* PTO d/e shift, in pixels (never modified by the library) */
  double brighten;         /* linear gain applied after interpolation (environment.h:1821) */
  /* derived */
  double x0, x1, y0, y1;   /* extent in model space */
  double step;
  double s, d, r_max, cap_radius;
  double shift_h, shift_v; /* h, v in model units (what facet_base::h/v hold after process_geometry) */
  int32_t has_shift, has_lcp, has_shear, has_2d_tf, has_translation;
  /* input, optional: a window of the total image ('W' clause of an i-line: the raster handed to
   * eu_source_upload is window_width x window_height, width/height/hfov describe the whole image).
   * Leave zero for uncropped facets; eu_facet_prepare then sets the window to the whole image. */
  int32_t window_width, window_height, window_x_offset, window_y_offset;
  /* input, optional: --mask_for (envutil_main.cc:999-1001,1077-1091; facet_spec::masked, masking.h:70-139). 0 = normal
   * operation; 1 = this facet is painted BLACK, 2 = WHITE (the reference's masked == 0 / 1): its colour channels are
   * replaced by the paint value - times the interpolated alpha for facets with an alpha channel - before brighten and
   * the synopsis. A masked facet whose channel count differs from the job's is converted by mono_t instead of repix_t
   * (environment.h:1325-1384), which exists for one- and two-channel jobs only. (Occupies what used to be padding:
   * the struct's size is unchanged.) */
  int32_t masked;
} eu_facet_t;

/* Mirror of the members of `arguments` that define the target and the job (reference
 * envutil_basic.h:633-701). */
typedef struct eu_target {
  /* input */
  int32_t projection;
  int32_t width, height;
  int32_t nchannels;
  double hfov;
  double yaw, pitch, roll;
  double gain;             /* --single: 1 / brighten of the facet whose geometry the target takes over
                              (work(), envutil_payload.cc:481-511); 0 or 1 = none */
  /* derived */
  double x0, x1, y0, y1;
  double step;
  /* cropped output (PTO p-line "S x0,x1,y0,y1", envutil_main.cc:615-627): crop_width > 0 renders the
   * window [crop_x0, crop_x0+crop_width) x [crop_y0, crop_y0+crop_height) of the width x height target
   * only - the output raster has the crop's size, rows of eu_render_rows count from the crop's top
   * (envutil_payload.cc:440-443,470-474: the discrete coordinates fed to the steppers are offset).
   * Not for cubemap / biatan6 targets. All zero = the whole target. */
  int32_t crop_x0, crop_y0, crop_width, crop_height;
  /* --single K (envutil_main.cc:1157-1178): the target has taken over the geometry of facets[K]; set
   * single = K + 1 (0 = not a 'single' job). The library needs the facet itself when it carries lens
   * correction, shift, shear or translation: the rays are then produced by the generic stepper through
   * the INVERSE of those transformations (tf_ex_facet, envutil_payload.cc:1841-1883). */
  int32_t single, reserved;
} eu_target_t;

typedef struct eu_opts {
  int32_t spline_degree;    /* --degree (default 1) */
  int32_t prefilter_degree; /* --prefilter (<0: same as spline_degree) */
  int32_t synopsis;         /* eu_synopsis_t */
  int32_t solo;             /* -1, or the single facet to show; forced to 0 for one facet */
  int32_t support_min;      /* cubemap IR support, default 8  (envutil_main.cc:458) */
  int32_t tile_size;        /* cubemap IR tile size, default 64 (envutil_main.cc:457) */
  int32_t reserved[2];      /* back-end options, 0 = defaults. [0]: texel layout of RGB sources in HBM - 0 = 16-byte texels
                               for degree <= 1 (one 128-bit load per tap), 12-byte texels otherwise; 1 = always 16-byte;
                               2 = always 12-byte. [1]: EU_OPT_* bits */
} eu_opts_t;
/* eu_opts_t.reserved[1] */
#define EU_OPT_NO_TILES 1      /* never stage the gather footprint in shared memory */
#define EU_OPT_NO_SHAPES 2     /* never use the kernels compiled for one job shape */
#define EU_OPT_NARROW_STORES 4 /* no 128-bit pixel stores into peer frames */
/* Arithmetic of the render kernels. Default (bit clear): no contraction anywhere - every product and sum is rounded
 * separately, the output is bit-identical to the reference built with -ffp-contract=off. With EU_OPT_CONTRACTED the
 * b-spline weights, the window sum (zimt/eval.h:903-1059) and the twining accumulation (twining.h:106-263) use fused
 * multiply-adds, as a reference built with g++'s default -ffp-contract=fast on FMA hardware does. Rays, source
 * coordinates, gates, window positions, face and facet indices are the same bits either way; pixel values move by a few
 * ulp (measured against the pinned reference build at BASELINE's full sizes: profiles/). */
#define EU_OPT_CONTRACTED 16

/* One twining tap: sub-pixel offset in units of the target's pixel step and weight
 * (args.twine_spread, reference envutil_main.cc:1253-1355; consumed at twining.h:106-121). */
typedef struct eu_tap {
  float x, y, w;
} eu_tap_t;

typedef struct eu_timing {
  float render_ms;  /* CUDA-event time of the render kernel(s) only */
  float h2d_ms;     /* host->device copies inside the call (0 for device entry points) */
  float d2h_ms;
  int32_t launches; /* kernels launched by the call */
  int32_t shape;    /* eu_render / eu_render_rows: index of the compiled-in job shape whose kernel rendered
                       (0: a general kernel); 0 from the other entry points */
} eu_timing_t;

/* Alpha from PTO: exclude masks (k-lines, variant t0) and the lens crop of an i-line (S clause)
 * make parts of a facet transparent (reference environment.h:703-890): a 0/1 plane is built
 * (polygons by scan-line fill envutil_basic.cc:236-321, crop rectangle, or ellipse for fisheye
 * facets), feathered with the 5-tap binomial (1 4 6 4 1)/16 along both axes, and multiplied into
 * every channel; facets without alpha get an alpha channel first (nchannels = native + 1). */
typedef struct eu_alpha_spec {
  int32_t native_nchannels; /* channels of the raster handed in; the facet's nchannels is this, or
                               this + 1 when an alpha channel has to be added (1 -> 2, 3 -> 4) */
  int32_t has_crop;
  int32_t crop_x0, crop_x1, crop_y0, crop_y1;
  int32_t n_masks;
  const int32_t* mask_sizes; /* vertices per polygon */
  const float* mask_xy;      /* all vertices, x y x y ..., polygon after polygon */
} eu_alpha_spec_t;

typedef struct eu_source* eu_source_h; /* opaque: a staged, braced, prefiltered source */

/* ---------------------------------------------------------------------------------------
 * Host-side set-up arithmetic (no GPU needed). Restates, in the reference's precision:
 *   get_vfov/get_step/get_extent      envutil_basic.cc:50-226
 *   facet set-up                      envutil_main.cc:935-976, envutil_basic.h:499-543
 *   target set-up                     envutil_main.cc:483-510,1199-1232
 *   rotate_3d / make_r3_t / rotate    envutil_payload.cc:136-218, geometry.h:79-97
 *   make_spread / twine_setup         envutil_main.cc:1253-1355,1405-1616
 *   metrics_t                         cubemap.h:233-400
 */
double eu_get_vfov(int projection, int width, int height, double hfov);
double eu_get_step(int projection, int width, int height, double hfov);
void eu_get_extent(int projection, int width, int height, double hfov, double ext[4]);
/* fills every derived member of *f; returns EU_ERR_ARGUMENT for nonsensical input */
int eu_facet_prepare(eu_facet_t* f);
/* applies the height rules (cubemap: 6*width; spherical with height 0: width/2 after rounding
 * width up to even; else height 0 -> width) and fills the derived members */
int eu_target_prepare(eu_target_t* t);
/* rows = images of e_x,e_y,e_z under the rotation (float quaternion, double rows) */
void eu_rotation_matrix(double roll, double pitch, double yaw, int inverse, double m[9]);
/* basis handed to the stepper for one facet: R_camera * R_facet^-1 (envutil_payload.cc:1923-1948) */
void eu_facet_basis(const eu_target_t* t, const eu_facet_t* f, double m[9]);
/* twining filter. twine > 0: twine*twine taps (box, or truncated gaussian if sigma > 0,
 * taps below threshold dropped, weights renormalised). twine < 0: automatic twining as
 * arguments::twine_setup does (it looks at o->spline_degree and o->solo). Writes at most max_taps taps; returns the tap count
 * (0 = twining off) or a negative status. *twine_out receives the effective twine factor. */
int eu_make_spread(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets, int twine,
                   double twine_width, double twine_density, double twine_sigma,
                   double twine_threshold, int twine_max, eu_tap_t* taps, int max_taps,
                   int* twine_out);
/* cubemap internal representation metrics (cubemap.h:233-400):
 * out[0]=section_px out[1]=frame_px (left margin) ; refc_md/model_to_px as the view uses them */
int eu_cubemap_metrics(int face_px, double hfov, int support_min, int tile_size, int32_t out_i[4],
                       double out_d[4]);

/* ---------------------------------------------------------------------------------------
 * Device side. One process drives one GPU (multi-GPU = one process per GPU, each rendering
 * a row band; see eu_render_rows). Single caller, blocking, like payload(): the library keeps one
 * context per process and is NOT thread-safe - all calls must come from one host thread at a time
 * (eu_last_error alone is per thread). Jobs on different CUDA streams are safe: every plan's tables
 * live in their own slot, reused only behind an event recorded after the kernels that read them. */
int eu_init(int device_id);
void eu_shutdown(void);
const char* eu_last_error(void);
int eu_device_count(void);
/* Number of arithmetics the render kernels are built in: 2 (exact and contracted, selected per job by
 * EU_OPT_CONTRACTED in eu_opts_t.reserved[1]). */
int eu_render_arithmetic(void);

/* Stage one source raster: upload, place into the braced container (lat/lon & mounted
 * images, reference environment.h:594-950) or the cubemap internal representation
 * (cubemap.h:548-946,1147-1233), run the b-spline prefilter, brace. pixels: host pointer,
 * f->width * f->height * f->nchannels floats (cubemaps: width x 6*width).
 * asset_key may be NULL; a non-NULL key makes the source findable by eu_source_find and
 * subject to the two-generation ageing of eu_cycle (environment.h:84-227). */
int eu_source_upload(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o,
                     const float* pixels, eu_source_h* out, eu_timing_t* t);
/* same, pixels already in device memory (contiguous interleaved float32) */
int eu_source_upload_device(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o,
                            const float* d_pixels, void* cuda_stream, eu_source_h* out,
                            eu_timing_t* t);
/* eu_source_upload for a facet with exclude masks and/or a lens crop: pixels has
 * a->native_nchannels channels, the staged source f->nchannels */
int eu_source_upload_alpha(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, const float* pixels,
                           const eu_alpha_spec_t* a, eu_source_h* out, eu_timing_t* t);
eu_source_h eu_source_find(const char* asset_key);
int eu_source_release(eu_source_h s);
int eu_cycle(void); /* conclude_cycle(): drop sources not used since the previous cycle */
/* copy the staged coefficient container back (tests): out must hold eu_source_container_floats */
size_t eu_source_container_floats(eu_source_h s, int32_t shape[4] /* w,h,left_x,left_y */);
int eu_source_download(eu_source_h s, float* out);

/* Render. facets[i] pairs with sources[i]. taps/n_taps: twining filter (n_taps 0 = off, the
 * reference's ninputs==3 path; otherwise the ninputs==9 path). out: host buffer of
 * width*height*nchannels floats. */
int eu_render(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
              const eu_source_h* sources, const eu_tap_t* taps, int n_taps, float* out,
              eu_timing_t* timing);
/* Rows [row0,row1) only, into device memory: d_out points at the first float of row `row0`
 * (a band buffer of (row1-row0)*width*nchannels floats). Asynchronous on cuda_stream when
 * timing is NULL. This is the multi-GPU work unit (one band per rank). */
int eu_render_rows(const eu_target_t* t, const eu_opts_t* o, int n_facets,
                   const eu_facet_t* facets, const eu_source_h* sources, const eu_tap_t* taps,
                   int n_taps, int row0, int row1, float* d_out, void* cuda_stream,
                   eu_timing_t* timing);
/* same, with the output rows `out_pitch_floats` apart (>= width*nchannels): renders into a larger
 * raster in place - e.g. into the core of a reserved source, below */
int eu_render_rows_pitched(const eu_target_t* t, const eu_opts_t* o, int n_facets,
                           const eu_facet_t* facets, const eu_source_h* sources, const eu_tap_t* taps,
                           int n_taps, int row0, int row1, float* d_out, int out_pitch_floats,
                           void* cuda_stream, eu_timing_t* timing);
/* same, restricted to the columns [col0, col1) of those rows (col0 a multiple of 32): d_out still points at the
 * first float of row `row0`, column 0. Pipelines that need only part of an intermediate raster (BASELINE
 * configs[4]: stage B samples a curved region of every merged image) render just the rectangles that cover it.
 * out_texel_floats: floats from one output pixel to the next - 0 or nchannels for a dense raster, 4 for an RGB job
 * that renders into a reserved source whose texels are 16 bytes (eu_source_reserve reports it). */
int eu_render_rect_pitched(const eu_target_t* t, const eu_opts_t* o, int n_facets,
                           const eu_facet_t* facets, const eu_source_h* sources, const eu_tap_t* taps,
                           int n_taps, int row0, int row1, int col0, int col1, float* d_out,
                           int out_pitch_floats, int out_texel_floats, void* cuda_stream, eu_timing_t* timing);
/* A source whose raster is produced on the device: two-stage jobs (BASELINE configs[4]: hdr_merge
 * of a position's brackets, then the panorama over the merged images) hand the first stage's result
 * to the second without an intermediate raster and without the placement copy of
 * eu_source_upload_device. eu_source_reserve allocates the braced container of a single image
 * (not a cubemap) and returns the device address of core texel (0,0), the row pitch and the texel stride in floats
 * (the container follows the library's layout rule: RGB texels of bilinear jobs are 16 bytes, eu_opts_t.reserved[0]);
 * the caller fills the rows (eu_render_rect_pitched with that address, pitch and texel stride), then
 * eu_source_commit - ordered after cuda_stream's work - prefilters and braces. The handle is
 * used and released like any other source. */
int eu_source_reserve(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, eu_source_h* out,
                      float** d_core, int* pitch_floats, int* texel_floats);
int eu_source_commit(eu_source_h s, const eu_facet_t* f, const eu_opts_t* o, void* cuda_stream,
                     eu_timing_t* t);
/* Fill part of a reserved source from a raster in host (page-locked, for a truly asynchronous copy) or device
 * memory: the rectangle rows [row0,row1) x columns [col0,col1) of the image; `pixels` points at its first texel,
 * rows src_pitch_floats apart. Enqueued on cuda_stream (asynchronous for page-locked host memory; pageable memory is
 * staged before the call returns): work enqueued on cuda_stream after the call sees the texels, and the host buffer
 * may be reused once cuda_stream has reached that point. A HOST raster must hold its final contents when the call is
 * made - the library may start reading it at once, on a copy stream of its own, so that the copies of consecutive
 * rectangles follow each other while kernels on cuda_stream widen them into 16-byte texels. Device rasters are read in
 * stream order.
 * Ranks of a multi-GPU job upload just the part of every source their band of the output can see. Rows and
 * columns never written read as zero. f and o given to eu_source_commit must be those given to
 * eu_source_reserve; with t == NULL eu_source_commit only enqueues (later work on cuda_stream is ordered after it). */
int eu_source_write_rect(eu_source_h s, const float* pixels, size_t src_pitch_floats, int row0, int row1, int col0,
                         int col1, void* cuda_stream);
/* Pipelined jobs. payload() is blocking, but a host that streams jobs (pipe mode, sequences) can
 * keep the PCIe links busy in both directions: eu_source_upload_async enqueues the H2D copy on an
 * upload stream and the staging kernels behind it, eu_render_async enqueues the render and - on a
 * download stream - the D2H copy of the result, and eu_job_wait blocks until that copy is done.
 * With two or three jobs in flight the upload of job n+1 overlaps the download of job n. `pixels`
 * and `out` must be page-locked host memory and stay valid until eu_job_wait returns; sources may
 * be released after the wait. At most EU_MAX_JOBS_IN_FLIGHT jobs may be pending. */
#define EU_MAX_JOBS_IN_FLIGHT 4
typedef struct eu_job* eu_job_h;
int eu_source_upload_async(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, const float* pixels,
                           eu_source_h* out);
int eu_render_async(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                    const eu_source_h* sources, const eu_tap_t* taps, int n_taps, float* out, eu_job_h* job);
int eu_job_wait(eu_job_h job, eu_timing_t* timing);

/* Frames shared between the GPUs of one box ("output bands gathered"). The rank that
 * wants the whole frame allocates it and exports a handle; the other ranks (one process per
 * GPU) open the handle and pass `frame + row0*width*nchannels` as d_out of eu_render_rows: the
 * render kernel then stores its band straight into the owner's HBM over NVLink (peer stores,
 * 128-bit for RGB) - rendering and gathering are one kernel, there is no band buffer and no
 * collective. The owner may read the frame once every rank has synchronised its stream (the
 * host's barrier). CUDA IPC underneath: the handle is EU_FRAME_HANDLE_BYTES opaque bytes that
 * travel through any host channel (torch.distributed, a pipe, MPI). */
#define EU_FRAME_HANDLE_BYTES 64
int eu_frame_alloc(size_t n_floats, float** d_frame);
int eu_frame_free(float* d_frame);
int eu_frame_export(const float* d_frame, unsigned char handle[EU_FRAME_HANDLE_BYTES]);
int eu_frame_open(const unsigned char handle[EU_FRAME_HANDLE_BYTES], float** d_frame);
int eu_frame_close(float* d_frame);

/* Tethered output: the frame as the reference hands it to its viewer (visor) instead of writing an image file -
 * one uint32 sRGBA value per pixel (A<<24 | B<<16 | G<<8 | R). Replaces `act + to_screen_t` in work()
 * (reference envutil_payload.cc:298-413,524-531; the frame buffer is args.p_screen_data, envutil_main.cc:1791):
 * every channel goes through a 256-knot table of 255 * sRGB(x) with linear interpolation (lut_based_tf, :243-283)
 * and is truncated; one channel = grey, opaque; two = grey + alpha; three = opaque RGB; four = RGBA. As in the
 * reference, a 'single' job's un-brighten gain (eu_target_t.gain) is NOT applied on this path (:491).
 * eu_render_screen = eu_render with that store: out is a host buffer of width*height uint32 - a third of the float
 * frame's bytes cross PCIe. eu_to_screen_device converts n_pixels of nchannels interleaved floats already in device
 * memory (asynchronous on cuda_stream). eu_screen_lut fills the table (256 knots + one brace value; host, no GPU). */
void eu_screen_lut(float lut[257]);
int eu_render_screen(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                     const eu_source_h* sources, const eu_tap_t* taps, int n_taps, uint32_t* out,
                     eu_timing_t* timing);
int eu_to_screen_device(const float* d_pixels, int nchannels, size_t n_pixels, uint32_t* d_out, void* cuda_stream);

/* Index plane for bit-exact parity checks: per target pixel the cube face hit (single cubemap
 * facet) or the winning facet of the panorama synopsis (-1: no facet hit). Host buffer w*h. */
int eu_debug_planes(const eu_target_t* t, const eu_opts_t* o, int n_facets,
                    const eu_facet_t* facets, const eu_source_h* sources, int32_t* index_out);

/* The tie band of the same job ("indices bit-exact except within a tie band ... mask those pixels,
 * report their count"): per target pixel 1 where the cube-face choice (ray_to_cubeface, geometry.h:1178-1289:
 * the two largest |components| of the ray) or the winning facet (_voronoi_syn, envutil_payload.cc:818-956: the
 * two best z * recip_step scores) is within `ulps` units in the last place of flipping, else 0. A build of the
 * reference with another math library may choose the other face / facet exactly there. Host buffer w*h bytes. */
int eu_debug_tie_plane(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                       const eu_source_h* sources, int ulps, unsigned char* tie_out);

#ifdef __cplusplus
}
#endif
#endif /* ENVUTIL_B200_H */
