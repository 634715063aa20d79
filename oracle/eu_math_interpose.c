/* TEST INFRASTRUCTURE. Link-time replacement of libm's binary32 elementary functions by the
 * back-end's specified ones (include/eu_math.h) for the "pinned math" build of the reference
 * (oracle/_ref/envutil_ref_pm). zimt's goading back-end calls std::sin/cos/tan/atan/atan2 on
 * float lanes (reference zimt/simd/vector_common.h:203-272), which resolve to these symbols;
 * defining them in the executable pre-empts libm's. Of the double-precision functions only atan is replaced (its one per-pixel call site is the
 * stereographic stepper, reference stepper.h:1146-1151).
 */
#include "eu_math.h"

float sinf(float x) { return eu_sinf(x); }
float cosf(float x) { return eu_cosf(x); }
void sincosf(float x, float* s, float* c) { eu_sincosf(x, s, c); }
float tanf(float x) { return eu_tanf(x); }
float atanf(float x) { return eu_atanf(x); }
float atan2f(float y, float x) { return eu_atan2f(y, x); }
double atan(double x) { return eu_atan(x); }
