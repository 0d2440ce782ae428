/* oracle/shim: stand-in for <gnuradio/math.h> (fast_atan2f, branchless_clip). TEST INFRASTRUCTURE. */
#ifndef BTB_SHIM_GR_MATH_H
#define BTB_SHIM_GR_MATH_H
#include "../../gr_arith.h"
namespace gr {
float fast_atan2f(float y, float x);
static inline float branchless_clip(float x, float clip) { return gra_branchless_clip(x, clip); }
}
#endif
