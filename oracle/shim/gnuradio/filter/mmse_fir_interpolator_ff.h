/* oracle/shim: stand-in for <gnuradio/filter/mmse_fir_interpolator_ff.h>. TEST INFRASTRUCTURE. */
#ifndef BTB_SHIM_GR_MMSE_H
#define BTB_SHIM_GR_MMSE_H
#include "../../../gr_arith.h"
namespace gr { namespace filter {
class mmse_fir_interpolator_ff {
public:
  mmse_fir_interpolator_ff();
  unsigned ntaps() const { return GRA_MMSE_NTAPS; }
  unsigned nsteps() const { return GRA_MMSE_NSTEPS; }
  float interpolate(const float input[], float mu) const;
};
}}
#endif
