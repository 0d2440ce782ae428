/* oracle/shim: stand-in for <gnuradio/filter/firdes.h>. TEST INFRASTRUCTURE. */
#ifndef BTB_SHIM_GR_FIRDES_H
#define BTB_SHIM_GR_FIRDES_H
#include <vector>
#include "../../../gr_arith.h"
namespace gr { namespace filter {
class firdes {
public:
  enum win_type { WIN_HAMMING = 0, WIN_HANN = 1, WIN_BLACKMAN = 2, WIN_RECTANGULAR = 3 };
  static std::vector<float> low_pass(double gain, double fs, double fc, double tw,
                                     win_type w = WIN_HAMMING, double beta = 6.76)
  {
    (void)beta;
    if (w != WIN_HANN) { fprintf(stderr, "shim firdes: only WIN_HANN restated\n"); abort(); }
    std::vector<float> taps((size_t)gra_lowpass_ntaps(fs, tw));
    gra_lowpass(gain, fs, fc, tw, taps.data());
    return taps;
  }
};
}}
#endif
