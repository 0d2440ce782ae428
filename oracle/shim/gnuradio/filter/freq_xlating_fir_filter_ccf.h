/* oracle/shim: stand-in for <gnuradio/filter/freq_xlating_fir_filter_ccf.h>. TEST INFRASTRUCTURE. */
#ifndef BTB_SHIM_GR_FXLAT_H
#define BTB_SHIM_GR_FXLAT_H
#include <gnuradio/sync_block.h>
#include "../../../gr_arith.h"
namespace gr { namespace filter {
class freq_xlating_fir_filter_ccf {
public:
  typedef boost::shared_ptr<freq_xlating_fir_filter_ccf> sptr;
  gra_fxlat d_f;
  int       d_id;            /* creation order: 2*ch_index (channel) / 2*ch_index+1 (noise) */
  double    d_center_freq;
  static sptr make(int decimation, const std::vector<float> &taps, double center_freq, double fs);
  freq_xlating_fir_filter_ccf(int decimation, const std::vector<float> &taps, double center_freq, double fs);
  ~freq_xlating_fir_filter_ccf();
  unsigned history() const { return (unsigned)d_f.ntaps; }
  int fixed_rate_ninput_to_noutput(int ninput) { return gra_fxlat_ninput_to_noutput(&d_f, ninput); }
  int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out);
};
}}
#endif
