/* oracle/shim: stand-in for <gnuradio/blocks/complex_to_mag_squared.h>. TEST INFRASTRUCTURE. */
#ifndef BTB_SHIM_GR_C2MAG2_H
#define BTB_SHIM_GR_C2MAG2_H
#include <gnuradio/sync_block.h>
namespace gr { namespace blocks {
class complex_to_mag_squared {
public:
  typedef boost::shared_ptr<complex_to_mag_squared> sptr;
  static sptr make(size_t vlen = 1) { (void)vlen; return sptr(new complex_to_mag_squared()); }
  int work(int noutput_items, gr_vector_const_void_star &in, gr_vector_void_star &out);
};
}}
#endif
