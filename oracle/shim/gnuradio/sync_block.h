/*
 * oracle/shim: stand-in for the slice of the GNU Radio 3.7 runtime that
 * gr-bluetooth's lib/*.cc touch.  TEST INFRASTRUCTURE:
 * it exists so the reference's own sources compile UNMODIFIED into oracle/_ref.
 */
#ifndef BTB_SHIM_GR_SYNC_BLOCK_H
#define BTB_SHIM_GR_SYNC_BLOCK_H

#include <complex>
#include <memory>
#include <string>
#include <map>
#include <iostream>
#include <vector>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef std::complex<float> gr_complex;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace boost {
using std::shared_ptr;
using std::dynamic_pointer_cast;
}

namespace gr {

class io_signature {
public:
  typedef boost::shared_ptr<io_signature> sptr;
  static sptr make(int, int, int) { return sptr(new io_signature()); }
};

class sync_block {
  unsigned d_history;
  std::string d_name;
public:
  sync_block() : d_history(1) {}
  sync_block(const std::string &name, io_signature::sptr, io_signature::sptr)
    : d_history(1), d_name(name) {}
  virtual ~sync_block() {}
  unsigned history() const { return d_history; }
  void set_history(unsigned h) { d_history = h; }
  virtual int work(int noutput_items, gr_vector_const_void_star &input_items,
                   gr_vector_void_star &output_items) = 0;
};

} // namespace gr

namespace gnuradio {
template <class T> boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
}

#endif
