// gr_compat: the few GNU Radio runtime types the blocks are written against, for builds on
// machines without GNU Radio (this repository's build container).  With GNU Radio installed,
// compile with -DBTB200_WITH_GNURADIO and the real <gnuradio/sync_block.h> is used instead;
// the block sources are identical in both cases.
#pragma once
#include <complex>
#include <memory>
#include <string>
#include <vector>

typedef std::complex<float> gr_complex;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace boost {
using std::shared_ptr;
using std::dynamic_pointer_cast;
}  // namespace boost

namespace gr {

class io_signature {
 public:
  typedef boost::shared_ptr<io_signature> sptr;
  static sptr make(int min_streams, int max_streams, int sizeof_stream_item) {
    return sptr(new io_signature(min_streams, max_streams, sizeof_stream_item));
  }
  int min_streams() const { return d_min; }
  int max_streams() const { return d_max; }
  int sizeof_stream_item() const { return d_size; }

 private:
  io_signature(int a, int b, int c) : d_min(a), d_max(b), d_size(c) {}
  int d_min, d_max, d_size;
};

// A sync block with history: the scheduler hands work() a pointer to history()-1 old items
// followed by the new ones and expects the number of items consumed back.
class sync_block {
 public:
  sync_block() {}
  sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
      : d_name(name), d_in(in), d_out(out) {}
  virtual ~sync_block() {}
  const std::string &name() const { return d_name; }
  unsigned history() const { return d_history; }
  void set_history(unsigned h) { d_history = h; }
  // gr::sync_block's default: ninput = noutput + history() - 1
  virtual void forecast(int noutput_items, std::vector<int> &ninput_items_required) {
    for (auto &n : ninput_items_required) n = noutput_items + (int)history() - 1;
  }
  virtual int work(int noutput_items, gr_vector_const_void_star &input_items,
                   gr_vector_void_star &output_items) = 0;

 private:
  std::string d_name;
  io_signature::sptr d_in, d_out;
  unsigned d_history = 1;
};

}  // namespace gr

namespace gnuradio {
template <class T>
boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
}  // namespace gnuradio
