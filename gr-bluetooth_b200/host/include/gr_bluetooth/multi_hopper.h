// gr::bluetooth::multi_hopper -- same factory as the reference (include/gr_bluetooth/multi_hopper.h:53):
// make(sample_rate, center_freq, squelch_threshold, LAP, aliased, tun)
#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_HOPPER_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_HOPPER_H
#include <gr_bluetooth/api.h>
#include "gr_bluetooth/multi_block.h"

namespace gr {
namespace bluetooth {

class GR_BLUETOOTH_API multi_hopper : virtual public multi_block {
 public:
  typedef boost::shared_ptr<multi_hopper> sptr;
  static sptr make(double sample_rate, double center_freq, double squelch_threshold, int LAP, bool aliased, bool tun);
};

}  // namespace bluetooth
}  // namespace gr
#endif
