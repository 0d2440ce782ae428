// gr::bluetooth::multi_LAP -- same factory as the reference (include/gr_bluetooth/multi_LAP.h:53)
#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_LAP_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_LAP_H
#include <gr_bluetooth/api.h>
#include "gr_bluetooth/multi_block.h"

namespace gr {
namespace bluetooth {

class GR_BLUETOOTH_API multi_LAP : virtual public multi_block {
 public:
  typedef boost::shared_ptr<multi_LAP> sptr;
  static sptr make(double sample_rate, double center_freq, double squelch_threshold);
};

}  // namespace bluetooth
}  // namespace gr
#endif
