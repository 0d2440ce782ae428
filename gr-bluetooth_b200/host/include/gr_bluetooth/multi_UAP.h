// gr::bluetooth::multi_UAP -- same factory as the reference (include/gr_bluetooth/multi_UAP.h:51):
// make(sample_rate, center_freq, squelch_threshold, LAP): determine the UAP of the piconet with this LAP
// from the headers of its packets.
#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_UAP_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_UAP_H
#include <gr_bluetooth/api.h>
#include "gr_bluetooth/multi_block.h"

namespace gr {
namespace bluetooth {

class GR_BLUETOOTH_API multi_UAP : virtual public multi_block {
 public:
  typedef boost::shared_ptr<multi_UAP> sptr;
  static sptr make(double sample_rate, double center_freq, double squelch_threshold, int LAP);
};

}  // namespace bluetooth
}  // namespace gr
#endif
