// gr::bluetooth::multi_sniffer -- same factory as the reference
// (include/gr_bluetooth/multi_sniffer.h:54): make(sample_rate, center_freq, squelch_threshold, tun)
#ifndef INCLUDED_GR_BLUETOOTH_B200_MULTI_SNIFFER_H
#define INCLUDED_GR_BLUETOOTH_B200_MULTI_SNIFFER_H
#include <gr_bluetooth/api.h>
#include "gr_bluetooth/multi_block.h"

namespace gr {
namespace bluetooth {

class GR_BLUETOOTH_API multi_sniffer : virtual public multi_block {
 public:
  typedef boost::shared_ptr<multi_sniffer> sptr;
  static sptr make(double sample_rate, double center_freq, double squelch_threshold, bool tun);
};

}  // namespace bluetooth
}  // namespace gr
#endif
