/* oracle/shim: stand-in for <gnuradio/attributes.h> (GNU Radio is absent). TEST INFRASTRUCTURE. */
#ifndef BTB_SHIM_GR_ATTRIBUTES_H
#define BTB_SHIM_GR_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
