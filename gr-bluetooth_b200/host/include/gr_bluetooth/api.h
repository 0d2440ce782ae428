#ifndef INCLUDED_GR_BLUETOOTH_B200_API_H
#define INCLUDED_GR_BLUETOOTH_B200_API_H
#include <gnuradio/attributes.h>
#ifdef gnuradio_bluetooth_EXPORTS
#define GR_BLUETOOTH_API __GR_ATTR_EXPORT
#else
#define GR_BLUETOOTH_API __GR_ATTR_IMPORT
#endif
#endif
