/* oracle/shim: stand-in for <gnuradio/io_signature.h>. TEST INFRASTRUCTURE. */
#ifndef BTB_SHIM_GR_IO_SIGNATURE_H
#define BTB_SHIM_GR_IO_SIGNATURE_H
#include <gnuradio/sync_block.h>
#endif
