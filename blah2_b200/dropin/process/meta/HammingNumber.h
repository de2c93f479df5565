// Drop-in for the reference's process/meta/HammingNumber.h (src/process/meta/HammingNumber.h:36):
// only the free function the hot path calls.  Implemented by libb200dd (host arithmetic).
#ifndef B200DD_DROPIN_HAMMINGNUMBER_H
#define B200DD_DROPIN_HAMMINGNUMBER_H

#include <stdint.h>

/// First 5-smooth ("Hamming") number strictly greater than value.
uint32_t next_hamming(uint32_t value);

#endif
