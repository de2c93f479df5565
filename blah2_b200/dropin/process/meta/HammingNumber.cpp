#include "HammingNumber.h"

#include "b200dd.h"

uint32_t next_hamming(uint32_t value) { return b200dd_next_hamming(value); }
