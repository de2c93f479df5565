// Pinned staging buffer for the drop-in classes (b200dd_host_alloc): contiguous complex128 the GPU can DMA at the
// full PCIe rate; falls back to ordinary memory if page-locking fails.
#ifndef B200DD_DROPIN_PINNEDBUFFER_H
#define B200DD_DROPIN_PINNEDBUFFER_H

#include "b200dd.h"

#include <complex>
#include <cstdlib>
#include <new>

class PinnedBuffer
{
public:
  PinnedBuffer() : ptr(nullptr), count(0), pinned(false) {}
  ~PinnedBuffer() { release(); }
  PinnedBuffer(const PinnedBuffer &) = delete;
  PinnedBuffer &operator=(const PinnedBuffer &) = delete;
  void resize(size_t n)
  {
    release();
    ptr = static_cast<std::complex<double> *>(b200dd_host_alloc(n * sizeof(std::complex<double>)));
    pinned = ptr != nullptr;
    if (!ptr) ptr = static_cast<std::complex<double> *>(std::malloc((n ? n : 1) * sizeof(std::complex<double>)));
    if (!ptr) throw std::bad_alloc();
    count = n;
  }
  std::complex<double> *data() { return ptr; }
  const std::complex<double> *data() const { return ptr; }
  size_t size() const { return count; }
  std::complex<double> &operator[](size_t i) { return ptr[i]; }
  const std::complex<double> &operator[](size_t i) const { return ptr[i]; }

private:
  void release()
  {
    if (ptr) { if (pinned) b200dd_host_free(ptr); else std::free(ptr); }
    ptr = nullptr;
    count = 0;
  }
  std::complex<double> *ptr;
  size_t count;
  bool pinned;
};

#endif
