// fastdiv.h -- unsigned 32-bit division by a run-time constant (Granlund-Montgomery round-up method): exact for every
// 32-bit numerator; 4 instructions instead of the ~20 (32-bit) / ~70 (64-bit) of an integer division on the GPU.
// Used by the 3-D push kernels (node id -> (x, line), item -> (bx, y, z)).  Plain C++ so that the host part is unit
// tested on the CPU (tests/test_fastdiv.py).
#pragma once
#include <stdint.h>

struct FastDiv {
  unsigned d = 1, magic = 0, shift = 0;
  FastDiv() {}
  explicit FastDiv(unsigned div) : d(div) {
    unsigned l = 0;
    while ((1ull << l) < div) ++l;  // l = ceil(log2 d)
    shift = l;
    magic = (unsigned)((((1ull << l) - div) << 32) / div + 1);
  }
  // the arithmetic of div(), spelled with a 64-bit product
  unsigned div_host(unsigned n) const {
    const unsigned t = (unsigned)(((uint64_t)magic * n) >> 32);
    return shift == 0 ? n : (t + ((n - t) >> 1)) >> (shift - 1);
  }
#if defined(__CUDACC__)
  __device__ __forceinline__ unsigned div(unsigned n) const {
    const unsigned t = __umulhi(magic, n);
    return shift == 0 ? n : (t + ((n - t) >> 1)) >> (shift - 1);
  }
  __device__ __forceinline__ void divmod(unsigned n, unsigned& q, unsigned& r) const {
    q = div(n);
    r = n - q * d;
  }
#endif
};
