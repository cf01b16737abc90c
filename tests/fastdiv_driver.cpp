// fastdiv_driver.cpp -- TEST INFRASTRUCTURE: exercises the host part of raytracer.jl_b200/csrc/fastdiv.h.
#include "../raytracer.jl_b200/csrc/fastdiv.h"

extern "C" long long fastdiv_mismatches(unsigned d, unsigned long long start, unsigned long long stop,
                                        unsigned long long step) {
  FastDiv f(d);
  long long bad = 0;
  for (unsigned long long n = start; n <= stop && n <= 0xffffffffull; n += step)
    if (f.div_host((unsigned)n) != (unsigned)n / d) ++bad;
  return bad;
}
