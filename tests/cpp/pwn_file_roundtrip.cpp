// pwn_file_roundtrip.cpp -- CPU-only: reads a .pwn cloud file with pwn::Cloud::load (include/pwn/pwn.h) and writes it back
// with pwn::Cloud::save.  tests/test_reference_pwn_core.py feeds it files written by the REFERENCE's Cloud::save and hands
// its output to the reference's Cloud::load (oracle/_ref/libpwn_core_ref.so).  No device context is created.
//   pwn_file_roundtrip in.pwn out.pwn binary(0|1)
#include <cstdio>
#include <cstdlib>

#include "pwn/pwn.h"

int main(int argc, char **argv) {
  if (argc < 4) return 2;
  pwn::Cloud c;
  pwn::Isometry3f T;
  if (!c.load(T, argv[1])) {
    std::fprintf(stderr, "load failed\n");
    return 1;
  }
  std::printf("%zu\n", c.points().size());
  if (!c.save(argv[2], T, 1, std::atoi(argv[3]) != 0)) {
    std::fprintf(stderr, "save failed\n");
    return 1;
  }
  return 0;
}
