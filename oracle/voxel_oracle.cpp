/*
 * voxel_oracle.cpp -- CPU restatement of pwn::VoxelCalculator::compute (voxelcalculator.cpp:15-73,
 * voxelcalculator.h:14-54).  TEST INFRASTRUCTURE ONLY, PARITY UNPINNED (see pwn_oracle.h).
 *
 * The reference keeps one representative (the FIRST point, by index) per occupied voxel in a
 * std::map keyed by the truncated voxel coordinates and emits the representatives in map order.
 * Its key comparator (voxelcalculator.h:40-46) is NOT a strict weak ordering: the third clause
 * compares indeces[2] when indeces[1] match without requiring indeces[0] to match.  With such a
 * comparator the content of a std::map depends on the tree shape, i.e. on the C++ library.  Two
 * flavours are therefore restated here, both through the very same container (libstdc++ std::map):
 *   strict = 0  the comparator exactly as written in the reference (what a libstdc++ build of the reference does)
 *   strict = 1  the lexicographic order the comparator evidently intends (what the CUDA path implements)
 * tests/ report how far the two are apart on the synthetic clouds.
 */
#include <cstddef>
#include <map>
#include <vector>

namespace {
struct KeyAsWritten {
  int v[3];
  bool operator<(const KeyAsWritten &s) const {
    if (v[0] < s.v[0]) return true;
    if (v[0] == s.v[0] && v[1] < s.v[1]) return true;
    if (v[1] == s.v[1] && v[2] < s.v[2]) return true;
    return false;
  }
};
struct KeyStrict {
  int v[3];
  bool operator<(const KeyStrict &s) const {
    if (v[0] != s.v[0]) return v[0] < s.v[0];
    if (v[1] != s.v[1]) return v[1] < s.v[1];
    return v[2] < s.v[2];
  }
};
template <typename Key>
int voxelize(const float *points, int n, float resolution, int *representatives) {
  std::map<Key, int> first;
  const float inverseResolution = 1.0f / resolution;
  for (int i = 0; i < n; i++) {
    Key k;
    for (int a = 0; a < 3; a++) k.v[a] = (int)(points[4 * (size_t)i + a] * inverseResolution);
    if (first.find(k) == first.end()) first.insert(std::make_pair(k, i));
  }
  int m = 0;
  for (typename std::map<Key, int>::const_iterator it = first.begin(); it != first.end(); ++it) representatives[m++] = it->second;
  return m;
}
}  // namespace

extern "C" int orc_voxelize(const float *points, int n, float resolution, int strict, int *representatives) {
  return strict ? voxelize<KeyStrict>(points, n, resolution, representatives)
                : voxelize<KeyAsWritten>(points, n, resolution, representatives);
}
