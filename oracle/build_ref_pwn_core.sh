#!/bin/sh
# Builds oracle/_ref/libpwn_core_ref.so: the REFERENCE'S OWN pwn_core sources (from $REF, default /root/reference)
# compiled against the Eigen / OpenCV stand-ins of oracle/shim/ behind the extern "C" face of oracle/ref_pwn_core.cpp.
# TEST INFRASTRUCTURE; run by oracle/Makefile where the reference tree exists (this container).
#
# The sources are compiled from a scratch copy under $TMPDIR (never from or into the repository) because one header
# needs a one-hunk fix before a present-day g++ accepts it: InformationMatrix::transformInPlace
# (informationmatrix.h:77-82) is a never-instantiated member template that is ill-formed on its face (`other.block<3,3>`
# without `template`, and `return *this` from a const member as a non-const reference); the compilers of 2013 did not
# look inside uninstantiated templates.  The hunk is deleted; nothing on the path calls it (Cloud::transformInPlace uses
# InformationMatrixVector::transformInPlace, informationmatrix.h:98-112).
# Flags: -DNDEBUG like the reference's release build (pwn_static.cpp:40,55 carry inverted asserts that fire on every valid
# input), -ffp-contract=off and no -march so that float32 arithmetic is evaluated as written, OpenMP on.
# Usage: build_ref_pwn_core.sh [all|demo]   (demo: only oracle/_ref/drop_in_demo, which also needs the CUDA library)
set -e
MODE=${1:-all}
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref/libpwn_core_ref.so"
CXX=/usr/bin/g++
[ -x "$CXX" ] || CXX=g++
SCRATCH=$(mktemp -d)
trap 'rm -rf "$SCRATCH"' EXIT
mkdir -p "$SCRATCH/g2o_frontend" "$HERE/_ref"
cp -r "$REF/g2o_frontend/pwn_core" "$REF/g2o_frontend/basemath" "$SCRATCH/g2o_frontend/"
H="$SCRATCH/g2o_frontend/pwn_core/informationmatrix.h"
sed -n '78p' "$H" | grep -q 'inline InformationMatrix& transformInPlace(const Eigen::MatrixBase<OtherDerived> &other) const' \
  || { echo "informationmatrix.h does not look like the expected revision" >&2; exit 1; }
sed -i '77,82d' "$H"
S="$SCRATCH/g2o_frontend/pwn_core"
SRCS="$S/pwn_static.cpp $S/pointprojector.cpp $S/pinholepointprojector.cpp $S/gaussian3.cpp $S/pointintegralimage.cpp
 $S/statscalculator.cpp $S/statscalculatorintegralimage.cpp $S/informationmatrixcalculator.cpp $S/cloud.cpp
 $S/depthimageconverter.cpp $S/depthimageconverterintegralimage.cpp $S/correspondencefinder.cpp $S/linearizer.cpp
 $S/se3_prior.cpp $S/aligner.cpp $S/merger.cpp $S/voxelcalculator.cpp $S/multipointprojector.cpp"
if [ "$MODE" != demo ]; then
$CXX -std=gnu++11 -fpermissive -w -O2 -DNDEBUG -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC \
  -I"$HERE/shim" -I"$SCRATCH" -I"$SCRATCH/g2o_frontend" -o "$OUT" $SRCS "$HERE/ref_pwn_core.cpp" \
  -L"$HERE/build" -loracle -Wl,-rpath,'$ORIGIN/../build' -lm
echo "built $OUT"
# performance flavour of the same library: the reference's own flags (-O3 -march + OpenMP, /root/reference/CMakeLists.txt:
# 135,145,171; x86-64-v3 instead of native so that it also runs on the GPU box's host), timed by bench.py next to the oracle port
$CXX -std=gnu++11 -fpermissive -w -O3 -march=x86-64-v3 -DNDEBUG -fopenmp -shared -fPIC \
  -I"$HERE/shim" -I"$SCRATCH" -I"$SCRATCH/g2o_frontend" -o "$HERE/_ref/libpwn_core_ref_fast.so" $SRCS "$HERE/ref_pwn_core.cpp" \
  -L"$HERE/build" -loracle -Wl,-rpath,'$ORIGIN/../build' -lm
echo "built $HERE/_ref/libpwn_core_ref_fast.so"
# the reference's own CLI drivers (pwn_core/pwn_simple_aligner.cpp = BASELINE config 0/1, frame-to-frame odometry;
# pwn_core/pwn_aligner.cpp = scene-based odometry with the local map, Merger and VoxelCalculator), unmodified
for drv in pwn_simple_aligner pwn_aligner; do
  $CXX -std=gnu++11 -fpermissive -w -O2 -DNDEBUG -ffp-contract=off -fno-fast-math -fopenmp \
    -I"$HERE/shim" -I"$SCRATCH" -I"$SCRATCH/g2o_frontend" -o "$HERE/_ref/${drv}_ref" $SRCS "$S/$drv.cpp" \
    -L"$HERE/build" -loracle -Wl,-rpath,'$ORIGIN/../build' -lm
  echo "built $HERE/_ref/${drv}_ref"
done
fi
# the drop-in demonstration (integration/drop_in_demo.cpp): the same reference sources + the option-A binding of
# integration/pwn_b200/b200_pwn.h + this repository's CUDA library, in one executable
REPO=$(cd "$HERE/.." && pwd)
if [ ! -f "$REPO/g2o_frontend_b200/lib/libnicp_b200.so" ]; then
  [ "$MODE" = demo ] && { echo "drop_in_demo needs g2o_frontend_b200/lib/libnicp_b200.so (make -C g2o_frontend_b200/csrc first)" >&2; exit 1; }
else
  $CXX -std=gnu++11 -fpermissive -w -O2 -DNDEBUG -ffp-contract=off -fno-fast-math -fopenmp \
    -I"$HERE/shim" -I"$SCRATCH" -I"$SCRATCH/g2o_frontend" -I"$REPO/include" -I"$REPO/integration" \
    -o "$HERE/_ref/drop_in_demo" $SRCS "$REPO/integration/drop_in_demo.cpp" \
    -L"$HERE/build" -loracle -L"$REPO/g2o_frontend_b200/lib" -lnicp_b200 \
    -Wl,-rpath,'$ORIGIN/../build' -Wl,-rpath,'$ORIGIN/../../g2o_frontend_b200/lib' -lm
  echo "built $HERE/_ref/drop_in_demo"
fi
