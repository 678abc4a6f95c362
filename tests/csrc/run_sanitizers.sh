#!/bin/bash
# The device code that is also host code (tb_math.cuh, tb_prior.cuh, tb_wcs.cuh) under
# AddressSanitizer + UndefinedBehaviorSanitizer: builds tests/csrc/host_math.cpp with the
# sanitizers, runs the CPU tests that drive it, restores the normal build.
#   bash tests/csrc/run_sanitizers.sh        (from the repository root; ~1 minute)
# Expected: all tests pass, no AddressSanitizer report, and ONE UndefinedBehaviorSanitizer line --
# tb_math.cuh `ir = jp + jm + 1` for the quaternion (s, s, 0, 0): the reference's own arithmetic at
# the exact south pole (see the comment there).
set -e
cd "$(dirname "$0")"
# (everything else is built first: nvcc does not run under a preloaded sanitizer runtime)
(cd ../.. && python -c "import __graft_entry__ as g; g.build()" > /dev/null)
GXX=/usr/bin/g++
ASAN=$($GXX -print-file-name=libasan.so)
UBSAN=$($GXX -print-file-name=libubsan.so)
$GXX -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -ffp-contract=off -fPIC -shared \
    -x c++ -o libhostmath.so host_math.cpp -lm
touch libhostmath.so
cd ..
ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=0 LD_PRELOAD="$ASAN $UBSAN" \
    python -m pytest test_host_math.py test_pixels_wcs.py test_offset_prior.py -q -s -m "not gpu" \
    -p no:cacheprovider 2>&1 | grep -E "runtime error|AddressSanitizer|passed|failed" | sort | uniq -c
rm -f csrc/libhostmath.so    # rebuilt without sanitizers by the next test run (helpers.host_math_lib)
