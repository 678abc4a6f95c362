#!/bin/bash
# Test infrastructure only.  Compiles the REFERENCE's own C++ hot-path kernels,
# unmodified and from where they lie under /root/reference, into
# oracle/_ref/_toast_oracle*.so (git-ignored; travels to the GPU box with gpurun).
# Flags matter for parity: baseline x86-64 (no -march=native => no FMA contraction),
# no -ffast-math.  See SURVEY.md 8(c).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${TOAST_REFERENCE:-/root/reference}"
R="$REF/src"; L="$R/toast/_libtoast"
OUT="$HERE/_ref"
if [ ! -d "$L" ]; then echo "reference not present at $REF; keeping prebuilt $OUT" >&2; exit 0; fi
mkdir -p "$OUT"
PY="${PYTHON:-python3}"
EXT=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
TARGET="$OUT/_toast_oracle$EXT"
if [ -f "$TARGET" ] && [ "${FORCE:-0}" != "1" ]; then echo "up to date: $TARGET"; exit 0; fi
INC="-I$L -I$R/libtoast/include -I$R/libtoast/src -I$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])") -I$($PY -c "import pybind11; print(pybind11.get_include())")"
# LAPACK for the reference's covariance eigen-inversion: scipy's bundled OpenBLAS (symbols
# prefixed scipy_, forwarded by ref_shim/lapack_shim.cpp).  Without it the reference still builds
# but cov_eigendecompose_diag throws at run time.
BLAS_SO=$($PY - <<'PYEOF'
import glob, os
try:
    import scipy
    c = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                               "libscipy_openblas*.so"))
    print(os.path.realpath(c[0]) if c else "")
except Exception:
    print("")
PYEOF
)
LAPACK_DEFS=""; LAPACK_LINK=""; LAPACK_SRC=""
if [ -n "$BLAS_SO" ]; then
  LAPACK_DEFS="-DHAVE_LAPACK=1 -DLAPACK_NAMES_UBACK=1"
  LAPACK_LINK="-L$(dirname "$BLAS_SO") -l:$(basename "$BLAS_SO") -Wl,-rpath,$(dirname "$BLAS_SO")"
  LAPACK_SRC="$HERE/ref_shim/lapack_shim.cpp"
fi
SRC="$LAPACK_SRC $HERE/ref_shim/mini_module.cpp $HERE/ref_shim/version.cpp $R/libtoast/src/toast_sys_utils.cpp $R/libtoast/src/toast_sys_environment.cpp $R/libtoast/src/toast_map_cov.cpp $R/libtoast/src/toast_math_linearalgebra.cpp"
for f in common intervals qarray_core accelerator ops_pointing_detector ops_stokes_weights ops_pixels_healpix ops_mapmaker_utils ops_noise_weight ops_scan_map template_offset map_cov pixels; do SRC="$SRC $L/$f.cpp"; done
OBJ=""
mkdir -p "$OUT/obj"
pids=()
for s in $SRC; do
  o="$OUT/obj/$(basename "$s" .cpp).o"; OBJ="$OBJ $o"
  g++ -O3 -fopenmp -foffload=disable -std=c++17 -fPIC $LAPACK_DEFS -c "$s" -o "$o" $INC &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -fopenmp -o "$TARGET" $OBJ $LAPACK_LINK
rm -rf "$OUT/obj"
echo "built $TARGET"
