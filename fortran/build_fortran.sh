#!/bin/bash
# Builds the reference's own driver (src/diagnose/main.f90 + its includes, UNCHANGED) against the GPU
# drop-in module instead of xtt-lib-fortran/elliptic_tools.f90.  Mirrors make-diagnosis.sh:10-11.
#   REF=/path/to/XLab-EE-fortran  fortran/build_fortran.sh [r8]
# Needs gfortran (absent from the build image of this repository).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
FC="${FC:-gfortran}"
command -v "$FC" >/dev/null || { echo "build_fortran.sh: no Fortran compiler ($FC) on PATH" >&2; exit 3; }
python -c "import sys; sys.path.insert(0,'$ROOT'); from xlab_ee_fortran_b200 import _lib; print(_lib.build())"
LIBDIR="$ROOT/xlab_ee_fortran_b200/lib"
mkdir -p "$ROOT/bin"
# r8: the whole driver promoted with -freal-4-real-8 (every real(4) is a real(8)); the shim is the same source with the
# _f64 entry points, c_double and real(8), generated here instead of being kept as a second file
if [ "${1:-}" = "r8" ]; then
  SHIM="$ROOT/bin/elliptic_tools_r8.f90"
  sed -e 's/_f32/_f64/g; s/c_float/c_double/g; s/real(4)/real(8)/g' "$HERE/elliptic_tools.f90" > "$SHIM"
  FLAGS="-O2 -freal-4-real-8"; OUT=diagnose_gpu_r8
else SHIM="$HERE/elliptic_tools.f90"; FLAGS="-O2"; OUT=diagnose_gpu; fi
LIBS=$(ls "$REF"/xtt-lib-fortran/*.f90 | grep -v elliptic_tools.f90)
cd "$ROOT/bin"
# modules first, then the program (the reference compiles twice for the same reason, make-diagnosis.sh:10-11)
$FC $FLAGS -c $LIBS "$SHIM"
$FC $FLAGS -I"$REF/src/diagnose" "$REF/src/diagnose/main.f90" ./*.o -L"$LIBDIR" -lxee_b200 -Wl,-rpath,"$LIBDIR" -o "$OUT"
rm -f ./*.o ./*.mod ./elliptic_tools_r8.f90
echo "built $ROOT/bin/$OUT  (run it exactly like bin/diagnose:  cd test/test1 && $ROOT/bin/$OUT < diag.txt)"
