#!/bin/bash
# ref_build.sh -- compile the UNMODIFIED reference sources where they lie under $REF against the PETSc API shim
# (include/petsc_shim, our code).  Outputs ONLY into oracle/_ref/ (git-ignored; travels to the GPU box).
#   libref_cpu.so            reference Discretization.c + SaddlePointProblem.c + Visulaization.c, shim, ORACLE back end
#                            (ctypes target of tests/test_oracle_vs_ref.py: pins the oracle's restatement)
#   libintended_coords.so    our GetElementCoords (the commented-out intent), interposed in front of libref_*.so
#   saddle_point_run_cpu     reference main.c, as written, CPU oracle back end         (BASELINE config 0, CPU leg)
#   libref_b200.so / saddle_point_run_b200[_intended]   the same objects over libb200sp (GPU): the drop-in demonstration
# -ffp-contract=off -O2, baseline x86-64: no FMA contraction, the rounding the oracle restates.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
[ -d "$REF/src" ] || { echo "ref_build.sh: $REF/src not found (nothing to do)"; exit 0; }
mkdir -p "$OUT"
CC=gcc   # not $CC: this image exports CC=/opt/gcc/bin/gcc, a wrapper without libgomp.spec
CFLAGS="-O2 -ffp-contract=off -fPIC -w -I$ROOT/include/petsc_shim -I$REF/include"
SHIM="$ROOT/saddle_point_petsc_b200/csrc/petsc_shim.c"
REFSRC="$REF/src/Discretization.c $REF/src/SaddlePointProblem.c $REF/src/Visulaization.c"

# reference objects (no main), position independent, semantic interposition allowed (GCC default with -fPIC)
for f in $REFSRC $REF/src/main.c; do $CC $CFLAGS -c "$f" -o "$OUT/$(basename "${f%.c}").o"; done
$CC $CFLAGS -c "$SHIM" -o "$OUT/petsc_shim.o"
$CC $CFLAGS -c "$HERE/intended_coords.c" -o "$OUT/intended_coords.o"
$CC -shared -o "$OUT/libintended_coords.so" "$OUT/intended_coords.o"

# --- CPU (oracle back end)
$CC $CFLAGS -fopenmp -c "$HERE/sp_oracle.c" -o "$OUT/sp_oracle.o"
$CC $CFLAGS -I"$HERE" -c "$HERE/shim_backend_oracle.c" -o "$OUT/shim_backend_oracle.o"
$CC -shared -fopenmp -o "$OUT/libref_cpu.so" "$OUT/Discretization.o" "$OUT/SaddlePointProblem.o" "$OUT/Visulaization.o" \
    "$OUT/petsc_shim.o" "$OUT/shim_backend_oracle.o" "$OUT/sp_oracle.o" -lm
$CC -fopenmp -o "$OUT/saddle_point_run_cpu" "$OUT/main.o" -L"$OUT" -lref_cpu -Wl,-rpath,'$ORIGIN' -lm
$CC -fopenmp -o "$OUT/saddle_point_run_cpu_intended" "$OUT/main.o" "$OUT/intended_coords.o" -L"$OUT" -lref_cpu -Wl,-rpath,'$ORIGIN' -lm

# --- GPU (libb200sp back end); only links, needs a B200 to run
LIBDIR="$ROOT/saddle_point_petsc_b200"
if [ -f "$LIBDIR/libb200sp.so" ]; then
  $CC $CFLAGS -c "$ROOT/saddle_point_petsc_b200/csrc/shim_backend_b200sp.c" -o "$OUT/shim_backend_b200sp.o"
  $CC -shared -o "$OUT/libref_b200.so" "$OUT/Discretization.o" "$OUT/SaddlePointProblem.o" "$OUT/Visulaization.o" \
      "$OUT/petsc_shim.o" "$OUT/shim_backend_b200sp.o" -L"$LIBDIR" -lb200sp -Wl,-rpath,'$ORIGIN/../../saddle_point_petsc_b200' -lm
  $CC -o "$OUT/saddle_point_run_b200" "$OUT/main.o" -L"$OUT" -lref_b200 -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../saddle_point_petsc_b200' -lm
  $CC -o "$OUT/saddle_point_run_b200_intended" "$OUT/main.o" "$OUT/intended_coords.o" -L"$OUT" -lref_b200 -Wl,-rpath,'$ORIGIN' \
      -Wl,-rpath,'$ORIGIN/../../saddle_point_petsc_b200' -lm
fi
rm -f "$OUT"/*.o
echo "ref_build.sh: built $(ls "$OUT" | tr '\n' ' ')"
