#!/usr/bin/env bash
# TEST INFRASTRUCTURE (not product code).
# Stands in for the two cmake configure_file() steps of the reference build
# (/root/reference/CMakeLists.txt, tpl/camp/CMakeLists.txt): turns the
# `config.hpp.in` templates WHERE THEY LIE under /root/reference into
# concrete headers under oracle/_ref/ (git-ignored).  Nothing from the
# reference tree is copied into tracked files.
#
#   usage: gen_ref_config.sh <reference_root> <out_include_dir> <host|cuda>
set -euo pipefail
REF="$1"; OUT="$2"; MODE="${3:-host}"
mkdir -p "$OUT/RAJA" "$OUT/camp"

on_raja="RAJA_ENABLE_OPENMP RAJA_USE_CHRONO RAJA_HAVE_POSIX_MEMALIGN"
on_camp="CAMP_ENABLE_OPENMP"
if [ "$MODE" = "cuda" ]; then
  on_raja="$on_raja RAJA_ENABLE_CUDA"
  on_camp="$on_camp CAMP_ENABLE_CUDA"
fi

gen() { # <in> <out> <enabled macro list>
  local in="$1" out="$2" on="$3"
  local script=()
  for m in $on; do
    script+=(-e "s|^#cmakedefine $m\$|#define $m|")
    script+=(-e "s|^#cmakedefine $m |#define $m |")
  done
  sed "${script[@]}" \
      -e 's|^#cmakedefine \([A-Za-z0-9_]*\).*$|/* #undef \1 */|' \
      -e 's|@RAJA_VERSION_MAJOR@|2025|' -e 's|@RAJA_VERSION_MINOR@|9|' \
      -e 's|@RAJA_VERSION_PATCHLEVEL@|1|' -e 's|@RAJA_DATA_ALIGN@|64|' \
      -e 's|@RAJA_FP@|RAJA_USE_DOUBLE|' -e 's|@RAJA_PTR@|RAJA_USE_BARE_PTR|' \
      -e 's|@camp_VERSION_MAJOR@|2025|' -e 's|@camp_VERSION_MINOR@|9|' \
      -e 's|@camp_VERSION_PATCH@|2|' \
      "$in" > "$out"
}
gen "$REF/include/RAJA/config.hpp.in"            "$OUT/RAJA/config.hpp" "$on_raja"
gen "$REF/tpl/camp/include/camp/config.in.hpp"   "$OUT/camp/config.hpp" "$on_camp"
if grep -n '@[A-Za-z_]*@' "$OUT/RAJA/config.hpp" "$OUT/camp/config.hpp"; then
  echo "unsubstituted @VAR@ left in generated config" >&2; exit 1
fi
