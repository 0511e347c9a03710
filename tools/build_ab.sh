#!/bin/bash
# A/B builds of libohp.so with one experimental feature compiled out each (build/ab/libohp_<name>.so, selected with OHP_LIB)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  rm -rf build/ab_obj_$name; mkdir -p build/ab_obj_$name
  make -s -C omegahpetsc_b200/csrc -j8 OBJDIR=../../build/ab_obj_$name LIBDIR=../../build/ab_lib_$name EXTRA="$flags"
  cp build/ab_lib_$name/libohp.so build/ab/libohp_$name.so
  rm -rf build/ab_obj_$name build/ab_lib_$name
done
ls -la build/ab
