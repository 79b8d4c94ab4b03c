#!/bin/bash
# build_variant.sh NAME "-DFLAGS..." : an experimental build of the library next to the default one
# (fjsph_b200/lib/var_NAME.so; select it with FJSPH_B200_LIB=...). Exploration helper.
set -e
cd "$(dirname "$0")/../fjsph_b200/csrc"
make -s -j8 OUT=../lib/var_$1.so OBJ=../lib/obj_$1 EXTRA="$2" ../lib/var_$1.so > /dev/null
grep -h -A2 "Compiling entry" ../lib/obj_$1/sweeps.ptxas.log | grep -o "k_[a-z0-9_]*I[Lb01E]*\|k_prestepE\|Used [0-9]* registers\|[0-9]* bytes spill stores" | paste - - - - | grep "force\|surf\|prestep" | awk '{print "  ",$1,$4,$5,$6,$7}'
