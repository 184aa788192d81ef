#!/bin/bash
# Build-container tool (needs /root/reference): compiles UNMODIFIED test programs of the reference
# (test/fnft__poly/*.c, test/fnft__akns_fscatter/*.c, ...; CMakeLists.txt:170-199 builds each of them as one
# executable linked against libfnft) and links them against the drop-in library libfnft_b200.so through its
# SONAME link fnft_b200/lib/libfnft.so.  The executables land in tests/reftests_bin/ (git-ignored; they travel
# to the GPU box with the snapshot, where tests/test_reference_programs.py runs them).  Sources are read where
# they lie; nothing is copied into the repository.
#   scripts/link_reference_tests.sh            build the default list
set -u
cd "$(dirname "$0")/.."
REF=${REF:-/root/reference}
[ -d "$REF/test" ] || { echo "no reference tree at $REF"; exit 0; }
make -j8 > /dev/null || exit 1
make -C oracle _ref/fnft_config.h > /dev/null
OUT=tests/reftests_bin
mkdir -p $OUT
rm -f $OUT/*_test* $OUT/LINKED.txt $OUT/NOT_LINKED.txt
INC="-Ioracle/_ref -I$REF/include -I$REF/include/private -I$REF/include/3rd_party/kiss_fft"
ok=0; bad=0
for src in $REF/test/fnft__poly/fnft__poly_fmult*.c $REF/test/fnft__poly/fnft__poly_chirpz_test.c \
           $REF/test/fnft__poly/fnft__poly_eval_test.c $REF/test/fnft__poly/fnft__poly_roots_fasteigen_test.c \
           $REF/test/fnft__akns_fscatter/*.c $REF/test/fnft__nse_scatter/fnft__nse_scatter_bound_states_test_bo.c \
           $REF/test/fnft_version_test.c $REF/test/fnft__nse_finvscatter/*.c $REF/test/fnft__poly/fnft__poly_specfact_test*.c \
           $REF/test/fnft_nsev_inverse/*.c $REF/test/fnft_nsev_inverse/*/*.c "$@"; do
  name=$(basename "$src" .c)
  if gcc -std=gnu99 -O1 -w $INC "$src" -Lfnft_b200/lib -lfnft -lm -Wl,-rpath,'$ORIGIN/../../fnft_b200/lib' \
         -o $OUT/$name 2> $OUT/$name.linkerr; then
    echo $name >> $OUT/LINKED.txt; rm -f $OUT/$name.linkerr; ok=$((ok+1))
  else
    echo "$name: $(grep -o 'undefined reference to `[^'"'"']*' $OUT/$name.linkerr | sort -u | sed 's/undefined reference to `//' | tr '\n' ' ')" >> $OUT/NOT_LINKED.txt
    rm -f $OUT/$name.linkerr $OUT/$name; bad=$((bad+1))
  fi
done
echo "linked $ok programs against libfnft_b200.so, $bad not linked"
[ -f $OUT/NOT_LINKED.txt ] && cat $OUT/NOT_LINKED.txt
exit 0
