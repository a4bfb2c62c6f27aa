#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- builds the UNMODIFIED reference C programs into oracle/_ref/.
#
# The reference sources are compiled from where they lie (/root/reference, read-only);
# nothing is copied into this repo.  The only change is to capacity #defines
# (MAX_COEF_NUMBER, MAX_MIXTURE_NUMBER, ...), rewritten on the fly by `sed` and piped
# straight into gcc's stdin, because the stock limits (D<=9, M<=3) are below
# BASELINE.json's configs.  The arithmetic is untouched.
#
# Outputs (git-ignored, but they DO travel to the GPU box):
#   oracle/_ref/hmm_fs_<tag>, rec_fs_<tag>          CLI binaries (CPU baseline, KATs)
#   oracle/_ref/libref_train_<tag>.so, libref_test_<tag>.so   function-level oracle
#   oracle/_ref/variants.txt                        the capacity limits per tag
# Usage: oracle/build_ref.sh [reference_root]
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
TSRC="$REF/train/source/hmm-fs/hmm_continuous_fs.c"
RSRC="$REF/test/source/recognition-fs/recognition_continuous_fs.c"
if [ ! -f "$TSRC" ] || [ ! -f "$RSRC" ]; then
  echo "build_ref: reference sources not found under $REF (fine on the GPU box: prebuilt files are used)"
  exit 0
fi
mkdir -p "$OUT"
CC="${CC:-gcc}"
CFLAGS="-O2 -w"
: > "$OUT/variants.txt"

patch_defs() {  # coef mix params words number_words
  sed -E \
    -e "s/^#define MAX_COEF_NUMBER[ \t]+[0-9]+/#define MAX_COEF_NUMBER $1/" \
    -e "s/^#define MAX_MIXTURE_NUMBER[ \t]+[0-9]+/#define MAX_MIXTURE_NUMBER $2/" \
    -e "s/^#define MAX_PARAMETERS_NUMBER[ \t]+[0-9]+/#define MAX_PARAMETERS_NUMBER $3/" \
    -e "s/^#define MAX_WORDS_NUMBER[ \t]+[0-9]+/#define MAX_WORDS_NUMBER $4/" \
    -e "s/^#define NUMBER_WORDS[ \t]+[0-9]+/#define NUMBER_WORDS $5/"
}

build_variant() {  # tag coef mix params words number_words
  local tag="$1" coef="$2" mix="$3" par="$4" words="$5" nw="$6"
  patch_defs "$coef" "$mix" "$par" "$words" "$nw" < "$TSRC" | $CC $CFLAGS -x c - -o "$OUT/hmm_fs_$tag" -lm
  patch_defs "$coef" "$mix" "$par" "$words" "$nw" < "$RSRC" | $CC $CFLAGS -x c - -o "$OUT/rec_fs_$tag" -lm
  patch_defs "$coef" "$mix" "$par" "$words" "$nw" < "$TSRC" | $CC $CFLAGS -fPIC -shared -Dmain=ref_train_main -x c - -o "$OUT/libref_train_$tag.so" -lm
  patch_defs "$coef" "$mix" "$par" "$words" "$nw" < "$RSRC" | $CC $CFLAGS -fPIC -shared -Dmain=ref_test_main -x c - -o "$OUT/libref_test_$tag.so" -lm
  echo "$tag MAX_COEF_NUMBER=$coef MAX_MIXTURE_NUMBER=$mix MAX_PARAMETERS_NUMBER=$par MAX_WORDS_NUMBER=$words NUMBER_WORDS=$nw" >> "$OUT/variants.txt"
}

# stock limits of the trainer are D<=9, M<=3, P<=6; the recogniser's are D<=16, M<=5.
# "stock" keeps the trainer's numbers on both sides (ctypes strides must agree) and NUMBER_WORDS=13
# (the shipped 13-word fixture set).
build_variant stock   9   3 1   50 13
build_variant d39m16  39  16 1 1000 10
build_variant d39m128 39 128 1 2000 10
# two feature streams (param_number = 2) at the stock sizes: pins the multi-stream restatement (tests/golden/synth_p2.npz)
build_variant p2      9   3 2   50 2
echo "build_ref: built $(ls "$OUT" | wc -l) files in $OUT"
