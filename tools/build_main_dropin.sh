#!/bin/bash
# Builds the reference's main.cpp twice:
#   oracle/_ref/recommendation_ref  = main.cpp + the reference's own headers            (the checker)
#   oracle/_ref/recommendation_crx  = main.cpp + include/crx/lib drop-in headers + libcrx.so (the product behind main.cpp)
# main.cpp includes "./lib/..." relative to itself, so the drop-in build compiles it from a scratch tree of symlinks:
# main.cpp and the non-hot-path files (arg_parser, tweet) point at the reference, the hot-path headers at
# include/crx/lib.  Nothing is copied; only the two binaries are kept.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${CRX_REF_DIR:-/root/reference}"
[ -d "$REF/lib" ] || { echo "build_main_dropin: $REF absent - keeping any prebuilt binaries"; exit 0; }
OUT="$REPO/oracle/_ref"
mkdir -p "$OUT"
T="$(mktemp -d)"
trap 'rm -rf "$T"' EXIT
FLAGS="-O2 -w -std=c++14 -ffp-contract=off"
# ---- reference build
g++ $FLAGS -I"$REF" "$REPO/oracle/main_ref_wrap.cpp" "$REF/lib/in_out/arg_parser.cpp" "$REF/lib/utils.cpp" \
    "$REF/lib/data_structures/tweet.cpp" -o "$OUT/recommendation_ref"
# ---- drop-in build
mkdir -p "$T/w/lib"
ln -s "$REPO/include/crx.h" "$T/crx.h"
ln -s "$REF/main.cpp" "$T/w/main.cpp"
ln -s "$REPO/include/crx/crx_shim.hpp" "$T/w/crx_shim.hpp"
ln -s "$REPO/oracle/main_ref_wrap.cpp" "$T/w/wrap.cpp"
cp -rs "$REPO/include/crx/lib/." "$T/w/lib/"
mkdir -p "$T/w/lib/in_out"
for f in in_out/arg_parser.cpp in_out/arg_parser.h data_structures/tweet.h data_structures/tweet.cpp; do
    ln -sf "$REF/lib/$f" "$T/w/lib/$f"
done
(cd "$T/w" && g++ $FLAGS -DCRX_DROPIN_BUILD wrap.cpp lib/in_out/arg_parser.cpp lib/data_structures/tweet.cpp \
    -L"$REPO/crypto_recommendation_b200" -lcrx -Wl,-rpath,'$ORIGIN/../../crypto_recommendation_b200' \
    -Wl,--allow-shlib-undefined -o "$OUT/recommendation_crx")
echo "built $OUT/recommendation_ref and $OUT/recommendation_crx"
