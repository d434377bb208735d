#!/bin/bash
# Times the reference's main.cpp built over its own headers (CPU) and over the drop-in headers (engine) on the same
# synthetic input files.  Usage: tools/time_main.sh [users]
set -u
REPO="$(cd "$(dirname "$0")/.." && pwd)"
U="${1:-6000}"
W="$(mktemp -d)"
python "$REPO/tools/make_main_inputs.py" "$W" --users "$U" > /dev/null
cd "$W"
for b in ref crx; do
    s=$(date +%s%N)
    CRX_FAKE_SEED=5 "$REPO/oracle/_ref/recommendation_$b" -d ./tweets.tsv -o "./out_$b.txt" > /dev/null
    e=$(date +%s%N)
    echo "$b: $(( (e - s) / 1000000 )) ms wall; stage times (ms): $(grep 'Execution Time' out_$b.txt | sed 's/Execution Time: //' | tr '\n' ' ')"
done
echo "differing lines: $(diff <(grep -v 'Execution Time' out_ref.txt) <(grep -v 'Execution Time' out_crx.txt) | grep -c '^[<>]') of $(grep -c . out_ref.txt)"
rm -rf "$W"
