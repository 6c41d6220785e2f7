#!/usr/bin/env bash
# Developer helper: builds tuning variants of the fit kernel (compile-time macros of sucre_b200/csrc/fit.cu) into
# variants/<name>.so so that ONE GPU call can time them all:
#     tools/fit_variants.sh base "ilp3 -DSUCRE_FIT_ILP=3" "c4_11_24 -DSUCRE_COST_BLOCK=4 -DSUCRE_COST_SEGMENT=11 -DSUCRE_COST_TILE=24"
#     gpurun -- 'python tools/quick_fit.py variants/*.so'
# The in-tree library is rebuilt with the default flags at the end.  (variants/ is not tracked: *.so is git-ignored.)
set -euo pipefail
root="$(cd "$(dirname "$0")/.." && pwd)"
cd "$root/sucre_b200/csrc"
mkdir -p "$root/variants"
for spec in "$@"; do
    name="${spec%% *}"
    flags=""
    [[ "$spec" == *" "* ]] && flags="${spec#* }"
    touch fit.cu
    make --no-print-directory EXTRA="$flags" >/dev/null
    cp ../libsucre_b200.so "$root/variants/$name.so"
    printf '%-24s %s | %s\n' "$name" "$flags" "$(grep -A2 'fit_kernelILi0ELi0ELb0' build/fit.ptxas.log | tail -1 | sed 's/ptxas info    : //')"
done
touch fit.cu
make --no-print-directory >/dev/null
echo "in-tree library rebuilt with default flags"
