#!/bin/sh
# Builds the CPU oracle (test infrastructure) into oracle/_build/libbpl_oracle.so.
# -ffp-contract=off: the fp32 restatement must round after every operation.
set -e
cd "$(dirname "$0")"
mkdir -p _build
OMP=""
if echo 'int main(void){return 0;}' | gcc -fopenmp -x c - -o _build/.omp_probe 2>/dev/null; then OMP="-fopenmp"; fi
rm -f _build/.omp_probe
gcc -O2 -std=c99 -fPIC -shared -ffp-contract=off -fno-fast-math -Wall -Wextra -Wno-unused-function $OMP \
    bpl_oracle.c -o _build/libbpl_oracle.so -lm
echo "built oracle/_build/libbpl_oracle.so (openmp: ${OMP:-no})"
