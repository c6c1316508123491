#!/bin/sh
# Builds the C part of the parity oracle (test infrastructure only; never used by the product path).
set -e
cd "$(dirname "$0")"
mkdir -p _build
gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -o _build/libpt3d_cpu.so pt3d_cpu.c geom_cpu.c -lm
echo "built oracle/_build/libpt3d_cpu.so"
