#!/bin/bash
# Where the C2 forward spends its time: build the library with one role's work compiled out (FFNO_KO bitmask, see
# csrc/umma_pipelined.cu) and time the 24-layer forward (tools/pipe_sweep.py persist0).  The results are WRONG numbers
# by construction — only the times mean anything.  Builds run here or on the GPU box; timing needs the GPU.
#   1    every kernel returns after its prologue (launch + barrier / TMEM set-up + weight image copy)
#   2    transform kernels: no global stores          4  converters do not read / split / store      8  one MMA pass of 3
#   16   FF: no residual loads, no global stores      32 FF: one MMA pass of 3 (both GEMMs)            64 FF loaders: no global loads
#   128  mix: no global stores                        256 mix: one MMA pass of 3                       512 mix loaders: zero-fill
# usage: tools/knockout_sweep.sh build | run
set -e
KOS="1 2 4 8 16 32 64 128 256 512 14 112 896 1022"
if [ "$1" = build ]; then
  for ko in $KOS; do FFNO_BUILD_TAG=ko$ko FFNO_BUILD_DEFS="-DFFNO_KO=$ko" python -m fourierflow_b200.build --force; done
else
  for ko in $KOS; do
    echo -n "KO $ko: "
    FFNO_B200_LIB=fourierflow_b200/lib/libffno_b200_ko$ko.so timeout 100 python tools/pipe_sweep.py persist0 2>&1 | tail -1
  done
  echo -n "product: "; timeout 100 python tools/pipe_sweep.py persist0 2>&1 | tail -1
fi
