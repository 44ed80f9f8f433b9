#!/bin/bash
# A/B timing of two builds of libsemic_b200.so on ONE box, alternating (power-cap clock drift hits both alike).
#   tools/ab_builds.sh <prev.so> <k> [rounds]      -> prints the mean ms of tools/ab.py for each build, per round
prev=$1; k=${2:-3}; rounds=${3:-2}
for r in $(seq 1 $rounds); do
  echo -n "round $r prev: "; SEMIC_B200_LIB=$prev python tools/ab.py 24x0 $k 1 6 | tail -1
  echo -n "round $r new:  "; python tools/ab.py 24x0 $k 1 6 | tail -1
done
