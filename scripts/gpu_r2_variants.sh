#!/bin/bash
set -u
for v in v00 v01 v11 v21 st1; do
  echo "== variant $v"; LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_$v.so timeout 120 python scripts/quick_time.py 2>&1 | tail -2
done
