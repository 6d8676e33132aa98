#!/bin/bash
D=$PWD/surface_vision_transformers_b200
for v in "" bwdv1 attnr1 gemmr1; do
  if [ -z "$v" ]; then unset SVIT_LIB; echo "== current"; else export SVIT_LIB=$D/libsvit_b200_$v.so; echo "== $v"; fi
  python -m pytest tests/test_gpu_model.py -q -s -k "bookkeeping" 2>&1 | grep -E "un-synchron|passed|failed" | grep -v print
done
unset SVIT_LIB
SVIT_NO_FUSE_LN=1 python -m pytest tests/test_gpu_model.py -q -s -k "bookkeeping" 2>&1 | grep -E "un-synchron|passed|failed" | grep -v print
