#!/bin/bash
# TEST INFRASTRUCTURE.  Stages the files of the UNMODIFIED reference that its model classes import (model/, hparams.py) into
# oracle/_ref/reference/ so that the gpu-marked patch() test can run the reference's own classes on the GPU box, where
# /root/reference does not exist.  oracle/_ref/ is git-ignored (nothing of the reference enters the history) but travels with
# the gpurun snapshot, like built .so files.  Usage: bash oracle/stage_reference.sh [/root/reference]
set -e
src=${1:-/root/reference}
dst=$(dirname "$0")/_ref/reference
rm -rf "$dst"; mkdir -p "$dst"
cp -r "$src/model" "$dst/model"
cp "$src/hparams.py" "$dst/hparams.py"
find "$dst" -name '__pycache__' -prune -exec rm -rf {} +
echo "staged $(find "$dst" -name '*.py' | wc -l) files under $dst"
