#!/bin/bash
# Same-box A/B of a kernel change (the pool's boxes differ by +-2 %, so only pairs measured in ONE gpurun call count):
# builds the committed sources (HEAD) into scratch_ab/old.so and leaves the working tree's build in place.  The
# scripts/gpu_r4e.sh / gpu_r4i.sh / gpu_r4k.sh runs then swap the two files under uegan_b200/ between bench runs.
# scratch_ab/ is git-ignored via *.so but travels to the GPU box; delete it afterwards (18 MiB per push).
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch_ab
python -m uegan_b200.build > /dev/null
cp uegan_b200/libuegan_sm100.so /tmp/uegan_new.so
git stash -q
python -m uegan_b200.build > /dev/null
cp uegan_b200/libuegan_sm100.so scratch_ab/old.so
git stash pop -q
cp /tmp/uegan_new.so uegan_b200/libuegan_sm100.so
echo "scratch_ab/old.so = HEAD, uegan_b200/libuegan_sm100.so = working tree"
