#!/bin/bash
for o in 2 3 4; do echo "== occ $o"; UEGAN_WGRAD_OCC=$o python scripts/layer_bench.py wgrad 2>&1 | grep -E "dec3|dec4|dec5|D.p|total"; done
