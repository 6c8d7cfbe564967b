#!/bin/bash
python scripts/layer_bench.py fprop "G.dec4" 2>&1 | tail -2
UEGAN_OCC2_MINST=2 python scripts/layer_bench.py fprop "G.dec4" 2>&1 | tail -2
