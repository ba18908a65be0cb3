#!/bin/bash
# one full-size tensor-core conv (tools/tc_time.py) under the kernel's experiment switches: where does a stage's time go?
run() { echo "== $*"; env "$@" python tools/tc_time.py 2>&1 | grep split=1; }
run AG3D_TC_SPLIT_OCC=2 AG3D_TC_DEBUG=35        # barriers only
run AG3D_TC_SPLIT_OCC=2 AG3D_TC_DEBUG=163       # + no proxy fence
run AG3D_TC_SPLIT_OCC=2 AG3D_TC_DEBUG=291       # + 4 arrivals instead of 128
run AG3D_TC_SPLIT_OCC=2 AG3D_TC_DEBUG=419       # + both
run AG3D_TC_SPLIT_OCC=2 AG3D_TC_DEBUG=99        # barriers only, no epilogue
run AG3D_TC_SPLIT_OCC=2 AG3D_TC_DEBUG=128       # full work, no proxy fence (wrong results possible)
