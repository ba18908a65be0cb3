#!/bin/bash
# one full-size tensor-core conv (tools/tc_time.py) under the kernel's experiment switches
run() { echo "== $*"; env "$@" timeout 120 python tools/tc_time.py 2>&1 | grep -E "split=1|Error|error" | head -3; }
run AG3D_TC_GATHER=cpasync
run AG3D_TC_GATHER=tma
run AG3D_TC_GATHER=tma AG3D_TC_DEBUG=1          # no MMAs
run AG3D_TC_GATHER=tma AG3D_TC_DEBUG=2          # no gathers
run AG3D_TC_GATHER=tma AG3D_TC_DEBUG=3
