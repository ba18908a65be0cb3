#!/bin/bash
# one full-size tensor-core conv (tools/tc_time.py) under the kernel's experiment switches: where does a stage's time go?
run() { echo "== $*"; env "$@" python tools/tc_time.py 2>&1 | grep split=1; }
run AG3D_TC_DEBUG=0
run AG3D_TC_DEBUG=1            # no MMAs (commits only)
run AG3D_TC_DEBUG=4            # one product instead of three
run AG3D_TC_DEBUG=2            # no gather copies
run AG3D_TC_DEBUG=3            # neither
run AG3D_TC_DEBUG=64           # no epilogue
run AG3D_TC_NA=2
run AG3D_TC_T=1
run AG3D_TC_ISSUERS=1
