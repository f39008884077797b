#!/bin/bash
# static SASS size per kernel of libmsim.so (I-cache budget check)
set -e
D=$(mktemp -d)
cd "$D"
cuobjdump -xelf all /root/repo/manusim_b200/libmsim.so >/dev/null 2>&1
nvdisasm -c msim_kernel.sm_100a.cubin 2>/dev/null | grep -E "^\.text\.|^\s+/\*[0-9a-f]{4,}\*/" | awk '/^\.text/{if(name)print n*16/1024" KB "name; name=$1; n=0; next} {n++} END{print n*16/1024" KB "name}'
