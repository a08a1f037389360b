#!/bin/bash
# rebuild the in-tree library (stale objects only)
cd /root/repo && python -c "
from pfs_neural_net_b200 import _abi
_abi.build_library(verbose=False)" 2>&1 | grep -v "^$" | grep -v "seg_rows" | grep -v "\^" | grep -v Remark | tail -${1:-15}
