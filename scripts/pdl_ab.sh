#!/bin/bash
# Dev tool: A/B of the PCG iteration with / without programmatic dependent launch, then the GPU tests with it on.
O=gpurun_out
for k in "pdl=0" "pdl=1" "pdl=0" "pdl=1"; do echo "== $k"; KNOBS=$k timeout 100 python scripts/order_sweep.py 7:32 2>&1 | grep "^N="; done | tee $O/p_ab2.log
timeout 200 python -m pytest tests -m gpu -x -q > $O/p_pytest2.log 2>&1; echo "pytest(pdl=1) rc=$?"; tail -2 $O/p_pytest2.log
