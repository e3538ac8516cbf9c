#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2d_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.txt 2>&1
tail -4 gpurun_out/r2d_pytest.txt; tail -2 gpurun_out/r2d_smoke.txt
