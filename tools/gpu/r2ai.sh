#!/bin/bash
# round 2, call AI: the whole GPU suite and smoke() on the tree with the detect stage, the new selection kernel and the
# upstream drivers
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ai_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2ai_smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2ai_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2ai_pytest.log
