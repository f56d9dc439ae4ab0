#!/bin/bash
# round 2: the -m gpu suite and smoke() on the library the round ends with
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > gpurun_out/r02aj_pytest.log
cat gpurun_out/r02aj_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
