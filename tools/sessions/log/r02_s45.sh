#!/bin/bash
# round 2, session 45: GPU suite and smoke of the last commit (ordered ray batches added)
( timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
