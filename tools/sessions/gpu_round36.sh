#!/bin/bash
# session 36 (short): counters in their own lines + pool up to 2^25 as defaults: full GPU suite, smoke, steady-state figure
( timeout 300 python -m pytest tests -m gpu -q ) 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | grep "smoke"
QB_NO_BATCH=1 QB_SCENES=cornell-box QB_SPP=128 python tools/quick_bench.py ploc8 2>&1
