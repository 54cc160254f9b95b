#!/bin/bash
timeout 200 python -m pytest tests/test_b200_parity.py -x -q -m gpu -k "resident or warptile-integrate or multi_pair" 2>&1 | tail -2
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
