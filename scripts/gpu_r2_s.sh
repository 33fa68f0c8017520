#!/bin/bash
# round 2, call S (1 GPU): last-minute smoke of the final tree (communicator creation with the tuning table)
timeout 50 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2 | cut -c1-200
