#!/bin/bash
timeout 120 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
