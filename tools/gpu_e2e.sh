#!/bin/bash
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/e2e_pageable.py 50000000
COUPE_B200_NO_STAGING=1 timeout 300 python tools/e2e_pageable.py 50000000
