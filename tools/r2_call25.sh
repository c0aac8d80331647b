#!/bin/bash
timeout 600 python -m pytest "tests/test_network_gpu.py" -m gpu -x -q -k "xresnet50" 2>&1 | grep -v "^$" | tail -45 | cut -c1-250
