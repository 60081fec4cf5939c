#!/bin/bash
timeout 300 python tools/gpu_probe.py wgbench 2>&1 | grep -E "BENCH|--|Error|error|Traceback"
