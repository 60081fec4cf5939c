#!/bin/bash
timeout 120 python tools/gpu_probe.py shift 2>&1 | grep -E "SHIFT|Error|error|Traceback" | head -60
