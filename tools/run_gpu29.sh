#!/bin/bash
timeout 120 python tools/gpu_probe.py conv 2>&1 | grep -E "FAIL|PASS|Error|error|Traceback|vqb\]" | head -40
