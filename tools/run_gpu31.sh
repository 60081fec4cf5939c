#!/bin/bash
timeout 200 python tools/gpu_probe.py conv 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head
timeout 300 python tools/gpu_probe.py halobench 2>&1 | grep -E "BENCH|Error|error|Traceback" | grep -E "128->128|Error|error" 
