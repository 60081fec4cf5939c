#!/bin/bash
timeout 300 python tools/gpu_probe.py conv 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head -30
timeout 300 python tools/gpu_probe.py conv2 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head -30
timeout 300 python tools/gpu_probe.py up 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head -30
timeout 300 python tools/gpu_probe.py stats 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head -30
timeout 300 python tools/gpu_probe.py halobench 2>&1 | grep -E "BENCH|Error|error|Traceback"
timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -1
