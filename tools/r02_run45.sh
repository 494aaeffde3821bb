#!/bin/bash
timeout -k 10 300 python tools/pipeline_breakdown.py 2>&1 | grep "^rep\|Error\|error" | head -12
