#!/bin/bash
timeout 900 python -m pytest tests/test_randaug_gpu.py tests/test_event_pipeline_gpu.py -q -x 2>&1 | grep -v Warning | tail -15
