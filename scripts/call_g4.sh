#!/bin/bash
python -m pytest tests/test_gpu_detect.py tests/test_gpu_lag_locate.py -x -q 2>&1 | tail -25
