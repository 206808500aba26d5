#!/bin/bash
python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -15
OFP_K6CC_WARP=1 python -m pytest tests/test_gpu_model.py -x -q -k cccnn 2>&1 | tail -3
