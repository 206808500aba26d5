#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputest_e1.log 2>&1; tail -30 gpurun_out/r02_gputest_e1.log
