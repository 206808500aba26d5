#!/bin/bash
./scripts/ubench/ffma2 | tee gpurun_out/r02_ubench_ffma2.txt
