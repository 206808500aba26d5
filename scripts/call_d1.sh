#!/bin/bash
scripts/k4_ab.sh > gpurun_out/r02_k4_ab.txt 2>&1; cat gpurun_out/r02_k4_ab.txt
