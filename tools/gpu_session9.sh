#!/bin/bash
mkdir -p gpurun_out
PYTHONPATH=. python tools/profile_e2e.py 1e6 > gpurun_out/profile_e2e.txt 2>&1
head -70 gpurun_out/profile_e2e.txt
