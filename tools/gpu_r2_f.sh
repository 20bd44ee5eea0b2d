#!/bin/bash
# round 2, call F: GPU-tail tests, application profile after, 2001^2 sweep
set -x
O=gpurun_out/r02
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_f.log 2>&1; echo "pytest rc=$?" >> $O/pytest_f.log
tail -30 $O/pytest_f.log
timeout 600 python tools/app_profile.py 1000003 > $O/app_profile_after.txt 2>&1; head -40 $O/app_profile_after.txt | cut -c1-170
