#!/bin/bash
# round 2, session ze: GPU test-suite at HEAD incl. the tests against the interpreted-Fortran golden vectors
TAG=r2ze
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s -k "golden or interpreted or regint" > gpurun_out/pytest_new_$TAG.log 2>&1; echo "new tests rc $?"; tail -6 gpurun_out/pytest_new_$TAG.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_$TAG.log
