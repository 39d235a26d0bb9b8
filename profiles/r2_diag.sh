#!/bin/bash
python tests/diag_grad.py baseline 2 192 640 0 1234 waves > gpurun_out/diag1.log 2>&1
python tests/diag_grad.py fm 2 96 320 16 1235 waves > gpurun_out/diag2.log 2>&1
python tests/diag_grad.py fm 1 192 640 64 1238 waves > gpurun_out/diag3.log 2>&1
grep -h "cells to drop\|flips" gpurun_out/diag1.log gpurun_out/diag2.log gpurun_out/diag3.log | cut -c1-400
