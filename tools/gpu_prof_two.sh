#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pt_extend_q -s 2 -c 1 -f -o gpurun_out/r2d_prof_c4_extend_two python tools/prof_once.py c4two > gpurun_out/r2d_prof.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2d_prof.log
