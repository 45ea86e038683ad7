#!/bin/bash
# usage: tools/prof.sh report.ncu-rep [kernel-substring] [top-lines]  — per-function and per-line breakdown of one ncu capture (in-tree .so must be the profiled build)
cd "$(dirname "$0")/.."
REP=$1; KEY=${2:-k_env_step_lsILb0ELb1}; TOP=${3:-70}
SO=mujoco_rl_manipulate_unknown_objects_b200/libgripper_sim_b200.so
ncu -i $REP --page source --csv > /tmp/src.csv 2>/dev/null
cuobjdump -elf $SO | grep -E "^\s+0x[0-9a-f]+\s+0x[0-9a-f]+\s+0x[0-9a-f]+.*$KEY" | awk '{print $2, $3, $NF}' > /tmp/syms.txt
python tools/prof_by_func.py /tmp/syms.txt /tmp/src.csv
cuobjdump -xelf all $SO > /dev/null 2>&1
