#!/bin/bash
# bring-up build with tracing compiled in (the shipped library is rebuilt afterwards without it)
make -C mirror_nerf_b200/csrc clean > /dev/null; make -C mirror_nerf_b200/csrc -j8 EXTRA=-DMNRF_TC_TRACE > /dev/null 2>&1
python tools/tc_trace.py tc3 > gpurun_out/trace_tc3.log 2>&1
python tools/tc_trace.py tc1 > gpurun_out/trace_tc1.log 2>&1
head -4 gpurun_out/trace_tc3.log
