#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/trace_step.py --out $O/r02_ingraph_trace_train_ds_b8.txt > $O/c33_trace.log 2>&1; tail -2 $O/c33_trace.log; head -2 $O/r02_ingraph_trace_train_ds_b8.txt; wc -l $O/r02_ingraph_trace_train_ds_b8_streams.txt
