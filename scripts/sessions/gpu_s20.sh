#!/bin/bash
O=gpurun_out/s20; mkdir -p $O
XA_CONV_IDX_HINT=0 timeout 300 python scripts/cnn_bench.py 2>&1 | tail -1 > $O/nohint.txt
XA_CONV_IDX_HINT=1 timeout 300 python scripts/cnn_bench.py 2>&1 | tail -1 > $O/hint.txt
cat $O/nohint.txt $O/hint.txt
