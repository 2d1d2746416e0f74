#!/bin/bash
O=gpurun_out/s35; mkdir -p $O
for i in 1 2 3; do for pair in 1 0; do
  XA_GEMM_PAIR=$pair timeout 300 python scripts/cnn_bench.py 2>&1 | grep "native plan" | sed "s/^/pair=$pair /" >> $O/cnn_ab.log
done; done
cat $O/cnn_ab.log
