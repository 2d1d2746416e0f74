#!/bin/bash
# full ncu capture (with source) of the Dense kernel as the network runs it: FC forward and FC data gradient
O=gpurun_out/s36; mkdir -p $O
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -f -o $O/fc_full python scripts/cnn_launches.py > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
