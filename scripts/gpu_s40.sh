#!/bin/bash
O=gpurun_out/s40; mkdir -p $O
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_flat_kernel -f -o $O/conv_full python scripts/cnn_launches.py > $O/ncu.log 2>&1
tail -2 $O/ncu.log; ls -la $O
