#!/bin/bash
O=gpurun_out/s14; mkdir -p $O
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -o $O/net_full -f python scripts/cnn_launches.py > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O
