#!/bin/bash
for i in 1 2 3; do timeout 300 python scripts/fit_step_times.py 6 2>&1 | tail -7; done
