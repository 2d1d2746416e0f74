#!/bin/bash
for i in 1 2; do timeout 300 python scripts/update_launches.py 2>&1 | tail -2; timeout 300 python scripts/update_launches.py generic-hooks 2>&1 | tail -2; done
