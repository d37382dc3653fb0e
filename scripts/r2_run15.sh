#!/bin/bash
set -u
for rep in 1 2; do for v in g2 g4 g4_mb7 g4_mb6; do echo -n "$v c2: "; python scripts/k1_time.py --lib kubegpu_b200/lib/variants/libkgpu_$v.so --config c2 --variants 5 --reps 10 | cut -c20-80; echo -n "$v c3: "; python scripts/k1_time.py --lib kubegpu_b200/lib/variants/libkgpu_$v.so --config c3 --variants 5 --reps 6 | cut -c20-80; done; done
