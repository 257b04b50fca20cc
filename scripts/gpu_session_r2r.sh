#!/bin/bash
# round 2, session r: pair-kernel variants (quick sphere test, 64-term chains, unroll 4, CTA -> i-tile mapping)
mkdir -p gpurun_out
timeout 1200 python scripts/variant_probe2.py gpurun_out/variant_probe_r2r.json it1b4t it1b4tq it1b4tf it1b4tu it1b4tqu it1b4t+x it1b4tq+x > gpurun_out/variant_probe_r2r.txt 2>&1
cat gpurun_out/variant_probe_r2r.txt
