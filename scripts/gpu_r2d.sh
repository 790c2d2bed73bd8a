#!/bin/bash
# round 2, visit D: droplet mode on texture-layout images: parity of the particle / checkpoint / golden tests + the droplet bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "particle or droplet or checkpoint or golden or host_driver or errors" 2>&1 | tail -15 > gpurun_out/pytest_drops.log
cat gpurun_out/pytest_drops.log
timeout 600 python scripts/particle_bench.py 60 2>&1 | tee gpurun_out/drops_bench.log
