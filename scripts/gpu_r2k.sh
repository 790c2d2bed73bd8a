#!/bin/bash
# round 2, queued thermal outflow in the droplet-mode tail: parity + timing
mkdir -p gpurun_out
for d in 3 0; do
HG_DROPS_VARIANT=$d timeout 600 python -m pytest tests/test_gpu_particles_slabs.py tests/test_ref_golden.py -m gpu -x -q 2>&1 | tail -3
HG_DROPS_VARIANT=$d timeout 300 python scripts/particle_bench.py 100 2>&1 | tee gpurun_out/q_drops_$d.log
done
