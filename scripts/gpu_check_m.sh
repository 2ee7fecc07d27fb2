mkdir -p gpurun_out
timeout 600 python scripts/flake_deferred.py 40 2>&1 | tail -30
