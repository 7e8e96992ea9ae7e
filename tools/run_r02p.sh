python -m pytest tests/test_gpu_large.py -q -x -s -k "lod" 2>&1 | grep "LOD\|passed\|failed" | tee gpurun_out/r02_lod.txt
