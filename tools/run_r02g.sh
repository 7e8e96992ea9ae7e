T=${1:-r02g}
python -m pytest tests/test_gpu_parity.py -x -q -k "batch" 2>&1 | tail -3
HORIZONATOR_TRACE_HOST=1 python tools/batch_sweep.py --reps 10 --batches 16,64,256 --out gpurun_out/${T}_sweep.jsonl "" "SETS=1" "SETS=3" "SETS=4" "SETS=4,LANES=8" "SETS=2,LANES=32" "SETS=2,LANES=8" "GRID_SCALE_BATCH=100" "GRID_SCALE_BATCH=400" "SETS=4,GRID_SCALE_BATCH=400" "SETS=3,GRID_SCALE_BATCH=300" "BANDS_BATCH=16:48:110" "BANDS_BATCH=12:32:80:160" 2>&1 | tail -16
