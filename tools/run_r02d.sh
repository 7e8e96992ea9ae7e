T=${1:-r02d}
ncu --set full --clock-control none --import-source on -k regex:"k_mesh|k_blocks|k_raster|k_tiles" -s 9 -c 9 -o gpurun_out/${T}_eye12km python tools/view_probe.py --ncu eye12km --reps 2 > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_big|k_raster" -s 6 -c 6 -o gpurun_out/${T}_zoom5 python tools/view_probe.py --ncu zoom5 --reps 2 > gpurun_out/${T}_ncu2.log 2>&1
ls -la gpurun_out/
