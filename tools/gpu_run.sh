mkdir -p gpurun_out
python tools/prof_stage.py gauss > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gauss_tile -s 0 -c 3 -o gpurun_out/gauss_r02 python tools/prof_stage.py gauss > gpurun_out/ncu_gauss.log 2>&1
echo ncu_rc=$?
