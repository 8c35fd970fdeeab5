mkdir -p gpurun_out
python tools/prof_stage.py hwc4 && \
ncu --set full --clock-control none --import-source on -k regex:warp_hwc4 -s 2 -c 1 -f -o gpurun_out/warp_hwc4_r02 python tools/prof_stage.py hwc4 > gpurun_out/ncu_warp4.log 2>&1
echo "rc $?"; tail -3 gpurun_out/ncu_warp4.log
