mkdir -p gpurun_out
for v in _wcp_b4 _wcp_b1 _wcp_nod; do
  RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge$v.so python tools/time_wc.py 2>&1 | tail -1 | tee -a gpurun_out/wcp2.txt
done
RVDD_WC_PLANES=0 RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_wcp_b4.so python tools/time_wc.py 2>&1 | tail -1 | tee -a gpurun_out/wcp2.txt
