python tools/time_single.py 2>&1 | tail -4 | cut -c1-40
RVDD_BRIDGE_LIB=rvdd-release_b200/lib/libBridge_red.so python tools/time_single.py 2>&1 | tail -4 | cut -c1-40
