timeout 600 python tools/time_groups.py 2>&1 | tail -9
