timeout 300 python tools/ab_basemul.py 2>&1 | grep -v DIFFER_IGNORED
