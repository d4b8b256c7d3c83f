for f in 0 1; do echo "UNIMP_XB_FLAGS=$f"; UNIMP_XB_FLAGS=$f timeout 200 python tools/xblock_check.py bench 2>&1 | grep -E "XB bench"; done
