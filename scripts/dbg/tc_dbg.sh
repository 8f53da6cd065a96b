for d in 0 1 2 3 4 8 15; do echo "== GAT_TC_DEBUG=$d"; GAT_TC_DEBUG=$d timeout 100 python scripts/dbg/tc_bench.py 2>&1 | grep -E '"K": (264|32|1024), "P": 1' | cut -c1-140; done
