for d in 0 1 2 4 8 9 3 7 15; do echo "== DBG=$d"; TSD_TAIL_Y_DBG=$d python tools/tail_bench.py 256 2>&1 | grep "co=3 64x64" | sed 's/max abs err.*|//'; done | tee gpurun_out/tail_dbg.txt
