#!/bin/sh
# Dev tool: DRAM bytes + duration of one train_tcp_kernel launch (cfg4, 480 members x 20 steps) per library variant.
for lib in "" W; do
  if [ -n "$lib" ]; then export NMB_LIB=multi_modal_normative_modeling_b200/lib/libnmb_$lib.so; else unset NMB_LIB; fi
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:train_tcp -s 3 -c 1 --csv python tools/run_tcp.py 24 20 5 2>/dev/null | grep -E "dram__bytes|gpu__time|hit_rate" | awk -F'","' '{print "'"variant=$lib"' ", $(NF-2), $(NF-1), $NF}'
done
