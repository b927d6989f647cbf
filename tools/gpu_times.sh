#!/bin/bash
# per-kernel durations of one training step under ncu (serialised, cold cache; relative numbers): bash tools/gpu_times.sh tag
OUT=gpurun_out/${1:-times}; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --kernel-name-base demangled -k regex:nerfca -s 7 -c 7 --csv --log-file $OUT/times.csv python tools/profile_step.py 1024 500 3 > $OUT/log.txt 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/times.csv")) if len(r)>10 and r[0].isdigit()]
cur=None
for r in rows:
    name=r[4].split('(')[0]
    if name!=cur: print(); print(name.ljust(28), end=' '); cur=name
    print(f"{r[-3].split('.')[0].replace('gpu__','').replace('sm__','')[:24]}={r[-1]} {r[-2]}", end='  ')
print()
PY
