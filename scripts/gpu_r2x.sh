#!/bin/bash
# per-CTA fixed cost vs per-query-block cost of the attention backward: 144 CTAs (one round), 1..8 query blocks each
mkdir -p gpurun_out
for L in 128 256 512 1024; do
  LK=1024 B200_FLASH_TAILSPLIT=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fl_$L.csv python scripts/one_flash.py $L 9 > /dev/null 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/fl_$L.csv")) if len(r) > 5 and r[0].isdigit()]
from collections import defaultdict
d = defaultdict(list)
for r in rows: d[r[4][:40]].append(float(r[-1]))
print("L=$L", {k: round(sum(v[-2:]) / 2 / 1000, 1) for k, v in d.items() if "flash" in k or "f32_to" in k})
PY
done
