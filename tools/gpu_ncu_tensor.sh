#!/bin/bash
o=gpurun_out; tag=r02f
for shape in 8192,10240,1280 16384,2560,2560; do
  nm=${shape//,/_}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_i8_persist -s 2 -c 1 \
    -o $o/${tag}_persist_$nm -f python tools/ncu_persist.py $shape > $o/${tag}_persist_$nm.log 2>&1
  ncu -i $o/${tag}_persist_$nm.ncu-rep --page raw --csv > $o/${tag}_persist_${nm}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $o/${tag}_persist_${nm}_raw.csv
done
for shape in 64,16,1280,1280 32,32,640,640 64,64,320,320; do
  nm=${shape//,/_}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_i8_persist -s 2 -c 1 \
    -o $o/${tag}_conv_$nm -f python tools/ncu_conv.py $shape > $o/${tag}_conv_$nm.log 2>&1
  ncu -i $o/${tag}_conv_$nm.ncu-rep --page raw --csv > $o/${tag}_conv_${nm}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $o/${tag}_conv_${nm}_raw.csv | tail -1
done
rm -f $o/${tag}_persist_*.ncu-rep $o/${tag}_conv_*.ncu-rep
