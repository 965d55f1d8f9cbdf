mkdir -p gpurun_out
for w in c2 c4; do
  if [ $w = c2 ]; then S=70; C=120; else S=100; C=220; fi
  timeout 1200 ncu --set full --clock-control none -s $S -c $C -o /tmp/ncu_full_$w -f python tools/launch_list.py $w > gpurun_out/r2_ncu_full_$w.log 2>&1; echo "ncu $w rc=$?"
  python tools/ncu_summary.py /tmp/ncu_full_$w.ncu-rep gpurun_out/r02_v7_ncu_full_$w.txt $w "one whole step (launch_list.py $w), final round-2 kernels" | tail -2
  ncu -i /tmp/ncu_full_$w.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > gpurun_out/r2_ncu_full_$w.csv.gz
done
ls -la gpurun_out | tail -8; du -sh gpurun_out
