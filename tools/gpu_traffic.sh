for nm in "reference:k_(fwd|adj)4:dg4" "closed_form:k_(fwd|adj)_cf2:cf2"; do
  IFS=: read num rx tag <<< "$nm"
  timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"$rx" -s 2 -c 2 -o gpurun_out/r2_traffic_$tag -f python tools/prof_step.py cfg5 $num 8 3 > gpurun_out/r2_traffic_$tag.log 2>&1
  tail -1 gpurun_out/r2_traffic_$tag.log
done
