mkdir -p gpurun_out/r3f
for v in lb8 lb10 lb12; do
  cp exp/lib_$v.so iskra_b200/libiskra_b200.so
  timeout 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:'k_mcc_collide' --launch-skip 8 -c 4 --csv --log-file gpurun_out/r3f/coll_$v.csv python bench.py --steps 4 --warmup 4 --no-cpu --no-e2e > gpurun_out/r3f/ncu_$v.log 2>&1
  grep -v "^==" gpurun_out/r3f/coll_$v.csv | awk -F'","' '{print "'$v'", $(NF-2), $NF}' | tail -8
done
