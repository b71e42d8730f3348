mkdir -p gpurun_out/ncu3
for k in texture_grad_fast; do
  timeout -k 5 150 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/ncu3/$k \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --batch 4 > gpurun_out/ncu3/$k.log 2>&1
done
