run() {
  B200DR_LIB=$2 timeout -k 5 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --batch $3 > gpurun_out/b_var.json 2>gpurun_out/b_var.err
  python -c "
import json; d=json.load(open('gpurun_out/b_var.json')); o=d['ops']; print('$1 batch $3', round(d['ms_per_step'],3), ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if 'tex_grad' in k.replace('texture','tex')))" || tail -5 gpurun_out/b_var.err
}
for b in 4 8 16 32; do
run base "" $b
for f in $(ls nvdiffrast_b200/variants/*.so 2>/dev/null); do n=$(basename $f .so); run ${n#libb200dr_} $PWD/$f $b; done
done
