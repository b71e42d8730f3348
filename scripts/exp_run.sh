# A/B tuning builds: per-op ms of the c3 bench for every library under nvdiffrast_b200/variants (plus the default build)
run() {
  B200DR_LIB=$2 timeout -k 5 100 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $3 > gpurun_out/b_var.json 2>gpurun_out/b_var.err
  python -c "
import json; d=json.load(open('gpurun_out/b_var.json')); o=d['ops']; print('$1', round(d['ms_per_step'],3), 'e2e', d.get('e2e',{}).get('ms_per_step'), ' '.join('%s=%.3f'%(k.replace('rasterize','ras').replace('interpolate','ip').replace('texture','tex').replace('antialias','aa'),v['ms']) for k,v in o.items() if v['ms']>0.1))" || tail -5 gpurun_out/b_var.err
}
run base "" --no-e2e
for f in $(ls nvdiffrast_b200/variants/*.so 2>/dev/null); do n=$(basename $f .so); run ${n#libb200dr_} $PWD/$f --no-e2e; done
